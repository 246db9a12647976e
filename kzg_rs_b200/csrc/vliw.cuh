// Cooperative pairing engine: a CTA executes the level-scheduled Fp programs of vliw_programs.cuh over a
// register file of Fp values in shared memory (instruction k of a level on thread k, a barrier per level).
// One thread runs Fp multiplications back to back at ~1 us each; the pairing check that closes every batch
// (reference src/pairings.rs:5-9 via src/kzg_proof.rs:436-441) is ~25 000 of them in sequence.  Here the 36
// independent dual products of an Fp12 multiplication run side by side, so the check is ~560 multiplication
// levels deep instead.  Host build: the same code runs the lanes one after the other (unit-tested on the CPU).
#pragma once
#include "pairing.cuh"
#include "vliw_programs.cuh"

namespace kzgb200 {
namespace vliw {

// program tables (global memory, or a shared-memory copy made by load_tables)
struct Tables {
    const uint16_t (*mul)[6];
    const uint32_t (*lin)[3];
    const uint16_t* term;
    const Level* level;
    const Program* prog;
};
struct Lanes {
    int tid, n;   // this thread's lane and the number of cooperating threads (host: 0, 1)
    Tables tab;
    KZG_HD void sync() const {
#ifdef __CUDA_ARCH__
        __syncthreads();
#endif
    }
};
KZG_HD Tables default_tables() {
#ifdef __CUDA_ARCH__
    return Tables{d_mul, d_lin, d_term, d_level, d_prog};
#else
    return Tables{h_mul, h_lin, h_term, h_level, h_prog};
#endif
}
// shared-memory image of the tables (instruction fetch becomes an LDS instead of a dependent global load)
struct SharedTables {
    uint16_t mul[kNumMul][6];
    uint32_t lin[kNumLin][3];
    uint16_t term[kNumTerm];
    Level level[kNumLevel];
    Program prog[kNumPrograms];
};
#ifdef __CUDACC__
__device__ __forceinline__ Tables load_tables(SharedTables* st, int tid, int n) {
    for (int i = tid; i < kNumMul * 6; i += n) (&st->mul[0][0])[i] = (&d_mul[0][0])[i];
    for (int i = tid; i < kNumLin * 3; i += n) (&st->lin[0][0])[i] = (&d_lin[0][0])[i];
    for (int i = tid; i < kNumTerm; i += n) st->term[i] = d_term[i];
    for (int i = tid; i < kNumLevel; i += n) st->level[i] = d_level[i];
    for (int i = tid; i < kNumPrograms; i += n) st->prog[i] = d_prog[i];
    __syncthreads();
    return Tables{st->mul, st->lin, st->term, st->level, st->prog};
}
#endif

KZG_HD void exec_mul(Fp* regs, const uint16_t* ins) {
    Fp a = regs[ins[1]], b = regs[ins[2]];
    if (ins[3] == 0xffff) {
        regs[ins[0]] = a.mul_inl(b);
    } else {
        Fp c = regs[ins[3]], d = regs[ins[4]];
        if (ins[5] & 1) c = Fp::zero().sub_inl(c);
        regs[ins[0]] = Fp::mul_dual_inl(a, b, c, d);
    }
}
// dst = sum of (+/-)(1|2) * src over up to 24 terms.  Everything is accumulated as a NON-NEGATIVE integer
// (a negative term contributes p - v), 12 limbs plus a small carry count, and reduced once at the end:
// quotient estimate from the top words, one multiply-subtract, at most three conditional subtractions.
KZG_HD void exec_lin(Fp* regs, const uint32_t* ins, const uint16_t* terms) {
    uint32_t acc[12];
#pragma unroll
    for (int i = 0; i < 12; i++) acc[i] = 0;
    uint32_t top = 0;                           // total = top * 2^384 + acc  <  48 p  <  2^387
    Fp p = Fp::modulus();
    const uint16_t* t = terms + ins[1];
    for (uint32_t k = 0; k < ins[2]; k++) {
        uint16_t e = t[k];
        Fp v = regs[e & 0x3fff];
        if (e & 0x4000) sub_n<12>(v.l, p.l, v.l);            // p - v  (v <= p, no borrow)
        top += add_n<12>(acc, acc, v.l);
        if (e & 0x8000) top += add_n<12>(acc, acc, v.l);     // doubled term
    }
    // q <= total / p, within 3 of it: top 64 bits of total over (top word of p) + 1
    uint64_t t64 = ((uint64_t)top << 32) | acc[11];
    uint32_t q = (uint32_t)(t64 / ((uint64_t)p.l[11] + 1));
    // total -= q * p   (q < 2^9)
    uint64_t carry = 0, bw = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        carry += (uint64_t)p.l[i] * q;
        uint64_t d = (uint64_t)acc[i] - (uint32_t)carry - bw;
        acc[i] = (uint32_t)d; bw = (d >> 32) & 1; carry >>= 32;
    }
    top -= (uint32_t)carry + (uint32_t)bw;
    // now 0 <= total < 4p
#pragma unroll
    for (int r = 0; r < 3; r++) {
        uint32_t tmp[12];
        uint32_t borrow = sub_n<12>(tmp, acc, p.l);
        if (top || !borrow) {
#pragma unroll
            for (int i = 0; i < 12; i++) acc[i] = tmp[i];
            top -= borrow;
        }
    }
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = acc[i];
    regs[ins[0]] = r;
}
// run one program; every cooperating thread must call it (barriers inside)
KZG_HD void run(int prog, Fp* regs, const Lanes& L) {
    const Program p = L.tab.prog[prog];
    for (int lv = p.first_level; lv < p.first_level + p.n_levels; lv++) {
        const Level lev = L.tab.level[lv];
        if (lev.kind == 1) { for (int k = L.tid; k < lev.count; k += L.n) exec_mul(regs, L.tab.mul[lev.first + k]); }
        else { for (int k = L.tid; k < lev.count; k += L.n) exec_lin(regs, L.tab.lin[lev.first + k], L.tab.term); }
        L.sync();
    }
}
// regs[dst .. dst+count) = regs[src ..)
KZG_HD void copy_regs(Fp* regs, int dst, int src, int count, const Lanes& L) {
    for (int k = L.tid; k < count * 12; k += L.n) regs[dst + k / 12].l[k % 12] = regs[src + k / 12].l[k % 12];
    L.sync();
}

// a^-1 for a != 0 by the binary extended Euclid algorithm on the raw limbs (variable time: inputs are public).
// Input and output in Montgomery form.  ~0.1 ms on one thread versus ~0.55 ms for a^(p-2).
KZG_NI Fp fp_inv_bingcd(const Fp& a_mont) {
    constexpr int N = 12;
    Fp p = Fp::modulus();
    uint32_t u[N], v[N], x1[N], x2[N], t[N];
    for (int i = 0; i < N; i++) { u[i] = a_mont.l[i]; v[i] = p.l[i]; x1[i] = 0; x2[i] = 0; }
    x1[0] = 1;
    if (a_mont.is_zero()) return a_mont;
    auto is_one = [](const uint32_t* w) { uint32_t o = w[0] ^ 1u; for (int i = 1; i < N; i++) o |= w[i]; return o == 0; };
    auto shr1 = [](uint32_t* w, uint32_t top) { for (int i = 0; i < N - 1; i++) w[i] = (w[i] >> 1) | (w[i + 1] << 31); w[N - 1] = (w[N - 1] >> 1) | (top << 31); };
    auto halve_mod = [&](uint32_t* w) {   // w/2 mod p
        if (w[0] & 1) { uint32_t c = add_n<N>(w, w, p.l); shr1(w, c); } else shr1(w, 0);
    };
    while (!is_one(u) && !is_one(v)) {
        while (!(u[0] & 1)) { shr1(u, 0); halve_mod(x1); }
        while (!(v[0] & 1)) { shr1(v, 0); halve_mod(x2); }
        uint32_t borrow = sub_n<N>(t, u, v);
        if (!borrow) {   // u >= v
            for (int i = 0; i < N; i++) u[i] = t[i];
            if (sub_n<N>(t, x1, x2)) add_n<N>(t, t, p.l);
            for (int i = 0; i < N; i++) x1[i] = t[i];
        } else {
            sub_n<N>(v, v, u);
            if (sub_n<N>(t, x2, x1)) add_n<N>(t, t, p.l);
            for (int i = 0; i < N; i++) x2[i] = t[i];
        }
    }
    Fp r;   // (aR)^-1 as a raw value -> a^-1 R needs two multiplications by R^2
    for (int i = 0; i < N; i++) r.l[i] = is_one(u) ? x1[i] : x2[i];
    Fp r2; for (int i = 0; i < N; i++) r2.l[i] = FpParams::r2(i);
    return (r * r2) * r2;
}

// Shared-memory register file: program registers [0, kMaxRegs) then saved Fp12 values.
constexpr int kSave0 = kMaxRegs;            // each save slot = 12 registers
constexpr int kNumSaves = 5;
constexpr int kTotalRegs = kMaxRegs + 12 * kNumSaves;

KZG_HD void load_lines(Fp* regs, const LineCoeffs* c1, const LineCoeffs* c2, int k, const Lanes& L) {
    // 6 Fp per line (A, B, C as Fp2) into regs[kRegLines + 6 j ..]
    for (int i = L.tid; i < 12 * 12; i += L.n) {
        int fe = i / 12, limb = i % 12, j = fe / 6, e = fe % 6;
        const LineCoeffs* src = j == 0 ? c1 : c2;
        if (!src) continue;
        const Fp2& f2 = e < 2 ? src[k].A : (e < 4 ? src[k].B : src[k].C);
        regs[kRegLines + fe].l[limb] = (e & 1) ? f2.c1.l[limb] : f2.c0.l[limb];
    }
    L.sync();
}
// F <- F^|x| conjugated (x < 0), base = save slot `base` (F is overwritten; G is used as the multiplier slot)
KZG_HD void exp_by_x_slot(Fp* regs, int base, const Lanes& L) {
    copy_regs(regs, kRegF, base, 12, L);
    for (int bit = 62; bit >= 0; bit--) {
        run(kProg_f12_sqr, regs, L);
        if ((KZG_BLS_X_ABS >> bit) & 1) { copy_regs(regs, kRegG, base, 12, L); run(kProg_f12_mul, regs, L); }
    }
    run(kProg_conj, regs, L);
}
// e(P1, Q1) e(P2, Q2) == 1 with the lines of Q1, Q2 precomputed (c1, c2).  All cooperating threads call it with
// the same arguments; returns the same verdict to all.  regs: kTotalRegs Fp values shared by the threads.
KZG_HD bool coop_pairing_product_is_one(Fp* regs, const G1Affine& P1, const LineCoeffs* c1, const G1Affine& P2, const LineCoeffs* c2,
                                        const Lanes& L) {
    bool live1 = !P1.inf, live2 = !P2.inf;
    if (!live1 && !live2) return true;
    // constants and the G1 arguments; with a single live pair it takes slot 0
    const G1Affine& Pa = live1 ? P1 : P2;
    const LineCoeffs* ca = live1 ? c1 : c2;
    const LineCoeffs* cb = (live1 && live2) ? c2 : nullptr;
    if (L.tid == 0) {
        const uint32_t g[10][12] = {KZG_FP_FROB6_1_C0_M, KZG_FP_FROB6_1_C1_M, KZG_FP_FROB6_2_C0_M, KZG_FP_FROB6_2_C1_M, KZG_FP_FROB6_3_C0_M,
                                    KZG_FP_FROB6_3_C1_M, KZG_FP_FROB6_4_C0_M, KZG_FP_FROB6_4_C1_M, KZG_FP_FROB6_5_C0_M, KZG_FP_FROB6_5_C1_M};
        for (int i = 0; i < 10; i++) regs[kRegConst + i] = fp_const(g[i]);
        regs[kRegP] = Pa.x; regs[kRegP + 1] = Pa.y;
        regs[kRegP + 2] = cb ? P2.x : Fp::zero(); regs[kRegP + 3] = cb ? P2.y : Fp::zero();
        for (int i = 0; i < 12; i++) regs[kRegF + i] = Fp::zero();
        regs[kRegF] = Fp::one();
    }
    L.sync();
    // Miller loop
    int k = 0;
    for (int bit = 62; bit >= 0; bit--) {
        load_lines(regs, ca, cb, k++, L);
        if (cb) run(kProg_sqr_lines, regs, L); else { run(kProg_f12_sqr, regs, L); run(kProg_line1, regs, L); }
        run(kProg_f12_mul, regs, L);
        if ((KZG_BLS_X_ABS >> bit) & 1) {
            load_lines(regs, ca, cb, k++, L);
            run(cb ? kProg_lines : kProg_line1, regs, L);
            run(kProg_f12_mul, regs, L);
        }
    }
    run(kProg_conj, regs, L);
    // final exponentiation, f^(3(p^12-1)/r):  easy part
    const int S0 = kSave0, S1 = kSave0 + 12, S2 = kSave0 + 24, S3 = kSave0 + 36, S4 = kSave0 + 48;
    copy_regs(regs, S0, kRegF, 12, L);                       // S0 = f0
    run(kProg_inv_prep, regs, L);
    if (L.tid == 0) regs[kRegH + 8] = fp_inv_bingcd(regs[kRegH + 8]);
    L.sync();
    run(kProg_inv_finish, regs, L);                          // F = f0^-1
    copy_regs(regs, kRegG, kRegF, 12, L);
    copy_regs(regs, kRegF, S0, 12, L);
    run(kProg_conj, regs, L);
    run(kProg_f12_mul, regs, L);                             // F = f0^(p^6-1)
    run(kProg_frob2, regs, L);                               // G = F^(p^2)
    run(kProg_f12_mul, regs, L);                             // F = f = f0^((p^6-1)(p^2+1))
    copy_regs(regs, S0, kRegF, 12, L);                       // S0 = f
    // hard part: (x-1)^2 (x+p)(x^2+p^2-1) + 3
    exp_by_x_slot(regs, S0, L);                              // F = f^x
    copy_regs(regs, kRegG, S0, 12, L); run(kProg_conj_g, regs, L); run(kProg_f12_mul, regs, L);     // f^(x-1)
    copy_regs(regs, S1, kRegF, 12, L);
    exp_by_x_slot(regs, S1, L);
    copy_regs(regs, kRegG, S1, 12, L); run(kProg_conj_g, regs, L); run(kProg_f12_mul, regs, L);     // a = f^((x-1)^2)
    copy_regs(regs, S1, kRegF, 12, L);                       // S1 = a
    exp_by_x_slot(regs, S1, L);                              // a^x
    copy_regs(regs, S2, kRegF, 12, L);
    copy_regs(regs, kRegF, S1, 12, L); run(kProg_frob, regs, L);                                     // G = a^p
    copy_regs(regs, kRegF, S2, 12, L); run(kProg_f12_mul, regs, L);                                  // b = a^(x+p)
    copy_regs(regs, S2, kRegF, 12, L);                       // S2 = b
    exp_by_x_slot(regs, S2, L);
    copy_regs(regs, S3, kRegF, 12, L);
    exp_by_x_slot(regs, S3, L);                              // b^(x^2)
    copy_regs(regs, S4, kRegF, 12, L);
    copy_regs(regs, kRegF, S2, 12, L); run(kProg_frob2, regs, L);                                    // G = b^(p^2)
    copy_regs(regs, kRegF, S4, 12, L); run(kProg_f12_mul, regs, L);
    copy_regs(regs, kRegG, S2, 12, L); run(kProg_conj_g, regs, L); run(kProg_f12_mul, regs, L);      // c = b^(x^2+p^2-1)
    copy_regs(regs, S4, kRegF, 12, L);
    copy_regs(regs, kRegF, S0, 12, L); run(kProg_f12_sqr, regs, L);
    copy_regs(regs, kRegG, S0, 12, L); run(kProg_f12_mul, regs, L);                                  // f^3
    copy_regs(regs, kRegG, S4, 12, L); run(kProg_f12_mul, regs, L);                                  // c f^3
    // == 1 ?
    bool ok = regs[kRegF] == Fp::one();
    for (int i = 1; i < 12; i++) ok = ok && regs[kRegF + i].is_zero();
    L.sync();
    return ok;
}

}  // namespace vliw
}  // namespace kzgb200
