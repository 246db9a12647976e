// Cooperative pairing engine, second generation: the level-scheduled programs of vliw29_programs.cuh over a shared-memory
// register file of 29-bit-limb values (fp29.cuh).  Same idea as vliw.cuh (instruction k of a level on thread k, a barrier per
// level; reference src/pairings.rs:5-9 via src/kzg_proof.rs:436-441), different arithmetic:
//   MUL  dst = (a b +- c d) / 2^406     392 carry-free IMAD.WIDE + 196 for the reduction (was 432 chained IMAD.WIDE.X)
//   LIN  dst = K p + sum +-(1|2) src    14 IMAD.WIDE per term into signed 64-bit columns, one carry pass; NO modular reduction
//                                       (headroom: 2^406 = 2^25.3 p) unless the generator flagged the sum, in which case
//                                       floor-estimate(v / p) p is subtracted first (value then below 6 p)
// Measured on a lone 64-thread CTA (round 2): a MUL level 1.96 -> ~0.8 us, a LIN level 1.42 -> ~0.3 us.
// Host build: the same code runs the lanes one after the other (tools/hosttest/vliw29_host.cu).
#pragma once
#include "vliw.cuh"
#include "fp29.cuh"
#include "vliw29_programs.cuh"

namespace kzgb200 {
namespace vliw29 {
using f29::F29;

struct Tables {
    const uint32_t (*mul)[4];
    const uint32_t (*lin)[4];
    const uint16_t* term;
    const Level* level;
    const Program* prog;
};
struct Lanes {
    int tid, n;                   // this thread's lane and the number of cooperating threads (host: 0, 1)
    Tables tab;
    long long* ticks = nullptr;   // optional: per-section clock64() stamps (profiling aid)
    KZG_HD void tick(int i) const {
#ifdef __CUDA_ARCH__
        if (ticks && tid == 0) ticks[i] = clock64();
#endif
    }
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
struct SharedTables {
    uint32_t mul[kNumMul][4];
    uint32_t lin[kNumLin][4];
    uint16_t term[kNumTerm];
    Level level[kNumLevel];
    Program prog[kNumPrograms];
};
#ifdef __CUDACC__
__device__ __forceinline__ Tables load_tables(SharedTables* st, int tid, int n) {
    const Tables src = default_tables();
    for (int i = tid; i < kNumMul * 4; i += n) (&st->mul[0][0])[i] = (&src.mul[0][0])[i];
    for (int i = tid; i < kNumLin * 4; i += n) (&st->lin[0][0])[i] = (&src.lin[0][0])[i];
    for (int i = tid; i < kNumTerm; i += n) st->term[i] = src.term[i];
    for (int i = tid; i < kNumLevel; i += n) st->level[i] = src.level[i];
    for (int i = tid; i < kNumPrograms; i += n) st->prog[i] = src.prog[i];
    __syncthreads();
    return Tables{st->mul, st->lin, st->term, st->level, st->prog};
}
#endif

// Frobenius coefficients xi^(k(p-1)/6), k = 1..5, (c0, c1) each, in the engine's representation
#define KZG29_FROB_TABLE {KZG29_FROB6_1_C0, KZG29_FROB6_1_C1, KZG29_FROB6_2_C0, KZG29_FROB6_2_C1, KZG29_FROB6_3_C0, KZG29_FROB6_3_C1, \
                          KZG29_FROB6_4_C0, KZG29_FROB6_4_C1, KZG29_FROB6_5_C0, KZG29_FROB6_5_C1}
#ifdef __CUDACC__
static __device__ const uint32_t d_frob29[10][14] = KZG29_FROB_TABLE;
#endif
KZG_HD F29 frob_const(int i) {
    F29 r;
#ifdef __CUDA_ARCH__
    for (int k = 0; k < 14; k++) r.l[k] = d_frob29[i][k];
#else
    static const uint32_t h[10][14] = KZG29_FROB_TABLE;
    for (int k = 0; k < 14; k++) r.l[k] = h[i][k];
#endif
    r.l[14] = r.l[15] = 0;
    return r;
}

// MUL row: {dst | a << 16, b | c << 16, d | flags << 16, kx}; flags bit 0 = dual product, bit 1 = the second product is subtracted
KZG_HD void exec_mul(F29* regs, const uint32_t* ins) {
    const uint32_t w0 = ins[0], w1 = ins[1], w2 = ins[2];
    const bool dual = (w2 >> 16) & 1u, neg = (w2 >> 17) & 1u;
    const F29 a = regs[w0 >> 16], b = regs[w1 & 0xffffu];
    const F29 c = regs[dual ? (w1 >> 16) : (w0 >> 16)], d = regs[dual ? (w2 & 0xffffu) : (w1 & 0xffffu)];
    F29 r;
    f29::mont_mul29(r.l, a.l, b.l, c.l, d.l, dual, neg, ins[3]);
    r.l[14] = 0; r.l[15] = 0;
    regs[w0 & 0xffffu] = r;
}
// LIN row: {dst, first term, term count, K | reduce << 31}; a term = reg | neg << 14 | dbl << 15
KZG_HD void exec_lin(F29* regs, const uint32_t* ins, const uint16_t* terms) {
    uint64_t t[f29::kN];   // signed column sums (two's complement)
#pragma unroll
    for (int i = 0; i < f29::kN; i++) t[i] = 0;
    const uint16_t* tt = terms + ins[1];
    const uint32_t count = ins[2];
#pragma unroll 1
    for (uint32_t k = 0; k < count; k++) {
        const uint32_t e = tt[k];
        const F29 v = regs[e & 0x3fffu];
        int32_t coef = (int32_t)((e >> 15) & 1u) + 1;
        if (e & 0x4000u) coef = -coef;
#pragma unroll
        for (int i = 0; i < f29::kN; i++) f29::madw_s(t[i], (int32_t)v.l[i], coef);
    }
    const int32_t K = (int32_t)(ins[3] & 0x7fffffffu);
#pragma unroll
    for (int i = 0; i < f29::kN; i++) f29::madw_s(t[i], (int32_t)f29::p29(i), K);
    if (ins[3] >> 31) {
        // v < 2^20 p: quotient estimate from the two top columns in float (relative error 2^-21 -> off by less than 1), minus 2
        const float vf = (float)(int64_t)t[13] * 536870912.0f + (float)(int64_t)t[12];
        const float pinv = 1.0f / ((float)f29::p29(13) * 536870912.0f + (float)f29::p29(12) + 1.0f);
        int32_t q = (int32_t)(vf * pinv) - 2;
        if (q < 0) q = 0;
#pragma unroll
        for (int i = 0; i < f29::kN; i++) f29::madw_s(t[i], (int32_t)f29::p29(i), -q);
    }
    F29 r;
#pragma unroll
    for (int i = 0; i < f29::kN - 1; i++) {
        r.l[i] = (uint32_t)t[i] & f29::kMask;
        t[i + 1] += (uint64_t)((int64_t)t[i] >> f29::kW);
    }
    r.l[13] = (uint32_t)t[13];
    r.l[14] = 0; r.l[15] = 0;
    regs[ins[0]] = r;
}
// run one program; every cooperating thread must call it (barriers inside)
KZG_NI void run(int prog, F29* regs, const Lanes& L) {
    const Program p = L.tab.prog[prog];
    for (int lv = p.first_level; lv < p.first_level + p.n_levels; lv++) {
        const Level lev = L.tab.level[lv];
#ifdef __CUDA_ARCH__
        long long c0 = L.ticks ? clock64() : 0;
#endif
        if (lev.kind == 1) { for (int k = L.tid; k < lev.count; k += L.n) exec_mul(regs, L.tab.mul[lev.first + k]); }
        else { for (int k = L.tid; k < lev.count; k += L.n) exec_lin(regs, L.tab.lin[lev.first + k], L.tab.term); }
#ifdef __CUDA_ARCH__
        long long c1 = L.ticks ? clock64() : 0;
#endif
        L.sync();
#ifdef __CUDA_ARCH__
        if (L.ticks && L.tid == 0) {   // profiling aid: body / barrier-wait cycles of lane 0 per level kind
            long long c2 = clock64();
            int b = lev.kind == 1 ? 10 : 8;
            L.ticks[b] += c1 - c0; L.ticks[b + 1] += c2 - c1; L.ticks[b == 10 ? 13 : 12] += 1;
        }
#endif
    }
}
// regs[dst .. dst+count) = regs[src ..)   (16-byte words)
KZG_NI void copy_regs(F29* regs, int dst, int src, int count, const Lanes& L) {
    uint4* d = reinterpret_cast<uint4*>(regs + dst);
    const uint4* s = reinterpret_cast<const uint4*>(regs + src);
    for (int j = L.tid; j < count * 4; j += L.n) d[j] = s[j];
    L.sync();
}

// Line tables in the engine's representation: per fixed G2 point, per Miller step, (A, B, C) as 6 values (setup: k_setup.cu)
struct LineCoeffs29 { F29 v[6]; };
KZG_NI void load_lines(F29* regs, const LineCoeffs29* c1, const LineCoeffs29* c2, int k, const Lanes& L) {
    uint4* d = reinterpret_cast<uint4*>(regs + kRegLines);
    for (int q = L.tid; q < 2 * 6 * 4; q += L.n) {
        const LineCoeffs29* src = q < 24 ? c1 : c2;
        if (!src) continue;
        d[q] = reinterpret_cast<const uint4*>(src[k].v)[q < 24 ? q : q - 24];
    }
    L.sync();
}

// v^-1 (both in the engine's representation); v != 0 mod p
KZG_NI F29 inv29(const F29& v) {
    Fp x = f29::canonical(v), y;
    vliw::fp_inv_raw(y.l, x.l);
    return f29::from_raw(y);
}

constexpr int kSave0 = kMaxRegs;            // each save slot = 12 registers
constexpr int kNumSaves = 5;
constexpr int kTotalRegs = kMaxRegs + 12 * kNumSaves;

KZG_HD void sqr_times(F29* regs, int k, const Lanes& L) {
    for (; k > 0; k--) run(kProg_cyc_sqr1, regs, L);
}
// F <- conj(base^|x|) for base in the cyclotomic subgroup (save slot `base`); G is the multiplier slot
KZG_HD void exp_by_x_slot(F29* regs, int base, const Lanes& L) {
    copy_regs(regs, kRegF, base, 12, L);
    int pending = 0;
    for (int bit = 62; bit >= 0; bit--) {
        pending++;
        if ((KZG_BLS_X_ABS >> bit) & 1) {
            sqr_times(regs, pending, L);
            pending = 0;
            copy_regs(regs, kRegG, base, 12, L);
            run(kProg_f12_mul, regs, L);
        }
    }
    sqr_times(regs, pending, L);
    run(kProg_conj, regs, L);
}
// e(P1, Q1) e(P2, Q2) == 1 with the lines of Q1, Q2 precomputed (c1, c2).  All cooperating threads call it with the same
// arguments; returns the same verdict to all.  regs: kTotalRegs values shared by the threads.
KZG_HD bool coop_pairing_product_is_one(F29* regs, const G1Affine& P1, const LineCoeffs29* c1, const G1Affine& P2, const LineCoeffs29* c2,
                                        const Lanes& L) {
    const bool live1 = !P1.inf, live2 = !P2.inf;
    if (!live1 && !live2) return true;
    const G1Affine& Pa = live1 ? P1 : P2;       // with a single live pair it takes slot 0
    const LineCoeffs29* ca = live1 ? c1 : c2;
    const LineCoeffs29* cb = (live1 && live2) ? c2 : nullptr;
    for (int i = L.tid; i < 26; i += L.n) {
        if (i < 10) regs[kRegConst + i] = frob_const(i);
        else if (i < 14) {
            const int j = i - 10;
            if (j < 2) regs[kRegP + j] = f29::from_fp(j == 0 ? Pa.x : Pa.y);
            else regs[kRegP + j] = cb ? f29::from_fp(j == 2 ? P2.x : P2.y) : f29::f29_zero();
        } else regs[kRegF + (i - 14)] = i == 14 ? f29::f29_one() : f29::f29_zero();
    }
    L.sync();
    L.tick(1);
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
    L.tick(2);
    // final exponentiation, f^(3(p^12-1)/r):  easy part
    const int S0 = kSave0, S1 = kSave0 + 12, S2 = kSave0 + 24, S3 = kSave0 + 36, S4 = kSave0 + 48;
    copy_regs(regs, S0, kRegF, 12, L);                       // S0 = f0
    run(kProg_inv_prep, regs, L);
    if (L.tid == 0) regs[kRegH + 8] = inv29(regs[kRegH + 8]);
    L.sync();
    run(kProg_inv_finish, regs, L);                          // F = f0^-1
    copy_regs(regs, kRegG, kRegF, 12, L);
    copy_regs(regs, kRegF, S0, 12, L);
    run(kProg_conj, regs, L);
    run(kProg_f12_mul, regs, L);                             // F = f0^(p^6-1)
    run(kProg_frob2, regs, L);                               // G = F^(p^2)
    run(kProg_f12_mul, regs, L);                             // F = f = f0^((p^6-1)(p^2+1))
    copy_regs(regs, S0, kRegF, 12, L);                       // S0 = f
    L.tick(3);
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
    L.tick(4);
    // == 1 ?  (each coefficient canonicalised by its own thread; the verdict words land in the G slot)
    for (int i = L.tid; i < 12; i += L.n) {
        Fp x = f29::canonical(regs[kRegF + i]);
        bool good = i == 0 ? (x.l[0] == 1u) : (x.l[0] == 0u);
        for (int w = 1; w < 12; w++) good = good && x.l[w] == 0u;
        regs[kRegG + i].l[15] = good ? 1u : 0u;
    }
    L.sync();
    bool ok = true;
    for (int i = 0; i < 12; i++) ok = ok && regs[kRegG + i].l[15] == 1u;
    L.sync();
    return ok;
}

}  // namespace vliw29
}  // namespace kzgb200
