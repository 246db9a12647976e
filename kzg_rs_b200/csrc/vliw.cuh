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
    const uint16_t (*mul)[19];
    const uint32_t (*lin)[3];
    const uint16_t* term;
    const Level* level;
    const Program* prog;
    bool plain;               // every multiplication operand is a single register (the `lat` programs)
};
struct Lanes {
    int tid, n;   // this thread's lane and the number of cooperating threads (host: 0, 1)
    Tables tab;
    long long* ticks = nullptr;   // optional: per-section clock64() stamps (profiling aid)
    bool warp = false;            // the cooperating threads are the 32 lanes of ONE warp (many independent checks per CTA)
    int groups = 1;               // independent register files run in lockstep through the same program: a level of k instructions
    int stride = 0;               // becomes groups * k independent ones spread over the threads.  Group g's file starts `stride` 32-bit
                                  // WORDS after group g-1's; an odd stride puts the same register of 32 groups on 32 different banks
                                  // (consecutive threads run the SAME instruction for consecutive groups: uniform control flow,
                                  // broadcast table reads, conflict-free register-file accesses)
    Fp* shared = nullptr;         // multi-group form: the registers every check reads alike -- line coefficients and Frobenius constants,
                                  // 22 values -- live ONCE here; the group files hold the rest, compacted (see remap_multi)
    KZG_HD Fp* file(Fp* regs, int g) const { return reinterpret_cast<Fp*>(reinterpret_cast<uint32_t*>(regs) + (size_t)g * stride); }
    KZG_HD void tick(int i) const {
#ifdef __CUDA_ARCH__
        if (ticks && tid == 0) ticks[i] = clock64();
#endif
    }
    KZG_HD void sync() const {
#ifdef __CUDA_ARCH__
        if (warp) __syncwarp(); else __syncthreads();
#endif
    }
};
// two program sets (tools/gen_vliw.py): `lat` for ONE check on one CTA, `thr` for many checks in lockstep
KZG_HD Tables default_tables() {
#ifdef __CUDA_ARCH__
    return Tables{lat::d_mul, lat::d_lin, lat::d_term, lat::d_level, lat::d_prog, true};
#else
    return Tables{lat::h_mul, lat::h_lin, lat::h_term, lat::h_level, lat::h_prog, true};
#endif
}
KZG_HD Tables throughput_tables() {
#ifdef __CUDA_ARCH__
    return Tables{thr::d_mul, thr::d_lin, thr::d_term, thr::d_level, thr::d_prog, false};
#else
    return Tables{thr::h_mul, thr::h_lin, thr::h_term, thr::h_level, thr::h_prog, false};
#endif
}
// shared-memory image of the tables (instruction fetch becomes an LDS instead of a dependent global load)
struct SharedTables {
    uint16_t mul[kNumMulMax][19];
    uint32_t lin[kNumLinMax][3];
    uint16_t term[kNumTermMax];
    Level level[kNumLevelMax];
    Program prog[kNumPrograms];
};
// Multi-group register numbering: logical registers [kRegLines, kRegP) are the shared ones (flag 0x2000 | offset), the logical
// registers above them move down by their count.  Applied to the table copy the multi-group kernels execute from.
constexpr int kNumSharedRegs = kRegP - kRegLines;
constexpr uint32_t kSharedFlag = 0x2000u;
KZG_HD uint32_t remap_multi(uint32_t idx) {
    if (idx == 0xffffu) return idx;
    if (idx >= (uint32_t)kRegLines && idx < (uint32_t)kRegP) return kSharedFlag | (idx - kRegLines);
    return idx >= (uint32_t)kRegP ? idx - kNumSharedRegs : idx;
}
KZG_HD int compact_multi(int idx, bool multi) { return (multi && idx >= kRegP) ? idx - kNumSharedRegs : idx; }
KZG_HD const Fp& rreg(const Fp* regs, const Fp* shared, uint32_t idx) { return (idx & kSharedFlag) ? shared[idx & 0x1fffu] : regs[idx]; }
// in-place remap of a throughput-table copy (mul: dst + 16 slots; lin: dst; term: register bits)
KZG_HD void remap_tables_multi(SharedTables* st, int tid, int n) {
    for (int i = tid; i < thr::kNumMul * 17; i += n) { uint16_t& v = st->mul[i / 17][i % 17]; v = (uint16_t)remap_multi(v); }
    for (int i = tid; i < thr::kNumLin; i += n) st->lin[i][0] = remap_multi(st->lin[i][0]);
    for (int i = tid; i < thr::kNumTerm; i += n) { uint16_t e = st->term[i]; st->term[i] = (uint16_t)((e & 0xc000u) | remap_multi(e & 0x3fffu)); }
}
#ifdef __CUDACC__
__device__ __forceinline__ Tables load_tables(SharedTables* st, int tid, int n, bool throughput = false) {
    const Tables src = throughput ? throughput_tables() : default_tables();
    const int nmul = throughput ? thr::kNumMul : lat::kNumMul, nlin = throughput ? thr::kNumLin : lat::kNumLin;
    const int nterm = throughput ? thr::kNumTerm : lat::kNumTerm, nlevel = throughput ? thr::kNumLevel : lat::kNumLevel;
    for (int i = tid; i < nmul * 19; i += n) (&st->mul[0][0])[i] = (&src.mul[0][0])[i];
    for (int i = tid; i < nlin * 3; i += n) (&st->lin[0][0])[i] = (&src.lin[0][0])[i];
    for (int i = tid; i < nterm; i += n) st->term[i] = src.term[i];
    for (int i = tid; i < nlevel; i += n) st->level[i] = src.level[i];
    for (int i = tid; i < kNumPrograms; i += n) st->prog[i] = src.prog[i];
    __syncthreads();
    return Tables{st->mul, st->lin, st->term, st->level, st->prog, !throughput};
}
#endif

// One multiplication operand = up to four registers with signs, summed by the multiplying thread itself (no LIN level, no
// barrier): lazily, sum(pos) + #neg * p - sum(neg) < 4p, then one conditional subtraction of 2p for three or four terms, so the
// operand is < 2p -- the Montgomery products tolerate that: a b + c d < 8 p^2 < R p (R = 2^384 = 9.8 p), result < p as usual.
KZG_HD Fp mul_operand(const Fp* regs, const uint16_t* slot, uint32_t signs, const Fp* shared = nullptr) {
    Fp x = rreg(regs, shared, slot[0]);                     // the first term is always positive
    if (slot[1] == 0xffff) return x;
    const Fp p = Fp::modulus();
    int n = 1;
#pragma unroll 1
    for (int j = 1; j < 4 && slot[j] != 0xffff; j++, n++) {
        Fp y = rreg(regs, shared, slot[j]);
        if ((signs >> j) & 1u) { add_n<12>(x.l, x.l, p.l); sub_n<12>(x.l, x.l, y.l); }
        else add_n<12>(x.l, x.l, y.l);
    }
    if (n > 2) {
        Fp p2, t;
        add_n<12>(p2.l, p.l, p.l);
        uint32_t borrow = sub_n<12>(t.l, x.l, p2.l);
#pragma unroll
        for (int i = 0; i < 12; i++) x.l[i] = borrow ? x.l[i] : t.l[i];
    }
    return x;
}
// ins: dst, 4 operands x 4 slots, sign bits (4 per operand), negate-second-product flag
KZG_HD void exec_mul(Fp* regs, const uint16_t* ins, bool plain, const Fp* shared = nullptr) {
    if (plain) {            // latency programs: every operand is one register
        Fp a = regs[ins[1]], b = regs[ins[5]];
        if (ins[9] == 0xffff) { regs[ins[0]] = a.mul_inl(b); return; }
        Fp c = regs[ins[9]], d = regs[ins[13]];
        if (ins[18] & 1) c = Fp::zero().sub_inl(c);
        regs[ins[0]] = Fp::mul_dual_inl(a, b, c, d);
        return;
    }
    const uint32_t signs = ins[17];
    Fp a = mul_operand(regs, ins + 1, signs, shared), b = mul_operand(regs, ins + 5, signs >> 4, shared);
    if (ins[9] == 0xffff) {
        regs[ins[0]] = a.mul_inl(b);
    } else {
        Fp c = mul_operand(regs, ins + 9, signs >> 8, shared), d = mul_operand(regs, ins + 13, signs >> 12, shared);
        if (ins[18] & 1) {                                  // - c d = (2p - c) d with 2p - c in (0, 2p]
            Fp p2, p = Fp::modulus();
            add_n<12>(p2.l, p.l, p.l);
            sub_n<12>(c.l, p2.l, c.l);
        }
        regs[ins[0]] = Fp::mul_dual_inl(a, b, c, d);
    }
}
// dst = sum of (+/-)(1|2) * src over up to 24 terms.  The lanes of a warp execute different LIN instructions, so the
// per-term code is branch-free: the term is shifted left by its "double" bit (funnel shifts), XOR-ed with its sign mask and
// added with the sign as carry-in to ONE signed 14-limb accumulator (two's complement subtraction through the adder).
// The signed total S in (-48p, 48p) is then made positive (D = S + 64p) and reduced once: quotient estimate from the top
// words, one multiply-subtract, the candidates D - p, D - 2p, D - 3p side by side.
KZG_HD void exec_lin(Fp* regs, const uint32_t* ins, const uint16_t* terms, const Fp* shared = nullptr) {
    uint32_t acc[12];
#pragma unroll
    for (int i = 0; i < 12; i++) acc[i] = 0;
    uint32_t hi = 0;                            // signed overflow word(s): total = (int32)hi * 2^384 + acc
    const uint16_t* t = terms + ins[1];
    for (uint32_t k = 0; k < ins[2]; k++) {
        uint16_t e = t[k];
        const Fp& v = rreg(regs, shared, e & 0x3fffu);
        uint32_t sh = (e >> 15) & 1u, m = (e & 0x4000) ? 0xffffffffu : 0u;
        uint32_t wv[12], prev = 0;
#pragma unroll
        for (int i = 0; i < 12; i++) {
            uint32_t cur = v.l[i];
#ifdef __CUDA_ARCH__
            wv[i] = __funnelshift_l(prev, cur, sh);      // one SHF: (cur << sh) | (prev >> (32 - sh)), sh in {0, 1}
#else
            wv[i] = (cur << sh) | ((prev >> 31) & sh);
#endif
            prev = cur;
        }
        uint32_t c = add_signed_12(acc, wv, m);
        hi += (uint32_t)c + (((prev >> 31) & sh) ^ m);
    }
    Fp p = Fp::modulus();
    // D = S + 64p  in (16p, 112p)
    uint32_t p64[12];
#pragma unroll
    for (int i = 0; i < 12; i++) p64[i] = (p.l[i] << 6) | (i ? (p.l[i - 1] >> 26) : 0u);
    uint32_t top = hi + (p.l[11] >> 26) + add_n<12>(acc, acc, p64);
    // quotient estimate q <= D / p (at most 3 short) from the top 64 bits of D and the top word of p, in float
    // (D < 112 p, so q < 128: the 2^-23 relative error of the float product is far below one unit; minus one for safety)
    uint64_t t64 = ((uint64_t)top << 32) | acc[11];
    float qf = (float)t64 * (1.0f / ((float)p.l[11] + 2.0f));
    uint32_t q = qf >= 1.0f ? (uint32_t)qf - 1u : 0u;
    uint64_t carry = 0, bw = 0;
#pragma unroll
    for (int i = 0; i < 12; i++) {
        carry += (uint64_t)p.l[i] * q;
        uint64_t d = (uint64_t)acc[i] - (uint32_t)carry - bw;
        acc[i] = (uint32_t)d; bw = (d >> 32) & 1; carry >>= 32;
    }
    top -= (uint32_t)carry + (uint32_t)bw;
    // now 0 <= D < 4p: the three candidates D - p, D - 2p, D - 3p are formed side by side (independent borrow chains)
    uint32_t p2[12], p3[12], r1[12], r2[12], r3[12];
    add_n<12>(p2, p.l, p.l);
    add_n<12>(p3, p2, p.l);
    uint32_t b1 = sub_n<12>(r1, acc, p.l), b2 = sub_n<12>(r2, acc, p2), b3 = sub_n<12>(r3, acc, p3);
    // top is 0 or 1 here (D < 4p < 2^384 * 0.41, so top == 0 in fact); candidate k is valid iff D >= k p
    bool ok1 = top || !b1, ok2 = top || !b2, ok3 = top || !b3;
#pragma unroll
    for (int i = 0; i < 12; i++) acc[i] = ok3 ? r3[i] : (ok2 ? r2[i] : (ok1 ? r1[i] : acc[i]));
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
#ifdef __CUDA_ARCH__
        long long c0 = L.ticks ? clock64() : 0;
#endif
        if (L.groups == 1) {
            if (lev.kind == 1) { for (int k = L.tid; k < lev.count; k += L.n) exec_mul(regs, L.tab.mul[lev.first + k], L.tab.plain); }
            else { for (int k = L.tid; k < lev.count; k += L.n) exec_lin(regs, L.tab.lin[lev.first + k], L.tab.term); }
        } else {
            const int total = lev.count * L.groups;
            if (lev.kind == 1) { for (int j = L.tid; j < total; j += L.n) { int k = j / L.groups, g = j - k * L.groups; exec_mul(L.file(regs, g), L.tab.mul[lev.first + k], L.tab.plain, L.shared); } }
            else { for (int j = L.tid; j < total; j += L.n) { int k = j / L.groups, g = j - k * L.groups; exec_lin(L.file(regs, g), L.tab.lin[lev.first + k], L.tab.term, L.shared); } }
        }
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
// regs[dst .. dst+count) = regs[src ..)
KZG_HD void copy_regs(Fp* regs, int dst, int src, int count, const Lanes& L) {
    dst = compact_multi(dst, L.shared != nullptr); src = compact_multi(src, L.shared != nullptr);     // logical -> group-file numbering
    const int per = count * 12;
    for (int j = L.tid; j < per * L.groups; j += L.n) {
        int k = j / L.groups, g = j - k * L.groups;
        Fp* r = L.file(regs, g);
        r[dst + k / 12].l[k % 12] = r[src + k / 12].l[k % 12];
    }
    L.sync();
}

// out = in^-1 mod p for a raw integer 0 < in < p, by the binary extended Euclid algorithm on the limbs (variable time: the
// inputs are public).  ~0.1 ms on one thread versus ~0.55 ms for a^(p-2).
KZG_NI void fp_inv_raw(uint32_t* out, const uint32_t* in) {
    constexpr int N = 12;
    Fp p = Fp::modulus();
    uint32_t u[N], v[N], x1[N], x2[N], t[N];
    for (int i = 0; i < N; i++) { u[i] = in[i]; v[i] = p.l[i]; x1[i] = 0; x2[i] = 0; }
    x1[0] = 1;
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
    for (int i = 0; i < N; i++) out[i] = is_one(u) ? x1[i] : x2[i];
}
// a^-1 for a != 0; input and output in Montgomery form
KZG_NI Fp fp_inv_bingcd(const Fp& a_mont) {
    if (a_mont.is_zero()) return a_mont;
    Fp r;   // (aR)^-1 as a raw value -> a^-1 R needs two multiplications by R^2
    fp_inv_raw(r.l, a_mont.l);
    Fp r2; for (int i = 0; i < 12; i++) r2.l[i] = FpParams::r2(i);
    return (r * r2) * r2;
}

// Shared-memory register file: program registers [0, kMaxRegs) then saved Fp12 values.
constexpr int kMaxRegs = lat::kMaxRegs;
constexpr int kSave0 = kMaxRegs;            // each save slot = 12 registers
constexpr int kNumSaves = 5;
constexpr int kTotalRegs = kMaxRegs + 12 * kNumSaves;
// register file of the lockstep (multi-group) form, throughput programs: logical numbering up to kSave0Thr + 60 with the 22
// shared registers (lines, constants) taken out of the group files
constexpr int kNumSavesThr = 3;      // the final exponentiation needs three saved Fp12 values at a time (see coop_pairing_multi)
constexpr int kSave0Thr = thr::kMaxRegs, kTotalRegsThr = thr::kMaxRegs - kNumSharedRegs + 12 * kNumSavesThr;

KZG_HD void load_lines(Fp* regs, const LineCoeffs* c1, const LineCoeffs* c2, int k, const Lanes& L) {
    if (L.shared) {          // multi-group form: the lines are the same for every check -- one copy
        for (int i = L.tid; i < 12 * 12; i += L.n) {
            int fe = i / 12, limb = i % 12, j = fe / 6, e = fe % 6;
            const LineCoeffs* src = j == 0 ? c1 : c2;
            const Fp2& f2 = e < 2 ? src[k].A : (e < 4 ? src[k].B : src[k].C);
            L.shared[fe].l[limb] = (e & 1) ? f2.c1.l[limb] : f2.c0.l[limb];
        }
        L.sync();
        return;
    }
    // 6 Fp per line (A, B, C as Fp2) into regs[kRegLines + 6 j ..]
    for (int q = L.tid; q < 12 * 12 * L.groups; q += L.n) {
        int i = q / L.groups, g = q - i * L.groups;
        int fe = i / 12, limb = i % 12, j = fe / 6, e = fe % 6;
        const LineCoeffs* src = j == 0 ? c1 : c2;
        if (!src) continue;
        const Fp2& f2 = e < 2 ? src[k].A : (e < 4 ? src[k].B : src[k].C);
        L.file(regs, g)[kRegLines + fe].l[limb] = (e & 1) ? f2.c1.l[limb] : f2.c0.l[limb];
    }
    L.sync();
}
// F <- F^|x| conjugated (x < 0), base = save slot `base` (F is overwritten; G is used as the multiplier slot)
// k squarings of F, F in the cyclotomic subgroup (everything after the easy part of the final exponentiation):
// Granger-Scott squarings (18 single products per squaring instead of 24 dual ones).  Measured: fusing consecutive
// squarings into one program (output level + next operand level) makes the sums longer and the chain SLOWER --
// a level costs time proportional to its longest sum.
KZG_HD void sqr_times(Fp* regs, int k, const Lanes& L) {
    for (; k > 0; k--) run(kProg_cyc_sqr1, regs, L);
}
KZG_HD void exp_by_x_slot(Fp* regs, int base, const Lanes& L) {
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
    // == 1 ?
    bool ok = regs[kRegF] == Fp::one();
    for (int i = 1; i < 12; i++) ok = ok && regs[kRegF + i].is_zero();
    L.sync();
    return ok;
}

// L.groups independent checks e(P1[g], Q1) e(P2[g], Q2) == 1 in lockstep (both points of every group must be finite: the program
// path then is the same for all groups; the caller routes the rare identity inputs to the single-group form).  ok[g] (memory
// all cooperating threads see) receives the verdicts.  regs: L.groups register files, L.stride words apart.
KZG_HD void coop_pairing_multi(Fp* regs, const G1Affine* P1, const LineCoeffs* c1, const G1Affine* P2, const LineCoeffs* c2, const Lanes& L,
                               uint8_t* ok) {
    // L.shared must be set: the Frobenius constants (and, per Miller step, the line coefficients) live once for all groups
    if (L.tid == 0) {
        const uint32_t g[10][12] = {KZG_FP_FROB6_1_C0_M, KZG_FP_FROB6_1_C1_M, KZG_FP_FROB6_2_C0_M, KZG_FP_FROB6_2_C1_M, KZG_FP_FROB6_3_C0_M,
                                    KZG_FP_FROB6_3_C1_M, KZG_FP_FROB6_4_C0_M, KZG_FP_FROB6_4_C1_M, KZG_FP_FROB6_5_C0_M, KZG_FP_FROB6_5_C1_M};
        for (int i = 0; i < 10; i++) L.shared[kRegConst - kRegLines + i] = fp_const(g[i]);
    }
    const int rp = compact_multi(kRegP, true);
    for (int gi = L.tid; gi < L.groups; gi += L.n) {
        Fp* r = L.file(regs, gi);
        r[rp] = P1[gi].x; r[rp + 1] = P1[gi].y; r[rp + 2] = P2[gi].x; r[rp + 3] = P2[gi].y;
        for (int i = 0; i < 12; i++) r[kRegF + i] = Fp::zero();
        r[kRegF] = Fp::one();
    }
    L.sync();
    int k = 0;
    for (int bit = 62; bit >= 0; bit--) {
        load_lines(regs, c1, c2, k++, L);
        run(kProg_sqr_lines, regs, L);
        run(kProg_f12_mul, regs, L);
        if ((KZG_BLS_X_ABS >> bit) & 1) {
            load_lines(regs, c1, c2, k++, L);
            run(kProg_lines, regs, L);
            run(kProg_f12_mul, regs, L);
        }
    }
    run(kProg_conj, regs, L);
    // three save slots: S0 = f to the end; S1 = f^(x-1), a; S2 = a^x, b; S3 (b^x) and S4 (b^(x^2), c) take S1's slot once a is dead
    const int S0 = kSave0Thr, S1 = kSave0Thr + 12, S2 = kSave0Thr + 24, S3 = S1, S4 = S1;
    copy_regs(regs, S0, kRegF, 12, L);
    run(kProg_inv_prep, regs, L);
    for (int gi = L.tid; gi < L.groups; gi += L.n) { Fp* r = L.file(regs, gi); r[kRegH + 8] = fp_inv_bingcd(r[kRegH + 8]); }
    L.sync();
    run(kProg_inv_finish, regs, L);
    copy_regs(regs, kRegG, kRegF, 12, L);
    copy_regs(regs, kRegF, S0, 12, L);
    run(kProg_conj, regs, L);
    run(kProg_f12_mul, regs, L);
    run(kProg_frob2, regs, L);
    run(kProg_f12_mul, regs, L);
    copy_regs(regs, S0, kRegF, 12, L);
    exp_by_x_slot(regs, S0, L);
    copy_regs(regs, kRegG, S0, 12, L); run(kProg_conj_g, regs, L); run(kProg_f12_mul, regs, L);
    copy_regs(regs, S1, kRegF, 12, L);
    exp_by_x_slot(regs, S1, L);
    copy_regs(regs, kRegG, S1, 12, L); run(kProg_conj_g, regs, L); run(kProg_f12_mul, regs, L);
    copy_regs(regs, S1, kRegF, 12, L);
    exp_by_x_slot(regs, S1, L);
    copy_regs(regs, S2, kRegF, 12, L);
    copy_regs(regs, kRegF, S1, 12, L); run(kProg_frob, regs, L);
    copy_regs(regs, kRegF, S2, 12, L); run(kProg_f12_mul, regs, L);
    copy_regs(regs, S2, kRegF, 12, L);
    exp_by_x_slot(regs, S2, L);
    copy_regs(regs, S3, kRegF, 12, L);
    exp_by_x_slot(regs, S3, L);
    copy_regs(regs, S4, kRegF, 12, L);
    copy_regs(regs, kRegF, S2, 12, L); run(kProg_frob2, regs, L);
    copy_regs(regs, kRegF, S4, 12, L); run(kProg_f12_mul, regs, L);
    copy_regs(regs, kRegG, S2, 12, L); run(kProg_conj_g, regs, L); run(kProg_f12_mul, regs, L);
    copy_regs(regs, S4, kRegF, 12, L);
    copy_regs(regs, kRegF, S0, 12, L); run(kProg_f12_sqr, regs, L);
    copy_regs(regs, kRegG, S0, 12, L); run(kProg_f12_mul, regs, L);
    copy_regs(regs, kRegG, S4, 12, L); run(kProg_f12_mul, regs, L);
    for (int gi = L.tid; gi < L.groups; gi += L.n) {
        const Fp* r = L.file(regs, gi);
        bool one = r[kRegF] == Fp::one();
        for (int i = 1; i < 12; i++) one = one && r[kRegF + i].is_zero();
        ok[gi] = one ? 1 : 0;
    }
    L.sync();
}

}  // namespace vliw
}  // namespace kzgb200
