// Fp in 14 limbs of 29 bits (406 bits, Montgomery radix 2^406) -- the representation of the cooperative pairing engine
// (vliw29.cuh).  Why a second representation beside field.cuh's 12 x 32:
//   * a 29 x 29-bit product plus a 64-bit column accumulator is ONE IMAD.WIDE with no carry in or out: the 196 (392 for the
//     fused dual product) partial products are independent column sums instead of one predicate-chained IMAD.WIDE.X sequence.
//     IMAD.WIDE issues at twice the rate of IMAD.WIDE.X (tools/microbench/intpipe.cu) and, what matters for the engine, a
//     lone warp is no longer latency-bound on the carry chain;
//   * 2^406 = 2^25.3 p: sums of dozens of terms and their negations need no modular reduction (the engine's LIN instruction
//     is 14 IMAD.WIDE per term and one carry pass); the static bounds are kept by tools/gen_vliw.py.
// Values are ANY representative below 2^406; limbs 0..12 below 2^29, limb 13 holds the rest.
#pragma once
#include "field.cuh"
#include "consts29.cuh"

namespace kzgb200 {
namespace f29 {

constexpr int kN = 14, kW = 29;
constexpr uint32_t kMask = 0x1fffffffu;
struct alignas(16) F29 { uint32_t l[16]; };   // l[14], l[15]: padding (registers are moved as four 16-byte words)

KZG_HD constexpr uint32_t p29(int i) { constexpr uint32_t m[14] = KZG29_P; return m[i]; }
KZG_HD uint32_t p29_rt(int i) { constexpr uint32_t m[14] = KZG29_P; return m[i]; }   // run-time index
KZG_HD constexpr uint32_t p2_29(int i) { constexpr uint32_t m[28] = KZG29_P2; return m[i]; }
KZG_HD F29 f29_const(const uint32_t (&v)[14]) { F29 r; for (int i = 0; i < 14; i++) r.l[i] = v[i]; r.l[14] = r.l[15] = 0; return r; }
KZG_HD F29 f29_zero() { F29 r; for (int i = 0; i < 16; i++) r.l[i] = 0; return r; }
KZG_HD F29 f29_one() { const uint32_t v[14] = KZG29_ONE; return f29_const(v); }

// acc += a * b as ONE IMAD.WIDE with the accumulator as its addend.  Written as the (mad.lo.cc, madc.hi) pair: ptxas fuses the
// pair into IMAD.WIDE Rd, Ra, Rb, Rd; given `acc += (uint64_t)a * b` or mad.wide it instead emits IMAD.WIDE Rt, Ra, Rb, RZ plus a
// three-input IADD3 / IADD3.X pair per two products -- twice the instructions, and a lone warp pays per instruction.
KZG_HD void madw(uint64_t& acc, uint32_t a, uint32_t b) {
#ifdef __CUDA_ARCH__
    asm("{\n\t.reg .u32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tmad.lo.cc.u32 lo, %1, %2, lo;\n\tmadc.hi.u32 hi, %1, %2, hi;\n\tmov.b64 %0, {lo, hi};\n\t}"
        : "+l"(acc) : "r"(a), "r"(b));
#else
    acc += (uint64_t)a * b;
#endif
}
KZG_HD void madw_s(uint64_t& acc, int32_t a, int32_t b) {
#ifdef __CUDA_ARCH__
    asm("{\n\t.reg .u32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tmad.lo.cc.s32 lo, %1, %2, lo;\n\tmadc.hi.s32 hi, %1, %2, hi;\n\tmov.b64 %0, {lo, hi};\n\t}"
        : "+l"(acc) : "r"(a), "r"(b));
#else
    acc += (uint64_t)((int64_t)a * b);
#endif
}
// t >> 29 for a column sum: logical, or arithmetic when the columns are signed (subtracted product)
KZG_HD uint64_t shr29(uint64_t v, uint64_t signed_mask) {
    return (v >> kW) | (((uint64_t)((int64_t)v >> 63) & signed_mask) << (64 - kW));
}

// r = (a b [+|-] c d) / 2^406 (+ a multiple of p).  Operand values a < A p, ... with A B + C D + 1 <= 2^23: result < 1.25 p.
// A subtracted product (neg) needs kx >= c d / p^2: kx p^2 is added so that the total stays non-negative; the columns are then
// signed 64-bit sums (|column| < 2^62.8), otherwise unsigned (< 2^63.4).
KZG_HD void mont_mul29(uint32_t* __restrict__ r, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d, bool dual, bool neg,
                       uint32_t kx) {
    uint64_t t[2 * kN];
#pragma unroll
    for (int k = 0; k < 2 * kN; k++) t[k] = 0;
#pragma unroll
    for (int i = 0; i < kN; i++)
#pragma unroll
        for (int j = 0; j < kN; j++) madw(t[i + j], a[i], b[j]);
    const uint64_t smask = neg ? ~0ull : 0ull;
    if (dual) {
        int32_t sd[kN];
#pragma unroll
        for (int j = 0; j < kN; j++) sd[j] = neg ? -(int32_t)d[j] : (int32_t)d[j];
#pragma unroll
        for (int i = 0; i < kN; i++)
#pragma unroll
            for (int j = 0; j < kN; j++) madw_s(t[i + j], (int32_t)c[i], sd[j]);
        if (neg) {
#pragma unroll
            for (int k = 0; k < 2 * kN; k++) madw(t[k], p2_29(k), kx);
        }
    }
#pragma unroll
    for (int i = 0; i < kN; i++) {
        uint32_t m = ((uint32_t)t[i] * KZG29_PINV) & kMask;
#pragma unroll
        for (int j = 0; j < kN; j++) madw(t[i + j], m, p29(j));
        t[i + 1] += shr29(t[i], smask);
    }
#pragma unroll
    for (int k = kN; k < 2 * kN - 1; k++) {
        r[k - kN] = (uint32_t)t[k] & kMask;
        t[k + 1] += shr29(t[k], smask);
    }
    r[kN - 1] = (uint32_t)t[2 * kN - 1];
}
KZG_HD F29 mul29(const F29& a, const F29& b) {
    F29 r; mont_mul29(r.l, a.l, b.l, a.l, b.l, false, false, 0); r.l[14] = r.l[15] = 0; return r;
}

// bit repacking between 12 x 32 and 14 x 29 (the value must be below 2^384 / 2^406)
KZG_HD F29 pack29(const uint32_t* x /* 12 limbs */) {
    F29 r;
#pragma unroll
    for (int i = 0; i < kN; i++) {
        int bit = kW * i, w = bit >> 5, s = bit & 31;
        uint64_t v = (uint64_t)(w < 12 ? x[w] : 0u) | ((uint64_t)(w + 1 < 12 ? x[w + 1] : 0u) << 32);
        r.l[i] = (uint32_t)(v >> s) & kMask;
    }
    r.l[14] = r.l[15] = 0;
    return r;
}
KZG_HD void unpack29(uint32_t* x /* 12 limbs */, const F29& v) {   // value < 2^384
#pragma unroll
    for (int w = 0; w < 12; w++) {
        // bits [32w, 32w+32): limbs i0 = floor(32w/29) and the next one or two
        int bit = 32 * w, i0 = bit / kW, s = bit - kW * i0;
        uint64_t acc = (uint64_t)v.l[i0] >> s;
        int have = kW - s;
        if (i0 + 1 < kN) acc |= (uint64_t)v.l[i0 + 1] << have;
        have += kW;
        if (have < 32 && i0 + 2 < kN) acc |= (uint64_t)v.l[i0 + 2] << have;
        x[w] = (uint32_t)acc;
    }
}
// field.cuh Montgomery form (x 2^384, < p) -> this representation (x 2^406, < 1.25 p)
KZG_HD F29 from_fp(const Fp& x) {
    const uint32_t k[14] = KZG29_FROM32;
    return mul29(pack29(x.l), f29_const(k));
}
// any representative -> the canonical integer of the field element it stands for (12 x 32, < p)
KZG_HD Fp canonical(const F29& v) {
    F29 one = f29_zero(); one.l[0] = 1;
    F29 u = mul29(v, one);                       // v / 2^406 + (multiple of p) < p + 1
    Fp x, s, m = Fp::modulus();
    unpack29(x.l, u);
    uint32_t borrow = sub_n<12>(s.l, x.l, m.l);
    Fp r; for (int i = 0; i < 12; i++) r.l[i] = borrow ? x.l[i] : s.l[i];
    return r;
}
// raw integer x < p -> x 2^406
KZG_HD F29 from_raw(const Fp& x) {
    const uint32_t k[14] = KZG29_R2;
    return mul29(pack29(x.l), f29_const(k));
}

}  // namespace f29
}  // namespace kzgb200
