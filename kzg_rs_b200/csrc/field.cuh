// Fr (255-bit, 8x32 limbs) and Fp (381-bit, 12x32 limbs) Montgomery arithmetic for sm_100a.
//
// Replaces the field layer of sp1_bls12_381 that kzg-rs calls (Scalar / Fp: reference call sites
// src/kzg_proof.rs:36,90,112,124,127-130,176,188,196-197).  Same Montgomery radix (R = 2^256 / 2^384), so
// the limb image of a value equals the reference's in-memory form on a little-endian host.
//
// Multiplication is a 64-bit-digit CIOS on two accumulators (even / odd limb positions): every
// 32x32->64 product is one IMAD.WIDE.U32 on an aligned register pair with the carry chained through a
// predicate (rows in bigint_rows.cuh), i.e. 2*N^2 FMA-pipe instructions per product-and-reduce and
// 3*N^2 for the fused dual product (a*b + c*d)/R used by the barycentric tree and by Fp2.
// All functions are __host__ __device__ so the layers above can be unit-tested on the CPU.
#pragma once
#include <stdint.h>
#include "bigint_rows.cuh"
#include "consts32.cuh"
#ifndef KZG_NI
#define KZG_NI __host__ __device__ __noinline__ inline
#endif

namespace kzgb200 {

// Device-side moduli: read through ld.global.nc (kept in registers / uniform registers by ptxas).  Note from
// the SASS: ptxas fuses (mad.lo.cc, madc.hi.cc) into IMAD.WIDE.U32.X for most rows and emits the equivalent
// IMAD.X + IMAD.HI.U32.X pair for some reduction rows of q; both forms are carry-chained FMA-pipe work.
#ifdef __CUDACC__
static __device__ const uint32_t d_fr_q[8] = KZG_FR_Q;
static __device__ const uint32_t d_fp_p[12] = KZG_FP_P;
#endif

struct FrParams {
    static constexpr int N = 8;
    static constexpr int EXTRA = 1;  // 3q*2^64 needs one more top word in the dual product
    static constexpr uint32_t INV = KZG_FR_INV32;
    KZG_HD static uint32_t mod(int i) {
#ifdef __CUDA_ARCH__
        return __ldg(&d_fr_q[i]);
#else
        constexpr uint32_t m[8] = KZG_FR_Q; return m[i];
#endif
    }
    KZG_HD static constexpr uint32_t one(int i) { constexpr uint32_t m[8] = KZG_FR_R; return m[i]; }
    KZG_HD static constexpr uint32_t r2(int i) { constexpr uint32_t m[8] = KZG_FR_R2; return m[i]; }
};
struct FpParams {
    static constexpr int N = 12;
    static constexpr int EXTRA = 0;
    static constexpr uint32_t INV = KZG_FP_INV32;
    KZG_HD static uint32_t mod(int i) {
#ifdef __CUDA_ARCH__
        return __ldg(&d_fp_p[i]);
#else
        constexpr uint32_t m[12] = KZG_FP_P; return m[i];
#endif
    }
    KZG_HD static constexpr uint32_t one(int i) { constexpr uint32_t m[12] = KZG_FP_R; return m[i]; }
    KZG_HD static constexpr uint32_t r2(int i) { constexpr uint32_t m[12] = KZG_FP_R2; return m[i]; }
};

// r = (a*b [+ c*d]) / R mod p.  Inputs < p (b and d may be any N-limb value); output < p.
template <class F, bool DUAL>
KZG_HD void mont_mul_impl(uint32_t* r, const uint32_t* a, const uint32_t* b, const uint32_t* c, const uint32_t* d) {
    constexpr int N = F::N, H = N / 2, X = F::EXTRA;
    uint32_t P[N];
#pragma unroll
    for (int i = 0; i < N; i++) P[i] = F::mod(i);
    // E[k] <-> limb position k ; O[k] <-> limb position k+1 ; s0 = a stray limb at position 0 (the odd
    // accumulator's second limb lands there after each 64-bit shift)
    uint32_t E[N + 3], O[N + 3], s0 = 0;
#pragma unroll
    for (int i = 0; i < N + 3; i++) { E[i] = 0; O[i] = 0; }
#pragma unroll
    for (int i = 0; i < N; i += 2) {
        mad_row<H, 2 + X>(E, a, b[i]);          // a_even * b0 -> positions 0..
        mad_row<H, 1 + X>(O, a + 1, b[i]);      // a_odd  * b0 -> positions 1..
        mad_row<H, 1 + X>(O, a, b[i + 1]);      // a_even * b1 -> positions 1..
        mad_row<H, X>(E + 2, a + 1, b[i + 1]);  // a_odd  * b1 -> positions 2..
        if (DUAL) {
            mad_row<H, 2 + X>(E, c, d[i]);
            mad_row<H, 1 + X>(O, c + 1, d[i]);
            mad_row<H, 1 + X>(O, c, d[i + 1]);
            mad_row<H, X>(E + 2, c + 1, d[i + 1]);
        }
        uint32_t m0 = (E[0] + s0) * F::INV;
        mad_row<H, 2 + X>(E, P, m0);                            // position 0: E[0] + s0 == 0 mod 2^32 ...
        mad_row_cin<H, 1 + X>(O, P + 1, m0, E[0], s0);          // ... its overflow carries into position 1
        uint32_t m1 = (E[1] + O[0]) * F::INV;
        mad_row<H, 1 + X>(O, P, m1);                            // position 1: E[1] + O[0] == 0 mod 2^32 ...
        mad_row_cin<H, X>(E + 2, P + 1, m1, E[1], O[0]);        // ... its overflow carries into position 2
        s0 = O[1];                                              // position 2 -> position 0 after the shift
#pragma unroll
        for (int k = 0; k <= N; k++) { E[k] = E[k + 2]; O[k] = O[k + 2]; }
        E[N + 1] = 0; E[N + 2] = 0; O[N + 1] = 0; O[N + 2] = 0;
    }
    // T = E + (O << 32) + s0 < 2p fits N limbs
    uint32_t Bv[N], t[N], s[N];
    Bv[0] = s0;
#pragma unroll
    for (int k = 1; k < N; k++) Bv[k] = O[k - 1];
    add_n<N>(t, E, Bv);
    uint32_t borrow = sub_n<N>(s, t, P);
#pragma unroll
    for (int k = 0; k < N; k++) r[k] = borrow ? t[k] : s[k];
}

template <class F>
struct Fe {
    uint32_t l[F::N];
    static constexpr int N = F::N;

    KZG_HD static Fe zero() { Fe r; for (int i = 0; i < N; i++) r.l[i] = 0; return r; }
    KZG_HD static Fe one() { Fe r; for (int i = 0; i < N; i++) r.l[i] = F::one(i); return r; }
    KZG_HD static Fe modulus() { Fe r; for (int i = 0; i < N; i++) r.l[i] = F::mod(i); return r; }
    KZG_HD bool is_zero() const { uint32_t o = 0; for (int i = 0; i < N; i++) o |= l[i]; return o == 0; }
    KZG_HD bool operator==(const Fe& b) const { uint32_t o = 0; for (int i = 0; i < N; i++) o |= l[i] ^ b.l[i]; return o == 0; }
    KZG_HD bool operator!=(const Fe& b) const { return !(*this == b); }
    // raw limbs >= modulus ?
    KZG_HD bool geq_modulus() const { Fe m = modulus(), t; return sub_n<N>(t.l, l, m.l) == 0; }

    KZG_NI Fe operator+(const Fe& b) const { return add_inl(b); }
    KZG_NI Fe operator-(const Fe& b) const { return sub_inl(b); }
    KZG_HD Fe add_inl(const Fe& b) const {
        Fe t, s, m = modulus();
        uint32_t carry = add_n<N>(t.l, l, b.l);
        uint32_t borrow = sub_n<N>(s.l, t.l, m.l);
        bool use_s = carry || !borrow;
        Fe r; for (int i = 0; i < N; i++) r.l[i] = use_s ? s.l[i] : t.l[i];
        return r;
    }
    KZG_HD Fe sub_inl(const Fe& b) const {
        Fe t, s, m = modulus();
        uint32_t borrow = sub_n<N>(t.l, l, b.l);
        add_n<N>(s.l, t.l, m.l);
        Fe r; for (int i = 0; i < N; i++) r.l[i] = borrow ? s.l[i] : t.l[i];
        return r;
    }
    KZG_HD Fe neg() const { return zero() - *this; }
    KZG_HD Fe dbl() const { return *this + *this; }
    // operator* / mul_dual are real calls (the tower / curve / pairing code above would otherwise inline
    // hundreds of 300-instruction bodies); hot kernels use the *_inl forms.
    KZG_NI Fe operator*(const Fe& b) const { Fe r; mont_mul_impl<F, false>(r.l, l, b.l, nullptr, nullptr); return r; }
    KZG_HD Fe mul_inl(const Fe& b) const { Fe r; mont_mul_impl<F, false>(r.l, l, b.l, nullptr, nullptr); return r; }
    KZG_HD Fe sqr() const { return *this * *this; }
    // (a*b + c*d) / R
    KZG_NI static Fe mul_dual(const Fe& a, const Fe& b, const Fe& c, const Fe& d) {
        Fe r; mont_mul_impl<F, true>(r.l, a.l, b.l, c.l, d.l); return r;
    }
    KZG_HD static Fe mul_dual_inl(const Fe& a, const Fe& b, const Fe& c, const Fe& d) {
        Fe r; mont_mul_impl<F, true>(r.l, a.l, b.l, c.l, d.l); return r;
    }
    // raw (non-Montgomery, any N-limb value) -> Montgomery form of value mod p
    KZG_HD static Fe from_raw(const Fe& raw) { Fe r2; for (int i = 0; i < N; i++) r2.l[i] = F::r2(i); return r2 * raw; }
    // Montgomery -> canonical limbs
    KZG_HD Fe to_raw() const { Fe o = zero(); o.l[0] = 1; return *this * o; }
    KZG_HD static Fe from_u32(uint32_t v) { Fe t = zero(); t.l[0] = v; return from_raw(t); }
    // exponent given as little-endian 32-bit limbs, scanned from bit (nbits-1)
    KZG_NI Fe pow(const uint32_t* e, int nbits) const {
        Fe acc = one();
        for (int i = nbits - 1; i >= 0; i--) {
            acc = acc.sqr();
            if ((e[i >> 5] >> (i & 31)) & 1) acc = acc * *this;
        }
        return acc;
    }
};

using Fr = Fe<FrParams>;
using Fp = Fe<FpParams>;

// a^(q-2) / a^(p-2)
KZG_HD Fr fr_inv(const Fr& a) { const uint32_t e[8] = KZG_FR_Q_MINUS_2; return a.pow(e, 255); }
KZG_HD Fp fp_inv(const Fp& a) { const uint32_t e[12] = KZG_FP_P_MINUS_2; return a.pow(e, 381); }
// candidate square root a^((p+1)/4); caller checks
KZG_HD Fp fp_sqrt_candidate(const Fp& a) { const uint32_t e[12] = KZG_FP_SQRT_EXP; return a.pow(e, 379); }
// canonical value > (p-1)/2
KZG_HD bool fp_lex_largest(const Fp& a) {
    Fp raw = a.to_raw(); const uint32_t h[12] = KZG_FP_P_MINUS_1_HALF; uint32_t t[12];
    // raw > h  <=>  h - raw borrows
    return sub_n<12>(t, h, raw.l) != 0;
}
KZG_HD Fp fp_const(const uint32_t (&v)[12]) { Fp r; for (int i = 0; i < 12; i++) r.l[i] = v[i]; return r; }
KZG_HD Fr fr_const(const uint32_t (&v)[8]) { Fr r; for (int i = 0; i < 8; i++) r.l[i] = v[i]; return r; }

// 32 big-endian bytes at p (4-byte aligned words already loaded) helpers are in the kernels.

}  // namespace kzgb200
