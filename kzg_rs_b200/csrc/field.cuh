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

// r = a*a / R mod p: the off-diagonal products once (N(N-1)/2 wide multiplications instead of N^2, in rows a_i * {a_j : j > i}
// split by the parity of i + j onto the even / odd accumulators), doubled, plus the diagonal, then the same word-serial
// Montgomery reduction as above on the low half with the high half added at the end.  N(N+1)/2 + N^2 FMA-pipe instructions
// against 2 N^2 for the general product (222 / 288 for Fp, 100 / 128 for Fr).
template <int N, int I>
struct SqrRows {
    KZG_HD static void run(uint32_t* E, uint32_t* O, const uint32_t* a) {
        constexpr int He = (N - 1 - I) / 2, Ho = (N - I) / 2;   // partners j > i with j - i even / odd
        if constexpr (He > 0) mad_row<He, 1>(E + 2 * I + 2, a + I + 2, a[I]);
        if constexpr (Ho > 0) mad_row<Ho, 1>(O + 2 * I, a + I + 1, a[I]);
        if constexpr (I + 2 < N) SqrRows<N, I + 1>::run(E, O, a);
    }
};
template <class F>
KZG_HD void mont_sqr_impl(uint32_t* r, const uint32_t* a) {
    constexpr int N = F::N, H = N / 2;
    uint32_t E[2 * N + 2], O[2 * N + 2];
#pragma unroll
    for (int i = 0; i < 2 * N + 2; i++) { E[i] = 0; O[i] = 0; }
    SqrRows<N, 0>::run(E, O, a);
    // T = E + (O << 32) = sum_{i<j} a_i a_j 2^(32(i+j)); then 2T + diagonal
    uint32_t T[2 * N], Bv[2 * N];
    Bv[0] = 0;
#pragma unroll
    for (int k = 1; k < 2 * N; k++) Bv[k] = O[k - 1];
    add_n<2 * N>(T, E, Bv);
#pragma unroll
    for (int k = 2 * N - 1; k >= 1; k--) {
#ifdef __CUDA_ARCH__
        T[k] = __funnelshift_l(T[k - 1], T[k], 1);
#else
        T[k] = (T[k] << 1) | (T[k - 1] >> 31);
#endif
    }
    T[0] <<= 1;
    sqr_diag<N>(T, a);
    // word-serial reduction of the low half (two 32-bit digits per round, as in mont_mul_impl), high half added afterwards
    uint32_t P[N];
#pragma unroll
    for (int i = 0; i < N; i++) P[i] = F::mod(i);
    uint32_t s0 = 0;
#pragma unroll
    for (int i = 0; i < N + 3; i++) { E[i] = i < N ? T[i] : 0u; O[i] = 0; }
#pragma unroll
    for (int i = 0; i < N; i += 2) {
        uint32_t m0 = (E[0] + s0) * F::INV;
        mad_row<H, 2>(E, P, m0);
        mad_row_cin<H, 1>(O, P + 1, m0, E[0], s0);
        uint32_t m1 = (E[1] + O[0]) * F::INV;
        mad_row<H, 1>(O, P, m1);
        mad_row_cin<H, 0>(E + 2, P + 1, m1, E[1], O[0]);
        s0 = O[1];
#pragma unroll
        for (int k = 0; k <= N; k++) { E[k] = E[k + 2]; O[k] = O[k + 2]; }
        E[N + 1] = 0; E[N + 2] = 0; O[N + 1] = 0; O[N + 2] = 0;
    }
    // (T_lo + m p) / R = E + (O << 32) + s0 <= p, plus T_hi < p^2 / R: the total is < 2p and fits N limbs
    uint32_t t[N], u[N], s[N];
    Bv[0] = s0;
#pragma unroll
    for (int k = 1; k < N; k++) Bv[k] = O[k - 1];
    add_n<N>(t, E, Bv);
    add_n<N>(u, t, T + N);
    uint32_t borrow = sub_n<N>(s, u, P);
#pragma unroll
    for (int k = 0; k < N; k++) r[k] = borrow ? u[k] : s[k];
}

template <class F>
struct Fe;
// The out-of-line forms take and return their operands BY VALUE: the CUDA ABI passes such aggregates in registers, so a
// multiplication called from curve / tower code costs the call and some moves; by reference (`this` included) every operand
// would have to live in the caller's stack frame (round 1: 2.7-3.2 KB frames in the G1 parsing kernels, 19.7 KB in the
// per-tuple pairing kernel).  The tower / curve / pairing code above would otherwise inline hundreds of 300-instruction bodies;
// hot kernels use the *_inl forms.
template <class F> KZG_NI Fe<F> fe_mul(Fe<F> a, Fe<F> b);
template <class F> KZG_NI Fe<F> fe_sqr(Fe<F> a);
template <class F> KZG_NI Fe<F> fe_mul_dual(Fe<F> a, Fe<F> b, Fe<F> c, Fe<F> d);
template <class F> KZG_NI Fe<F> fe_add(Fe<F> a, Fe<F> b);
template <class F> KZG_NI Fe<F> fe_sub(Fe<F> a, Fe<F> b);

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

    KZG_HD Fe operator+(const Fe& b) const { return fe_add<F>(*this, b); }
    KZG_HD Fe operator-(const Fe& b) const { return fe_sub<F>(*this, b); }
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
    KZG_HD Fe operator*(const Fe& b) const { return fe_mul<F>(*this, b); }
    KZG_HD Fe mul_inl(const Fe& b) const { Fe r; mont_mul_impl<F, false>(r.l, l, b.l, nullptr, nullptr); return r; }
    KZG_HD Fe sqr() const { return fe_sqr<F>(*this); }
    KZG_HD Fe sqr_inl() const { Fe r; mont_sqr_impl<F>(r.l, l); return r; }
    // (a*b + c*d) / R
    KZG_HD static Fe mul_dual(const Fe& a, const Fe& b, const Fe& c, const Fe& d) { return fe_mul_dual<F>(a, b, c, d); }
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
    // the same with a 4-bit fixed window: nbits squarings + nbits/4 multiplications (+ 14 for the table) instead of ~nbits/2
    KZG_NI Fe pow_w4(const uint32_t* e, int nbits) const {
        Fe tab[16];
        tab[0] = one(); tab[1] = *this;
        for (int i = 2; i < 16; i++) tab[i] = (i & 1) ? tab[i - 1] * *this : tab[i >> 1].sqr();
        Fe acc = one();
        for (int i = ((nbits + 3) / 4) * 4 - 4; i >= 0; i -= 4) {
            acc = acc.sqr().sqr().sqr().sqr();
            uint32_t d = (e[i >> 5] >> (i & 31)) & 15u;
            if (d) acc = acc * tab[d];
        }
        return acc;
    }
};
template <class F> KZG_NI Fe<F> fe_mul(Fe<F> a, Fe<F> b) { Fe<F> r; mont_mul_impl<F, false>(r.l, a.l, b.l, nullptr, nullptr); return r; }
template <class F> KZG_NI Fe<F> fe_sqr(Fe<F> a) { Fe<F> r; mont_sqr_impl<F>(r.l, a.l); return r; }
template <class F> KZG_NI Fe<F> fe_mul_dual(Fe<F> a, Fe<F> b, Fe<F> c, Fe<F> d) { Fe<F> r; mont_mul_impl<F, true>(r.l, a.l, b.l, c.l, d.l); return r; }
template <class F> KZG_NI Fe<F> fe_add(Fe<F> a, Fe<F> b) { return a.add_inl(b); }
template <class F> KZG_NI Fe<F> fe_sub(Fe<F> a, Fe<F> b) { return a.sub_inl(b); }

using Fr = Fe<FrParams>;
using Fp = Fe<FpParams>;

// a^(q-2) / a^(p-2)
KZG_HD Fr fr_inv(const Fr& a) { const uint32_t e[8] = KZG_FR_Q_MINUS_2; return a.pow(e, 255); }
KZG_HD Fp fp_inv(const Fp& a) { const uint32_t e[12] = KZG_FP_P_MINUS_2; return a.pow(e, 381); }
// candidate square root a^((p+1)/4); caller checks
KZG_HD Fp fp_sqrt_candidate(const Fp& a) { const uint32_t e[12] = KZG_FP_SQRT_EXP; return a.pow_w4(e, 379); }
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
