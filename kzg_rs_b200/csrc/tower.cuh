// Fp2 / Fp6 / Fp12 tower for the BLS12-381 pairing on sm_100a:
//   Fp2 = Fp[u]/(u^2+1), Fp6 = Fp2[v]/(v^3-(1+u)), Fp12 = Fp6[w]/(w^2-v).
// Replaces the extension-field layer of sp1_bls12_381 behind pairings_verify
// (reference src/pairings.rs:5-9).  Fp2 products use the fused dual Montgomery product
// (a0*b0 + (-a1)*b1, a0*b1 + a1*b0): two reductions instead of Karatsuba's three.
#pragma once
#include "field.cuh"

namespace kzgb200 {

struct Fp2 {
    Fp c0, c1;
    KZG_HD static Fp2 zero() { return {Fp::zero(), Fp::zero()}; }
    KZG_HD static Fp2 one() { return {Fp::one(), Fp::zero()}; }
    KZG_HD bool is_zero() const { return c0.is_zero() && c1.is_zero(); }
    KZG_HD bool operator==(const Fp2& b) const { return c0 == b.c0 && c1 == b.c1; }
    KZG_HD Fp2 operator+(const Fp2& b) const { return {c0 + b.c0, c1 + b.c1}; }
    KZG_HD Fp2 operator-(const Fp2& b) const { return {c0 - b.c0, c1 - b.c1}; }
    KZG_HD Fp2 neg() const { return {c0.neg(), c1.neg()}; }
    KZG_HD Fp2 dbl() const { return {c0.dbl(), c1.dbl()}; }
    KZG_HD Fp2 conj() const { return {c0, c1.neg()}; }
    KZG_NI Fp2 operator*(const Fp2& b) const {
        Fp n1 = c1.neg();
        return {Fp::mul_dual(c0, b.c0, n1, b.c1), Fp::mul_dual(c0, b.c1, c1, b.c0)};
    }
    KZG_NI Fp2 sqr() const {
        Fp s = c0 + c1, d = c0 - c1, m = c0 * c1;
        return {s * d, m.dbl()};
    }
    KZG_HD Fp2 mul_fp(const Fp& k) const { return {c0 * k, c1 * k}; }
    // times xi = 1 + u
    KZG_HD Fp2 mul_xi() const { return {c0 - c1, c0 + c1}; }
    KZG_NI Fp2 inv() const {
        Fp n = fp_inv(Fp::mul_dual(c0, c0, c1, c1));
        return {c0 * n, (c1 * n).neg()};
    }
};

struct Fp6 {
    Fp2 c0, c1, c2;
    KZG_HD static Fp6 zero() { return {Fp2::zero(), Fp2::zero(), Fp2::zero()}; }
    KZG_HD static Fp6 one() { return {Fp2::one(), Fp2::zero(), Fp2::zero()}; }
    KZG_HD bool operator==(const Fp6& b) const { return c0 == b.c0 && c1 == b.c1 && c2 == b.c2; }
    KZG_NI Fp6 operator+(const Fp6& b) const { return {c0 + b.c0, c1 + b.c1, c2 + b.c2}; }
    KZG_NI Fp6 operator-(const Fp6& b) const { return {c0 - b.c0, c1 - b.c1, c2 - b.c2}; }
    KZG_HD Fp6 neg() const { return {c0.neg(), c1.neg(), c2.neg()}; }
    KZG_NI Fp6 operator*(const Fp6& b) const {
        Fp2 t0 = c0 * b.c0, t1 = c1 * b.c1, t2 = c2 * b.c2;
        Fp2 r0 = t0 + ((c1 + c2) * (b.c1 + b.c2) - t1 - t2).mul_xi();
        Fp2 r1 = (c0 + c1) * (b.c0 + b.c1) - t0 - t1 + t2.mul_xi();
        Fp2 r2 = (c0 + c2) * (b.c0 + b.c2) - t0 - t2 + t1;
        return {r0, r1, r2};
    }
    KZG_HD Fp6 mul_v() const { return {c2.mul_xi(), c0, c1}; }
    // times (b0 + b1 v)
    KZG_NI Fp6 mul_by_01(const Fp2& b0, const Fp2& b1) const {
        Fp2 t0 = c0 * b0, t1 = c1 * b1;
        Fp2 r0 = t0 + (c2 * b1).mul_xi();
        Fp2 r1 = (c0 + c1) * (b0 + b1) - t0 - t1;
        Fp2 r2 = c2 * b0 + t1;
        return {r0, r1, r2};
    }
    // times (b1 v)
    KZG_NI Fp6 mul_by_1(const Fp2& b1) const { return {(c2 * b1).mul_xi(), c0 * b1, c1 * b1}; }
    KZG_NI Fp6 inv() const {
        Fp2 a0 = c0.sqr() - (c1 * c2).mul_xi();
        Fp2 a1 = c2.sqr().mul_xi() - c0 * c1;
        Fp2 a2 = c1.sqr() - c0 * c2;
        Fp2 t = (c0 * a0 + (c2 * a1 + c1 * a2).mul_xi()).inv();
        return {a0 * t, a1 * t, a2 * t};
    }
};

struct Fp12 {
    Fp6 c0, c1;
    KZG_HD static Fp12 one() { return {Fp6::one(), Fp6::zero()}; }
    KZG_HD bool operator==(const Fp12& b) const { return c0 == b.c0 && c1 == b.c1; }
    KZG_NI Fp12 operator*(const Fp12& b) const {
        Fp6 t0 = c0 * b.c0, t1 = c1 * b.c1;
        Fp6 m = (c0 + c1) * (b.c0 + b.c1) - t0 - t1;
        return {t0 + t1.mul_v(), m};
    }
    KZG_NI Fp12 sqr() const {
        Fp6 ab = c0 * c1;
        Fp6 s = (c0 + c1) * (c0 + c1.mul_v()) - ab - ab.mul_v();
        return {s, ab + ab};
    }
    KZG_HD Fp12 conj() const { return {c0, c1.neg()}; }
    KZG_NI Fp12 inv() const {
        Fp6 t = (c0 * c0 - (c1 * c1).mul_v()).inv();
        return {c0 * t, (c1 * t).neg()};
    }
    // times the sparse line value  A + B v + C vw  (coefficients at c0.c0, c0.c1, c1.c1)
    KZG_NI Fp12 mul_by_014(const Fp2& A, const Fp2& B, const Fp2& C) const {
        Fp6 t0 = c0.mul_by_01(A, B);
        Fp6 t1 = c1.mul_by_1(C);
        Fp6 m = (c0 + c1).mul_by_01(A, B + C) - t0 - t1;
        return {t0 + t1.mul_v(), m};
    }
    // x -> x^p : conjugate each Fp2 coefficient, scale the w^k coefficient by xi^(k(p-1)/6)
    KZG_NI Fp12 frob() const {
        const uint32_t g10[12] = KZG_FP_FROB6_1_C0_M, g11[12] = KZG_FP_FROB6_1_C1_M;
        const uint32_t g20[12] = KZG_FP_FROB6_2_C0_M, g21[12] = KZG_FP_FROB6_2_C1_M;
        const uint32_t g30[12] = KZG_FP_FROB6_3_C0_M, g31[12] = KZG_FP_FROB6_3_C1_M;
        const uint32_t g40[12] = KZG_FP_FROB6_4_C0_M, g41[12] = KZG_FP_FROB6_4_C1_M;
        const uint32_t g50[12] = KZG_FP_FROB6_5_C0_M, g51[12] = KZG_FP_FROB6_5_C1_M;
        Fp2 g1 = {fp_const(g10), fp_const(g11)}, g2 = {fp_const(g20), fp_const(g21)}, g3 = {fp_const(g30), fp_const(g31)},
            g4 = {fp_const(g40), fp_const(g41)}, g5 = {fp_const(g50), fp_const(g51)};
        Fp12 r;
        r.c0.c0 = c0.c0.conj();
        r.c0.c1 = c0.c1.conj() * g2;
        r.c0.c2 = c0.c2.conj() * g4;
        r.c1.c0 = c1.c0.conj() * g1;
        r.c1.c1 = c1.c1.conj() * g3;
        r.c1.c2 = c1.c2.conj() * g5;
        return r;
    }
};

}  // namespace kzgb200
