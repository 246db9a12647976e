// G1 / G2 group law (Jacobian, a = 0), ZCash compressed decoding with the endomorphism subgroup check.
// Replaces G1Affine::from_compressed / G1Projective arithmetic / G2Affine::from_compressed_unchecked of
// sp1_bls12_381 as called from reference src/kzg_proof.rs:17-25 (safe_g1_affine_from_bytes), :419-433
// (linear combinations) and build.rs:68,73 (setup points).
#pragma once
#include "tower.cuh"

namespace kzgb200 {

template <class FE>
struct Affine {
    FE x, y;
    uint32_t inf;  // 1 = identity
};

template <class FE>
struct Jac;
// out-of-line group law, operands by value (registers, see field.cuh)
template <class FE> KZG_NI Jac<FE> jac_dbl(Jac<FE> p);
template <class FE> KZG_NI Jac<FE> jac_add_mixed(Jac<FE> p, Affine<FE> q);
template <class FE> KZG_NI Jac<FE> jac_add(Jac<FE> p, Jac<FE> q);

template <class FE>
struct Jac {
    FE x, y, z;  // z == 0 <=> identity
    KZG_HD static Jac identity() { return {FE::one(), FE::one(), FE::zero()}; }
    KZG_HD bool is_identity() const { return z.is_zero(); }
    KZG_HD static Jac from_affine(const Affine<FE>& a) {
        if (a.inf) return identity();
        return {a.x, a.y, FE::one()};
    }
    KZG_HD Jac neg() const { return {x, y.neg(), z}; }
    KZG_HD Jac dbl() const { return jac_dbl<FE>(*this); }
    KZG_HD Jac add_mixed(const Affine<FE>& q) const { return jac_add_mixed<FE>(*this, q); }
    KZG_HD Jac add(const Jac& q) const { return jac_add<FE>(*this, q); }
    // projective equality
    KZG_NI bool equals(const Jac& b) const {
        bool ia = is_identity(), ib = b.is_identity();
        if (ia || ib) return ia && ib;
        FE za = z.sqr(), zb = b.z.sqr();
        if (!(x * zb == b.x * za)) return false;
        return y * zb * b.z == b.y * za * z;
    }
};
// dbl-2009-l (a = 0): 2M + 5S
template <class FE> KZG_NI Jac<FE> jac_dbl(Jac<FE> p) {
    if (p.is_identity()) return p;
    FE A = p.x.sqr(), B = p.y.sqr(), C = B.sqr();
    FE D = ((p.x + B).sqr() - A - C).dbl();
    FE E = A.dbl() + A, F = E.sqr();
    FE z3 = (p.y * p.z).dbl();
    FE x3 = F - D.dbl();
    FE y3 = E * (D - x3) - C.dbl().dbl().dbl();
    return {x3, y3, z3};
}
// 8M + 3S
template <class FE> KZG_NI Jac<FE> jac_add_mixed(Jac<FE> p, Affine<FE> q) {
    if (q.inf) return p;
    if (p.is_identity()) return Jac<FE>::from_affine(q);
    FE Z2 = p.z.sqr(), U2 = q.x * Z2, S2 = q.y * Z2 * p.z;
    FE H = U2 - p.x, R = S2 - p.y;
    if (H.is_zero()) return R.is_zero() ? jac_dbl<FE>(p) : Jac<FE>::identity();
    FE H2 = H.sqr(), H3 = H2 * H, XH2 = p.x * H2;
    FE x3 = R.sqr() - H3 - XH2.dbl();
    FE y3 = R * (XH2 - x3) - p.y * H3;
    return {x3, y3, p.z * H};
}
// 12M + 4S
template <class FE> KZG_NI Jac<FE> jac_add(Jac<FE> p, Jac<FE> q) {
    if (p.is_identity()) return q;
    if (q.is_identity()) return p;
    FE Z1Z1 = p.z.sqr(), Z2Z2 = q.z.sqr();
    FE U1 = p.x * Z2Z2, U2 = q.x * Z1Z1;
    FE S1 = p.y * Z2Z2 * q.z, S2 = q.y * Z1Z1 * p.z;
    FE H = U2 - U1, R = S2 - S1;
    if (H.is_zero()) return R.is_zero() ? jac_dbl<FE>(p) : Jac<FE>::identity();
    FE H2 = H.sqr(), H3 = H2 * H, UH2 = U1 * H2;
    FE x3 = R.sqr() - H3 - UH2.dbl();
    FE y3 = R * (UH2 - x3) - S1 * H3;
    return {x3, y3, p.z * q.z * H};
}

using G1Affine = Affine<Fp>;
using G1 = Jac<Fp>;
using G2Affine = Affine<Fp2>;
using G2 = Jac<Fp2>;

// [k]P, k given as little-endian 32-bit limbs; plain left-to-right double-and-add
template <class FE>
KZG_NI Jac<FE> scalar_mul(const Jac<FE>& p, const uint32_t* k, int nbits) {
    Jac<FE> acc = Jac<FE>::identity();
    for (int i = nbits - 1; i >= 0; i--) {
        acc = acc.dbl();
        if ((k[i >> 5] >> (i & 31)) & 1) acc = acc.add(p);
    }
    return acc;
}
template <class FE>
KZG_NI Jac<FE> scalar_mul_affine(const Affine<FE>& p, const uint32_t* k, int nbits) {
    Jac<FE> acc = Jac<FE>::identity();
    for (int i = nbits - 1; i >= 0; i--) {
        acc = acc.dbl();
        if ((k[i >> 5] >> (i & 31)) & 1) acc = acc.add_mixed(p);
    }
    return acc;
}

KZG_HD G1Affine g1_generator() {
    const uint32_t gx[12] = KZG_FP_G1X_M, gy[12] = KZG_FP_G1Y_M;
    return {fp_const(gx), fp_const(gy), 0};
}
KZG_HD G2Affine g2_generator() {
    const uint32_t x0[12] = KZG_FP_G2X0_M, x1[12] = KZG_FP_G2X1_M, y0[12] = KZG_FP_G2Y0_M, y1[12] = KZG_FP_G2Y1_M;
    return {{fp_const(x0), fp_const(x1)}, {fp_const(y0), fp_const(y1)}, 0};
}
KZG_NI G1Affine g1_to_affine(const G1& p) {
    if (p.is_identity()) return {Fp::zero(), Fp::zero(), 1};
    Fp zi = fp_inv(p.z), zi2 = zi.sqr();
    return {p.x * zi2, p.y * zi2 * zi, 0};
}
KZG_NI G2Affine g2_to_affine(const G2& p) {
    if (p.is_identity()) return {Fp2::zero(), Fp2::zero(), 1};
    Fp2 zi = p.z.inv(), zi2 = zi.sqr();
    return {p.x * zi2, p.y * zi2 * zi, 0};
}

// P in the prime-order subgroup  <=>  phi(P) == -[x^2]P with phi(x,y) = (beta x, y)
// (Scott, eprint 2021/1130 sec. 6 -- the test is_torsion_free evaluates in the reference's dependency).
KZG_NI bool g1_in_subgroup(const G1Affine& a) {
    if (a.inf) return true;
    const uint32_t xabs[2] = {(uint32_t)(KZG_BLS_X_ABS & 0xffffffffu), (uint32_t)(KZG_BLS_X_ABS >> 32)};
    G1 t = scalar_mul_affine(a, xabs, 64);
    t = scalar_mul(t, xabs, 64).neg();
    const uint32_t beta[12] = KZG_FP_BETA_M;
    G1 e = {a.x * fp_const(beta), a.y, Fp::one()};
    return e.equals(t);
}

// 48 big-endian bytes (flag bits already stripped from b[0]) -> raw limbs
KZG_HD void be48_to_limbs(uint32_t* l, const uint8_t* b) {
    for (int i = 0; i < 12; i++) {
        const uint8_t* p = b + 4 * (11 - i);
        l[i] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
    }
}
// G1Affine::from_compressed: returns false for anything the reference rejects
// (uncompressed flag, bad infinity encoding, x >= p, x^3+4 not a square, not in the subgroup).
KZG_NI bool g1_from_compressed(G1Affine& out, const uint8_t* b, bool check_subgroup) {
    uint32_t comp = (b[0] >> 7) & 1, inf = (b[0] >> 6) & 1, sort = (b[0] >> 5) & 1;
    uint8_t xb[48];
    for (int i = 0; i < 48; i++) xb[i] = b[i];
    xb[0] &= 0x1f;
    Fp raw;
    be48_to_limbs(raw.l, xb);
    out = {Fp::zero(), Fp::zero(), 1};
    if (!comp) return false;
    if (inf) return raw.is_zero() && !sort;
    if (raw.geq_modulus()) return false;
    Fp x = Fp::from_raw(raw);
    Fp rhs = x.sqr() * x + Fp::from_u32(4);
    Fp y = fp_sqrt_candidate(rhs);
    if (!(y.sqr() == rhs)) return false;
    if (fp_lex_largest(y) != (sort != 0)) y = y.neg();
    out = {x, y, 0};
    if (check_subgroup && !g1_in_subgroup(out)) return false;
    return true;
}

// G1Affine::to_compressed (ZCash format): 48 big-endian bytes of x with the compression / infinity / sign flags
KZG_NI void g1_to_compressed(uint8_t* out, const G1Affine& a) {
    if (a.inf) { for (int i = 0; i < 48; i++) out[i] = 0; out[0] = 0xc0; return; }
    Fp x = a.x.to_raw();
    for (int i = 0; i < 12; i++) {
        uint8_t* p = out + 4 * (11 - i);
        p[0] = (uint8_t)(x.l[i] >> 24); p[1] = (uint8_t)(x.l[i] >> 16); p[2] = (uint8_t)(x.l[i] >> 8); p[3] = (uint8_t)x.l[i];
    }
    out[0] |= 0x80;
    if (fp_lex_largest(a.y)) out[0] |= 0x20;
}

// Fp2 square root for p = 3 mod 4 (complex method); false if not a square
KZG_NI bool fp2_sqrt(Fp2& out, const Fp2& a) {
    if (a.is_zero()) { out = a; return true; }
    const uint32_t e1[12] = KZG_FP_P_MINUS_3_DIV_4, e2[12] = KZG_FP_P_MINUS_1_HALF;
    auto pw = [](const Fp2& b, const uint32_t* e, int nbits) {
        Fp2 acc = Fp2::one();
        for (int i = nbits - 1; i >= 0; i--) {
            acc = acc.sqr();
            if ((e[i >> 5] >> (i & 31)) & 1) acc = acc * b;
        }
        return acc;
    };
    Fp2 a1 = pw(a, e1, 379), alpha = a1.sqr() * a, x0 = a1 * a, r;
    if (alpha == Fp2::one().neg()) r = {x0.c1.neg(), x0.c0};
    else r = pw(Fp2::one() + alpha, e2, 380) * x0;
    out = r;
    return r.sqr() == a;
}
KZG_HD bool fp2_lex_largest(const Fp2& a) { return a.c1.is_zero() ? fp_lex_largest(a.c0) : fp_lex_largest(a.c1); }

// G2Affine::from_compressed_unchecked (trusted-setup points; no subgroup check, as build.rs:73)
KZG_NI bool g2_from_compressed_unchecked(G2Affine& out, const uint8_t* b) {
    uint32_t comp = (b[0] >> 7) & 1, inf = (b[0] >> 6) & 1, sort = (b[0] >> 5) & 1;
    uint8_t xb[48];
    for (int i = 0; i < 48; i++) xb[i] = b[i];
    xb[0] &= 0x1f;
    Fp r1, r0;
    be48_to_limbs(r1.l, xb);
    be48_to_limbs(r0.l, b + 48);
    out = {Fp2::zero(), Fp2::zero(), 1};
    if (!comp) return false;
    if (inf) return true;
    if (r1.geq_modulus() || r0.geq_modulus()) return false;
    Fp2 x = {Fp::from_raw(r0), Fp::from_raw(r1)};
    Fp four = Fp::from_u32(4);
    Fp2 rhs = x.sqr() * x + Fp2{four, four}, y;
    if (!fp2_sqrt(y, rhs)) return false;
    if (fp2_lex_largest(y) != (sort != 0)) y = y.neg();
    out = {x, y, 0};
    return true;
}

}  // namespace kzgb200
