// GLV split of an Fr scalar for G1 of BLS12-381.  The endomorphism phi(x, y) = (beta x, y) acts on G1 as
// [-x^2] (x = the curve parameter; the identity the subgroup check already uses), so with X2 = x^2 (128 bits)
//     k = k1 + k2 * X2   (plain Euclidean division, k1 < X2, k2 < 2^128)     =>     [k]P = [k1]P + [k2](-phi(P)),
// -phi(P) = (beta x, -y).  Halves the number of 8-bit windows (32 -> 16) and with it the serial doubling chain
// of the Horner recombination.
#pragma once
#include "curve.cuh"

namespace kzgb200 {

// k: 8 canonical limbs (< q).  k1, k2: 4 limbs each.
KZG_HD void glv_split(const uint32_t* k, uint32_t* k1, uint32_t* k2) {
    const uint32_t X2[4] = {0x00000000u, 0x00000001u, 0x0001a402u, 0xac45a401u};                   // x^2
    const uint32_t M[8] = {0x40c5f204u, 0xd0d4396bu, 0x93d6e013u, 0x01a75a5cu, 0x7b67f717u, 0xb1fb7291u, 0xf00fd56eu, 0xbe35f678u};  // floor(2^383 / x^2)
    // prod = k * M (16 limbs)
    uint32_t prod[16];
    for (int i = 0; i < 16; i++) prod[i] = 0;
    for (int i = 0; i < 8; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 8; j++) { c += (uint64_t)k[i] * M[j] + prod[i + j]; prod[i + j] = (uint32_t)c; c >>= 32; }
        prod[i + 8] = (uint32_t)c;
    }
    // q = prod >> 383  (limb 11 bit 31 upward); q < 2^128 + small
    uint32_t q[5];
    for (int i = 0; i < 5; i++) q[i] = (prod[11 + i] >> 31) | (i + 12 < 16 ? (prod[12 + i] << 1) : 0u);
    // rem = k - q * X2  (low 8 limbs are enough: 0 <= rem < 3 X2)
    uint32_t qx[9];
    for (int i = 0; i < 9; i++) qx[i] = 0;
    for (int i = 0; i < 5; i++) {
        uint64_t c = 0;
        for (int j = 0; j < 4 && i + j < 9; j++) { c += (uint64_t)q[i] * X2[j] + qx[i + j]; qx[i + j] = (uint32_t)c; c >>= 32; }
        if (i + 4 < 9) qx[i + 4] = (uint32_t)c;
    }
    uint32_t rem[8];
    uint64_t bw = 0;
    for (int i = 0; i < 8; i++) { uint64_t d = (uint64_t)k[i] - qx[i] - bw; rem[i] = (uint32_t)d; bw = (d >> 32) & 1; }
    // at most two corrections
    for (int r = 0; r < 2; r++) {
        bool ge = rem[4] | rem[5] | rem[6] | rem[7];
        if (!ge) {
            ge = true;
            for (int i = 3; i >= 0; i--) { if (rem[i] != X2[i]) { ge = rem[i] > X2[i]; break; } }
        }
        if (ge) {
            uint64_t b2 = 0;
            for (int i = 0; i < 8; i++) { uint64_t d = (uint64_t)rem[i] - (i < 4 ? X2[i] : 0u) - b2; rem[i] = (uint32_t)d; b2 = (d >> 32) & 1; }
            uint64_t c = 1;
            for (int i = 0; i < 5; i++) { c += q[i]; q[i] = (uint32_t)c; c >>= 32; }
        }
    }
    for (int i = 0; i < 4; i++) { k1[i] = rem[i]; k2[i] = q[i]; }
}
// -phi(P) = (beta x, -y)
KZG_HD G1Affine glv_endo_neg(const G1Affine& p) {
    if (p.inf) return p;
    const uint32_t beta[12] = KZG_FP_BETA_M;
    return {p.x * fp_const(beta), p.y.neg(), 0};
}

}  // namespace kzgb200
