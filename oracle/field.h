/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/README.md).  Not part of the product path.
 *
 * Fp (381-bit, 6x64 limbs) and Fr (255-bit, 4x64 limbs) Montgomery arithmetic for BLS12-381.
 * Restates the field layer of the un-vendored dependency sp1_bls12_381 =0.8.0-sp1-6.0.0
 * (/root/reference/Cargo.toml:13; call sites kzg_proof.rs:36,90,112,124,127-130,176,188,196-197):
 * same representation (R = 2^384 / 2^256, little-endian u64 limbs) so Montgomery forms are
 * byte-identical to the reference's in-memory Scalar / Fp.
 */
#ifndef KZG_ORACLE_FIELD_H
#define KZG_ORACLE_FIELD_H
#include <stdint.h>
#include <string.h>
#include "consts64.h"

typedef unsigned __int128 u128;
typedef struct { uint64_t l[6]; } fp;
typedef struct { uint64_t l[4]; } fr;

/* ---- generic n-limb helpers (n is a compile-time constant at every call site) ---- */
static inline int bn_geq(const uint64_t *a, const uint64_t *b, int n) {
    for (int i = n - 1; i >= 0; i--) { if (a[i] != b[i]) return a[i] > b[i]; }
    return 1;
}
static inline uint64_t bn_add(uint64_t *r, const uint64_t *a, const uint64_t *b, int n) {
    u128 c = 0;
    for (int i = 0; i < n; i++) { c += (u128)a[i] + b[i]; r[i] = (uint64_t)c; c >>= 64; }
    return (uint64_t)c;
}
static inline uint64_t bn_sub(uint64_t *r, const uint64_t *a, const uint64_t *b, int n) {
    uint64_t bw = 0;
    for (int i = 0; i < n; i++) {
        u128 d = (u128)a[i] - b[i] - bw; r[i] = (uint64_t)d; bw = (uint64_t)(d >> 64) & 1;
    }
    return bw;
}
static inline int bn_is_zero(const uint64_t *a, int n) {
    uint64_t o = 0; for (int i = 0; i < n; i++) o |= a[i]; return o == 0;
}
static inline void mod_add(uint64_t *r, const uint64_t *a, const uint64_t *b, const uint64_t *m, int n) {
    uint64_t t[6]; uint64_t c = bn_add(t, a, b, n);
    if (c || bn_geq(t, m, n)) bn_sub(r, t, m, n); else memcpy(r, t, 8 * n);
}
static inline void mod_sub(uint64_t *r, const uint64_t *a, const uint64_t *b, const uint64_t *m, int n) {
    uint64_t t[6]; uint64_t bw = bn_sub(t, a, b, n);
    if (bw) bn_add(r, t, m, n); else memcpy(r, t, 8 * n);
}
/* Montgomery product r = a*b/2^(64n) mod m, CIOS with the "no final carry word" shortcut that is
 * valid because both moduli leave the top bit of the top limb clear (p < 2^381, q < 2^255). */
#define DEFINE_MONT_MUL(NAME, N)                                                              \
    static inline void NAME(uint64_t *r, const uint64_t *a, const uint64_t *b, const uint64_t *m, \
                            uint64_t inv) {                                                    \
        uint64_t t[N] = {0};                                                                   \
        _Pragma("GCC unroll 8") for (int i = 0; i < N; i++) {                                  \
            u128 A = (u128)a[0] * b[i] + t[0];                                                 \
            uint64_t q = (uint64_t)A * inv;                                                    \
            u128 C = (u128)q * m[0] + (uint64_t)A;                                             \
            _Pragma("GCC unroll 8") for (int j = 1; j < N; j++) {                              \
                A = (u128)a[j] * b[i] + t[j] + (uint64_t)(A >> 64);                            \
                C = (u128)q * m[j] + (uint64_t)A + (uint64_t)(C >> 64);                        \
                t[j - 1] = (uint64_t)C;                                                        \
            }                                                                                  \
            t[N - 1] = (uint64_t)(A >> 64) + (uint64_t)(C >> 64);                              \
        }                                                                                      \
        if (bn_geq(t, m, N)) bn_sub(r, t, m, N); else memcpy(r, t, 8 * N);                     \
    }
DEFINE_MONT_MUL(mont_mul6, 6)
DEFINE_MONT_MUL(mont_mul4, 4)

/* ---- Fp ---- */
static inline void fp_add(fp *r, const fp *a, const fp *b) { mod_add(r->l, a->l, b->l, FP_P, 6); }
static inline void fp_sub(fp *r, const fp *a, const fp *b) { mod_sub(r->l, a->l, b->l, FP_P, 6); }
static inline void fp_mul(fp *r, const fp *a, const fp *b) { mont_mul6(r->l, a->l, b->l, FP_P, FP_INV); }
static inline void fp_sqr(fp *r, const fp *a) { mont_mul6(r->l, a->l, a->l, FP_P, FP_INV); }
static inline int fp_is_zero(const fp *a) { return bn_is_zero(a->l, 6); }
static inline int fp_eq(const fp *a, const fp *b) { return memcmp(a, b, sizeof(fp)) == 0; }
static inline void fp_neg(fp *r, const fp *a) {
    if (fp_is_zero(a)) *r = *a; else bn_sub(r->l, FP_P, a->l, 6);
}
static inline void fp_dbl(fp *r, const fp *a) { fp_add(r, a, a); }
static inline void fp_set_zero(fp *r) { memset(r, 0, sizeof(fp)); }
static inline void fp_set_one(fp *r) { memcpy(r->l, FP_R, 48); }
static inline void fp_pow(fp *r, const fp *a, const uint64_t *e, int nlimbs) {
    fp acc; fp_set_one(&acc);
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
        fp_sqr(&acc, &acc);
        if ((e[i >> 6] >> (i & 63)) & 1) fp_mul(&acc, &acc, a);
    }
    *r = acc;
}
static inline void fp_inv(fp *r, const fp *a) { fp_pow(r, a, FP_P_MINUS_2, 6); }
/* returns 1 and r = sqrt(a) if a is a square */
static inline int fp_sqrt(fp *r, const fp *a) {
    fp s, t; fp_pow(&s, a, FP_SQRT_EXP, 6); fp_sqr(&t, &s);
    *r = s; return fp_eq(&t, a);
}
static inline void fp_from_mont(uint64_t out[6], const fp *a) {
    fp one = {{1, 0, 0, 0, 0, 0}}, t; fp_mul(&t, a, &one); memcpy(out, t.l, 48);
}
/* 48 big-endian bytes -> Montgomery; returns 0 if the value is not < p */
static inline int fp_from_bytes_be(fp *r, const uint8_t b[48]) {
    fp t, r2;
    for (int i = 0; i < 6; i++) {
        uint64_t w = 0; for (int j = 0; j < 8; j++) w = (w << 8) | b[(5 - i) * 8 + j];
        t.l[i] = w;
    }
    if (bn_geq(t.l, FP_P, 6)) return 0;
    memcpy(r2.l, FP_R2, 48); fp_mul(r, &t, &r2); return 1;
}
static inline void fp_to_bytes_be(uint8_t b[48], const fp *a) {
    uint64_t t[6]; fp_from_mont(t, a);
    for (int i = 0; i < 6; i++) for (int j = 0; j < 8; j++) b[(5 - i) * 8 + j] = (uint8_t)(t[i] >> (56 - 8 * j));
}
/* lexicographically largest: canonical value > (p-1)/2 */
static inline int fp_lex_largest(const fp *a) {
    uint64_t t[6]; fp_from_mont(t, a);
    for (int i = 5; i >= 0; i--) { if (t[i] != FP_P_MINUS_1_HALF[i]) return t[i] > FP_P_MINUS_1_HALF[i]; }
    return 0;
}

/* ---- Fr ---- */
static inline void fr_add(fr *r, const fr *a, const fr *b) { mod_add(r->l, a->l, b->l, FR_Q, 4); }
static inline void fr_sub(fr *r, const fr *a, const fr *b) { mod_sub(r->l, a->l, b->l, FR_Q, 4); }
static inline void fr_mul(fr *r, const fr *a, const fr *b) { mont_mul4(r->l, a->l, b->l, FR_Q, FR_INV); }
static inline void fr_sqr(fr *r, const fr *a) { mont_mul4(r->l, a->l, a->l, FR_Q, FR_INV); }
static inline int fr_is_zero(const fr *a) { return bn_is_zero(a->l, 4); }
static inline int fr_eq(const fr *a, const fr *b) { return memcmp(a, b, sizeof(fr)) == 0; }
static inline void fr_set_zero(fr *r) { memset(r, 0, sizeof(fr)); }
static inline void fr_set_one(fr *r) { memcpy(r->l, FR_R, 32); }
static inline void fr_pow(fr *r, const fr *a, const uint64_t *e, int nlimbs) {
    fr acc; fr_set_one(&acc);
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
        fr_sqr(&acc, &acc);
        if ((e[i >> 6] >> (i & 63)) & 1) fr_mul(&acc, &acc, a);
    }
    *r = acc;
}
static inline void fr_inv(fr *r, const fr *a) { fr_pow(r, a, FR_Q_MINUS_2, 4); }
/* Scalar::from_raw [dep]: raw little-endian limbs (any 256-bit value) -> value*R mod q */
static inline void fr_from_raw(fr *r, const uint64_t raw[4]) {
    fr t, r2; memcpy(t.l, raw, 32); memcpy(r2.l, FR_R2, 32); fr_mul(r, &r2, &t);  /* unreduced operand goes second (see mont_mul) */
}
static inline void fr_from_u64(fr *r, uint64_t v) { uint64_t raw[4] = {v, 0, 0, 0}; fr_from_raw(r, raw); }
static inline void fr_to_raw(uint64_t out[4], const fr *a) {
    fr one = {{1, 0, 0, 0}}, t; fr_mul(&t, a, &one); memcpy(out, t.l, 32);
}
/* 32 big-endian bytes -> limbs (no reduction) */
static inline void be32_to_limbs(uint64_t out[4], const uint8_t b[32]) {
    for (int i = 0; i < 4; i++) {
        uint64_t w = 0; for (int j = 0; j < 8; j++) w = (w << 8) | b[(3 - i) * 8 + j];
        out[i] = w;
    }
}
static inline void fr_to_bytes_be(uint8_t b[32], const fr *a) {
    uint64_t t[4]; fr_to_raw(t, a);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) b[(3 - i) * 8 + j] = (uint8_t)(t[i] >> (56 - 8 * j));
}
/* Scalar::to_bytes [dep]: canonical little-endian */
static inline void fr_to_bytes_le(uint8_t b[32], const fr *a) {
    uint64_t t[4]; fr_to_raw(t, a);
    for (int i = 0; i < 4; i++) for (int j = 0; j < 8; j++) b[i * 8 + j] = (uint8_t)(t[i] >> (8 * j));
}
#endif
