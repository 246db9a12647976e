/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  G1 / G2 of BLS12-381, ZCash compressed encodings,
 * subgroup checks, Pippenger MSM.  Restates G1Affine::from_compressed / to_compressed,
 * G1Projective::msm_variable_base, G2Affine::from_compressed_unchecked of sp1_bls12_381 [dep]
 * (call sites kzg_proof.rs:18,61,316,331,419,429,430; build.rs:68,73). */
#ifndef KZG_ORACLE_CURVE_H
#define KZG_ORACLE_CURVE_H
#include <stdlib.h>
#include "tower.h"

#define FE fp
#define PT g1
#define FN(x) g1_##x
#define g1_fe_set_one fp_set_one
#define g1_fe_set_zero fp_set_zero
#define g1_fe_is_zero fp_is_zero
#define g1_fe_inv fp_inv
#define g1_fe_sqr fp_sqr
#define g1_fe_mul fp_mul
#define g1_fe_add fp_add
#define g1_fe_sub fp_sub
#define g1_fe_neg fp_neg
#define g1_fe_eq fp_eq
#include "curve_tmpl.h"
#undef FE
#undef PT
#undef FN

#define FE fp2
#define PT g2
#define FN(x) g2_##x
#define g2_fe_set_one fp2_set_one
#define g2_fe_set_zero fp2_set_zero
#define g2_fe_is_zero fp2_is_zero
#define g2_fe_inv fp2_inv
#define g2_fe_sqr fp2_sqr
#define g2_fe_mul fp2_mul
#define g2_fe_add fp2_add
#define g2_fe_sub fp2_sub
#define g2_fe_neg fp2_neg
#define g2_fe_eq fp2_eq
#include "curve_tmpl.h"
#undef FE
#undef PT
#undef FN

static inline void g1_generator(g1_aff *r) { memcpy(r->x.l, FP_G1X_M, 48); memcpy(r->y.l, FP_G1Y_M, 48); r->inf = 0; }
static inline void g2_generator(g2_aff *r) {
    fp2_load_m(&r->x, FP_G2X0_M, FP_G2X1_M); fp2_load_m(&r->y, FP_G2Y0_M, FP_G2Y1_M); r->inf = 0;
}
static inline void fp_set_u64(fp *r, uint64_t v) {
    fp t = {{v, 0, 0, 0, 0, 0}}, r2; memcpy(r2.l, FP_R2, 48); fp_mul(r, &t, &r2);
}
static inline int g1_aff_on_curve(const g1_aff *a) {
    if (a->inf) return 1;
    fp l, r, four; fp_sqr(&l, &a->y); fp_sqr(&r, &a->x); fp_mul(&r, &r, &a->x);
    fp_set_u64(&four, 4); fp_add(&r, &r, &four); return fp_eq(&l, &r);
}
/* [q]P == O : definition of the prime-order subgroup (slow path, used to cross-check the fast one) */
static inline int g1_in_subgroup_naive(const g1_aff *a) {
    g1 p, r; g1_from_aff(&p, a); g1_mul(&r, &p, FR_Q, 4); return g1_is_inf(&r);
}
/* Endomorphism test (Scott, eprint 2021/1130 sec. 6; what is_torsion_free [dep] evaluates):
 * phi(x,y) = (beta x, y) acts on G1 as [-x^2]; P in G1  <=>  phi(P) == -[x^2]P. */
static inline int g1_in_subgroup(const g1_aff *a) {
    if (a->inf) return 1;
    g1 p, t; g1_from_aff(&p, a);
    uint64_t x = BLS_X_ABS;
    g1_mul(&t, &p, &x, 1); g1_mul(&t, &t, &x, 1); g1_neg(&t, &t);
    g1 e = p; fp beta; memcpy(beta.l, FP_BETA_M, 48); fp_mul(&e.x, &e.x, &beta);
    return g1_eq(&e, &t);
}
/* G1Affine::from_compressed [dep]; returns 1 on success */
static inline int g1_from_compressed(g1_aff *r, const uint8_t b[48], int check_subgroup) {
    int comp = (b[0] >> 7) & 1, inf = (b[0] >> 6) & 1, sort = (b[0] >> 5) & 1;
    uint8_t xb[48]; memcpy(xb, b, 48); xb[0] &= 0x1f;
    if (!comp) return 0;
    if (inf) {
        for (int i = 0; i < 48; i++) if (xb[i]) return 0;
        if (sort) return 0;
        r->inf = 1; fp_set_zero(&r->x); fp_set_zero(&r->y); return 1;
    }
    fp x, y, t, four;
    if (!fp_from_bytes_be(&x, xb)) return 0;
    fp_sqr(&t, &x); fp_mul(&t, &t, &x); fp_set_u64(&four, 4); fp_add(&t, &t, &four);
    if (!fp_sqrt(&y, &t)) return 0;
    if (fp_lex_largest(&y) != sort) fp_neg(&y, &y);
    r->x = x; r->y = y; r->inf = 0;
    if (check_subgroup && !g1_in_subgroup(r)) return 0;
    return 1;
}
static inline void g1_to_compressed(uint8_t b[48], const g1_aff *a) {
    if (a->inf) { memset(b, 0, 48); b[0] = 0xc0; return; }
    fp_to_bytes_be(b, &a->x); b[0] |= 0x80;
    if (fp_lex_largest(&a->y)) b[0] |= 0x20;
}
/* G2Affine::from_compressed_unchecked [dep] (build.rs:73): no subgroup check */
static inline int g2_from_compressed_unchecked(g2_aff *r, const uint8_t b[96]) {
    int comp = (b[0] >> 7) & 1, inf = (b[0] >> 6) & 1, sort = (b[0] >> 5) & 1;
    uint8_t xb[48]; memcpy(xb, b, 48); xb[0] &= 0x1f;
    if (!comp) return 0;
    if (inf) { r->inf = 1; fp2_set_zero(&r->x); fp2_set_zero(&r->y); return 1; }
    fp2 x, y, t, b2;
    if (!fp_from_bytes_be(&x.c1, xb)) return 0;
    if (!fp_from_bytes_be(&x.c0, b + 48)) return 0;
    fp2_sqr(&t, &x); fp2_mul(&t, &t, &x); fp_set_u64(&b2.c0, 4); b2.c1 = b2.c0; fp2_add(&t, &t, &b2);
    if (!fp2_sqrt(&y, &t)) return 0;
    if (fp2_lex_largest(&y) != sort) fp2_neg(&y, &y);
    r->x = x; r->y = y; r->inf = 0; return 1;
}

/* Pippenger bucket MSM: sum scalars[i] * points[i]; scalars are canonical (non-Montgomery) 4x64 limbs */
static inline void g1_msm(g1 *out, const g1_aff *pts, const uint64_t (*sc)[4], size_t n) {
    g1 acc; g1_set_inf(&acc);
    if (n == 0) { *out = acc; return; }
    int c = n < 8 ? 2 : n < 64 ? 4 : n < 1024 ? 7 : n < 8192 ? 9 : 11;
    int nwin = (255 + c - 1) / c;
    size_t nb = ((size_t)1 << c) - 1;
    g1 *buckets = (g1 *)malloc(nb * sizeof(g1));
    for (int w = nwin - 1; w >= 0; w--) {
        for (int k = 0; k < c; k++) g1_dbl(&acc, &acc);
        for (size_t b = 0; b < nb; b++) g1_set_inf(&buckets[b]);
        int lo = w * c;
        for (size_t i = 0; i < n; i++) {
            uint64_t d = sc[i][lo >> 6] >> (lo & 63);
            if ((lo & 63) + c > 64 && (lo >> 6) + 1 < 4) d |= sc[i][(lo >> 6) + 1] << (64 - (lo & 63));
            d &= nb;
            if (d) g1_add_mixed(&buckets[d - 1], &buckets[d - 1], &pts[i]);
        }
        g1 run, sum; g1_set_inf(&run); g1_set_inf(&sum);
        for (size_t b = nb; b-- > 0;) { g1_add(&run, &run, &buckets[b]); g1_add(&sum, &sum, &run); }
        g1_add(&acc, &acc, &sum);
    }
    free(buckets);
    *out = acc;
}
#endif
