/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Fp2 / Fp6 / Fp12 tower for BLS12-381:
 *   Fp2 = Fp[u]/(u^2+1),  Fp6 = Fp2[v]/(v^3 - (1+u)),  Fp12 = Fp6[w]/(w^2 - v).
 * Restates the extension-field layer of sp1_bls12_381 [dep] that pairings.rs:5-9 relies on
 * (multi_miller_loop / final_exponentiation / Gt::identity). */
#ifndef KZG_ORACLE_TOWER_H
#define KZG_ORACLE_TOWER_H
#include "field.h"

typedef struct { fp c0, c1; } fp2;
typedef struct { fp2 c0, c1, c2; } fp6;
typedef struct { fp6 c0, c1; } fp12;

static inline void fp2_add(fp2 *r, const fp2 *a, const fp2 *b) { fp_add(&r->c0, &a->c0, &b->c0); fp_add(&r->c1, &a->c1, &b->c1); }
static inline void fp2_sub(fp2 *r, const fp2 *a, const fp2 *b) { fp_sub(&r->c0, &a->c0, &b->c0); fp_sub(&r->c1, &a->c1, &b->c1); }
static inline void fp2_neg(fp2 *r, const fp2 *a) { fp_neg(&r->c0, &a->c0); fp_neg(&r->c1, &a->c1); }
static inline void fp2_dbl(fp2 *r, const fp2 *a) { fp2_add(r, a, a); }
static inline void fp2_conj(fp2 *r, const fp2 *a) { r->c0 = a->c0; fp_neg(&r->c1, &a->c1); }
static inline int fp2_is_zero(const fp2 *a) { return fp_is_zero(&a->c0) && fp_is_zero(&a->c1); }
static inline int fp2_eq(const fp2 *a, const fp2 *b) { return fp_eq(&a->c0, &b->c0) && fp_eq(&a->c1, &b->c1); }
static inline void fp2_set_zero(fp2 *r) { fp_set_zero(&r->c0); fp_set_zero(&r->c1); }
static inline void fp2_set_one(fp2 *r) { fp_set_one(&r->c0); fp_set_zero(&r->c1); }
static inline void fp2_mul(fp2 *r, const fp2 *a, const fp2 *b) {
    fp t0, t1, s0, s1, m;
    fp_mul(&t0, &a->c0, &b->c0); fp_mul(&t1, &a->c1, &b->c1);
    fp_add(&s0, &a->c0, &a->c1); fp_add(&s1, &b->c0, &b->c1); fp_mul(&m, &s0, &s1);
    fp_sub(&r->c0, &t0, &t1); fp_sub(&m, &m, &t0); fp_sub(&r->c1, &m, &t1);
}
static inline void fp2_sqr(fp2 *r, const fp2 *a) {
    fp s, d, m;
    fp_add(&s, &a->c0, &a->c1); fp_sub(&d, &a->c0, &a->c1); fp_mul(&m, &a->c0, &a->c1);
    fp_mul(&r->c0, &s, &d); fp_dbl(&r->c1, &m);
}
static inline void fp2_mul_fp(fp2 *r, const fp2 *a, const fp *k) { fp_mul(&r->c0, &a->c0, k); fp_mul(&r->c1, &a->c1, k); }
/* multiply by xi = 1 + u */
static inline void fp2_mul_xi(fp2 *r, const fp2 *a) {
    fp t0, t1; fp_sub(&t0, &a->c0, &a->c1); fp_add(&t1, &a->c0, &a->c1); r->c0 = t0; r->c1 = t1;
}
static inline void fp2_inv(fp2 *r, const fp2 *a) {
    fp n, t; fp_sqr(&n, &a->c0); fp_sqr(&t, &a->c1); fp_add(&n, &n, &t); fp_inv(&n, &n);
    fp_mul(&r->c0, &a->c0, &n); fp_mul(&t, &a->c1, &n); fp_neg(&r->c1, &t);
}
static inline void fp2_pow(fp2 *r, const fp2 *a, const uint64_t *e, int nlimbs) {
    fp2 acc; fp2_set_one(&acc);
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
        fp2_sqr(&acc, &acc);
        if ((e[i >> 6] >> (i & 63)) & 1) fp2_mul(&acc, &acc, a);
    }
    *r = acc;
}
/* sqrt in Fp2 for p = 3 mod 4; returns 1 on success */
static inline int fp2_sqrt(fp2 *r, const fp2 *a) {
    if (fp2_is_zero(a)) { *r = *a; return 1; }
    fp2 a1, alpha, x0, t, minus_one;
    fp2_pow(&a1, a, FP_P_MINUS_3_DIV_4, 6);
    fp2_sqr(&alpha, &a1); fp2_mul(&alpha, &alpha, a);
    fp2_mul(&x0, &a1, a);
    fp2_set_one(&minus_one); fp2_neg(&minus_one, &minus_one);
    if (fp2_eq(&alpha, &minus_one)) {
        /* multiply x0 by u */
        fp_neg(&t.c0, &x0.c1); t.c1 = x0.c0;
    } else {
        fp2 b; fp2_set_one(&b); fp2_add(&b, &b, &alpha);
        fp2_pow(&b, &b, FP_P_MINUS_1_HALF, 6);
        fp2_mul(&t, &b, &x0);
    }
    fp2 chk; fp2_sqr(&chk, &t); *r = t;
    return fp2_eq(&chk, a);
}
static inline int fp2_lex_largest(const fp2 *a) {
    if (!fp_is_zero(&a->c1)) return fp_lex_largest(&a->c1);
    return fp_lex_largest(&a->c0);
}

/* ---- Fp6 ---- */
static inline void fp6_add(fp6 *r, const fp6 *a, const fp6 *b) { fp2_add(&r->c0, &a->c0, &b->c0); fp2_add(&r->c1, &a->c1, &b->c1); fp2_add(&r->c2, &a->c2, &b->c2); }
static inline void fp6_sub(fp6 *r, const fp6 *a, const fp6 *b) { fp2_sub(&r->c0, &a->c0, &b->c0); fp2_sub(&r->c1, &a->c1, &b->c1); fp2_sub(&r->c2, &a->c2, &b->c2); }
static inline void fp6_neg(fp6 *r, const fp6 *a) { fp2_neg(&r->c0, &a->c0); fp2_neg(&r->c1, &a->c1); fp2_neg(&r->c2, &a->c2); }
static inline void fp6_set_zero(fp6 *r) { fp2_set_zero(&r->c0); fp2_set_zero(&r->c1); fp2_set_zero(&r->c2); }
static inline void fp6_set_one(fp6 *r) { fp2_set_one(&r->c0); fp2_set_zero(&r->c1); fp2_set_zero(&r->c2); }
static inline int fp6_eq(const fp6 *a, const fp6 *b) { return fp2_eq(&a->c0, &b->c0) && fp2_eq(&a->c1, &b->c1) && fp2_eq(&a->c2, &b->c2); }
static inline void fp6_mul(fp6 *r, const fp6 *a, const fp6 *b) {
    fp2 t0, t1, t2, s, u, m, c0, c1, c2;
    fp2_mul(&t0, &a->c0, &b->c0); fp2_mul(&t1, &a->c1, &b->c1); fp2_mul(&t2, &a->c2, &b->c2);
    fp2_add(&s, &a->c1, &a->c2); fp2_add(&u, &b->c1, &b->c2); fp2_mul(&m, &s, &u);
    fp2_sub(&m, &m, &t1); fp2_sub(&m, &m, &t2); fp2_mul_xi(&m, &m); fp2_add(&c0, &t0, &m);
    fp2_add(&s, &a->c0, &a->c1); fp2_add(&u, &b->c0, &b->c1); fp2_mul(&m, &s, &u);
    fp2_sub(&m, &m, &t0); fp2_sub(&m, &m, &t1); fp2_mul_xi(&s, &t2); fp2_add(&c1, &m, &s);
    fp2_add(&s, &a->c0, &a->c2); fp2_add(&u, &b->c0, &b->c2); fp2_mul(&m, &s, &u);
    fp2_sub(&m, &m, &t0); fp2_sub(&m, &m, &t2); fp2_add(&c2, &m, &t1);
    r->c0 = c0; r->c1 = c1; r->c2 = c2;
}
static inline void fp6_mul_v(fp6 *r, const fp6 *a) {
    fp2 t; fp2_mul_xi(&t, &a->c2); fp2 a0 = a->c0, a1 = a->c1; r->c0 = t; r->c1 = a0; r->c2 = a1;
}
static inline void fp6_inv(fp6 *r, const fp6 *a) {
    fp2 c0, c1, c2, t, u;
    fp2_sqr(&c0, &a->c0); fp2_mul(&t, &a->c1, &a->c2); fp2_mul_xi(&t, &t); fp2_sub(&c0, &c0, &t);
    fp2_sqr(&c1, &a->c2); fp2_mul_xi(&c1, &c1); fp2_mul(&t, &a->c0, &a->c1); fp2_sub(&c1, &c1, &t);
    fp2_sqr(&c2, &a->c1); fp2_mul(&t, &a->c0, &a->c2); fp2_sub(&c2, &c2, &t);
    fp2_mul(&t, &a->c2, &c1); fp2_mul(&u, &a->c1, &c2); fp2_add(&t, &t, &u); fp2_mul_xi(&t, &t);
    fp2_mul(&u, &a->c0, &c0); fp2_add(&t, &t, &u); fp2_inv(&t, &t);
    fp2_mul(&r->c0, &c0, &t); fp2_mul(&r->c1, &c1, &t); fp2_mul(&r->c2, &c2, &t);
}

/* ---- Fp12 ---- */
static inline void fp12_set_one(fp12 *r) { fp6_set_one(&r->c0); fp6_set_zero(&r->c1); }
static inline int fp12_eq(const fp12 *a, const fp12 *b) { return fp6_eq(&a->c0, &b->c0) && fp6_eq(&a->c1, &b->c1); }
static inline void fp12_mul(fp12 *r, const fp12 *a, const fp12 *b) {
    fp6 t0, t1, s, u, m;
    fp6_mul(&t0, &a->c0, &b->c0); fp6_mul(&t1, &a->c1, &b->c1);
    fp6_add(&s, &a->c0, &a->c1); fp6_add(&u, &b->c0, &b->c1); fp6_mul(&m, &s, &u);
    fp6_sub(&m, &m, &t0); fp6_sub(&m, &m, &t1);
    fp6_mul_v(&t1, &t1); fp6_add(&r->c0, &t0, &t1); r->c1 = m;
}
static inline void fp12_sqr(fp12 *r, const fp12 *a) {
    /* (a0 + a1 w)^2 = (a0^2 + v a1^2) + 2 a0 a1 w, via (a0+a1)(a0+v a1) */
    fp6 ab, s, t, va1;
    fp6_mul(&ab, &a->c0, &a->c1);
    fp6_add(&s, &a->c0, &a->c1); fp6_mul_v(&va1, &a->c1); fp6_add(&t, &a->c0, &va1);
    fp6_mul(&s, &s, &t); fp6_sub(&s, &s, &ab); fp6_mul_v(&t, &ab); fp6_sub(&r->c0, &s, &t);
    fp6_add(&r->c1, &ab, &ab);
}
static inline void fp12_conj(fp12 *r, const fp12 *a) { r->c0 = a->c0; fp6_neg(&r->c1, &a->c1); }
static inline void fp12_inv(fp12 *r, const fp12 *a) {
    fp6 t0, t1;
    fp6_mul(&t0, &a->c0, &a->c0); fp6_mul(&t1, &a->c1, &a->c1); fp6_mul_v(&t1, &t1); fp6_sub(&t0, &t0, &t1);
    fp6_inv(&t0, &t0);
    fp6_mul(&r->c0, &a->c0, &t0); fp6_mul(&t1, &a->c1, &t0); fp6_neg(&r->c1, &t1);
}
/* line value A + B v + C vw (sparse in positions c0.c0, c0.c1, c1.c1) times f */
static inline void fp12_mul_by_014(fp12 *r, const fp12 *f, const fp2 *A, const fp2 *B, const fp2 *C) {
    fp12 l; fp6_set_zero(&l.c0); fp6_set_zero(&l.c1);
    l.c0.c0 = *A; l.c0.c1 = *B; l.c1.c1 = *C;
    fp12_mul(r, f, &l);
}
static inline void fp2_load_m(fp2 *r, const uint64_t *c0, const uint64_t *c1) { memcpy(r->c0.l, c0, 48); memcpy(r->c1.l, c1, 48); }
/* Frobenius x -> x^p: conjugate every Fp2 coefficient and scale the coefficient of w^k by xi^(k(p-1)/6) */
static inline void fp12_frob(fp12 *r, const fp12 *a) {
    fp2 g1, g2, g3, g4, g5, t;
    fp2_load_m(&g1, FP_FROB6_1_C0_M, FP_FROB6_1_C1_M); fp2_load_m(&g2, FP_FROB6_2_C0_M, FP_FROB6_2_C1_M);
    fp2_load_m(&g3, FP_FROB6_3_C0_M, FP_FROB6_3_C1_M); fp2_load_m(&g4, FP_FROB6_4_C0_M, FP_FROB6_4_C1_M);
    fp2_load_m(&g5, FP_FROB6_5_C0_M, FP_FROB6_5_C1_M);
    fp2_conj(&r->c0.c0, &a->c0.c0);
    fp2_conj(&t, &a->c0.c1); fp2_mul(&r->c0.c1, &t, &g2);
    fp2_conj(&t, &a->c0.c2); fp2_mul(&r->c0.c2, &t, &g4);
    fp2_conj(&t, &a->c1.c0); fp2_mul(&r->c1.c0, &t, &g1);
    fp2_conj(&t, &a->c1.c1); fp2_mul(&r->c1.c1, &t, &g3);
    fp2_conj(&t, &a->c1.c2); fp2_mul(&r->c1.c2, &t, &g5);
}
#endif
