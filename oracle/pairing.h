/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Optimal-ate pairing check on BLS12-381.
 * Restates multi_miller_loop + final_exponentiation + "== Gt::identity()" of sp1_bls12_381 [dep] as
 * used by /root/reference/src/pairings.rs:5-9.  Only the boolean e(..)e(..) == 1 is ever compared,
 * so any correct pairing (here: f^(3(p^12-1)/r)) gives the reference's verdict. */
#ifndef KZG_ORACLE_PAIRING_H
#define KZG_ORACLE_PAIRING_H
#include "curve.h"

/* f *= line(P), line = A + (B xP) v + (C yP) vw */
static inline void ell(fp12 *f, const fp2 *A, const fp2 *B, const fp2 *C, const g1_aff *P) {
    fp2 b, c; fp2_mul_fp(&b, B, &P->x); fp2_mul_fp(&c, C, &P->y);
    fp12_mul_by_014(f, f, A, &b, &c);
}
/* tangent at T (Jacobian on the twist): A = 3X^3 - 2Y^2, B = -3X^2 Z^2, C = 2YZ * Z^2 ; T <- 2T */
static inline void miller_dbl_step(g2 *T, fp2 *A, fp2 *B, fp2 *C) {
    fp2 X2, Z2, t, u;
    fp2_sqr(&X2, &T->x); fp2_sqr(&Z2, &T->z);
    fp2_add(&t, &X2, &X2); fp2_add(&t, &t, &X2);          /* 3X^2 */
    fp2_mul(A, &t, &T->x); fp2_sqr(&u, &T->y); fp2_add(&u, &u, &u); fp2_sub(A, A, &u);
    fp2_mul(B, &t, &Z2); fp2_neg(B, B);
    g2_dbl(T, T);
    fp2_mul(C, &T->z, &Z2);
}
/* chord through T and affine Q: theta = yQ Z^3 - Y, H = xQ Z^2 - X; A = theta xQ - yQ Z H, B = -theta, C = Z H */
static inline void miller_add_step(g2 *T, const g2_aff *Q, fp2 *A, fp2 *B, fp2 *C) {
    fp2 Z2, th, H, t;
    fp2_sqr(&Z2, &T->z);
    fp2_mul(&th, &Q->y, &Z2); fp2_mul(&th, &th, &T->z); fp2_sub(&th, &th, &T->y);
    fp2_mul(&H, &Q->x, &Z2); fp2_sub(&H, &H, &T->x);
    fp2_mul(C, &T->z, &H);
    fp2_mul(A, &th, &Q->x); fp2_mul(&t, &Q->y, C); fp2_sub(A, A, &t);
    fp2_neg(B, &th);
    g2_add_mixed(T, T, Q);
}
/* product of Miller functions; pairs with an identity member contribute 1 (as [dep] does) */
static inline void multi_miller_loop(fp12 *out, const g1_aff *Ps, const g2_aff *Qs, int n) {
    fp12 f; fp12_set_one(&f);
    g2 T[4]; int live[4];
    for (int i = 0; i < n; i++) { live[i] = !(Ps[i].inf || Qs[i].inf); if (live[i]) g2_from_aff(&T[i], &Qs[i]); }
    fp2 A, B, C;
    for (int bit = 62; bit >= 0; bit--) {
        fp12_sqr(&f, &f);
        for (int i = 0; i < n; i++) if (live[i]) { miller_dbl_step(&T[i], &A, &B, &C); ell(&f, &A, &B, &C, &Ps[i]); }
        if ((BLS_X_ABS >> bit) & 1)
            for (int i = 0; i < n; i++) if (live[i]) { miller_add_step(&T[i], &Qs[i], &A, &B, &C); ell(&f, &A, &B, &C, &Ps[i]); }
    }
    fp12_conj(out, &f);   /* x < 0 */
}
static inline void fp12_exp_by_x(fp12 *r, const fp12 *a) {
    fp12 acc = *a;
    for (int bit = 62; bit >= 0; bit--) {
        fp12_sqr(&acc, &acc);
        if ((BLS_X_ABS >> bit) & 1) fp12_mul(&acc, &acc, a);
    }
    fp12_conj(r, &acc);
}
/* f^(3 (p^12-1)/r); hard part from 3(p^4-p^2+1)/r = (x-1)^2 (x+p)(x^2+p^2-1) + 3 */
static inline void final_exponentiation(fp12 *r, const fp12 *f0) {
    fp12 f, t, a, b, c, u;
    fp12_conj(&t, f0); fp12_inv(&f, f0); fp12_mul(&f, &t, &f);        /* ^(p^6-1) */
    fp12_frob(&t, &f); fp12_frob(&t, &t); fp12_mul(&f, &t, &f);        /* ^(p^2+1) */
    fp12_exp_by_x(&a, &f); fp12_conj(&t, &f); fp12_mul(&a, &a, &t);    /* f^(x-1) */
    fp12_exp_by_x(&u, &a); fp12_conj(&t, &a); fp12_mul(&a, &u, &t);    /* ^(x-1) */
    fp12_exp_by_x(&b, &a); fp12_frob(&t, &a); fp12_mul(&b, &b, &t);    /* ^(x+p) */
    fp12_exp_by_x(&c, &b); fp12_exp_by_x(&c, &c);
    fp12_frob(&t, &b); fp12_frob(&t, &t); fp12_mul(&c, &c, &t);
    fp12_conj(&t, &b); fp12_mul(&c, &c, &t);                           /* b^(x^2+p^2-1) */
    fp12_sqr(&t, &f); fp12_mul(&t, &t, &f); fp12_mul(r, &c, &t);       /* * f^3 */
}
/* pairings.rs:5-9 : e(-a1, a2) * e(b1, b2) == 1 */
static inline int pairings_verify(const g1_aff *a1, const g2_aff *a2, const g1_aff *b1, const g2_aff *b2) {
    g1_aff Ps[2] = {*a1, *b1}; g2_aff Qs[2] = {*a2, *b2};
    if (!Ps[0].inf) fp_neg(&Ps[0].y, &Ps[0].y);
    fp12 f, one; multi_miller_loop(&f, Ps, Qs, 2); final_exponentiation(&f, &f);
    fp12_set_one(&one); return fp12_eq(&f, &one);
}
#endif
