/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Short-Weierstrass (a = 0) group law in Jacobian
 * coordinates, instantiated twice by curve.h: G1 over Fp and G2 over Fp2.  Restates the group layer
 * of sp1_bls12_381 [dep] (G1Affine/G1Projective/G2Affine/G2Projective; call sites
 * kzg_proof.rs:210-214,385-389,419-433).  Include with FE, FN(x), PT defined. */
typedef struct { FE x, y, z; } PT;        /* z == 0 <=> identity */
typedef struct { FE x, y; int inf; } FN(aff);

static inline void FN(set_inf)(PT *r) { FN(fe_set_one)(&r->x); FN(fe_set_one)(&r->y); FN(fe_set_zero)(&r->z); }
static inline int FN(is_inf)(const PT *a) { return FN(fe_is_zero)(&a->z); }
static inline void FN(from_aff)(PT *r, const FN(aff) *a) {
    if (a->inf) { FN(set_inf)(r); return; }
    r->x = a->x; r->y = a->y; FN(fe_set_one)(&r->z);
}
static inline void FN(to_aff)(FN(aff) *r, const PT *a) {
    if (FN(is_inf)(a)) { r->inf = 1; FN(fe_set_zero)(&r->x); FN(fe_set_zero)(&r->y); return; }
    FE zi, zi2; FN(fe_inv)(&zi, &a->z); FN(fe_sqr)(&zi2, &zi);
    FN(fe_mul)(&r->x, &a->x, &zi2); FN(fe_mul)(&zi2, &zi2, &zi); FN(fe_mul)(&r->y, &a->y, &zi2); r->inf = 0;
}
static inline void FN(neg)(PT *r, const PT *a) { r->x = a->x; r->z = a->z; FN(fe_neg)(&r->y, &a->y); }
static inline void FN(dbl)(PT *r, const PT *p) {
    if (FN(is_inf)(p)) { *r = *p; return; }
    FE A, B, C, D, E, F, t, z3;
    FN(fe_sqr)(&A, &p->x); FN(fe_sqr)(&B, &p->y); FN(fe_sqr)(&C, &B);
    FN(fe_add)(&t, &p->x, &B); FN(fe_sqr)(&t, &t); FN(fe_sub)(&t, &t, &A); FN(fe_sub)(&t, &t, &C); FN(fe_add)(&D, &t, &t);
    FN(fe_add)(&E, &A, &A); FN(fe_add)(&E, &E, &A); FN(fe_sqr)(&F, &E);
    FN(fe_mul)(&z3, &p->y, &p->z); FN(fe_add)(&z3, &z3, &z3);
    FN(fe_sub)(&r->x, &F, &D); FN(fe_sub)(&r->x, &r->x, &D);
    FN(fe_sub)(&t, &D, &r->x); FN(fe_mul)(&t, &t, &E);
    FN(fe_add)(&C, &C, &C); FN(fe_add)(&C, &C, &C); FN(fe_add)(&C, &C, &C);
    FN(fe_sub)(&r->y, &t, &C); r->z = z3;
}
/* r = p + (x2, y2) with (x2,y2) affine, not the identity */
static inline void FN(add_mixed)(PT *r, const PT *p, const FN(aff) *q) {
    if (q->inf) { *r = *p; return; }
    if (FN(is_inf)(p)) { FN(from_aff)(r, q); return; }
    FE Z2, U2, S2, H, R, H2, H3, t, u;
    FN(fe_sqr)(&Z2, &p->z); FN(fe_mul)(&U2, &q->x, &Z2);
    FN(fe_mul)(&S2, &q->y, &Z2); FN(fe_mul)(&S2, &S2, &p->z);
    FN(fe_sub)(&H, &U2, &p->x); FN(fe_sub)(&R, &S2, &p->y);
    if (FN(fe_is_zero)(&H)) { if (FN(fe_is_zero)(&R)) FN(dbl)(r, p); else FN(set_inf)(r); return; }
    FN(fe_sqr)(&H2, &H); FN(fe_mul)(&H3, &H2, &H);
    FN(fe_mul)(&t, &p->x, &H2);                 /* X H^2 */
    FN(fe_sqr)(&u, &R); FN(fe_sub)(&u, &u, &H3); FN(fe_sub)(&u, &u, &t); FN(fe_sub)(&u, &u, &t);  /* X3 */
    FN(fe_sub)(&t, &t, &u); FN(fe_mul)(&t, &t, &R); FN(fe_mul)(&H3, &H3, &p->y);
    FN(fe_mul)(&r->z, &p->z, &H); r->x = u; FN(fe_sub)(&r->y, &t, &H3);
}
static inline void FN(add)(PT *r, const PT *p, const PT *q) {
    if (FN(is_inf)(p)) { *r = *q; return; }
    if (FN(is_inf)(q)) { *r = *p; return; }
    FE Z1Z1, Z2Z2, U1, U2, S1, S2, H, R, H2, H3, t, u, z3;
    FN(fe_sqr)(&Z1Z1, &p->z); FN(fe_sqr)(&Z2Z2, &q->z);
    FN(fe_mul)(&U1, &p->x, &Z2Z2); FN(fe_mul)(&U2, &q->x, &Z1Z1);
    FN(fe_mul)(&S1, &p->y, &Z2Z2); FN(fe_mul)(&S1, &S1, &q->z);
    FN(fe_mul)(&S2, &q->y, &Z1Z1); FN(fe_mul)(&S2, &S2, &p->z);
    FN(fe_sub)(&H, &U2, &U1); FN(fe_sub)(&R, &S2, &S1);
    if (FN(fe_is_zero)(&H)) { if (FN(fe_is_zero)(&R)) FN(dbl)(r, p); else FN(set_inf)(r); return; }
    FN(fe_sqr)(&H2, &H); FN(fe_mul)(&H3, &H2, &H);
    FN(fe_mul)(&t, &U1, &H2);
    FN(fe_sqr)(&u, &R); FN(fe_sub)(&u, &u, &H3); FN(fe_sub)(&u, &u, &t); FN(fe_sub)(&u, &u, &t);
    FN(fe_sub)(&t, &t, &u); FN(fe_mul)(&t, &t, &R); FN(fe_mul)(&H3, &H3, &S1);
    FN(fe_mul)(&z3, &p->z, &q->z); FN(fe_mul)(&r->z, &z3, &H); r->x = u; FN(fe_sub)(&r->y, &t, &H3);
}
/* [k]p, k = nlimbs little-endian u64 limbs; plain double-and-add as Mul<Scalar> [dep] does */
static inline void FN(mul)(PT *r, const PT *p, const uint64_t *k, int nlimbs) {
    PT acc; FN(set_inf)(&acc);
    for (int i = nlimbs * 64 - 1; i >= 0; i--) {
        FN(dbl)(&acc, &acc);
        if ((k[i >> 6] >> (i & 63)) & 1) FN(add)(&acc, &acc, p);
    }
    *r = acc;
}
static inline int FN(eq)(const PT *a, const PT *b) {
    int ia = FN(is_inf)(a), ib = FN(is_inf)(b);
    if (ia || ib) return ia && ib;
    FE za, zb, t, u;
    FN(fe_sqr)(&za, &a->z); FN(fe_sqr)(&zb, &b->z);
    FN(fe_mul)(&t, &a->x, &zb); FN(fe_mul)(&u, &b->x, &za);
    if (!FN(fe_eq)(&t, &u)) return 0;
    FN(fe_mul)(&za, &za, &a->z); FN(fe_mul)(&zb, &zb, &b->z);
    FN(fe_mul)(&t, &a->y, &zb); FN(fe_mul)(&u, &b->y, &za);
    return FN(fe_eq)(&t, &u);
}
