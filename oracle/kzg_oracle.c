/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY.
 *
 * A C restatement of the EIP-4844 verification path of succinctlabs/kzg-rs v0.2.8, used as the
 * bit-exact checker for the CUDA library and as the timed CPU baseline (bench.py cpu_baseline leg /
 * --impl reference).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may
 * load it; the product (kzg_rs_b200/, libkzgb200.so) never does.
 *
 * Follows, function by function:
 *   /root/reference/src/kzg_proof.rs:17-25    safe_g1_affine_from_bytes
 *   /root/reference/src/kzg_proof.rs:27-43    safe_scalar_affine_from_bytes
 *   /root/reference/src/kzg_proof.rs:46-72    compute_challenge
 *   /root/reference/src/kzg_proof.rs:74-91    scalar_from_bytes_unchecked
 *   /root/reference/src/kzg_proof.rs:94-133   evaluate_polynomial_in_evaluation_form
 *   /root/reference/src/kzg_proof.rs:155-201  batch_inversion
 *   /root/reference/src/kzg_proof.rs:203-223  verify_kzg_proof_impl
 *   /root/reference/src/kzg_proof.rs:251-277  compute_challenges_and_evaluate_polynomial
 *   /root/reference/src/kzg_proof.rs:279-348  compute_powers / compute_r_powers
 *   /root/reference/src/kzg_proof.rs:353-525  KzgProof::{verify_kzg_proof, verify_kzg_proof_batch,
 *                                             verify_blob_kzg_proof, verify_blob_kzg_proof_batch}
 *   /root/reference/src/pairings.rs:5-9       pairings_verify
 *   /root/reference/src/dtypes.rs:48-57       Blob::as_polynomial
 *   /root/reference/build.rs:89-105,131-170   bit-reversal order of the roots of unity / G1 points
 * The arithmetic of the un-vendored dependency sp1_bls12_381 =0.8.0-sp1-6.0.0 is restated in
 * field.h / tower.h / curve.h / pairing.h, SHA-256 (sha2 0.10.9) in sha256.h.
 *
 * Parity pin: tests/test_oracle_fixtures.py runs all 122 + 29 + 24 c-kzg-4844 vectors shipped in
 * /root/reference/tests and both known-answer tests of kzg_proof.rs:739-778 through this library.
 *
 * Additionally provides the harness-side commit/prove path (EIP-4844 blob_to_kzg_commitment /
 * compute_blob_kzg_proof; kzg-rs has none) used to make synthetic test data.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include "pairing.h"
#include "sha256.h"

#define N_FE 4096
#define BLOB_BYTES 131072
#define CHALLENGE_INPUT_SIZE 131152            /* consts.rs:12-13 */
enum { KZG_OK = 0, KZG_BADARGS = 1, KZG_INTERNAL = 2, KZG_BADLEN = 3, KZG_BADSETUP = 5 };

static struct {
    int ready, g1_ready;
    fr roots[N_FE];              /* roots_of_unity[i] = omega^bitrev12(i), Montgomery (build.rs:131-170) */
    fr inv_n;
    uint8_t g1_bytes[N_FE][48];  /* Lagrange G1 points, bit-reversal permuted (build.rs:79) */
    g1_aff g1_lagrange[N_FE];
    g2_aff g2_gen, tau_g2;       /* g2_points[0], g2_points[1] */
    pthread_mutex_t lock;
} S = {.lock = PTHREAD_MUTEX_INITIALIZER};

static uint32_t bitrev12(uint32_t i) {
    uint32_t r = 0; for (int b = 0; b < 12; b++) r |= ((i >> b) & 1) << (11 - b); return r;
}

/* setup_bin = "KZGS" | u32 n1 | u32 n2 | n1*48 | n2*96 (tests/golden/make_fixtures.py) */
int kzgo_init(const uint8_t *setup_bin, size_t len) {
    if (len < 12 || memcmp(setup_bin, "KZGS", 4)) return KZG_BADSETUP;
    uint32_t n1, n2; memcpy(&n1, setup_bin + 4, 4); memcpy(&n2, setup_bin + 8, 4);
    if (n1 != N_FE || n2 < 2 || len != 12 + (size_t)n1 * 48 + (size_t)n2 * 96) return KZG_BADSETUP;
    fr w, acc; memcpy(w.l, FR_OMEGA_M, 32); fr_set_one(&acc);
    for (uint32_t i = 0; i < N_FE; i++) { S.roots[bitrev12(i)] = acc; fr_mul(&acc, &acc, &w); }
    memcpy(S.inv_n.l, FR_INV4096_M, 32);
    for (uint32_t i = 0; i < N_FE; i++) memcpy(S.g1_bytes[i], setup_bin + 12 + 48 * (size_t)bitrev12(i), 48);
    const uint8_t *g2 = setup_bin + 12 + (size_t)n1 * 48;
    if (!g2_from_compressed_unchecked(&S.g2_gen, g2)) return KZG_BADSETUP;
    if (!g2_from_compressed_unchecked(&S.tau_g2, g2 + 96)) return KZG_BADSETUP;
    S.ready = 1; S.g1_ready = 0;
    return KZG_OK;
}
static int ensure_g1(void) {
    pthread_mutex_lock(&S.lock);
    if (!S.g1_ready) {
        for (int i = 0; i < N_FE; i++)
            if (!g1_from_compressed(&S.g1_lagrange[i], S.g1_bytes[i], 0)) { pthread_mutex_unlock(&S.lock); return KZG_BADSETUP; }
        S.g1_ready = 1;
    }
    pthread_mutex_unlock(&S.lock);
    return KZG_OK;
}

/* kzg_proof.rs:27-43 */
static int safe_scalar_from_bytes(fr *out, const uint8_t b[32]) {
    uint64_t raw[4]; be32_to_limbs(raw, b);
    if (bn_geq(raw, FR_Q, 4)) return KZG_BADARGS;
    fr_from_raw(out, raw); return KZG_OK;
}
/* kzg_proof.rs:17-25 */
static int safe_g1_from_bytes(g1_aff *out, const uint8_t b[48]) {
    return g1_from_compressed(out, b, 1) ? KZG_OK : KZG_BADARGS;
}
/* kzg_proof.rs:74-91 : big-endian 256-bit value reduced mod q via from_raw */
static void scalar_from_bytes_unchecked(fr *out, const uint8_t b[32]) {
    uint64_t raw[4]; be32_to_limbs(raw, b); fr_from_raw(out, raw);
}
/* dtypes.rs:48-57 */
static int blob_as_polynomial(fr *poly, const uint8_t *blob) {
    for (int i = 0; i < N_FE; i++)
        if (safe_scalar_from_bytes(&poly[i], blob + 32 * i)) return KZG_BADARGS;
    return KZG_OK;
}
/* kzg_proof.rs:46-72 */
static void compute_challenge(fr *z, const uint8_t *blob, const g1_aff *commitment) {
    uint8_t *buf = (uint8_t *)malloc(CHALLENGE_INPUT_SIZE);
    memcpy(buf, "FSBLOBVERIFY_V1_", 16);
    memset(buf + 16, 0, 16); buf[30] = 0x10;              /* u64be 0 | u64be 4096 */
    memcpy(buf + 32, blob, BLOB_BYTES);
    g1_to_compressed(buf + 32 + BLOB_BYTES, commitment);
    uint8_t dig[32]; sha256(dig, buf, CHALLENGE_INPUT_SIZE); free(buf);
    scalar_from_bytes_unchecked(z, dig);
}
/* kzg_proof.rs:155-201 */
static int batch_inversion(fr *out, const fr *a, int n) {
    fr acc; fr_set_one(&acc);
    for (int i = 0; i < n; i++) { out[i] = acc; fr_mul(&acc, &acc, &a[i]); }
    if (fr_is_zero(&acc)) return KZG_BADARGS;
    fr_inv(&acc, &acc);
    for (int i = n - 1; i >= 0; i--) { fr_mul(&out[i], &out[i], &acc); fr_mul(&acc, &acc, &a[i]); }
    return KZG_OK;
}
/* kzg_proof.rs:94-133 */
static int evaluate_polynomial_in_evaluation_form(fr *y, const fr *poly, const fr *x) {
    fr *in = (fr *)malloc(2 * N_FE * sizeof(fr)), *inv = in + N_FE;
    for (int i = 0; i < N_FE; i++) {
        if (fr_eq(x, &S.roots[i])) { *y = poly[i]; free(in); return KZG_OK; }
        fr_sub(&in[i], x, &S.roots[i]);
    }
    int rc = batch_inversion(inv, in, N_FE);
    if (rc) { free(in); return rc; }
    fr out, t; fr_set_zero(&out);
    for (int i = 0; i < N_FE; i++) { fr_mul(&t, &inv[i], &S.roots[i]); fr_mul(&t, &t, &poly[i]); fr_add(&out, &out, &t); }
    fr n; fr_from_u64(&n, N_FE); fr_inv(&n, &n); fr_mul(&out, &out, &n);          /* :127-129 */
    uint64_t e[4] = {N_FE, 0, 0, 0}; fr_pow(&t, x, e, 4);                          /* :130 */
    fr one; fr_set_one(&one); fr_sub(&t, &t, &one); fr_mul(y, &out, &t);
    free(in); return KZG_OK;
}
/* kzg_proof.rs:203-223 and :385-396 */
static int verify_kzg_proof_impl(const g1_aff *C, const fr *z, const fr *y, const g1_aff *proof) {
    uint64_t zr[4], yr[4]; fr_to_raw(zr, z); fr_to_raw(yr, y);
    g2 gx, tau, xmz; g2_from_aff(&gx, &S.g2_gen); g2_mul(&gx, &gx, zr, 4);
    g2_neg(&gx, &gx); g2_from_aff(&tau, &S.tau_g2); g2_add(&xmz, &tau, &gx);
    g1_aff gen; g1_generator(&gen);
    g1 gy, c, pmy; g1_from_aff(&gy, &gen); g1_mul(&gy, &gy, yr, 4); g1_neg(&gy, &gy);
    g1_from_aff(&c, C); g1_add(&pmy, &c, &gy);
    g1_aff pmy_a; g2_aff xmz_a; g1_to_aff(&pmy_a, &pmy); g2_to_aff(&xmz_a, &xmz);
    return pairings_verify(&pmy_a, &S.g2_gen, proof, &xmz_a);
}

/* kzg_proof.rs:353-397 ; parse order z, y, commitment, proof */
int kzgo_verify_kzg_proof(const uint8_t *c48, const uint8_t *z32, const uint8_t *y32, const uint8_t *p48, int *ok) {
    fr z, y; g1_aff C, pr;
    if (safe_scalar_from_bytes(&z, z32)) return KZG_BADARGS;
    if (safe_scalar_from_bytes(&y, y32)) return KZG_BADARGS;
    if (safe_g1_from_bytes(&C, c48)) return KZG_BADARGS;
    if (safe_g1_from_bytes(&pr, p48)) return KZG_BADARGS;
    *ok = verify_kzg_proof_impl(&C, &z, &y, &pr);
    return KZG_OK;
}

/* kzg_proof.rs:446-470 ; error order commitment, blob, proof */
int kzgo_verify_blob_kzg_proof(const uint8_t *blob, const uint8_t *c48, const uint8_t *p48, int *ok,
                               uint8_t *z_out, uint8_t *y_out) {
    g1_aff C, pr; fr z, y;
    if (safe_g1_from_bytes(&C, c48)) return KZG_BADARGS;
    fr *poly = (fr *)malloc(N_FE * sizeof(fr));
    if (blob_as_polynomial(poly, blob)) { free(poly); return KZG_BADARGS; }
    if (safe_g1_from_bytes(&pr, p48)) { free(poly); return KZG_BADARGS; }
    compute_challenge(&z, blob, &C);
    int rc = evaluate_polynomial_in_evaluation_form(&y, poly, &z);
    free(poly);
    if (rc) return rc;
    if (z_out) fr_to_bytes_be(z_out, &z);
    if (y_out) fr_to_bytes_be(y_out, &y);
    *ok = verify_kzg_proof_impl(&C, &z, &y, &pr);
    return KZG_OK;
}

/* ---- small thread pool helper: run fn(i) for i in [0,n) on nthreads threads ---- */
typedef struct { void (*fn)(size_t, void *); void *arg; size_t n; size_t *next; pthread_mutex_t *m; } job_t;
static void *job_worker(void *p) {
    job_t *j = (job_t *)p;
    for (;;) {
        pthread_mutex_lock(j->m); size_t i = (*j->next)++; pthread_mutex_unlock(j->m);
        if (i >= j->n) break;
        j->fn(i, j->arg);
    }
    return NULL;
}
static void parallel_for(size_t n, int nthreads, void (*fn)(size_t, void *), void *arg) {
    if (nthreads <= 1 || n <= 1) { for (size_t i = 0; i < n; i++) fn(i, arg); return; }
    if ((size_t)nthreads > n) nthreads = (int)n;
    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nthreads);
    size_t next = 0; pthread_mutex_t m = PTHREAD_MUTEX_INITIALIZER;
    job_t j = {fn, arg, n, &next, &m};
    for (int t = 0; t < nthreads; t++) pthread_create(&th[t], NULL, job_worker, &j);
    for (int t = 0; t < nthreads; t++) pthread_join(th[t], NULL);
    free(th);
}

typedef struct {
    const uint8_t *blobs, *cb, *pb;
    g1_aff *C, *P; fr *z, *y, *rpow; g1 *cmy; int *err;
} batch_t;
static void job_parse_c(size_t i, void *a) { batch_t *b = (batch_t *)a; b->err[i] = safe_g1_from_bytes(&b->C[i], b->cb + 48 * i); }
static void job_parse_p(size_t i, void *a) { batch_t *b = (batch_t *)a; b->err[i] = safe_g1_from_bytes(&b->P[i], b->pb + 48 * i); }
/* body of the loop at kzg_proof.rs:261-273 */
static void job_blob(size_t i, void *a) {
    batch_t *b = (batch_t *)a;
    fr *poly = (fr *)malloc(N_FE * sizeof(fr));
    const uint8_t *blob = b->blobs + (size_t)BLOB_BYTES * i;
    b->err[i] = blob_as_polynomial(poly, blob);
    if (!b->err[i]) {
        compute_challenge(&b->z[i], blob, &b->C[i]);
        b->err[i] = evaluate_polynomial_in_evaluation_form(&b->y[i], poly, &b->z[i]);
    }
    free(poly);
}
/* body of the loop at kzg_proof.rs:422-426: c_minus_y[i] = C_i - [y_i]G */
static void job_cmy(size_t i, void *a) {
    batch_t *b = (batch_t *)a;
    uint64_t yr[4]; fr_to_raw(yr, &b->y[i]);
    g1_aff gen; g1_generator(&gen);
    g1 gy, c; g1_from_aff(&gy, &gen); g1_mul(&gy, &gy, yr, 4); g1_neg(&gy, &gy);
    g1_from_aff(&c, &b->C[i]); g1_add(&b->cmy[i], &c, &gy);
}

/* kzg_proof.rs:291-348 (r) + :279-289 (powers).  z, y are hashed little-endian (Scalar::to_bytes). */
static void compute_r_powers(fr *rp, fr *r_out, const g1_aff *C, const fr *z, const fr *y, const g1_aff *P, size_t n) {
    size_t len = 32 + n * 160;
    uint8_t *buf = (uint8_t *)malloc(len);
    memcpy(buf, "RCKZGBATCH___V1_", 16);
    memset(buf + 16, 0, 16); buf[22] = 0x10;                      /* u64be 4096 */
    for (int i = 0; i < 8; i++) buf[24 + i] = (uint8_t)((uint64_t)n >> (56 - 8 * i));   /* usize n, BE */
    uint8_t *p = buf + 32;
    for (size_t i = 0; i < n; i++) {
        g1_to_compressed(p, &C[i]); p += 48;
        fr_to_bytes_le(p, &z[i]); p += 32;
        fr_to_bytes_le(p, &y[i]); p += 32;
        g1_to_compressed(p, &P[i]); p += 48;
    }
    uint8_t dig[32]; sha256(dig, buf, len); free(buf);
    fr r; scalar_from_bytes_unchecked(&r, dig); if (r_out) *r_out = r;
    if (n) fr_set_one(&rp[0]);
    for (size_t i = 1; i < n; i++) fr_mul(&rp[i], &rp[i - 1], &r);
}

/* trace layout (optional, 128 bytes): r (32 BE) | proof_lincomb (48 compressed) | rhs_g1 (48 compressed) */
static int verify_kzg_proof_batch(batch_t *b, size_t n, int nthreads, uint8_t *trace) {
    fr *rp = (fr *)malloc(sizeof(fr) * n), r;
    compute_r_powers(rp, &r, b->C, b->z, b->y, b->P, n);                       /* :413 */
    uint64_t (*sc)[4] = (uint64_t (*)[4])malloc(32 * n);
    uint64_t (*scz)[4] = (uint64_t (*)[4])malloc(32 * n);
    for (size_t i = 0; i < n; i++) {
        fr rz; fr_mul(&rz, &rp[i], &b->z[i]);                                  /* :425 */
        fr_to_raw(sc[i], &rp[i]); fr_to_raw(scz[i], &rz);
    }
    g1 proof_lincomb, proof_z_lincomb, cmy_lincomb, rhs;
    g1_msm(&proof_lincomb, b->P, sc, n);                                       /* :419 */
    b->cmy = (g1 *)malloc(sizeof(g1) * n);
    parallel_for(n, nthreads, job_cmy, b);                                     /* :422-424 */
    g1_aff *cmy_a = (g1_aff *)malloc(sizeof(g1_aff) * n);
    for (size_t i = 0; i < n; i++) g1_to_aff(&cmy_a[i], &b->cmy[i]);
    g1_msm(&proof_z_lincomb, b->P, scz, n);                                    /* :429 */
    g1_msm(&cmy_lincomb, cmy_a, sc, n);                                        /* :430 */
    g1_add(&rhs, &cmy_lincomb, &proof_z_lincomb);                              /* :433 */
    g1_aff pl_a, rhs_a; g1_to_aff(&pl_a, &proof_lincomb); g1_to_aff(&rhs_a, &rhs);
    if (trace) { fr_to_bytes_be(trace, &r); g1_to_compressed(trace + 32, &pl_a); g1_to_compressed(trace + 80, &rhs_a); }
    int ok = pairings_verify(&pl_a, &S.tau_g2, &rhs_a, &S.g2_gen);             /* :436-441 */
    free(rp); free(sc); free(scz); free(b->cmy); free(cmy_a);
    return ok;
}

/* kzg_proof.rs:472-525.  n_blobs/n_c/n_p are the three vector lengths.  nthreads = 1 reproduces the
 * reference's sequential execution; > 1 runs the per-blob loops blob-parallel (same results). */
int kzgo_verify_blob_kzg_proof_batch(const uint8_t *blobs, size_t n_blobs, const uint8_t *cb, size_t n_c,
                                     const uint8_t *pb, size_t n_p, int nthreads, int *ok,
                                     uint8_t *z_out, uint8_t *y_out, uint8_t *trace) {
    if (n_blobs == 0) { *ok = 1; return KZG_OK; }                                           /* :478-480 */
    if (n_blobs == 1) return kzgo_verify_blob_kzg_proof(blobs, cb, pb, ok, z_out, y_out);   /* :482-489 */
    if (n_blobs != n_c) return KZG_BADLEN;                                                  /* :491-495 */
    if (n_blobs != n_p) return KZG_BADLEN;                                                  /* :497-501 */
    size_t n = n_blobs; int rc = KZG_OK;
    batch_t b = {blobs, cb, pb, NULL, NULL, NULL, NULL, NULL, NULL, NULL};
    b.C = (g1_aff *)malloc(sizeof(g1_aff) * n); b.P = (g1_aff *)malloc(sizeof(g1_aff) * n);
    b.z = (fr *)malloc(sizeof(fr) * n); b.y = (fr *)malloc(sizeof(fr) * n); b.err = (int *)calloc(n, sizeof(int));
    parallel_for(n, nthreads, job_parse_c, &b);                                             /* :503-506 */
    for (size_t i = 0; i < n && !rc; i++) rc = b.err[i];
    if (!rc) { parallel_for(n, nthreads, job_parse_p, &b); for (size_t i = 0; i < n && !rc; i++) rc = b.err[i]; }  /* :508-511 */
    /* validate_batched_input (:225-249) cannot fail after from_compressed */
    if (!rc) { parallel_for(n, nthreads, job_blob, &b); for (size_t i = 0; i < n && !rc; i++) rc = b.err[i]; }     /* :515-516 */
    if (!rc) {
        for (size_t i = 0; i < n; i++) {
            if (z_out) fr_to_bytes_be(z_out + 32 * i, &b.z[i]);
            if (y_out) fr_to_bytes_be(y_out + 32 * i, &b.y[i]);
        }
        *ok = verify_kzg_proof_batch(&b, n, nthreads, trace);                               /* :518-524 */
    }
    free(b.C); free(b.P); free(b.z); free(b.y); free(b.err);
    return rc;
}

/* ---- exposed protocol pieces (KAT surface of kzg_proof.rs:739-778) ---- */
int kzgo_compute_challenge(const uint8_t *blob, const uint8_t *c48, uint8_t *z_out) {
    g1_aff C; if (safe_g1_from_bytes(&C, c48)) return KZG_BADARGS;
    fr z; compute_challenge(&z, blob, &C); fr_to_bytes_be(z_out, &z); return KZG_OK;
}
/* z32 is reduced mod q like scalar_from_bytes_unchecked (the reference KAT passes it that way) */
int kzgo_evaluate_polynomial(const uint8_t *blob, const uint8_t *z32, uint8_t *y_out) {
    fr *poly = (fr *)malloc(N_FE * sizeof(fr)), z, y;
    if (blob_as_polynomial(poly, blob)) { free(poly); return KZG_BADARGS; }
    scalar_from_bytes_unchecked(&z, z32);
    int rc = evaluate_polynomial_in_evaluation_form(&y, poly, &z); free(poly);
    if (!rc) fr_to_bytes_be(y_out, &y);
    return rc;
}
/* r and the first n powers for already-parsed inputs (bytes in, 32-byte BE scalars out) */
int kzgo_compute_r_powers(const uint8_t *cb, const uint8_t *zb, const uint8_t *yb, const uint8_t *pb, size_t n,
                          uint8_t *r_out, uint8_t *powers_out) {
    g1_aff *C = (g1_aff *)malloc(sizeof(g1_aff) * (n + 1)), *P = (g1_aff *)malloc(sizeof(g1_aff) * (n + 1));
    fr *z = (fr *)malloc(sizeof(fr) * (n + 1)), *y = (fr *)malloc(sizeof(fr) * (n + 1)), *rp = (fr *)malloc(sizeof(fr) * (n + 1)), r;
    int rc = KZG_OK;
    for (size_t i = 0; i < n && !rc; i++) {
        if (safe_g1_from_bytes(&C[i], cb + 48 * i) || safe_g1_from_bytes(&P[i], pb + 48 * i) ||
            safe_scalar_from_bytes(&z[i], zb + 32 * i) || safe_scalar_from_bytes(&y[i], yb + 32 * i)) rc = KZG_BADARGS;
    }
    if (!rc) {
        compute_r_powers(rp, &r, C, z, y, P, n);
        fr_to_bytes_be(r_out, &r);
        if (powers_out) for (size_t i = 0; i < n; i++) fr_to_bytes_be(powers_out + 32 * i, &rp[i]);
    }
    free(C); free(P); free(z); free(y); free(rp);
    return rc;
}

/* ---- group-level helpers for tests ---- */
/* 0 = accepted, 1 = rejected; *fast/naive subgroup agreement is asserted by the tests */
int kzgo_g1_check(const uint8_t *b48, int *in_subgroup_fast, int *in_subgroup_naive) {
    g1_aff a; if (!g1_from_compressed(&a, b48, 0)) return 1;
    *in_subgroup_fast = g1_in_subgroup(&a); *in_subgroup_naive = g1_in_subgroup_naive(&a);
    return 0;
}
/* sum scalars[i]*points[i]; points compressed (subgroup-checked), scalars 32-byte BE canonical */
int kzgo_g1_lincomb(const uint8_t *pts48, const uint8_t *sc32, size_t n, int use_msm, uint8_t *out48) {
    g1_aff *P = (g1_aff *)malloc(sizeof(g1_aff) * (n + 1));
    uint64_t (*sc)[4] = (uint64_t (*)[4])malloc(32 * (n + 1));
    int rc = KZG_OK;
    for (size_t i = 0; i < n && !rc; i++) {
        if (safe_g1_from_bytes(&P[i], pts48 + 48 * i)) rc = KZG_BADARGS;
        be32_to_limbs(sc[i], sc32 + 32 * i);
    }
    if (!rc) {
        g1 acc; g1_set_inf(&acc);
        if (use_msm) g1_msm(&acc, P, sc, n);
        else for (size_t i = 0; i < n; i++) { g1 t; g1_from_aff(&t, &P[i]); g1_mul(&t, &t, sc[i], 4); g1_add(&acc, &acc, &t); }
        g1_aff a; g1_to_aff(&a, &acc); g1_to_compressed(out48, &a);
    }
    free(P); free(sc); return rc;
}
/* e(-a1, a2) e(b1, b2) == 1 with a2, b2 chosen from {0: G2 generator, 1: [tau]G2} */
int kzgo_pairings_verify(const uint8_t *a1, int a2_idx, const uint8_t *b1, int b2_idx, int *ok) {
    g1_aff A, B; if (safe_g1_from_bytes(&A, a1) || safe_g1_from_bytes(&B, b1)) return KZG_BADARGS;
    *ok = pairings_verify(&A, a2_idx ? &S.tau_g2 : &S.g2_gen, &B, b2_idx ? &S.tau_g2 : &S.g2_gen);
    return KZG_OK;
}

/* ---- harness-side commit / prove (EIP-4844 semantics; not part of kzg-rs) ---- */
static int lagrange_msm(uint8_t *out48, const fr *coeffs) {
    int rc = ensure_g1(); if (rc) return rc;
    uint64_t (*sc)[4] = (uint64_t (*)[4])malloc(32 * N_FE);
    for (int i = 0; i < N_FE; i++) fr_to_raw(sc[i], &coeffs[i]);
    g1 acc; g1_msm(&acc, S.g1_lagrange, sc, N_FE); free(sc);
    g1_aff a; g1_to_aff(&a, &acc); g1_to_compressed(out48, &a); return KZG_OK;
}
int kzgo_blob_to_kzg_commitment(const uint8_t *blob, uint8_t *c48) {
    fr *poly = (fr *)malloc(N_FE * sizeof(fr));
    int rc = blob_as_polynomial(poly, blob);
    if (!rc) rc = lagrange_msm(c48, poly);
    free(poly); return rc;
}
/* proof for p(z) = y: quotient in evaluation form, with the in-domain special case */
int kzgo_compute_kzg_proof(const uint8_t *blob, const uint8_t *z32, uint8_t *p48, uint8_t *y32) {
    fr *poly = (fr *)malloc(3 * N_FE * sizeof(fr)), *q = poly + N_FE, *den = q + N_FE, z, y, t;
    int rc = blob_as_polynomial(poly, blob);
    if (!rc) rc = safe_scalar_from_bytes(&z, z32);
    if (!rc) rc = evaluate_polynomial_in_evaluation_form(&y, poly, &z);
    if (!rc) {
        int m = -1;
        for (int i = 0; i < N_FE; i++) {
            fr_sub(&den[i], &S.roots[i], &z);
            if (fr_is_zero(&den[i])) { m = i; fr_set_one(&den[i]); }
        }
        fr *inv = (fr *)malloc(N_FE * sizeof(fr)); batch_inversion(inv, den, N_FE);
        for (int i = 0; i < N_FE; i++) { fr_sub(&t, &poly[i], &y); fr_mul(&q[i], &t, &inv[i]); }
        if (m >= 0) {   /* q_m = sum_{i != m} (f_i - y) w_i / (z (z - w_i)) */
            fr acc, zi; fr_set_zero(&acc); fr_inv(&zi, &z);
            for (int i = 0; i < N_FE; i++) {
                if (i == m) continue;
                fr_mul(&t, &q[i], &S.roots[i]); fr_mul(&t, &t, &zi);   /* q_i = (f_i-y)/(w_i-z) ; term = -q_i w_i / z */
                fr_sub(&acc, &acc, &t);
            }
            q[m] = acc;
        }
        free(inv);
        rc = lagrange_msm(p48, q);
        if (!rc && y32) fr_to_bytes_be(y32, &y);
    }
    free(poly); return rc;
}
int kzgo_compute_blob_kzg_proof(const uint8_t *blob, const uint8_t *c48, uint8_t *p48) {
    uint8_t z[32]; int rc = kzgo_compute_challenge(blob, c48, z);
    if (rc) return rc;
    return kzgo_compute_kzg_proof(blob, z, p48, NULL);
}
/* [tau^j]G1 = sum_i w_i^j L_i(tau) G  (commitment to X^j) */
int kzgo_tau_power_g1(unsigned j, uint8_t *out48) {
    fr *c = (fr *)malloc(N_FE * sizeof(fr)); uint64_t e[1] = {j};
    for (int i = 0; i < N_FE; i++) fr_pow(&c[i], &S.roots[i], e, 1);
    int rc = lagrange_msm(out48, c); free(c); return rc;
}
int kzgo_sha256_uses_shani(void) { return !sha256_force_portable && sha256_have_shani(); }
void kzgo_sha256_force_portable(int v) { sha256_force_portable = v; }
void kzgo_sha256(const uint8_t *msg, size_t len, uint8_t *out32) { sha256(out32, msg, len); }
