/* kzgb200 -- C ABI of the B200-native (sm_100a) replacement for the EIP-4844 verification hot path of
 * succinctlabs/kzg-rs v0.2.8.  This is the boundary a Rust (extern "C"), cgo, or ctypes binding binds;
 * INTEGRATION.md shows the Rust shim that keeps kzg-rs's public API unchanged on top of it.
 *
 * Conventions
 *   - plain pointers and sizes only; the library never frees or retains caller memory past return;
 *   - "host" entry points take host memory (pageable or pinned) and do their own staging;
 *     "_device" entry points take device pointers on the context's GPU (inputs already resident in HBM).  The library
 *     works on its own CUDA streams: device inputs must be complete before the call (synchronise the stream that
 *     produced them); every entry point returns only after its outputs are complete;
 *   - return codes map 1:1 onto kzg-rs's KzgError (reference src/enums.rs:6-18):
 *         KZGB200_OK                 Ok(verdict), verdict in *ok (1 = true, 0 = false)
 *         KZGB200_BAD_ARGS           Err(KzgError::BadArgs)            -- unparsable scalar / G1 point
 *         KZGB200_INTERNAL_ERROR     Err(KzgError::InternalError)      -- CUDA failure
 *         KZGB200_INVALID_LENGTH     Err(KzgError::InvalidBytesLength) -- vector length mismatch
 *         KZGB200_INVALID_SETUP      Err(KzgError::InvalidTrustedSetup)
 *   - a context is bound to one GPU; calls on one context are serialised by an internal lock, use one
 *     context per thread (or per GPU) for concurrency.  There is no CPU fallback: every entry point fails
 *     with KZGB200_INTERNAL_ERROR when no sm_100 device / kernel image is available.
 */
#ifndef KZGB200_H
#define KZGB200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define KZGB200_OK 0
#define KZGB200_BAD_ARGS 1
#define KZGB200_INTERNAL_ERROR 2
#define KZGB200_INVALID_LENGTH 3
#define KZGB200_INVALID_SETUP 5

#define KZGB200_BYTES_PER_BLOB 131072       /* reference src/consts.rs:8 */
#define KZGB200_BYTES_PER_COMMITMENT 48     /* src/consts.rs:9 */
#define KZGB200_BYTES_PER_PROOF 48          /* src/consts.rs:10 */
#define KZGB200_BYTES_PER_FIELD_ELEMENT 32  /* src/consts.rs:3 */
#define KZGB200_PARTIAL_BYTES 352           /* per-rank partial of the sharded batch (see below) */

typedef struct kzgb200_ctx kzgb200_ctx;

/* Replaces KzgSettings::load_trusted_setup_file (reference src/trusted_setup.rs:94-98) + the table building
 * of build.rs:131-170: uploads / derives the device-resident tables (roots of unity in Montgomery form,
 * Miller-loop line coefficients of g2_points[0] and g2_points[1]).  g2_points = the first two G2 points of
 * the trusted setup, ZCash-compressed, 2 x 96 bytes (trusted_setup.txt lines 4099-4100).  The verification
 * path reads nothing else from the setup. */
int kzgb200_create(kzgb200_ctx** out, int device, const uint8_t* g2_points, size_t g2_points_len);
void kzgb200_destroy(kzgb200_ctx* ctx);
/* last CUDA error string of the context ("" if none); valid until the next call */
const char* kzgb200_last_error(const kzgb200_ctx* ctx);

/* KzgProof::verify_kzg_proof (reference src/kzg_proof.rs:353-397). */
int kzgb200_verify_kzg_proof(kzgb200_ctx* ctx, const uint8_t* commitment48, const uint8_t* z32, const uint8_t* y32,
                             const uint8_t* proof48, int* ok);

/* KzgProof::verify_blob_kzg_proof (reference src/kzg_proof.rs:446-470).  z_out / y_out (32 bytes, big-endian,
 * nullable) receive the Fiat-Shamir challenge and the evaluation -- intermediates that must be bit-exact. */
int kzgb200_verify_blob_kzg_proof(kzgb200_ctx* ctx, const uint8_t* blob, const uint8_t* commitment48,
                                  const uint8_t* proof48, int* ok, uint8_t* z_out, uint8_t* y_out);

/* KzgProof::verify_blob_kzg_proof_batch (reference src/kzg_proof.rs:472-525).  The three lengths are the
 * lengths of the three Vec arguments; n == 0 -> Ok(true); n == 1 -> single path; mismatch -> INVALID_LENGTH.
 * blobs = n_blobs x 131072 bytes contiguous (a Vec<Blob> is exactly that), commitments / proofs = n x 48.
 * z_out / y_out: n_blobs x 32 bytes big-endian, nullable. */
int kzgb200_verify_blob_kzg_proof_batch(kzgb200_ctx* ctx, const uint8_t* blobs, size_t n_blobs,
                                        const uint8_t* commitments, size_t n_commitments,
                                        const uint8_t* proofs, size_t n_proofs, int* ok,
                                        uint8_t* z_out, uint8_t* y_out);
/* Same, inputs (and z_out / y_out, nullable) are device pointers on the context's GPU; n >= 1. */
int kzgb200_verify_blob_kzg_proof_batch_device(kzgb200_ctx* ctx, const uint8_t* d_blobs, const uint8_t* d_commitments,
                                               const uint8_t* d_proofs, size_t n, int* ok,
                                               uint8_t* d_z_out, uint8_t* d_y_out);

/* m independent verify_kzg_proof tuples (BASELINE config 5); verdicts[i] = 0 false, 1 true, 2 BadArgs.
 * Host pointers. */
int kzgb200_verify_kzg_proof_many(kzgb200_ctx* ctx, const uint8_t* commitments, const uint8_t* zs, const uint8_t* ys,
                                  const uint8_t* proofs, size_t m, uint8_t* verdicts);

/* ---- sharded batch: one context (= one rank) per GPU, blobs partitioned by contiguous ranges -------------
 * The batch is verified as   phase 1 (per rank)  -> exchange (z,y)  -> r  -> phase 2 (per rank partial sums)
 * -> allgather of KZGB200_PARTIAL_BYTES per rank -> one final pairing check.  All pointers are device
 * pointers on the rank's GPU; the caller moves the small payloads between ranks (NCCL allgather).
 *
 * phase 1: parse C/pi, canonicity, z_i, y_i for this rank's n_local blobs.  d_zy_out = n_local x 64 bytes:
 *          z_i then y_i as 32-byte little-endian canonical scalars (the byte order the batch transcript
 *          hashes, reference src/kzg_proof.rs:320-328). */
int kzgb200_shard_evaluate(kzgb200_ctx* ctx, const uint8_t* d_blobs, const uint8_t* d_commitments,
                           const uint8_t* d_proofs, size_t n_local, uint8_t* d_zy_out);
/* phase 1 with the shard's inputs in HOST memory (pageable or pinned): the blobs are copied in chunks while earlier
 * chunks are hashed / evaluated.  The device copies of the commitments / proofs (n_local x 48 each) and zy are written
 * to the given device buffers for the exchange. */
int kzgb200_shard_evaluate_host(kzgb200_ctx* ctx, const uint8_t* blobs, const uint8_t* commitments, const uint8_t* proofs,
                                size_t n_local, uint8_t* d_commitments_out, uint8_t* d_proofs_out, uint8_t* d_zy_out);
/* r = SHA-256 transcript over ALL n_total blobs in global order (reference src/kzg_proof.rs:291-348), from the
 * gathered commitments (n_total x 48), zy (n_total x 64, as produced by phase 1) and proofs (n_total x 48). */
int kzgb200_shard_challenge(kzgb200_ctx* ctx, const uint8_t* d_all_commitments, const uint8_t* d_all_zy,
                            const uint8_t* d_all_proofs, size_t n_total);
/* phase 2: this rank's partial sums with r_i = r^(global_offset + i); writes KZGB200_PARTIAL_BYTES to d_partial_out. */
int kzgb200_shard_lincomb(kzgb200_ctx* ctx, size_t global_offset, uint8_t* d_partial_out);
/* final: sum the gathered partials (n_ranks x KZGB200_PARTIAL_BYTES) and run the single pairing check.  The subgroup checks of
 * this rank's own points run beside the tail (they may still be running when the partial is exported), so a rank whose shard
 * holds a point outside the subgroup learns it here: it returns KZGB200_BAD_ARGS while the other ranks see the flags only if they
 * were known at export time -- the caller combines the return codes of all ranks (kzg_rs_b200/sharded.py: one 4-byte all-reduce). */
int kzgb200_shard_finalize(kzgb200_ctx* ctx, const uint8_t* d_partials, size_t n_ranks, int* ok);

/* ---- transcript mode of the batch challenge r -----------------------------------------------------------------
 * EXACT (default): r = SHA-256 over the serial transcript exactly as compute_r_powers (reference
 *   src/kzg_proof.rs:291-348) -- r, its powers and both MSM sums are bit-identical to kzg-rs.  The hash is one
 *   serial chain of 2.5 SHA-256 blocks per blob.
 * TREE (opt-in): same transcript bytes hashed as a three-level tree (16-entry leaves, then 32 digests per middle hash, in parallel; then the root).  r differs from
 *   kzg-rs's; verdicts do not (z, y are unaffected).  For throughput at large n / many GPUs. */
#define KZGB200_TRANSCRIPT_EXACT 0
#define KZGB200_TRANSCRIPT_TREE 1
int kzgb200_set_transcript_mode(kzgb200_ctx* ctx, int mode);
/* canonical big-endian r of the last batch verified on this context (n >= 2) */
int kzgb200_last_r(kzgb200_ctx* ctx, uint8_t* r_out32);

/* raw partial of the last single-GPU batch (n >= 2), KZGB200_PARTIAL_BYTES: Jacobian A = sum r_i pi_i and
 * B' = sum r_i C_i + r_i z_i pi_i as 3 x 12 little-endian u32 Montgomery limbs each (R = 2^384), then
 * s = sum r_i y_i (8 limbs, canonical), then the error flags.  For parity tests of the MSM intermediates. */
int kzgb200_last_partial(kzgb200_ctx* ctx, uint8_t* out352);

/* ---- harness (test / bench data; kzg-rs itself has no commit/prove path) ---------------------------------
 * Fills device buffers with n synthetic blobs (evaluation form of seeded random polynomials of degree <
 * `degree`, 2 <= degree <= 16) and their valid commitments and proofs over the trusted setup whose
 * [tau^j]G1, j < degree, are given compressed in tau_powers48 (host, degree x 48 bytes). */
int kzgb200_harness_generate(kzgb200_ctx* ctx, uint64_t seed, size_t n, int degree, const uint8_t* tau_powers48,
                             uint8_t* d_blobs, uint8_t* d_commitments, uint8_t* d_proofs);
/* ---- commit / prove (EIP-4844 blob_to_kzg_commitment / compute_blob_kzg_proof) -------------------
 * kzg-rs has no commit/prove path (its g1_points are loaded but never read); these are the companion operations used
 * to make test data for arbitrary blobs, pinned by the commitment / proof bytes of the reference's valid vectors.
 * kzgb200_load_g1_lagrange: the setup's 4096 Lagrange G1 points, compressed, FILE order (host, 4096 x 48 bytes);
 * builds a 4.8 GB fixed-base window table on the device (once per context).  The batch calls take DEVICE pointers. */
int kzgb200_load_g1_lagrange(kzgb200_ctx* ctx, const uint8_t* g1_lagrange, size_t n_points);
int kzgb200_blob_to_kzg_commitment_batch(kzgb200_ctx* ctx, const uint8_t* d_blobs, size_t n, uint8_t* d_commitments_out);
int kzgb200_compute_blob_kzg_proof_batch(kzgb200_ctx* ctx, const uint8_t* d_blobs, const uint8_t* d_commitments, size_t n,
                                         uint8_t* d_proofs_out);

/* per-phase device timing of single-GPU batch calls (CUDA events on the context stream).  out7 = milliseconds of
 * {parse G1, challenge, evaluate, transcript r, lincomb terms, reduce, final pairing} of the last call. */
int kzgb200_set_profiling(kzgb200_ctx* ctx, int on);
int kzgb200_get_phase_ms(kzgb200_ctx* ctx, float* out7);
/* profiling aid: SM clock stamps of the sections of the last single-GPU final pairing kernel */
int kzgb200_debug_final_ticks(kzgb200_ctx* ctx, long long* out14);
/* the cudaStream_t all work of this context is issued on (for CUDA-event timing by the caller) */
void* kzgb200_stream(kzgb200_ctx* ctx);

/* ---- streaming front-end: several batches in flight on one GPU --------------------------------
 * A batch is a blob-streaming head (challenges, evaluations, G1 parsing: all SMs busy) followed by a latency-bound tail
 * (transcript hash, MSM reduction, ONE pairing: a handful of CTAs).  A pipeline owns `depth` (1..8; 2 is enough)
 * independent contexts with one host worker thread each, so that the tail of one batch runs under the head -- and,
 * for host buffers, under the PCIe copy -- of the next.  Each ticket is exactly one kzgb200_verify_blob_kzg_proof_batch
 * (host pointers) or ..._batch_device (device pointers) call: same verdicts, same return codes, same z / y.
 * submit returns once the batch is queued (it blocks while `depth` batches are already in flight); the buffers must
 * stay valid and unmodified until the ticket has been waited for.  wait returns that batch's return code and verdict;
 * a ticket can be waited for once (KZGB200_BAD_ARGS otherwise).  Tickets complete in any order; wait in any order.
 * kzgb200_pipeline_context(p, slot) exposes the contexts for the tuning calls (e.g. kzgb200_set_transcript_mode). */
typedef struct kzgb200_pipeline kzgb200_pipeline;
int kzgb200_pipeline_create(kzgb200_pipeline** out, int device, const uint8_t* g2_points, size_t g2_points_len, int depth);
void kzgb200_pipeline_destroy(kzgb200_pipeline* p);   /* finishes what was submitted, then frees */
int kzgb200_pipeline_depth(const kzgb200_pipeline* p);
kzgb200_ctx* kzgb200_pipeline_context(kzgb200_pipeline* p, int slot);
int kzgb200_pipeline_submit(kzgb200_pipeline* p, const uint8_t* blobs, size_t n_blobs, const uint8_t* commitments,
                            size_t n_commitments, const uint8_t* proofs, size_t n_proofs, uint8_t* z_out, uint8_t* y_out,
                            uint64_t* ticket);
int kzgb200_pipeline_submit_device(kzgb200_pipeline* p, const uint8_t* d_blobs, const uint8_t* d_commitments,
                                   const uint8_t* d_proofs, size_t n, uint8_t* d_z_out, uint8_t* d_y_out, uint64_t* ticket);
int kzgb200_pipeline_wait(kzgb200_pipeline* p, uint64_t ticket, int* ok);

/* pinned host memory helpers for callers that want full-rate host->device copies */
void* kzgb200_alloc_pinned(size_t bytes);
void kzgb200_free_pinned(void* p);

#ifdef __cplusplus
}
#endif
#endif
