// libkzgb200.so -- host runtime + C ABI (include/kzgb200.h) over the sm_100a kernels.
// Mirrors the orchestration of KzgProof::{verify_kzg_proof, verify_blob_kzg_proof,
// verify_blob_kzg_proof_batch} (reference src/kzg_proof.rs:353-525): argument checks and phase ordering on
// the host, every arithmetic step in a kernel.  No CPU fallback: CUDA failures surface as
// KZGB200_INTERNAL_ERROR.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include "../../include/kzgb200.h"
#include "kernels.cuh"

using namespace kzgb200;

static_assert(sizeof(Partial) == KZGB200_PARTIAL_BYTES, "Partial layout is part of the ABI");
static_assert(sizeof(ZY) == 64, "ZY layout is part of the ABI");

struct kzgb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    DeviceTables* tables = nullptr;
    // workspace, sized for `cap` blobs
    size_t cap = 0, blob_cap = 0, many_cap = 0;
    uint8_t *d_blobs = nullptr, *d_c = nullptr, *d_p = nullptr;     // staging of host inputs
    Fr* d_z_mont = nullptr;
    ZY* d_zy = nullptr;
    G1Affine *d_C = nullptr, *d_P = nullptr;
    uint32_t* d_status = nullptr;
    Fr *d_ry = nullptr, *d_r = nullptr;
    uint8_t* d_digits = nullptr;            // [2][32][cap]
    uint32_t *d_order = nullptr, *d_start = nullptr;
    G1 *d_buckets = nullptr, *d_windows = nullptr;
    Partial* d_partial = nullptr;
    uint32_t* d_result = nullptr;
    uint8_t *d_zout = nullptr, *d_yout = nullptr;
    uint8_t *d_many = nullptr;
    uint32_t* d_wk = nullptr;       // transcript W+K words, 64 per SHA block
    size_t wk_cap = 0;
    uint32_t* h_result = nullptr;   // pinned
    // inputs of the current shard (device pointers owned by the caller or by the staging buffers)
    const uint8_t *cur_c = nullptr, *cur_p = nullptr;
    size_t cur_n = 0;
    // optional per-phase timing (CUDA events on the context stream)
    int transcript_mode = KZGB200_TRANSCRIPT_EXACT;
    bool profile = false;
    cudaEvent_t ev[9] = {nullptr};
    int ev_used = 0;
    float phase_ms[8] = {0};
    std::mutex lock;
    char err[256] = {0};
};
enum Phase { kPhParse = 0, kPhChallenge, kPhEval, kPhTranscript, kPhLincomb, kPhReduce, kPhFinal, kPhCount };
static void mark(kzgb200_ctx* ctx, int idx) {
    if (ctx->profile && ctx->ev[idx]) { cudaEventRecord(ctx->ev[idx], ctx->stream); if (idx + 1 > ctx->ev_used) ctx->ev_used = idx + 1; }
}

#define CK(expr)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (expr);                                                                   \
        if (e_ != cudaSuccess) {                                                                   \
            snprintf(ctx->err, sizeof(ctx->err), "%s: %s (%s:%d)", #expr, cudaGetErrorString(e_), __FILE__, __LINE__); \
            return KZGB200_INTERNAL_ERROR;                                                         \
        }                                                                                          \
    } while (0)

template <class T>
static cudaError_t regrow(T*& p, size_t count) {
    if (p) cudaFree(p);
    p = nullptr;
    return cudaMalloc(reinterpret_cast<void**>(&p), count * sizeof(T));
}

static int ensure_capacity(kzgb200_ctx* ctx, size_t n, bool need_blob_staging) {
    if (n > ctx->cap) {
        size_t c = n;
        CK(regrow(ctx->d_c, c * 48)); CK(regrow(ctx->d_p, c * 48));
        CK(regrow(ctx->d_z_mont, c)); CK(regrow(ctx->d_zy, c));
        CK(regrow(ctx->d_C, c)); CK(regrow(ctx->d_P, c));
        CK(regrow(ctx->d_status, c)); CK(regrow(ctx->d_ry, c));
        CK(regrow(ctx->d_digits, c * 2 * kWindows)); CK(regrow(ctx->d_order, c * 2 * kWindows));
        CK(regrow(ctx->d_zout, c * 32)); CK(regrow(ctx->d_yout, c * 32));
        ctx->cap = c;
    }
    if (need_blob_staging && n > ctx->blob_cap) {
        CK(regrow(ctx->d_blobs, n * (size_t)kBytesPerBlob));
        ctx->blob_cap = n;
    }
    return KZGB200_OK;
}

extern "C" int kzgb200_create(kzgb200_ctx** out, int device, const uint8_t* g2_points, size_t g2_points_len) {
    if (!out) return KZGB200_BAD_ARGS;
    *out = nullptr;
    if (!g2_points || g2_points_len != 192) return KZGB200_INVALID_SETUP;
    kzgb200_ctx* ctx = new (std::nothrow) kzgb200_ctx();
    if (!ctx) return KZGB200_INTERNAL_ERROR;
    ctx->device = device;
    auto fail = [&](int rc) { fprintf(stderr, "kzgb200_create: %s\n", ctx->err); kzgb200_destroy(ctx); return rc; };
    int rc = [&]() -> int {
        CK(cudaSetDevice(device));
        CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
        CK(cudaMalloc(&ctx->tables, sizeof(DeviceTables)));
        CK(cudaMalloc(&ctx->d_r, sizeof(Fr)));
        CK(cudaMalloc(&ctx->d_partial, sizeof(Partial)));
        CK(cudaMalloc(&ctx->d_result, 16));
        CK(cudaMalloc(&ctx->d_start, 2 * kWindows * (kBuckets + 1) * sizeof(uint32_t)));
        CK(cudaMalloc(&ctx->d_buckets, kMsmSets * kWindows * kBuckets * sizeof(G1)));
        CK(cudaMalloc(&ctx->d_windows, kMsmSets * kWindows * sizeof(G1)));
        CK(cudaMallocHost(&ctx->h_result, 16));
        uint8_t* d_g2 = nullptr;
        CK(cudaMalloc(&d_g2, 192));
        CK(cudaMemcpyAsync(d_g2, g2_points, 192, cudaMemcpyHostToDevice, ctx->stream));
        setup_tables_kernel<<<(4096 + 960 + 127) / 128, 128, 0, ctx->stream>>>(ctx->tables, d_g2);
        CK(cudaGetLastError());
        uint32_t ok = 0;
        CK(cudaMemcpyAsync(&ok, &ctx->tables->setup_ok, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        cudaFree(d_g2);
        return ok ? KZGB200_OK : KZGB200_INVALID_SETUP;
    }();
    if (rc != KZGB200_OK) return fail(rc);
    *out = ctx;
    return KZGB200_OK;
}

extern "C" void kzgb200_destroy(kzgb200_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    void* ptrs[] = {ctx->tables, ctx->d_blobs, ctx->d_c, ctx->d_p, ctx->d_z_mont, ctx->d_zy, ctx->d_C, ctx->d_P, ctx->d_status,
                    ctx->d_ry, ctx->d_r, ctx->d_partial, ctx->d_result, ctx->d_zout, ctx->d_yout, ctx->d_many, ctx->d_wk,
                    ctx->d_digits, ctx->d_order, ctx->d_start, ctx->d_buckets, ctx->d_windows};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (ctx->h_result) cudaFreeHost(ctx->h_result);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* kzgb200_last_error(const kzgb200_ctx* ctx) { return ctx ? ctx->err : "null context"; }

// ---- phases (all asynchronous on ctx->stream) ---------------------------------------------------------------
// phase 1: K4 (parse C, pi) + K2 (challenge) + K1/K3 (canonicity + evaluation)
static int launch_phase1(kzgb200_ctx* ctx, const uint8_t* d_blobs, const uint8_t* d_c, const uint8_t* d_p, size_t n) {
    int ni = (int)n;
    CK(cudaMemsetAsync(ctx->d_status, 0, n * sizeof(uint32_t), ctx->stream));
    ctx->ev_used = 0;
    mark(ctx, 0);
    g1_parse_kernel<<<(2 * ni + 127) / 128, 128, 0, ctx->stream>>>(d_c, d_p, ni, ctx->d_C, ctx->d_P, ctx->d_status);
    mark(ctx, 1);
    challenge_kernel<<<(ni + 63) / 64, 64, 0, ctx->stream>>>(d_blobs, d_c, ni, ctx->d_z_mont, ctx->d_zy);
    mark(ctx, 2);
    eval_kernel<<<ni, kEvalThreads, 0, ctx->stream>>>(d_blobs, ni, ctx->d_z_mont, ctx->tables, ctx->d_zy, ctx->d_status);
    mark(ctx, 3);
    CK(cudaGetLastError());
    ctx->cur_c = d_c; ctx->cur_p = d_p; ctx->cur_n = n;
    return KZGB200_OK;
}
// K5
static int launch_transcript(kzgb200_ctx* ctx, const uint8_t* d_all_c, const ZY* d_all_zy, const uint8_t* d_all_p, size_t n_total) {
    if (ctx->transcript_mode == KZGB200_TRANSCRIPT_TREE) {
        size_t ngroups = (n_total + kTreeGroup - 1) / kTreeGroup;
        if (ngroups * 8 > ctx->wk_cap * 64) { CK(regrow(ctx->d_wk, ngroups * 8 + 64)); ctx->wk_cap = (ngroups * 8 + 64) / 64; }
        transcript_tree_leaf_kernel<<<(unsigned)((ngroups + 63) / 64), 64, 0, ctx->stream>>>(d_all_c, d_all_zy, d_all_p, (uint64_t)n_total, ctx->d_wk);
        transcript_tree_root_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_wk, (uint64_t)n_total, ctx->d_r);
        mark(ctx, 4);
        CK(cudaGetLastError());
        return KZGB200_OK;
    }
    size_t nblk = (32 + n_total * 160 + 9 + 63) / 64;
    if (nblk > ctx->wk_cap) { CK(regrow(ctx->d_wk, nblk * 64)); ctx->wk_cap = nblk; }
    transcript_schedule_kernel<<<(unsigned)((nblk + 127) / 128), 128, 0, ctx->stream>>>(d_all_c, d_all_zy, d_all_p, (uint64_t)n_total, ctx->d_wk);
    transcript_chain_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_wk, (uint64_t)n_total, ctx->d_r);
    mark(ctx, 4);
    CK(cudaGetLastError());
    return KZGB200_OK;
}
// K6: per-blob terms, tree sum, partial
static int launch_lincomb(kzgb200_ctx* ctx, size_t offset, Partial* d_out) {
    int n = (int)ctx->cur_n;
    msm_scalars_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_z_mont, ctx->d_zy, ctx->d_r, (uint64_t)offset, n, ctx->d_digits, ctx->d_ry);
    msm_sort_kernel<<<dim3(kWindows, 2), 256, 0, ctx->stream>>>(ctx->d_digits, n, ctx->d_order, ctx->d_start);
    msm_bucket_kernel<<<(kMsmSets * kWindows * kBuckets + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_C, ctx->d_P, n, ctx->d_order, ctx->d_start, ctx->d_buckets);
    mark(ctx, 5);
    msm_window_kernel<<<kMsmSets * kWindows, 32, 0, ctx->stream>>>(ctx->d_buckets, ctx->d_windows);
    msm_combine_kernel<<<1, 256, 0, ctx->stream>>>(ctx->d_windows, ctx->d_ry, ctx->d_status, n, d_out);
    mark(ctx, 6);
    CK(cudaGetLastError());
    return KZGB200_OK;
}
static int read_result(kzgb200_ctx* ctx, int* ok) {
    CK(cudaMemcpyAsync(ctx->h_result, ctx->d_result, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_result[0] == kBadArgs) return KZGB200_BAD_ARGS;
    *ok = ctx->h_result[0] == kTrue ? 1 : 0;
    return KZGB200_OK;
}
static int export_zy(kzgb200_ctx* ctx, size_t n, uint8_t* d_z, uint8_t* d_y) {
    if (!d_z && !d_y) return KZGB200_OK;
    export_scalars_kernel<<<((int)n + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_zy, (int)n, d_z, d_y);
    CK(cudaGetLastError());
    return KZGB200_OK;
}
// whole batch on one GPU, device-resident inputs, n >= 1
static int batch_device_locked(kzgb200_ctx* ctx, const uint8_t* d_blobs, const uint8_t* d_c, const uint8_t* d_p, size_t n, int* ok,
                               uint8_t* d_z_out, uint8_t* d_y_out) {
    int rc = launch_phase1(ctx, d_blobs, d_c, d_p, n);
    if (rc) return rc;
    if ((rc = export_zy(ctx, n, d_z_out, d_y_out))) return rc;
    if (n == 1) {   // single path (reference src/kzg_proof.rs:482-489)
        single_final_kernel<<<1, kFinalThreads, 0, ctx->stream>>>(ctx->d_C, ctx->d_P, ctx->d_zy, ctx->d_status, ctx->tables, ctx->d_result);
        CK(cudaGetLastError());
        return read_result(ctx, ok);
    }
    if ((rc = launch_transcript(ctx, d_c, ctx->d_zy, d_p, n))) return rc;
    if ((rc = launch_lincomb(ctx, 0, ctx->d_partial))) return rc;
    batch_final_kernel<<<1, kFinalThreads, 0, ctx->stream>>>(ctx->d_partial, 1, ctx->tables, ctx->d_result);
    mark(ctx, 7);
    CK(cudaGetLastError());
    rc = read_result(ctx, ok);
    if (ctx->profile && ctx->ev_used == 8)
        for (int i = 0; i < 7; i++) cudaEventElapsedTime(&ctx->phase_ms[i], ctx->ev[i], ctx->ev[i + 1]);
    return rc;
}

extern "C" int kzgb200_verify_blob_kzg_proof_batch_device(kzgb200_ctx* ctx, const uint8_t* d_blobs, const uint8_t* d_commitments,
                                                          const uint8_t* d_proofs, size_t n, int* ok, uint8_t* d_z_out, uint8_t* d_y_out) {
    if (!ctx || !ok || n == 0 || n > 0x7fffffff / 2) return KZGB200_BAD_ARGS;
    std::lock_guard<std::mutex> g(ctx->lock);
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_capacity(ctx, n, false);
    if (rc) return rc;
    return batch_device_locked(ctx, d_blobs, d_commitments, d_proofs, n, ok, d_z_out, d_y_out);
}

extern "C" int kzgb200_verify_blob_kzg_proof_batch(kzgb200_ctx* ctx, const uint8_t* blobs, size_t n_blobs, const uint8_t* commitments,
                                                   size_t n_commitments, const uint8_t* proofs, size_t n_proofs, int* ok,
                                                   uint8_t* z_out, uint8_t* y_out) {
    if (!ctx || !ok) return KZGB200_BAD_ARGS;
    if (n_blobs == 0) { *ok = 1; return KZGB200_OK; }                 // reference src/kzg_proof.rs:478-480
    if (n_blobs == 1) {                                                // :482-489, before the length checks
        if (n_commitments < 1 || n_proofs < 1) return KZGB200_BAD_ARGS;   // the reference would index out of bounds (panic)
    } else {
        if (n_blobs != n_commitments) return KZGB200_INVALID_LENGTH;   // :491-495
        if (n_blobs != n_proofs) return KZGB200_INVALID_LENGTH;        // :497-501
    }
    if (n_blobs > 0x7fffffff / 2) return KZGB200_BAD_ARGS;
    std::lock_guard<std::mutex> g(ctx->lock);
    CK(cudaSetDevice(ctx->device));
    size_t n = n_blobs;
    int rc = ensure_capacity(ctx, n, true);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->d_c, commitments, n * 48, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_p, proofs, n * 48, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_blobs, blobs, n * (size_t)kBytesPerBlob, cudaMemcpyHostToDevice, ctx->stream));
    rc = batch_device_locked(ctx, ctx->d_blobs, ctx->d_c, ctx->d_p, n, ok, z_out ? ctx->d_zout : nullptr, y_out ? ctx->d_yout : nullptr);
    if (rc == KZGB200_OK) {
        if (z_out) CK(cudaMemcpyAsync(z_out, ctx->d_zout, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
        if (y_out) CK(cudaMemcpyAsync(y_out, ctx->d_yout, n * 32, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    return rc;
}

extern "C" int kzgb200_verify_blob_kzg_proof(kzgb200_ctx* ctx, const uint8_t* blob, const uint8_t* commitment48, const uint8_t* proof48,
                                             int* ok, uint8_t* z_out, uint8_t* y_out) {
    return kzgb200_verify_blob_kzg_proof_batch(ctx, blob, 1, commitment48, 1, proof48, 1, ok, z_out, y_out);
}

extern "C" int kzgb200_verify_kzg_proof_many(kzgb200_ctx* ctx, const uint8_t* commitments, const uint8_t* zs, const uint8_t* ys,
                                             const uint8_t* proofs, size_t m, uint8_t* verdicts) {
    if (!ctx || (m && (!commitments || !zs || !ys || !proofs || !verdicts))) return KZGB200_BAD_ARGS;
    if (m == 0) return KZGB200_OK;
    std::lock_guard<std::mutex> g(ctx->lock);
    CK(cudaSetDevice(ctx->device));
    if (m > ctx->many_cap) { CK(regrow(ctx->d_many, m * 161)); ctx->many_cap = m; }
    uint8_t *dc = ctx->d_many, *dz = dc + m * 48, *dy = dz + m * 32, *dp = dy + m * 32, *dv = dp + m * 48;
    CK(cudaMemcpyAsync(dc, commitments, m * 48, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dz, zs, m * 32, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dy, ys, m * 32, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(dp, proofs, m * 48, cudaMemcpyHostToDevice, ctx->stream));
    verify_many_kernel<<<(unsigned)((m + 63) / 64), 64, 0, ctx->stream>>>(dc, dz, dy, dp, m, ctx->tables, dv);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(verdicts, dv, m, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return KZGB200_OK;
}

extern "C" int kzgb200_verify_kzg_proof(kzgb200_ctx* ctx, const uint8_t* commitment48, const uint8_t* z32, const uint8_t* y32,
                                        const uint8_t* proof48, int* ok) {
    if (!ok) return KZGB200_BAD_ARGS;
    uint8_t v = 0;
    int rc = kzgb200_verify_kzg_proof_many(ctx, commitment48, z32, y32, proof48, 1, &v);
    if (rc) return rc;
    if (v == kBadArgs) return KZGB200_BAD_ARGS;
    *ok = v == kTrue;
    return KZGB200_OK;
}

// ---- sharded batch ------------------------------------------------------------------------------------------
extern "C" int kzgb200_shard_evaluate(kzgb200_ctx* ctx, const uint8_t* d_blobs, const uint8_t* d_commitments, const uint8_t* d_proofs,
                                      size_t n_local, uint8_t* d_zy_out) {
    if (!ctx || n_local == 0 || n_local > 0x7fffffff / 2) return KZGB200_BAD_ARGS;
    std::lock_guard<std::mutex> g(ctx->lock);
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_capacity(ctx, n_local, false);
    if (rc) return rc;
    if ((rc = launch_phase1(ctx, d_blobs, d_commitments, d_proofs, n_local))) return rc;
    if (d_zy_out) CK(cudaMemcpyAsync(d_zy_out, ctx->d_zy, n_local * sizeof(ZY), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return KZGB200_OK;
}
extern "C" int kzgb200_shard_challenge(kzgb200_ctx* ctx, const uint8_t* d_all_commitments, const uint8_t* d_all_zy,
                                       const uint8_t* d_all_proofs, size_t n_total) {
    if (!ctx || n_total == 0) return KZGB200_BAD_ARGS;
    std::lock_guard<std::mutex> g(ctx->lock);
    CK(cudaSetDevice(ctx->device));
    int rc = launch_transcript(ctx, d_all_commitments, reinterpret_cast<const ZY*>(d_all_zy), d_all_proofs, n_total);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return KZGB200_OK;
}
extern "C" int kzgb200_shard_lincomb(kzgb200_ctx* ctx, size_t global_offset, uint8_t* d_partial_out) {
    if (!ctx || !d_partial_out || ctx->cur_n == 0) return KZGB200_BAD_ARGS;
    std::lock_guard<std::mutex> g(ctx->lock);
    CK(cudaSetDevice(ctx->device));
    int rc = launch_lincomb(ctx, global_offset, reinterpret_cast<Partial*>(d_partial_out));
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return KZGB200_OK;
}
extern "C" int kzgb200_shard_finalize(kzgb200_ctx* ctx, const uint8_t* d_partials, size_t n_ranks, int* ok) {
    if (!ctx || !d_partials || !ok || n_ranks == 0) return KZGB200_BAD_ARGS;
    std::lock_guard<std::mutex> g(ctx->lock);
    CK(cudaSetDevice(ctx->device));
    batch_final_kernel<<<1, kFinalThreads, 0, ctx->stream>>>(reinterpret_cast<const Partial*>(d_partials), (int)n_ranks, ctx->tables, ctx->d_result);
    CK(cudaGetLastError());
    return read_result(ctx, ok);
}

// ---- harness: synthetic workload with valid commitments / proofs (device outputs) ---------------------------
extern "C" int kzgb200_harness_generate(kzgb200_ctx* ctx, uint64_t seed, size_t n, int degree, const uint8_t* tau_powers48,
                                        uint8_t* d_blobs, uint8_t* d_commitments, uint8_t* d_proofs) {
    if (!ctx || n == 0 || degree < 2 || degree > kHarnessMaxDegree || !tau_powers48) return KZGB200_BAD_ARGS;
    std::lock_guard<std::mutex> g(ctx->lock);
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_capacity(ctx, n, false);
    if (rc) return rc;
    uint8_t* d_bytes = nullptr; G1Affine* d_M = nullptr; uint32_t* d_bad = nullptr; uint32_t bad = 0;
    CK(cudaMalloc(&d_bytes, degree * 48)); CK(cudaMalloc(&d_M, degree * sizeof(G1Affine))); CK(cudaMalloc(&d_bad, 4));
    CK(cudaMemsetAsync(d_bad, 0, 4, ctx->stream));
    CK(cudaMemcpyAsync(d_bytes, tau_powers48, degree * 48, cudaMemcpyHostToDevice, ctx->stream));
    harness_parse_points_kernel<<<1, 32, 0, ctx->stream>>>(d_bytes, degree, d_M, d_bad);
    int ni = (int)n;
    harness_blob_kernel<<<ni, 128, 0, ctx->stream>>>(seed, ni, degree, ctx->tables, d_blobs);
    harness_commit_kernel<<<(ni + 63) / 64, 64, 0, ctx->stream>>>(seed, ni, degree, d_M, nullptr, d_commitments, 0);
    challenge_kernel<<<(ni + 63) / 64, 64, 0, ctx->stream>>>(d_blobs, d_commitments, ni, ctx->d_z_mont, ctx->d_zy);
    harness_commit_kernel<<<(ni + 63) / 64, 64, 0, ctx->stream>>>(seed, ni, degree, d_M, ctx->d_z_mont, d_proofs, 1);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_bytes); cudaFree(d_M); cudaFree(d_bad);
    return bad ? KZGB200_BAD_ARGS : KZGB200_OK;
}
extern "C" int kzgb200_set_transcript_mode(kzgb200_ctx* ctx, int mode) {
    if (!ctx || (mode != KZGB200_TRANSCRIPT_EXACT && mode != KZGB200_TRANSCRIPT_TREE)) return KZGB200_BAD_ARGS;
    std::lock_guard<std::mutex> g(ctx->lock);
    ctx->transcript_mode = mode;
    return KZGB200_OK;
}
// the canonical big-endian r of the last batch (exact mode: bit-identical to compute_r_powers' r)
extern "C" int kzgb200_last_r(kzgb200_ctx* ctx, uint8_t* r_out32) {
    if (!ctx || !r_out32) return KZGB200_BAD_ARGS;
    std::lock_guard<std::mutex> g(ctx->lock);
    CK(cudaSetDevice(ctx->device));
    ZY* tmp = nullptr;
    CK(cudaMalloc(&tmp, sizeof(ZY)));
    r_to_raw_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_r, tmp);
    uint8_t* d_out = nullptr;
    CK(cudaMalloc(&d_out, 32));
    export_scalars_kernel<<<1, 1, 0, ctx->stream>>>(tmp, 1, d_out, nullptr);
    CK(cudaMemcpyAsync(r_out32, d_out, 32, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(tmp); cudaFree(d_out);
    return KZGB200_OK;
}
// raw per-rank partial of the last single-GPU batch (Partial struct: Jacobian A, B in Montgomery limbs, sum r_i y_i, flags)
extern "C" int kzgb200_last_partial(kzgb200_ctx* ctx, uint8_t* out352) {
    if (!ctx || !out352) return KZGB200_BAD_ARGS;
    std::lock_guard<std::mutex> g(ctx->lock);
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpyAsync(out352, ctx->d_partial, sizeof(Partial), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return KZGB200_OK;
}
// per-phase device times of the last single-GPU batch call: parse, challenge, eval, transcript, lincomb, reduce, final
extern "C" int kzgb200_set_profiling(kzgb200_ctx* ctx, int on) {
    if (!ctx) return KZGB200_BAD_ARGS;
    std::lock_guard<std::mutex> g(ctx->lock);
    CK(cudaSetDevice(ctx->device));
    if (on) for (auto& e : ctx->ev) if (!e) CK(cudaEventCreate(&e));
    ctx->profile = on != 0;
    return KZGB200_OK;
}
extern "C" int kzgb200_get_phase_ms(kzgb200_ctx* ctx, float* out7) {
    if (!ctx || !out7) return KZGB200_BAD_ARGS;
    for (int i = 0; i < 7; i++) out7[i] = ctx->phase_ms[i];
    return KZGB200_OK;
}
// the stream every call of this context is issued on (cudaStream_t), for event timing by the caller
extern "C" void* kzgb200_stream(kzgb200_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" void* kzgb200_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    return cudaMallocHost(&p, bytes) == cudaSuccess ? p : nullptr;
}
extern "C" void kzgb200_free_pinned(void* p) { if (p) cudaFreeHost(p); }
