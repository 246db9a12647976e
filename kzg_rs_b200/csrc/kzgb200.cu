// libkzgb200.so -- host runtime + C ABI (include/kzgb200.h) over the sm_100a kernels.
// Mirrors the orchestration of KzgProof::{verify_kzg_proof, verify_blob_kzg_proof,
// verify_blob_kzg_proof_batch} (reference src/kzg_proof.rs:353-525): argument checks and phase ordering on
// the host, every arithmetic step in a kernel.  No CPU fallback: CUDA failures surface as
// KZGB200_INTERNAL_ERROR.
#include <cuda_runtime.h>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <new>
#include "../../include/kzgb200.h"
#include "common.cuh"

using namespace kzgb200;

static_assert(sizeof(Partial) == KZGB200_PARTIAL_BYTES, "Partial layout is part of the ABI");
static_assert(sizeof(ZY) == 64, "ZY layout is part of the ABI");

constexpr int kTailSms = 8;                    // SMs kept free of deferred subgroup checks for the latency-bound tail kernels
constexpr int kTailHogSmem = 200 * 1024;       // dynamic shared memory of a subgroup-check CTA in deferred mode (never touched)
constexpr int kTailPadSmem = 28 * 1024;        // ... and of the tail kernels: 200 KB + 28 KB do not fit one SM
static_assert(sizeof(FinalSmem) >= (size_t)kTailPadSmem, "the pairing kernel must not fit beside a subgroup-check CTA");
struct kzgb200_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    DeviceTables* tables = nullptr;
    // workspace, sized for `cap` blobs
    size_t cap = 0, blob_cap = 0, many_cap = 0;
    uint8_t *d_blobs = nullptr, *d_c = nullptr, *d_p = nullptr;     // staging of host inputs
    Fr* d_z_mont = nullptr;
    Fr* d_zpow = nullptr;           // z^(2^k), k = 0..12, per blob (K2 -> K1/K3)
    ZY* d_zy = nullptr;
    G1Affine *d_C = nullptr, *d_P = nullptr;
    uint32_t* d_status = nullptr;
    Fr *d_ry = nullptr, *d_r = nullptr;
    uint8_t* d_digits = nullptr;            // [4*16][cap]
    uint32_t *d_order = nullptr, *d_start = nullptr;
    G1 *d_buckets = nullptr, *d_windows = nullptr;
    Partial* d_partial = nullptr;
    uint32_t* d_result = nullptr;
    uint8_t *d_zout = nullptr, *d_yout = nullptr;
    uint8_t *d_many = nullptr;
    G1* d_lag_table = nullptr;      // [4096][32][255] window table of the Lagrange G1 points (commit / prove only)
    Fr* d_scalars = nullptr;
    uint32_t* d_wk = nullptr;       // transcript W+K words, 64 per SHA block
    size_t wk_cap = 0;
    uint32_t* h_result = nullptr;   // pinned
    // inputs of the current shard (device pointers owned by the caller or by the staging buffers)
    const uint8_t *cur_c = nullptr, *cur_p = nullptr;
    size_t cur_n = 0;
    // optional per-phase timing (CUDA events on the context stream)
    int transcript_mode = KZGB200_TRANSCRIPT_EXACT;
    int num_sms = 148;
    cudaEvent_t ev_sha0 = nullptr;
    int parse_fused = 0;            // tuning: decompression + subgroup check in one kernel (env KZGB200_PARSE_FUSED; measured slower)
    int defer_subgroup = 1;         // single-GPU batches: subgroup checks run beside the latency-bound tail on their own SMs (env KZGB200_DEFER_SUBGROUP)
    bool subgroup_pending = false;
    cudaEvent_t ev_bucket = nullptr;
    int parse_first = 0;            // tuning: launch G1 parsing before the first hash launch (env KZGB200_PARSE_FIRST)
    cudaStream_t s_aux = nullptr, s_copy = nullptr, s_work[4] = {nullptr, nullptr, nullptr, nullptr};
    cudaEvent_t ev_begin = nullptr, ev_parse = nullptr, ev_decomp = nullptr, ev_h2d[64] = {nullptr}, ev_zy[64] = {nullptr};
    uint32_t* d_chain_state = nullptr;
    uint8_t* d_scratch = nullptr;   // 256 bytes for small exports
    size_t tr_done = 0;             // transcript blocks (exact) / leaf groups (tree) already hashed
    // optional per-phase timing (CUDA event pairs on the stream each phase runs on)
    bool profile = false;
    cudaEvent_t ev_s[8] = {nullptr}, ev_e[8] = {nullptr};
    bool ph_started[8] = {false};
    float phase_ms[8] = {0};
    std::mutex lock;
    char err[256] = {0};
};
enum Phase { kPhParse = 0, kPhChallenge, kPhEval, kPhTranscript, kPhLincomb, kPhReduce, kPhFinal, kPhCount };
constexpr int kMaxChunks = 64, kWorkStreams = 4;
constexpr size_t kMinChunk = 1024;   // blobs per host->device chunk (128 MiB)

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
        CK(regrow(ctx->d_z_mont, c)); CK(regrow(ctx->d_zy, c)); CK(regrow(ctx->d_zpow, c * 13));
        CK(regrow(ctx->d_C, c)); CK(regrow(ctx->d_P, c));
        CK(regrow(ctx->d_status, c)); CK(regrow(ctx->d_ry, c));
        CK(regrow(ctx->d_digits, c * kDigitRows)); CK(regrow(ctx->d_order, c * kDigitRows));
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
        CK(cudaDeviceGetAttribute(&ctx->num_sms, cudaDevAttrMultiProcessorCount, device));
        { int lo = 0, hi = 0; CK(cudaDeviceGetStreamPriorityRange(&lo, &hi)); CK(cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, hi)); }
        int prio_lo = 0, prio_hi = 0;
        CK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
        // the hash chains are the long pole of phase 1: their CTAs go first, G1 parsing fills the rest of the machine
        { const char* v = getenv("KZGB200_PARSE_PRIO"); CK(cudaStreamCreateWithPriority(&ctx->s_aux, cudaStreamNonBlocking, v && atoi(v) ? prio_hi : prio_lo)); }
        if (const char* v = getenv("KZGB200_PARSE_FIRST")) ctx->parse_first = atoi(v);
        if (const char* v = getenv("KZGB200_PARSE_FUSED")) ctx->parse_fused = atoi(v);
        if (const char* v = getenv("KZGB200_DEFER_SUBGROUP")) ctx->defer_subgroup = atoi(v);
        CK(cudaEventCreateWithFlags(&ctx->ev_bucket, cudaEventDisableTiming));
        CK(cudaFuncSetAttribute(g1_subgroup_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kTailHogSmem));
        CK(cudaStreamCreateWithFlags(&ctx->s_copy, cudaStreamNonBlocking));
        for (auto& w : ctx->s_work) CK(cudaStreamCreateWithPriority(&w, cudaStreamNonBlocking, prio_hi));
        CK(cudaEventCreateWithFlags(&ctx->ev_begin, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ev_parse, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&ctx->ev_decomp, cudaEventDisableTiming));
        for (auto& e : ctx->ev_h2d) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : ctx->ev_zy) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CK(cudaMalloc(&ctx->d_chain_state, 32));
        CK(cudaMalloc(&ctx->d_scratch, 512));
        CK(cudaFuncSetAttribute(batch_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FinalSmem)));
        CK(cudaFuncSetAttribute(single_final_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(FinalSmem)));
        CK(cudaEventCreateWithFlags(&ctx->ev_sha0, cudaEventDisableTiming));
        CK(cudaMalloc(&ctx->tables, sizeof(DeviceTables)));
        CK(cudaMalloc(&ctx->d_r, sizeof(Fr)));
        CK(cudaMalloc(&ctx->d_partial, sizeof(Partial)));
        CK(cudaMalloc(&ctx->d_result, 16));
        CK(cudaMalloc(&ctx->d_start, kDigitRows * (kBuckets + 1) * sizeof(uint32_t)));
        CK(cudaMalloc(&ctx->d_buckets, kMsmSets * kWindows * kBuckets * sizeof(G1)));
        CK(cudaMalloc(&ctx->d_windows, kMsmSets * kWindows * sizeof(G1)));
        CK(cudaMallocHost(&ctx->h_result, 16));
        uint8_t* d_g2 = nullptr;
        CK(cudaMalloc(&d_g2, 192));
        CK(cudaMemcpyAsync(d_g2, g2_points, 192, cudaMemcpyHostToDevice, ctx->stream));
        setup_tables_kernel<<<(8192 + 4096) / 128, 128, 0, ctx->stream>>>(ctx->tables, d_g2);
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
    cudaDeviceSynchronize();
    for (cudaStream_t st : {ctx->s_aux, ctx->s_copy, ctx->s_work[0], ctx->s_work[1], ctx->s_work[2], ctx->s_work[3]}) if (st) cudaStreamDestroy(st);
    for (cudaEvent_t e : {ctx->ev_begin, ctx->ev_parse, ctx->ev_decomp, ctx->ev_sha0, ctx->ev_bucket}) if (e) cudaEventDestroy(e);
    for (auto e : ctx->ev_h2d) if (e) cudaEventDestroy(e);
    for (auto e : ctx->ev_zy) if (e) cudaEventDestroy(e);
    for (auto e : ctx->ev_s) if (e) cudaEventDestroy(e);
    for (auto e : ctx->ev_e) if (e) cudaEventDestroy(e);
    if (ctx->d_chain_state) cudaFree(ctx->d_chain_state);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    void* ptrs[] = {ctx->tables, ctx->d_blobs, ctx->d_c, ctx->d_p, ctx->d_z_mont, ctx->d_zy, ctx->d_C, ctx->d_P, ctx->d_status,
                    ctx->d_ry, ctx->d_r, ctx->d_partial, ctx->d_result, ctx->d_zout, ctx->d_yout, ctx->d_many, ctx->d_wk,
                    ctx->d_digits, ctx->d_order, ctx->d_start, ctx->d_buckets, ctx->d_windows, ctx->d_lag_table, ctx->d_scalars, ctx->d_zpow};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (ctx->h_result) cudaFreeHost(ctx->h_result);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx;
}

extern "C" const char* kzgb200_last_error(const kzgb200_ctx* ctx) { return ctx ? ctx->err : "null context"; }

// ---- phases ---------------------------------------------------------------------------------------------------
// Streams: `stream` (main: transcript, MSM, pairing, result), `s_aux` (G1 decompression beside the hashing; the subgroup
// checks beside the tail, see launch_lincomb),
// `s_work[k]` (per-chunk challenge -> evaluation), `s_copy` (host->device blob chunks).  A device-resident batch is one
// chunk (the per-blob SHA-256 chain is latency-bound: splitting it buys nothing); a host batch is cut into chunks so
// that hashing / evaluation / the serial transcript chain of chunk c overlap the PCIe copy of chunk c+1.
static void phase_begin(kzgb200_ctx* ctx, int ph, cudaStream_t st) { if (ctx->profile && !ctx->ph_started[ph]) { cudaEventRecord(ctx->ev_s[ph], st); ctx->ph_started[ph] = true; } }
static void phase_end(kzgb200_ctx* ctx, int ph, cudaStream_t st) { if (ctx->profile) cudaEventRecord(ctx->ev_e[ph], st); }

// transcript blocks / tree groups that only depend on entries < avail
static int advance_transcript(kzgb200_ctx* ctx, const uint8_t* d_c, const ZY* d_zy, const uint8_t* d_p, size_t n, size_t avail) {
    if (ctx->transcript_mode == KZGB200_TRANSCRIPT_TREE) {
        size_t ngroups = (n + kTreeGroup - 1) / kTreeGroup;
        size_t ready = avail >= n ? ngroups : avail / kTreeGroup;
        if (ready > ctx->tr_done) {
            phase_begin(ctx, kPhTranscript, ctx->stream);
            size_t cnt = ready - ctx->tr_done;
            // entries of the newly complete groups -> word image -> leaf digests
            size_t e0 = ctx->tr_done * kTreeGroup, e1 = ready * kTreeGroup < n ? ready * kTreeGroup : n;
            uint32_t* words = ctx->d_wk + ((n + kTreeGroup - 1) / kTreeGroup) * 8 + 64;
            transcript_words_kernel<<<(unsigned)(((e1 - e0) * 40 + 255) / 256), 256, 0, ctx->stream>>>(d_c, d_zy, d_p, e0, e1 - e0, words);
            transcript_tree_leaf_words_kernel<<<(unsigned)((cnt + 63) / 64), 64, 0, ctx->stream>>>(words, (uint64_t)n, ctx->d_wk, ctx->tr_done, cnt);
            ctx->tr_done = ready;
        }
        if (avail >= n) {
            transcript_tree_root_kernel<<<1, 256, 0, ctx->stream>>>(ctx->d_wk, (uint64_t)n, ctx->d_wk + ((n + kTreeGroup - 1) / kTreeGroup) * 8 + 64 + n * 40 + 64,
                                                                   ctx->d_r);
            phase_end(ctx, kPhTranscript, ctx->stream);
        }
    } else {
        size_t nblk = (32 + n * 160 + 9 + 63) / 64;
        size_t ready = avail >= n ? nblk : (32 + avail * 160) / 64;
        if (ready > ctx->tr_done) {
            phase_begin(ctx, kPhTranscript, ctx->stream);
            size_t cnt = ready - ctx->tr_done;
            transcript_schedule_kernel<<<(unsigned)((cnt + 127) / 128), 128, 0, ctx->stream>>>(d_c, d_zy, d_p, (uint64_t)n, ctx->d_wk, ctx->tr_done, cnt);
            transcript_chain_kernel<<<1, 32, 0, ctx->stream>>>(ctx->d_wk, (uint64_t)n, ctx->d_r, ctx->d_chain_state, ctx->tr_done, cnt);
            ctx->tr_done = ready;
        }
        if (avail >= n) phase_end(ctx, kPhTranscript, ctx->stream);
    }
    CK(cudaGetLastError());
    return KZGB200_OK;
}
static int reserve_transcript(kzgb200_ctx* ctx, size_t n) {
    size_t words = ctx->transcript_mode == KZGB200_TRANSCRIPT_TREE ? ((n + kTreeGroup - 1) / kTreeGroup) * 8 + 64 + n * 40 + 64 + ((n + kTreeGroup * kTreeMid - 1) / (kTreeGroup * kTreeMid)) * 8 + 8
                                                                   : ((32 + n * 160 + 9 + 63) / 64) * 64;
    if (words > ctx->wk_cap) { CK(regrow(ctx->d_wk, words)); ctx->wk_cap = words; }
    ctx->tr_done = 0;
    return KZGB200_OK;
}
// phase 1 for blobs [0, n): K4 on s_aux; per chunk K2 -> K1/K3 on a work stream; optionally the transcript advances
// on the main stream as chunks complete.  h_blobs != nullptr: the blobs are copied chunk by chunk from the host.
static int launch_phase1(kzgb200_ctx* ctx, const uint8_t* d_blobs, const uint8_t* h_blobs, const uint8_t* d_c, const uint8_t* d_p, size_t n,
                         bool with_transcript, bool defer_subgroup = false) {
    for (int i = 0; i < kPhCount; i++) ctx->ph_started[i] = false;
    size_t chunk = n;
    if (h_blobs) { chunk = (n + kMaxChunks - 1) / kMaxChunks; if (chunk < kMinChunk) chunk = kMinChunk; }
    else if (with_transcript && ctx->transcript_mode == KZGB200_TRANSCRIPT_EXACT && n >= 4 * kMinChunk) {
        // resident batch, serial transcript: a few chunks on separate streams let the chain start as soon as the
        // first chunk's z, y exist instead of after the evaluation of the whole batch
        chunk = (n + 7) / 8; if (chunk < kMinChunk) chunk = kMinChunk;
    }
    size_t nchunks = (n + chunk - 1) / chunk;
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_parse, 0));    // a previous call's deferred subgroup checks still write status
    CK(cudaMemsetAsync(ctx->d_status, 0, n * sizeof(uint32_t), ctx->stream));
    CK(cudaEventRecord(ctx->ev_begin, ctx->stream));
    if (with_transcript) { int rc = reserve_transcript(ctx, n); if (rc) return rc; }
    if (h_blobs) CK(cudaStreamWaitEvent(ctx->s_copy, ctx->ev_begin, 0));
    auto launch_parse = [&]() -> int {
        CK(cudaStreamWaitEvent(ctx->s_aux, ctx->ev_begin, 0));
        phase_begin(ctx, kPhParse, ctx->s_aux);
        g1_decompress_kernel<<<(2 * (int)n + 127) / 128, 128, 0, ctx->s_aux>>>(d_c, d_p, (int)n, ctx->d_C, ctx->d_P, ctx->d_status, ctx->parse_fused != 0);
        CK(cudaEventRecord(ctx->ev_decomp, ctx->s_aux));
        if (defer_subgroup && !ctx->parse_fused) { ctx->subgroup_pending = true; return KZGB200_OK; }   // launched by launch_lincomb
        if (!ctx->parse_fused) g1_subgroup_kernel<<<(2 * (int)n + 255) / 256, 256, 0, ctx->s_aux>>>(ctx->d_C, ctx->d_P, (int)n, ctx->d_status);
        phase_end(ctx, kPhParse, ctx->s_aux);
        CK(cudaEventRecord(ctx->ev_parse, ctx->s_aux));
        return KZGB200_OK;
    };
    if (ctx->parse_first) { int rc = launch_parse(); if (rc) return rc; }
    for (size_t c = 0; c < nchunks; c++) {
        size_t lo = c * chunk, cnt = n - lo < chunk ? n - lo : chunk;
        cudaStream_t sw = ctx->s_work[c % kWorkStreams];
        if (h_blobs) {
            CK(cudaMemcpyAsync(const_cast<uint8_t*>(d_blobs) + lo * kBytesPerBlob, h_blobs + lo * kBytesPerBlob, cnt * (size_t)kBytesPerBlob,
                               cudaMemcpyHostToDevice, ctx->s_copy));
            CK(cudaEventRecord(ctx->ev_h2d[c], ctx->s_copy));
            CK(cudaStreamWaitEvent(sw, ctx->ev_h2d[c], 0));
        }
        CK(cudaStreamWaitEvent(sw, ctx->ev_begin, 0));
        phase_begin(ctx, kPhChallenge, sw);
        challenge_kernel<<<((int)cnt + kShaThreads - 1) / kShaThreads, kShaThreads, 0, sw>>>(d_blobs + lo * kBytesPerBlob, d_c + lo * 48, (int)cnt, ctx->d_z_mont + lo, ctx->d_zy + lo,
                                                              ctx->d_zpow + lo * 13, 1u);
        phase_end(ctx, kPhChallenge, sw);
        if (c == 0) CK(cudaEventRecord(ctx->ev_sha0, sw));
        phase_begin(ctx, kPhEval, sw);
        eval_kernel<<<(int)cnt, kEvalThreads, 0, sw>>>(d_blobs + lo * kBytesPerBlob, (int)cnt, ctx->d_zpow + lo * 13, ctx->tables, ctx->d_zy + lo,
                                                       ctx->d_status + lo);
        phase_end(ctx, kPhEval, sw);
        CK(cudaEventRecord(ctx->ev_zy[c], sw));
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_zy[c], 0));
        if (c == 0 && !ctx->parse_first) { int rc = launch_parse(); if (rc) return rc; }   // G1 parsing queued behind the first hash launch
        if (with_transcript && n >= 2) { int rc = advance_transcript(ctx, d_c, ctx->d_zy, d_p, n, lo + cnt); if (rc) return rc; }
    }
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_decomp, 0));   // the points exist; the subgroup verdicts (ev_parse) are awaited
    CK(cudaGetLastError());                                     // only where the error flags are consumed
    ctx->cur_c = d_c; ctx->cur_p = d_p; ctx->cur_n = n;
    return KZGB200_OK;
}
// K5 over gathered arrays (sharded path: every rank derives the same r)
static int launch_transcript(kzgb200_ctx* ctx, const uint8_t* d_all_c, const ZY* d_all_zy, const uint8_t* d_all_p, size_t n_total) {
    int rc = reserve_transcript(ctx, n_total);
    if (rc) return rc;
    return advance_transcript(ctx, d_all_c, d_all_zy, d_all_p, n_total, n_total);
}
// K6: digits, counting sort, buckets, window sums, Horner -> partial
static int launch_lincomb(kzgb200_ctx* ctx, size_t offset, Partial* d_out, bool wait_subgroup) {
    int n = (int)ctx->cur_n;
    phase_begin(ctx, kPhLincomb, ctx->stream);
    msm_scalars_kernel<<<(n + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_z_mont, ctx->d_zy, ctx->d_r, (uint64_t)offset, n, ctx->d_digits, ctx->d_ry);
    msm_sort_kernel<<<kDigitRows, 256, 0, ctx->stream>>>(ctx->d_digits, n, ctx->d_order, ctx->d_start);
    msm_bucket_kernel<<<(kMsmSets * kWindows * kBuckets * kBucketSplit + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_C, ctx->d_P, n, ctx->d_order, ctx->d_start, ctx->d_buckets);
    phase_end(ctx, kPhLincomb, ctx->stream);
    if (ctx->subgroup_pending) {
        // Deferred subgroup checks: from here on the batch is latency-bound (window sums, Horner combination, one pairing: a
        // few CTAs), so the checks get the rest of the machine.  They run one CTA per SM on all but kTailSms SMs -- each CTA
        // asks for kTailHogSmem of shared memory it never touches, and the tail kernels ask for kTailPadSmem, so that the two
        // cannot share an SM: the tail keeps SMs of its own and is not slowed down (sharing SMs cost more than the deferral saved).
        CK(cudaEventRecord(ctx->ev_bucket, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->s_aux, ctx->ev_bucket, 0));
        g1_subgroup_kernel<<<ctx->num_sms - kTailSms, 256, kTailHogSmem, ctx->s_aux>>>(ctx->d_C, ctx->d_P, n, ctx->d_status);
        phase_end(ctx, kPhParse, ctx->s_aux);
        CK(cudaEventRecord(ctx->ev_parse, ctx->s_aux));
        ctx->subgroup_pending = false;
    }
    phase_begin(ctx, kPhReduce, ctx->stream);
    msm_window_kernel<<<kMsmSets * kWindows, kWinLanes, kTailPadSmem, ctx->stream>>>(ctx->d_buckets, ctx->d_windows);
    if (wait_subgroup) CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_parse, 0));    // combine ORs the per-blob error flags
    msm_combine_kernel<<<1, 256, kTailPadSmem, ctx->stream>>>(ctx->d_windows, ctx->d_ry, ctx->d_status, n, d_out);
    phase_end(ctx, kPhReduce, ctx->stream);
    CK(cudaGetLastError());
    return KZGB200_OK;
}
static int read_result(kzgb200_ctx* ctx, int* ok) {
    CK(cudaMemcpyAsync(ctx->h_result, ctx->d_result, 12, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (ctx->h_result[0] == kBadArgs || ctx->h_result[2]) return KZGB200_BAD_ARGS;
    *ok = ctx->h_result[0] == kTrue ? 1 : 0;
    return KZGB200_OK;
}
static int export_zy(kzgb200_ctx* ctx, size_t n, uint8_t* d_z, uint8_t* d_y) {
    if (!d_z && !d_y) return KZGB200_OK;
    export_scalars_kernel<<<((int)n + 127) / 128, 128, 0, ctx->stream>>>(ctx->d_zy, (int)n, d_z, d_y);
    CK(cudaGetLastError());
    return KZGB200_OK;
}
// whole batch on one GPU, n >= 1; blobs either resident (h_blobs == nullptr) or streamed from the host
static int batch_locked(kzgb200_ctx* ctx, const uint8_t* d_blobs, const uint8_t* h_blobs, const uint8_t* d_c, const uint8_t* d_p, size_t n, int* ok,
                        uint8_t* d_z_out, uint8_t* d_y_out) {
    const bool defer = n >= 2 && ctx->defer_subgroup;
    int rc = launch_phase1(ctx, d_blobs, h_blobs, d_c, d_p, n, true, defer);
    if (rc) return rc;
    if ((rc = export_zy(ctx, n, d_z_out, d_y_out))) return rc;
    phase_begin(ctx, kPhFinal, ctx->stream);
    if (n == 1) {   // single path (reference src/kzg_proof.rs:482-489)
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_parse, 0));
        single_final_kernel<<<1, kFinalThreads, sizeof(FinalSmem), ctx->stream>>>(ctx->d_C, ctx->d_P, ctx->d_zy, ctx->d_status, ctx->tables, ctx->d_result);
    } else {
        ctx->ph_started[kPhFinal] = false;
        if ((rc = launch_lincomb(ctx, 0, ctx->d_partial, !defer))) return rc;   // deferred: the flags are merged by status_or below
        phase_begin(ctx, kPhFinal, ctx->stream);
        batch_final_kernel<<<1, kFinalThreads, sizeof(FinalSmem), ctx->stream>>>(ctx->d_partial, 1, ctx->tables, ctx->d_result, reinterpret_cast<long long*>(ctx->d_scratch + 128));
    }
    phase_end(ctx, kPhFinal, ctx->stream);
    // the deferred subgroup checks may still be running beside the pairing: their flags are merged last
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_parse, 0));
    status_or_kernel<<<1, 256, 0, ctx->stream>>>(ctx->d_status, (int)n, ctx->d_result + 2);
    CK(cudaGetLastError());
    rc = read_result(ctx, ok);
    if (ctx->profile) {
        cudaDeviceSynchronize();
        for (int i = 0; i < kPhCount; i++) { ctx->phase_ms[i] = 0; if (ctx->ph_started[i]) cudaEventElapsedTime(&ctx->phase_ms[i], ctx->ev_s[i], ctx->ev_e[i]); }
    }
    return rc;
}

extern "C" int kzgb200_verify_blob_kzg_proof_batch_device(kzgb200_ctx* ctx, const uint8_t* d_blobs, const uint8_t* d_commitments,
                                                          const uint8_t* d_proofs, size_t n, int* ok, uint8_t* d_z_out, uint8_t* d_y_out) {
    if (!ctx || !ok || n == 0 || n > 0x7fffffff / 2) return KZGB200_BAD_ARGS;
    std::lock_guard<std::mutex> g(ctx->lock);
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_capacity(ctx, n, false);
    if (rc) return rc;
    return batch_locked(ctx, d_blobs, nullptr, d_commitments, d_proofs, n, ok, d_z_out, d_y_out);
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
    rc = batch_locked(ctx, ctx->d_blobs, blobs, ctx->d_c, ctx->d_p, n, ok, z_out ? ctx->d_zout : nullptr, y_out ? ctx->d_yout : nullptr);
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
    if ((rc = launch_phase1(ctx, d_blobs, nullptr, d_commitments, d_proofs, n_local, false, ctx->defer_subgroup != 0))) return rc;
    if (d_zy_out) CK(cudaMemcpyAsync(d_zy_out, ctx->d_zy, n_local * sizeof(ZY), cudaMemcpyDeviceToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return KZGB200_OK;
}
// same with the shard's inputs in host memory: chunked host->device copies overlapped with hashing / evaluation
extern "C" int kzgb200_shard_evaluate_host(kzgb200_ctx* ctx, const uint8_t* blobs, const uint8_t* commitments, const uint8_t* proofs,
                                           size_t n_local, uint8_t* d_commitments_out, uint8_t* d_proofs_out, uint8_t* d_zy_out) {
    if (!ctx || n_local == 0 || n_local > 0x7fffffff / 2) return KZGB200_BAD_ARGS;
    std::lock_guard<std::mutex> g(ctx->lock);
    CK(cudaSetDevice(ctx->device));
    int rc = ensure_capacity(ctx, n_local, true);
    if (rc) return rc;
    CK(cudaMemcpyAsync(ctx->d_c, commitments, n_local * 48, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemcpyAsync(ctx->d_p, proofs, n_local * 48, cudaMemcpyHostToDevice, ctx->stream));
    if ((rc = launch_phase1(ctx, ctx->d_blobs, blobs, ctx->d_c, ctx->d_p, n_local, false, ctx->defer_subgroup != 0))) return rc;
    if (d_commitments_out) CK(cudaMemcpyAsync(d_commitments_out, ctx->d_c, n_local * 48, cudaMemcpyDeviceToDevice, ctx->stream));
    if (d_proofs_out) CK(cudaMemcpyAsync(d_proofs_out, ctx->d_p, n_local * 48, cudaMemcpyDeviceToDevice, ctx->stream));
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
    // deferred subgroup checks (launched in here, beside the tail): the partial then carries only the flags known so far;
    // kzgb200_shard_finalize merges this rank's late flags into its own return code
    int rc = launch_lincomb(ctx, global_offset, reinterpret_cast<Partial*>(d_partial_out), !ctx->subgroup_pending);
    if (rc) return rc;
    CK(cudaStreamSynchronize(ctx->stream));
    return KZGB200_OK;
}
extern "C" int kzgb200_shard_finalize(kzgb200_ctx* ctx, const uint8_t* d_partials, size_t n_ranks, int* ok) {
    if (!ctx || !d_partials || !ok || n_ranks == 0) return KZGB200_BAD_ARGS;
    std::lock_guard<std::mutex> g(ctx->lock);
    CK(cudaSetDevice(ctx->device));
    batch_final_kernel<<<1, kFinalThreads, sizeof(FinalSmem), ctx->stream>>>(reinterpret_cast<const Partial*>(d_partials), (int)n_ranks, ctx->tables, ctx->d_result, nullptr);
    CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_parse, 0));
    status_or_kernel<<<1, 256, 0, ctx->stream>>>(ctx->d_status, (int)ctx->cur_n, ctx->d_result + 2);
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
    challenge_kernel<<<(ni + kShaThreads - 1) / kShaThreads, kShaThreads, 0, ctx->stream>>>(d_blobs, d_commitments, ni, ctx->d_z_mont, ctx->d_zy, ctx->d_zpow, 1u);
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
    // scratch: the first ZY slot of the (idle) z/y export buffers
    ZY* tmp = reinterpret_cast<ZY*>(ctx->d_scratch);
    uint8_t* d_out = ctx->d_scratch + sizeof(ZY);
    r_to_raw_kernel<<<1, 1, 0, ctx->stream>>>(ctx->d_r, tmp);
    export_scalars_kernel<<<1, 1, 0, ctx->stream>>>(tmp, 1, d_out, nullptr);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(r_out32, d_out, 32, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
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
    if (on) { for (auto& e : ctx->ev_s) if (!e) CK(cudaEventCreate(&e)); for (auto& e : ctx->ev_e) if (!e) CK(cudaEventCreate(&e)); }
    ctx->profile = on != 0;
    return KZGB200_OK;
}
extern "C" int kzgb200_get_phase_ms(kzgb200_ctx* ctx, float* out7) {
    if (!ctx || !out7) return KZGB200_BAD_ARGS;
    for (int i = 0; i < 7; i++) out7[i] = ctx->phase_ms[i];
    return KZGB200_OK;
}
// ---- commit / prove ------------------------------------------------------------------------------
extern "C" int kzgb200_load_g1_lagrange(kzgb200_ctx* ctx, const uint8_t* g1_lagrange, size_t n_points) {
    if (!ctx || !g1_lagrange || n_points != (size_t)kFieldElementsPerBlob) return KZGB200_INVALID_SETUP;
    std::lock_guard<std::mutex> g(ctx->lock);
    CK(cudaSetDevice(ctx->device));
    if (ctx->d_lag_table) return KZGB200_OK;
    uint8_t* d_bytes = nullptr; G1Affine* d_L = nullptr; uint32_t* d_bad = nullptr; uint32_t bad = 0;
    CK(cudaMalloc(&d_bytes, n_points * 48)); CK(cudaMalloc(&d_L, n_points * sizeof(G1Affine))); CK(cudaMalloc(&d_bad, 4));
    CK(cudaMalloc(&ctx->d_lag_table, (size_t)kFieldElementsPerBlob * kLagWindows * kLagEntries * sizeof(G1)));
    CK(cudaMemsetAsync(d_bad, 0, 4, ctx->stream));
    CK(cudaMemcpyAsync(d_bytes, g1_lagrange, n_points * 48, cudaMemcpyHostToDevice, ctx->stream));
    lag_parse_kernel<<<(kFieldElementsPerBlob + 127) / 128, 128, 0, ctx->stream>>>(d_bytes, d_L, d_bad);
    lag_table_kernel<<<(kFieldElementsPerBlob * kLagWindows + 127) / 128, 128, 0, ctx->stream>>>(d_L, ctx->d_lag_table);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(&bad, d_bad, 4, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(d_bytes); cudaFree(d_L); cudaFree(d_bad);
    if (bad) { cudaFree(ctx->d_lag_table); ctx->d_lag_table = nullptr; return KZGB200_INVALID_SETUP; }
    return KZGB200_OK;
}
// shared driver: want_proof == 0 -> commitments of the blobs; 1 -> proofs at the Fiat-Shamir challenge of (blob, commitment)
static int commit_or_prove(kzgb200_ctx* ctx, const uint8_t* d_blobs, const uint8_t* d_commitments, size_t n, uint8_t* d_out, int want_proof) {
    if (!ctx->d_lag_table) return KZGB200_INVALID_SETUP;
    const size_t kChunk = 1024;
    int rc = ensure_capacity(ctx, n < kChunk ? n : kChunk, false);
    if (rc) return rc;
    if (!ctx->d_scalars) CK(cudaMalloc(&ctx->d_scalars, kChunk * (size_t)kFieldElementsPerBlob * sizeof(Fr)));
    uint32_t any_bad = 0;
    for (size_t lo = 0; lo < n; lo += kChunk) {
        int cnt = (int)(n - lo < kChunk ? n - lo : kChunk);
        const uint8_t* blobs = d_blobs + lo * kBytesPerBlob;
        CK(cudaMemsetAsync(ctx->d_status, 0, cnt * sizeof(uint32_t), ctx->stream));
        if (!want_proof) {
            blob_scalars_kernel<<<(unsigned)(((size_t)cnt * kFieldElementsPerBlob + 127) / 128), 128, 0, ctx->stream>>>(blobs, cnt, ctx->d_scalars, ctx->d_status);
        } else {
            challenge_kernel<<<(cnt + kShaThreads - 1) / kShaThreads, kShaThreads, 0, ctx->stream>>>(blobs, d_commitments + lo * 48, cnt, ctx->d_z_mont, ctx->d_zy, ctx->d_zpow, 1u);
            eval_kernel<<<cnt, kEvalThreads, 0, ctx->stream>>>(blobs, cnt, ctx->d_zpow, ctx->tables, ctx->d_zy, ctx->d_status);
            quotient_kernel<<<cnt, kEvalThreads, 0, ctx->stream>>>(blobs, cnt, ctx->d_z_mont, ctx->d_zy, ctx->tables, ctx->d_scalars);
        }
        lag_msm_kernel<<<cnt, 256, 0, ctx->stream>>>(ctx->d_scalars, cnt, ctx->d_lag_table, d_out + lo * 48);
        status_or_kernel<<<1, 256, 0, ctx->stream>>>(ctx->d_status, cnt, ctx->d_result);
        CK(cudaGetLastError());
        CK(cudaMemcpyAsync(ctx->h_result, ctx->d_result, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
        any_bad |= ctx->h_result[0];
    }
    return any_bad ? KZGB200_BAD_ARGS : KZGB200_OK;
}
extern "C" int kzgb200_blob_to_kzg_commitment_batch(kzgb200_ctx* ctx, const uint8_t* d_blobs, size_t n, uint8_t* d_commitments_out) {
    if (!ctx || !d_blobs || !d_commitments_out || n == 0) return KZGB200_BAD_ARGS;
    std::lock_guard<std::mutex> g(ctx->lock);
    CK(cudaSetDevice(ctx->device));
    return commit_or_prove(ctx, d_blobs, nullptr, n, d_commitments_out, 0);
}
extern "C" int kzgb200_compute_blob_kzg_proof_batch(kzgb200_ctx* ctx, const uint8_t* d_blobs, const uint8_t* d_commitments, size_t n,
                                                    uint8_t* d_proofs_out) {
    if (!ctx || !d_blobs || !d_commitments || !d_proofs_out || n == 0) return KZGB200_BAD_ARGS;
    std::lock_guard<std::mutex> g(ctx->lock);
    CK(cudaSetDevice(ctx->device));
    return commit_or_prove(ctx, d_blobs, d_commitments, n, d_proofs_out, 1);
}
// clock64() stamps of the last single-GPU batch_final_kernel: start, prelude end, Miller loop end, easy part end, hard part end, done
extern "C" int kzgb200_debug_final_ticks(kzgb200_ctx* ctx, long long* out14) {
    if (!ctx || !out14) return KZGB200_BAD_ARGS;
    std::lock_guard<std::mutex> g(ctx->lock);
    CK(cudaSetDevice(ctx->device));
    CK(cudaMemcpy(out14, ctx->d_scratch + 128, 112, cudaMemcpyDeviceToHost));
    return KZGB200_OK;
}
// the stream every call of this context is issued on (cudaStream_t), for event timing by the caller
extern "C" void* kzgb200_stream(kzgb200_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" void* kzgb200_alloc_pinned(size_t bytes) {
    void* p = nullptr;
    return cudaMallocHost(&p, bytes) == cudaSuccess ? p : nullptr;
}
extern "C" void kzgb200_free_pinned(void* p) { if (p) cudaFreeHost(p); }
