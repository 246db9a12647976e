// Blob-sharded batches over several GPUs behind the C ABI (include/kzgb200.h "multi-GPU").
//
// The batch path shards by blob: rank k owns a contiguous range and never sees other ranks' blobs.  All blobs meet in two
// places only (reference src/kzg_proof.rs:291-348 and :419-441):
//   1. the transcript hash that yields r -- one serial SHA-256 over 160 bytes per blob, hashed by the LEADER's host thread
//      behind the evaluation kernels of all ranks: every rank publishes its (C, z, y, pi) entries chunk by chunk into a shared
//      host block as they leave its GPU, the leader hashes them in global order and publishes the 32-byte digest;
//   2. the partial sums A_k, B_k, s_k (352 bytes per rank): the LAST STORE of each rank's msm_combine_kernel goes over NVLink
//      straight into the leader GPU's exchange buffer (peer-mapped: cudaDeviceEnablePeerAccess in one process, CUDA IPC across
//      processes), followed by a flag; the leader's stream waits on the flags in a tiny kernel and runs the single pairing
//      check.  No host synchronisation and no collective library in between.  (Fallback when the peer mapping is refused:
//      the partials travel through the shared host block.)
// Two ways to form a group: kzgb200_group_create (one process drives n GPUs) and kzgb200_group_join (one process per GPU, e.g.
// under torchrun; the ranks meet in a POSIX shared-memory segment named by the session string).  Both run the same protocol.
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <atomic>
#include <chrono>
#include <functional>
#include <new>
#include <string>
#include "runtime.cuh"

using namespace kzgb200;

struct kzgb200_group;

namespace {

constexpr int kMaxRanks = 16;
constexpr uint32_t kMagic = 0x4b5a4732;   // "KZG2"
constexpr double kWaitSeconds = 30.0;

// Shared HOST block (heap when one process drives all GPUs, POSIX shm across processes).  Epoch-stamped fields: a collective
// call number, incremented in lockstep by every rank, so no field needs resetting between calls.
struct XShared {
    std::atomic<uint32_t> magic;
    uint32_t world;
    uint64_t cap;                                   // blobs per rank the entry arrays are sized for
    std::atomic<uint32_t> attached;
    std::atomic<uint32_t> abort_epoch;              // a rank that fails mid-protocol stamps the epoch here; waiters give up
    std::atomic<uint64_t> n_local[kMaxRanks];       // epoch << 32 | blobs of this rank's shard
    std::atomic<uint64_t> ready[kMaxRanks];         // epoch << 32 | blobs whose transcript payload is in the block
    std::atomic<uint32_t> r_epoch;
    uint8_t r_digest[32];
    std::atomic<uint32_t> partial_epoch[kMaxRanks];
    uint8_t partials[kMaxRanks][KZGB200_PARTIAL_BYTES];
    std::atomic<uint32_t> late_epoch[kMaxRanks];    // this rank's subgroup checks / canonicity flags, known after its tail
    uint32_t late_err[kMaxRanks];
    std::atomic<uint32_t> verdict_epoch;
    uint32_t verdict_rc, verdict_ok;
    std::atomic<uint32_t> path[kMaxRanks];          // how rank k delivers its partial: 1 = peer store into the leader GPU, 2 = through this block
    std::atomic<uint32_t> ipc_ready;                // 1: handle valid, 2: the leader could not export one
    unsigned char ipc_handle[64];
    // then: entries[world][cap][160] -- each rank's transcript entries already in the byte order compute_r_powers hashes
    // (C_i | z_i LE | y_i LE | pi_i), so that the leader hashes them in place (tree mode: 32-byte leaf digests at the start
    // of the rank's region)
};
static_assert(sizeof(cudaIpcMemHandle_t) <= 64, "IPC handle size");
size_t shared_bytes(int world, size_t cap) { return ((sizeof(XShared) + 63) & ~(size_t)63) + (size_t)world * cap * 160; }

// Exchange buffer in the LEADER GPU's memory, written by the other GPUs' combine kernels
struct XDev {
    Partial partials[kMaxRanks];
    uint32_t flags[kMaxRanks];
    uint32_t timed_out;
};

struct Member {
    kzgb200_ctx* ctx = nullptr;
    int rank = 0;
    XDev* leader_x = nullptr;       // peer-mapped (or local, for the leader) exchange buffer; nullptr: host path
    void* ipc_mapping = nullptr;
    // current call
    const uint8_t *hc = nullptr, *hp = nullptr;
    size_t n = 0, offset = 0;
    kzgb200_group* g = nullptr;
};

bool wait_until(const std::function<bool()>& pred, double seconds = kWaitSeconds) {
    auto t0 = std::chrono::steady_clock::now();
    for (unsigned spins = 0;; spins++) {
        if (pred()) return true;
        if ((spins & 1023) == 1023) {
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > seconds) return false;
            sched_yield();
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
}

}  // namespace

struct kzgb200_group {
    int world = 0;
    std::vector<Member> local;          // members driven by this process (all of them, or one)
    XShared* sh = nullptr;
    size_t sh_bytes = 0;
    bool shm = false, owns_ctx = true;
    std::string shm_name;
    uint32_t epoch = 0;
    XDev* d_x = nullptr;                // leader only
    uint32_t* h_flag = nullptr;         // pinned epoch value for host-path flag writes
    HostSha256 sha;
    size_t hashed[kMaxRanks] = {0};
    std::mutex lock;
    char err[256] = {0};
    uint8_t* entries(int rank) const {          // rank's region of the shared block: cap x 160 bytes
        return reinterpret_cast<uint8_t*>(sh) + ((sizeof(XShared) + 63) & ~(size_t)63) + (size_t)rank * sh->cap * 160;
    }
    bool leader_here() const { return !local.empty() && local[0].rank == 0; }
};

namespace {

#define GFAIL(msg)                                                               \
    do {                                                                         \
        snprintf(g->err, sizeof(g->err), "%s (%s:%d)", msg, __FILE__, __LINE__); \
        g->sh->abort_epoch.store(g->epoch);                                      \
        return KZGB200_INTERNAL_ERROR;                                           \
    } while (0)

// a chunk of this member's transcript payload has reached its pinned host buffers: move it into the shared block and publish
void publish_chunk(void* arg, int c) {
    Member* m = static_cast<Member*>(arg);
    kzgb200_group* g = m->g;
    kzgb200_ctx* ctx = m->ctx;
    size_t lo = ctx->chunks[c].lo, cnt = ctx->chunks[c].cnt, end = lo + cnt;
    if (ctx->transcript_mode == KZGB200_TRANSCRIPT_TREE) {
        size_t ngroups = (m->n + kTreeGroup - 1) / kTreeGroup;
        size_t g0 = lo / kTreeGroup, g1 = end >= m->n ? ngroups : end / kTreeGroup;
        memcpy(g->entries(m->rank) + g0 * 32, ctx->h_zy + g0 * 32, (g1 - g0) * 32);     // leaf digests
    } else {
        uint8_t* e = g->entries(m->rank) + lo * 160;       // interleaved by the publishing rank: the leader's hash stays one pass
        for (size_t i = lo; i < end; i++, e += 160) {
            memcpy(e, ctx->tr_c + 48 * i, 48);
            memcpy(e + 48, ctx->h_zy + 64 * i, 64);
            memcpy(e + 112, ctx->tr_p + 48 * i, 48);
        }
    }
    g->sh->ready[m->rank].store(((uint64_t)g->epoch << 32) | (uint64_t)end, std::memory_order_release);
}

// the leader hashes whatever has been published, in global order (rank by rank, entry by entry)
bool leader_hash_available(kzgb200_group* g, int aw, const size_t* n_of, bool tree) {
    bool all = true;
    for (int k = 0; k < aw; k++) {
        if (g->hashed[k] < n_of[k]) {
            uint64_t v = g->sh->ready[k].load(std::memory_order_acquire);
            size_t avail = (uint32_t)(v >> 32) == g->epoch ? (size_t)(uint32_t)v : 0;
            if (avail > g->hashed[k]) {
                if (tree) {
                    size_t ng = (n_of[k] + kTreeGroup - 1) / kTreeGroup;
                    size_t g0 = g->hashed[k] / kTreeGroup, g1 = avail >= n_of[k] ? ng : avail / kTreeGroup;
                    host_sha256_update(&g->sha, g->entries(k) + g0 * 32, (g1 - g0) * 32);
                } else {
                    host_sha256_update(&g->sha, g->entries(k) + g->hashed[k] * 160, (avail - g->hashed[k]) * 160);
                }
                g->hashed[k] = avail;
            }
            if (g->hashed[k] < n_of[k]) { all = false; break; }     // global order: later ranks wait for this one
        }
    }
    return all;
}

struct ShardArgs {
    const uint8_t *blobs, *commitments, *proofs;   // this member's shard
    size_t n;
    bool device;
    uint8_t *z_out, *y_out;                        // same memory space as the inputs; nullable
};

// One collective batch.  `args[i]` belongs to g->local[i]; aw = number of participating ranks (the first aw).
int group_run(kzgb200_group* g, const ShardArgs* args, int aw, int* ok) {
    XShared* sh = g->sh;
    const uint32_t epoch = ++g->epoch;
    const int nloc = (int)g->local.size();
    auto aborted = [&] { return sh->abort_epoch.load() == epoch; };
    // ---- phase 1 on every local member (asynchronous) -------------------------------------------------------------
    for (int i = 0; i < nloc; i++) {
        Member& m = g->local[i];
        if (m.rank >= aw) continue;
        kzgb200_ctx* ctx = m.ctx;
        const ShardArgs& a = args[i];
        if (a.n == 0 || a.n > sh->cap) GFAIL("shard size out of range (every participating rank needs 1..max_blobs_per_rank blobs)");
        CK(cudaSetDevice(ctx->device));
        int rc = ensure_capacity(ctx, a.n, !a.device);
        if (rc) { snprintf(g->err, sizeof(g->err), "%s", ctx->err); sh->abort_epoch.store(epoch); return rc; }
        m.n = a.n; m.g = g;
        sh->n_local[m.rank].store(((uint64_t)epoch << 32) | (uint64_t)a.n, std::memory_order_release);
        ctx->chunk_sink = publish_chunk; ctx->chunk_sink_arg = &m;
        const uint8_t *d_b = a.blobs, *d_c = a.commitments, *d_p = a.proofs;
        if (!a.device) {
            CK(cudaMemcpyAsync(ctx->d_c, a.commitments, a.n * 48, cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(ctx->d_p, a.proofs, a.n * 48, cudaMemcpyHostToDevice, ctx->stream));
            d_b = ctx->d_blobs; d_c = ctx->d_c; d_p = ctx->d_p;
        }
        if (ctx->transcript_mode == KZGB200_TRANSCRIPT_EXACT_DEVICE) ctx->transcript_mode = KZGB200_TRANSCRIPT_EXACT;
        rc = launch_phase1(ctx, d_b, a.device ? nullptr : a.blobs, d_c, d_p, a.n, true, ctx->defer_subgroup != 0, a.device ? nullptr : a.commitments,
                           a.device ? nullptr : a.proofs);
        if (!rc) rc = export_zy(ctx, a.n, a.z_out ? (a.device ? a.z_out : ctx->d_zout) : nullptr, a.y_out ? (a.device ? a.y_out : ctx->d_yout) : nullptr);
        if (rc) { snprintf(g->err, sizeof(g->err), "%s", ctx->err); sh->abort_epoch.store(epoch); ctx->chunk_sink = nullptr; return rc; }
    }
    // ---- shard sizes of all ranks -> offsets ------------------------------------------------------------------------
    size_t n_of[kMaxRanks], total = 0;
    for (int k = 0; k < aw; k++) {
        if (!wait_until([&] { return (uint32_t)(sh->n_local[k].load(std::memory_order_acquire) >> 32) == epoch || aborted(); }) || aborted())
            GFAIL("a rank did not enter the collective call");
        n_of[k] = (size_t)(uint32_t)sh->n_local[k].load();
        total += n_of[k];
    }
    const bool tree = g->local[0].ctx->transcript_mode == KZGB200_TRANSCRIPT_TREE;
    for (Member& m : g->local) {
        m.offset = 0;
        for (int k = 0; k < m.rank && k < aw; k++) m.offset += n_of[k];
        if (tree && m.rank < aw && m.offset % kTreeGroup) GFAIL("tree transcript: every shard but the last must hold a multiple of 16 blobs");
    }
    // ---- transcript: members publish, the leader hashes ----------------------------------------------------------------
    const bool lead = g->leader_here();
    if (lead) {
        hash_transcript_header(&g->sha, total);
        for (int k = 0; k < aw; k++) g->hashed[k] = 0;
    }
    {
        bool ok_wait = wait_until([&] {
            bool mine_done = true;
            for (Member& m : g->local) {
                if (m.rank >= aw) continue;
                cudaSetDevice(m.ctx->device);
                transcript_progress(m.ctx, false);
                if (m.ctx->tr_next_chunk < m.ctx->nchunks) mine_done = false;
            }
            bool hash_done = lead ? leader_hash_available(g, aw, n_of, tree) : true;
            return (mine_done && hash_done) || aborted();
        });
        if (!ok_wait || aborted()) GFAIL("timed out waiting for the transcript entries of all ranks");
    }
    if (lead) {
        host_sha256_final(&g->sha, sh->r_digest);
        sh->r_epoch.store(epoch, std::memory_order_release);
    }
    if (!wait_until([&] { return sh->r_epoch.load(std::memory_order_acquire) == epoch || aborted(); }) || aborted()) GFAIL("timed out waiting for r");
    // ---- phase 2: partial sums; the gather is the combine kernel's last store ----------------------------------------
    for (Member& m : g->local) {
        if (m.rank >= aw) continue;
        kzgb200_ctx* ctx = m.ctx;
        CK(cudaSetDevice(ctx->device));
        int rc = upload_r_digest(ctx, sh->r_digest);
        phase_end(ctx, kPhTranscript, ctx->stream);
        ctx->tr_active = false; ctx->chunk_sink = nullptr;
        if (rc) { snprintf(g->err, sizeof(g->err), "%s", ctx->err); sh->abort_epoch.store(epoch); return rc; }
        Partial* out = m.leader_x ? &m.leader_x->partials[m.rank] : ctx->d_partial;
        uint32_t* flag = m.leader_x ? &m.leader_x->flags[m.rank] : nullptr;
        rc = launch_lincomb(ctx, m.offset, out, !ctx->subgroup_pending, flag, epoch);
        if (rc) { snprintf(g->err, sizeof(g->err), "%s", ctx->err); sh->abort_epoch.store(epoch); return rc; }
        if (!m.leader_x) CK(cudaMemcpyAsync(ctx->h_partial, ctx->d_partial, sizeof(Partial), cudaMemcpyDeviceToHost, ctx->stream));
        if (m.rank != 0) {      // late flags of a non-leader: after its deferred subgroup checks
            CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_parse, 0));
            status_or_kernel<<<1, 256, 0, ctx->stream>>>(ctx->d_status, (int)m.n, ctx->d_result + 3);
            CK(cudaMemcpyAsync(ctx->h_result + 3, ctx->d_result + 3, 4, cudaMemcpyDeviceToHost, ctx->stream));
        }
    }
    // host-path partials (no peer mapping): through the shared block
    for (Member& m : g->local) {
        if (m.rank >= aw || m.leader_x) continue;
        kzgb200_ctx* ctx = m.ctx;
        CK(cudaSetDevice(ctx->device));
        CK(cudaStreamSynchronize(ctx->stream));
        memcpy(sh->partials[m.rank], ctx->h_partial, sizeof(Partial));
        sh->partial_epoch[m.rank].store(epoch, std::memory_order_release);
    }
    int rc_final = KZGB200_OK, verdict = 0;
    if (lead) {
        Member& L = g->local[0];
        kzgb200_ctx* ctx = L.ctx;
        CK(cudaSetDevice(ctx->device));
        *g->h_flag = epoch;
        // ranks without a peer mapping: their partials arrive in the shared block; the leader forwards them to its exchange buffer
        for (int k = 1; k < aw; k++) {
            if (sh->path[k].load() != 2) continue;
            if (!wait_until([&] { return sh->partial_epoch[k].load(std::memory_order_acquire) == epoch || aborted(); }) || aborted()) GFAIL("timed out waiting for a rank's partial");
            CK(cudaMemcpyAsync(&g->d_x->partials[k], sh->partials[k], sizeof(Partial), cudaMemcpyHostToDevice, ctx->stream));
            CK(cudaMemcpyAsync(&g->d_x->flags[k], g->h_flag, 4, cudaMemcpyHostToDevice, ctx->stream));
        }
        wait_flags_kernel<<<1, 32, 0, ctx->stream>>>(g->d_x->flags, aw, epoch, &g->d_x->timed_out);
        phase_begin(ctx, kPhFinal, ctx->stream);
        launch_batch_final(ctx->stream, g->d_x->partials, aw, ctx->tables, ctx->d_result, ctx->d_scratch, false);
        phase_end(ctx, kPhFinal, ctx->stream);
        CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_parse, 0));
        status_or_kernel<<<1, 256, 0, ctx->stream>>>(ctx->d_status, (int)L.n, ctx->d_result + 2);
        CK(cudaMemcpyAsync(ctx->h_result + 3, &g->d_x->timed_out, 4, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaGetLastError());
        rc_final = read_result(ctx, &verdict);
        if (ctx->h_result[3]) GFAIL("timed out on the device waiting for the partial sums of all ranks");
    }
    // late flags of the non-leaders
    for (Member& m : g->local) {
        if (m.rank == 0 || m.rank >= aw) continue;
        kzgb200_ctx* ctx = m.ctx;
        CK(cudaSetDevice(ctx->device));
        CK(cudaStreamSynchronize(ctx->stream));
        sh->late_err[m.rank] = ctx->h_result[3];
        sh->late_epoch[m.rank].store(epoch, std::memory_order_release);
    }
    if (lead) {
        for (int k = 1; k < aw; k++) {
            if (!wait_until([&] { return sh->late_epoch[k].load(std::memory_order_acquire) == epoch || aborted(); }) || aborted()) GFAIL("timed out waiting for a rank's flags");
            if (sh->late_err[k] && rc_final == KZGB200_OK) rc_final = KZGB200_BAD_ARGS;    // Err on any rank is Err for the batch (:503-516)
        }
        sh->verdict_rc = (uint32_t)rc_final; sh->verdict_ok = (uint32_t)verdict;
        sh->verdict_epoch.store(epoch, std::memory_order_release);
    }
    if (!wait_until([&] { return sh->verdict_epoch.load(std::memory_order_acquire) == epoch || aborted(); }) || aborted()) GFAIL("timed out waiting for the verdict");
    *ok = (int)sh->verdict_ok;
    int rc = (int)sh->verdict_rc;
    // z / y of host-input shards back to the caller
    for (int i = 0; i < nloc; i++) {
        Member& m = g->local[i];
        if (m.rank >= aw || args[i].device) continue;
        kzgb200_ctx* ctx = m.ctx;
        CK(cudaSetDevice(ctx->device));
        if (args[i].z_out) CK(cudaMemcpyAsync(args[i].z_out, ctx->d_zout, m.n * 32, cudaMemcpyDeviceToHost, ctx->stream));
        if (args[i].y_out) CK(cudaMemcpyAsync(args[i].y_out, ctx->d_yout, m.n * 32, cudaMemcpyDeviceToHost, ctx->stream));
        CK(cudaStreamSynchronize(ctx->stream));
    }
    for (Member& m : g->local) if (m.rank < aw) collect_phase_times(m.ctx);
    return rc;
}

int alloc_leader_buffers(kzgb200_group* g) {
    kzgb200_ctx* ctx = g->local[0].ctx;
    CK(cudaSetDevice(ctx->device));
    CK(cudaMalloc(&g->d_x, sizeof(XDev)));
    CK(cudaMemset(g->d_x, 0, sizeof(XDev)));
    CK(cudaMallocHost(&g->h_flag, 4));
    g->local[0].leader_x = g->d_x;
    return KZGB200_OK;
}

}  // namespace

extern "C" int kzgb200_group_create(kzgb200_group** out, const int* device_ids, int n_devices, const uint8_t* g2_points, size_t g2_points_len,
                                    size_t max_blobs_per_device) {
    DeviceGuard dev;
    if (!out) return KZGB200_BAD_ARGS;
    *out = nullptr;
    if (!device_ids || n_devices < 1 || n_devices > kMaxRanks || max_blobs_per_device == 0) return KZGB200_BAD_ARGS;
    kzgb200_group* g = new (std::nothrow) kzgb200_group();
    if (!g) return KZGB200_INTERNAL_ERROR;
    g->world = n_devices;
    g->local.resize(n_devices);
    for (int k = 0; k < n_devices; k++) {
        g->local[k].rank = k; g->local[k].g = g;
        int rc = kzgb200_create(&g->local[k].ctx, device_ids[k], g2_points, g2_points_len);
        if (rc) { kzgb200_group_destroy(g); return rc; }
    }
    g->sh_bytes = shared_bytes(n_devices, max_blobs_per_device);
    void* p = nullptr;
    if (cudaMallocHost(&p, g->sh_bytes) != cudaSuccess) { kzgb200_group_destroy(g); return KZGB200_INTERNAL_ERROR; }
    memset(p, 0, g->sh_bytes);
    g->sh = new (p) XShared();
    g->sh->world = (uint32_t)n_devices; g->sh->cap = max_blobs_per_device; g->sh->magic.store(kMagic);
    kzgb200_ctx* ctx = g->local[0].ctx;
    int rc = alloc_leader_buffers(g);
    if (rc) { kzgb200_group_destroy(g); return rc; }
    // peer mappings: the other GPUs store their partials straight into the leader's exchange buffer
    for (int k = 1; k < n_devices; k++) {
        int dk = device_ids[k], d0 = device_ids[0], can = dk == d0;
        if (!can) {
            cudaDeviceCanAccessPeer(&can, dk, d0);
            if (can) {
                cudaSetDevice(dk);
                cudaError_t e = cudaDeviceEnablePeerAccess(d0, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) can = 0;
                cudaGetLastError();
            }
        }
        if (const char* v = getenv("KZGB200_GROUP_NO_P2P")) if (atoi(v)) can = 0;
        g->local[k].leader_x = can ? g->d_x : nullptr;
        g->sh->path[k].store(can ? 1 : 2);
    }
    g->sh->path[0].store(1);
    (void)ctx;
    *out = g;
    return KZGB200_OK;
}

extern "C" int kzgb200_group_join(kzgb200_group** out, const char* session, int rank, int world, int device, const uint8_t* g2_points,
                                  size_t g2_points_len, size_t max_blobs_per_rank) {
    DeviceGuard dev;
    if (!out) return KZGB200_BAD_ARGS;
    *out = nullptr;
    if (!session || !*session || world < 1 || world > kMaxRanks || rank < 0 || rank >= world || max_blobs_per_rank == 0) return KZGB200_BAD_ARGS;
    kzgb200_group* g = new (std::nothrow) kzgb200_group();
    if (!g) return KZGB200_INTERNAL_ERROR;
    g->world = world; g->shm = true;
    g->shm_name = std::string("/kzgb200_") + session;
    g->local.resize(1);
    g->local[0].rank = rank; g->local[0].g = g;
    int rc = kzgb200_create(&g->local[0].ctx, device, g2_points, g2_points_len);
    if (rc) { kzgb200_group_destroy(g); return rc; }
    g->sh_bytes = shared_bytes(world, max_blobs_per_rank);
    int fd = -1;
    if (rank == 0) {
        shm_unlink(g->shm_name.c_str());      // a stale segment of a crashed run
        fd = shm_open(g->shm_name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)g->sh_bytes) != 0) { if (fd >= 0) close(fd); kzgb200_group_destroy(g); return KZGB200_INTERNAL_ERROR; }
    } else {
        bool got = wait_until([&] {
            fd = shm_open(g->shm_name.c_str(), O_RDWR, 0600);
            if (fd < 0) { usleep(1000); return false; }
            struct stat st;
            if (fstat(fd, &st) == 0 && (size_t)st.st_size >= g->sh_bytes) return true;
            close(fd); fd = -1; usleep(1000);
            return false;
        }, 120.0);
        if (!got) { kzgb200_group_destroy(g); return KZGB200_INTERNAL_ERROR; }
    }
    void* p = mmap(nullptr, g->sh_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) { kzgb200_group_destroy(g); return KZGB200_INTERNAL_ERROR; }
    g->sh = static_cast<XShared*>(p);
    kzgb200_ctx* ctx = g->local[0].ctx;
    if (rank == 0) {
        g->sh->world = (uint32_t)world; g->sh->cap = max_blobs_per_rank;
        rc = alloc_leader_buffers(g);
        if (rc) { kzgb200_group_destroy(g); return rc; }
        cudaIpcMemHandle_t h;
        bool exported = cudaIpcGetMemHandle(&h, g->d_x) == cudaSuccess;
        if (const char* v = getenv("KZGB200_GROUP_NO_P2P")) if (atoi(v)) exported = false;
        if (exported) memcpy(g->sh->ipc_handle, &h, sizeof(h)); else cudaGetLastError();
        g->sh->ipc_ready.store(exported ? 1 : 2);
        g->sh->magic.store(kMagic, std::memory_order_release);
    } else {
        if (!wait_until([&] { return g->sh->magic.load(std::memory_order_acquire) == kMagic; }, 120.0) || g->sh->world != (uint32_t)world ||
            g->sh->cap != max_blobs_per_rank) { kzgb200_group_destroy(g); return KZGB200_BAD_ARGS; }
        if (g->sh->ipc_ready.load() == 1) {
            cudaIpcMemHandle_t h;
            memcpy(&h, g->sh->ipc_handle, sizeof(h));
            cudaSetDevice(ctx->device);
            void* mapped = nullptr;
            if (cudaIpcOpenMemHandle(&mapped, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess) {
                g->local[0].ipc_mapping = mapped;
                g->local[0].leader_x = static_cast<XDev*>(mapped);
            } else {
                cudaGetLastError();
            }
        }
    }
    g->sh->path[rank].store(g->local[0].leader_x ? 1 : 2);
    // everybody attached -> the name can go (the mapping lives on); nothing is left behind if a rank dies later
    g->sh->attached.fetch_add(1);
    if (!wait_until([&] { return g->sh->attached.load() >= (uint32_t)world; }, 120.0)) { kzgb200_group_destroy(g); return KZGB200_INTERNAL_ERROR; }
    if (rank == 0) shm_unlink(g->shm_name.c_str());
    *out = g;
    return KZGB200_OK;
}

extern "C" void kzgb200_group_destroy(kzgb200_group* g) {
    DeviceGuard dev;
    if (!g) return;
    for (Member& m : g->local) {
        if (m.ctx) { cudaSetDevice(m.ctx->device); cudaDeviceSynchronize(); }
        if (m.ipc_mapping) cudaIpcCloseMemHandle(m.ipc_mapping);
    }
    if (g->d_x) { cudaSetDevice(g->local[0].ctx->device); cudaFree(g->d_x); }
    if (g->h_flag) cudaFreeHost(g->h_flag);
    for (Member& m : g->local) if (m.ctx && g->owns_ctx) kzgb200_destroy(m.ctx);
    if (g->sh) { if (g->shm) munmap(g->sh, g->sh_bytes); else cudaFreeHost(g->sh); }
    delete g;
}

extern "C" int kzgb200_group_size(const kzgb200_group* g) { return g ? g->world : 0; }
extern "C" int kzgb200_group_local_members(const kzgb200_group* g) { return g ? (int)g->local.size() : 0; }
extern "C" kzgb200_ctx* kzgb200_group_context(kzgb200_group* g, int local_index) {
    return (g && local_index >= 0 && (size_t)local_index < g->local.size()) ? g->local[(size_t)local_index].ctx : nullptr;
}
extern "C" const char* kzgb200_group_last_error(const kzgb200_group* g) { return g ? g->err : "null group"; }
// 1 if local member `local_index` stores its partial straight into the leader GPU's memory (peer mapping), 0 = host path
extern "C" int kzgb200_group_uses_peer_stores(const kzgb200_group* g, int local_index) {
    return (g && local_index >= 0 && (size_t)local_index < g->local.size() && g->local[(size_t)local_index].leader_x) ? 1 : 0;
}

extern "C" int kzgb200_group_verify_shards(kzgb200_group* g, const uint8_t* const* blobs, const uint8_t* const* commitments,
                                           const uint8_t* const* proofs, const size_t* n_local, int device_pointers, int* ok,
                                           uint8_t* const* z_out, uint8_t* const* y_out) {
    DeviceGuard dev;
    if (!g || !blobs || !commitments || !proofs || !n_local || !ok) return KZGB200_BAD_ARGS;
    std::lock_guard<std::mutex> lk(g->lock);
    ShardArgs a[kMaxRanks];
    int nloc = (int)g->local.size();
    for (int i = 0; i < nloc; i++) a[i] = {blobs[i], commitments[i], proofs[i], n_local[i], device_pointers != 0, z_out ? z_out[i] : nullptr, y_out ? y_out[i] : nullptr};
    return group_run(g, a, g->world, ok);
}

// the whole batch in host memory, one process driving all GPUs: contiguous ranges, rank k gets blobs [k * per, (k + 1) * per)
extern "C" int kzgb200_group_verify_blob_kzg_proof_batch(kzgb200_group* g, const uint8_t* blobs, size_t n_blobs, const uint8_t* commitments,
                                                         size_t n_commitments, const uint8_t* proofs, size_t n_proofs, int* ok, uint8_t* z_out,
                                                         uint8_t* y_out) {
    DeviceGuard dev;
    if (!g || !ok || g->shm) return KZGB200_BAD_ARGS;
    if (n_blobs == 0) { *ok = 1; return KZGB200_OK; }                 // reference src/kzg_proof.rs:478-480
    if (n_blobs < 32 || g->world == 1)                                  // too small to shard (and the n = 1 dispatch of :482-489)
        return kzgb200_verify_blob_kzg_proof_batch(g->local[0].ctx, blobs, n_blobs, commitments, n_commitments, proofs, n_proofs, ok, z_out, y_out);
    if (n_blobs != n_commitments || n_blobs != n_proofs) return KZGB200_INVALID_LENGTH;   // :491-501
    std::lock_guard<std::mutex> lk(g->lock);
    int aw = g->world;
    if ((size_t)aw > n_blobs / 16) aw = (int)(n_blobs / 16);
    size_t per = ((n_blobs + aw - 1) / aw + kTreeGroup - 1) / kTreeGroup * kTreeGroup;
    while ((size_t)(aw - 1) * per >= n_blobs) aw--;
    ShardArgs a[kMaxRanks];
    for (int k = 0; k < aw; k++) {
        size_t lo = (size_t)k * per, cnt = n_blobs - lo < per ? n_blobs - lo : per;
        a[k] = {blobs + lo * kBytesPerBlob, commitments + lo * 48, proofs + lo * 48, cnt, false, z_out ? z_out + lo * 32 : nullptr, y_out ? y_out + lo * 32 : nullptr};
    }
    for (int k = aw; k < g->world; k++) a[k] = {nullptr, nullptr, nullptr, 0, false, nullptr, nullptr};
    return group_run(g, a, aw, ok);
}

// the gathered partials of the last collective call (leader only): world x KZGB200_PARTIAL_BYTES, for parity tests
extern "C" int kzgb200_group_last_partials(kzgb200_group* g, uint8_t* out, size_t n_ranks) {
    DeviceGuard dev;
    if (!g || !out || !g->d_x || n_ranks > (size_t)kMaxRanks) return KZGB200_BAD_ARGS;
    std::lock_guard<std::mutex> lk(g->lock);
    cudaSetDevice(g->local[0].ctx->device);
    return cudaMemcpy(out, g->d_x->partials, n_ranks * sizeof(Partial), cudaMemcpyDeviceToHost) == cudaSuccess ? KZGB200_OK : KZGB200_INTERNAL_ERROR;
}

// Host-side protocol self-test (no GPU): every rank passes the transcript entries of its shard (as the kernels would have produced
// them); the entries go through the shared block, the leader hashes them in global order and every rank returns the same digest.
// Used by the world-size-2 CPU tests of the multi-GPU path.
extern "C" int kzgb200_group_host_protocol_test(const char* session, int rank, int world, const uint8_t* commitments, const uint8_t* zy,
                                                const uint8_t* proofs, size_t n_local, size_t chunk, uint8_t* digest_out32) {
    if (!session || world < 1 || world > kMaxRanks || rank < 0 || rank >= world || !digest_out32 || !chunk) return KZGB200_BAD_ARGS;
    kzgb200_group G;
    kzgb200_group* g = &G;
    g->world = world; g->shm = true;
    std::string name = std::string("/kzgb200_") + session;
    size_t cap = 1 << 16;
    g->sh_bytes = shared_bytes(world, cap);
    int fd = -1;
    if (rank == 0) {
        shm_unlink(name.c_str());
        fd = shm_open(name.c_str(), O_CREAT | O_EXCL | O_RDWR, 0600);
        if (fd < 0 || ftruncate(fd, (off_t)g->sh_bytes) != 0) return KZGB200_INTERNAL_ERROR;
    } else if (!wait_until([&] {
                   fd = shm_open(name.c_str(), O_RDWR, 0600);
                   if (fd < 0) { usleep(1000); return false; }
                   struct stat st;
                   if (fstat(fd, &st) == 0 && (size_t)st.st_size >= g->sh_bytes) return true;
                   close(fd); fd = -1; usleep(1000);
                   return false;
               }, 60.0)) return KZGB200_INTERNAL_ERROR;
    void* p = mmap(nullptr, g->sh_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    close(fd);
    if (p == MAP_FAILED) return KZGB200_INTERNAL_ERROR;
    g->sh = static_cast<XShared*>(p);
    XShared* sh = g->sh;
    if (rank == 0) { sh->world = (uint32_t)world; sh->cap = cap; sh->magic.store(kMagic, std::memory_order_release); }
    int rc = KZGB200_OK;
    if (!wait_until([&] { return sh->magic.load(std::memory_order_acquire) == kMagic; }, 60.0)) rc = KZGB200_INTERNAL_ERROR;
    if (!rc) {
        sh->attached.fetch_add(1);
        if (!wait_until([&] { return sh->attached.load() >= (uint32_t)world; }, 60.0)) rc = KZGB200_INTERNAL_ERROR;
    }
    if (rank == 0) shm_unlink(name.c_str());
    const uint32_t epoch = g->epoch = 1;
    if (!rc && n_local <= cap) {
        sh->n_local[rank].store(((uint64_t)epoch << 32) | n_local, std::memory_order_release);
        size_t n_of[kMaxRanks], total = 0;
        for (int k = 0; k < world && !rc; k++) {
            if (!wait_until([&] { return (uint32_t)(sh->n_local[k].load(std::memory_order_acquire) >> 32) == epoch; }, 60.0)) rc = KZGB200_INTERNAL_ERROR;
            n_of[k] = (size_t)(uint32_t)sh->n_local[k].load();
            total += n_of[k];
        }
        if (rank == 0) { hash_transcript_header(&g->sha, total); for (int k = 0; k < world; k++) g->hashed[k] = 0; }
        size_t pub = 0;
        if (!rc && !wait_until([&] {
                if (pub < n_local) {     // publish the next chunk
                    size_t cnt = n_local - pub < chunk ? n_local - pub : chunk;
                    uint8_t* e = g->entries(rank) + pub * 160;
                    for (size_t i = pub; i < pub + cnt; i++, e += 160) {
                        memcpy(e, commitments + 48 * i, 48);
                        memcpy(e + 48, zy + 64 * i, 64);
                        memcpy(e + 112, proofs + 48 * i, 48);
                    }
                    pub += cnt;
                    sh->ready[rank].store(((uint64_t)epoch << 32) | pub, std::memory_order_release);
                }
                bool hd = rank == 0 ? leader_hash_available(g, world, n_of, false) : true;
                return pub == n_local && hd;
            }, 60.0)) rc = KZGB200_INTERNAL_ERROR;
        if (!rc && rank == 0) { host_sha256_final(&g->sha, sh->r_digest); sh->r_epoch.store(epoch, std::memory_order_release); }
        if (!rc && !wait_until([&] { return sh->r_epoch.load(std::memory_order_acquire) == epoch; }, 60.0)) rc = KZGB200_INTERNAL_ERROR;
        if (!rc) memcpy(digest_out32, sh->r_digest, 32);
        // leave together (the leader's mapping must outlive the readers' last access: each rank has its own mapping, so this is only tidy)
        sh->late_epoch[rank].store(epoch);
    } else if (!rc) rc = KZGB200_BAD_ARGS;
    munmap(p, g->sh_bytes);
    g->sh = nullptr;
    return rc;
}
