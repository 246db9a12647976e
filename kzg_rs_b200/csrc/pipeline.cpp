// Streaming front-end over the blocking C ABI: several batches in flight on ONE GPU.
//
// A batch has a blob-streaming head (SHA-256 challenges, barycentric evaluation, G1 parsing: every SM busy) and a
// latency-bound tail (transcript hash, MSM reduction, one pairing: a handful of CTAs).  A pipeline owns `depth`
// independent device contexts (own streams, own workspace) with one host worker thread each, so the tail of batch i runs
// under the head -- and, for host buffers, under the PCIe copy -- of batch i+1.  Verdicts, error codes and z / y are
// exactly those of kzgb200_verify_blob_kzg_proof_batch[_device]: each ticket is one such call on one of the contexts.
#include <condition_variable>
#include <cstdint>
#include <deque>
#include <map>
#include <mutex>
#include <new>
#include <set>
#include <thread>
#include <vector>
#include "../../include/kzgb200.h"

namespace {
struct Job {
    uint64_t ticket;
    const uint8_t *blobs, *commitments, *proofs;
    size_t n_blobs, n_commitments, n_proofs;
    uint8_t *z_out, *y_out;
    bool device;
};
struct Done { int rc, ok; };
}  // namespace

struct kzgb200_pipeline {
    std::vector<kzgb200_ctx*> ctx;
    std::vector<std::thread> workers;
    std::mutex m;
    std::condition_variable cv_job, cv_done;
    std::deque<Job> queue;              // submitted, not yet picked up
    std::map<uint64_t, Done> done;      // finished, not yet waited for
    std::set<uint64_t> live;            // submitted, not yet waited for
    uint64_t next_ticket = 1;
    size_t in_flight = 0;               // queued + running
    size_t waiters = 0;                 // threads inside wait() / blocked in submit()
    bool stop = false;

    void work(size_t slot) {
        for (;;) {
            Job j;
            {
                std::unique_lock<std::mutex> lk(m);
                cv_job.wait(lk, [&] { return stop || !queue.empty(); });
                if (queue.empty()) return;
                j = queue.front();
                queue.pop_front();
            }
            Done d{KZGB200_INTERNAL_ERROR, 0};
            if (j.device)
                d.rc = kzgb200_verify_blob_kzg_proof_batch_device(ctx[slot], j.blobs, j.commitments, j.proofs, j.n_blobs, &d.ok, j.z_out, j.y_out);
            else
                d.rc = kzgb200_verify_blob_kzg_proof_batch(ctx[slot], j.blobs, j.n_blobs, j.commitments, j.n_commitments, j.proofs, j.n_proofs,
                                                           &d.ok, j.z_out, j.y_out);
            {
                std::lock_guard<std::mutex> lk(m);
                done[j.ticket] = d;
                in_flight--;
            }
            cv_done.notify_all();
        }
    }
};

extern "C" int kzgb200_pipeline_create(kzgb200_pipeline** out, int device, const uint8_t* g2_points, size_t g2_points_len, int depth) {
    if (!out || depth < 1 || depth > 8) return KZGB200_BAD_ARGS;
    *out = nullptr;
    kzgb200_pipeline* p = new (std::nothrow) kzgb200_pipeline();
    if (!p) return KZGB200_INTERNAL_ERROR;
    for (int i = 0; i < depth; i++) {
        kzgb200_ctx* c = nullptr;
        int rc = kzgb200_create(&c, device, g2_points, g2_points_len);
        if (rc == KZGB200_OK && depth > 1) kzgb200_set_slab_tail(c, 0);     // the next call's copies already hide this call's tail
        if (rc != KZGB200_OK) {
            for (kzgb200_ctx* x : p->ctx) kzgb200_destroy(x);
            delete p;
            return rc;
        }
        p->ctx.push_back(c);
    }
    for (int i = 0; i < depth; i++) p->workers.emplace_back([p, i] { p->work((size_t)i); });
    *out = p;
    return KZGB200_OK;
}

extern "C" void kzgb200_pipeline_destroy(kzgb200_pipeline* p) {
    if (!p) return;
    {
        std::lock_guard<std::mutex> lk(p->m);
        p->stop = true;
    }
    p->cv_job.notify_all();
    p->cv_done.notify_all();                         // submitters blocked on back-pressure give up
    for (std::thread& t : p->workers) t.join();     // workers drain the queue first
    {
        std::unique_lock<std::mutex> lk(p->m);       // threads still inside wait() / submit() leave before the memory goes
        p->cv_done.notify_all();
        p->cv_done.wait(lk, [&] { return p->waiters == 0; });
    }
    for (kzgb200_ctx* c : p->ctx) kzgb200_destroy(c);
    delete p;
}

extern "C" int kzgb200_pipeline_depth(const kzgb200_pipeline* p) { return p ? (int)p->ctx.size() : 0; }

extern "C" kzgb200_ctx* kzgb200_pipeline_context(kzgb200_pipeline* p, int slot) {
    return (p && slot >= 0 && (size_t)slot < p->ctx.size()) ? p->ctx[(size_t)slot] : nullptr;
}

static int submit(kzgb200_pipeline* p, Job j, uint64_t* ticket) {
    if (!p || !ticket) return KZGB200_BAD_ARGS;
    {
        std::unique_lock<std::mutex> lk(p->m);
        if (p->stop) return KZGB200_BAD_ARGS;
        // back-pressure: at most `depth` batches queued or running (each context holds one batch of workspace)
        p->waiters++;
        p->cv_done.wait(lk, [&] { return p->stop || p->in_flight < p->ctx.size(); });
        p->waiters--;
        if (p->stop) { p->cv_done.notify_all(); return KZGB200_BAD_ARGS; }      // the pipeline is being destroyed
        j.ticket = *ticket = p->next_ticket++;
        p->queue.push_back(j);
        p->live.insert(j.ticket);
        p->in_flight++;
    }
    p->cv_job.notify_one();
    return KZGB200_OK;
}

extern "C" int kzgb200_pipeline_submit(kzgb200_pipeline* p, const uint8_t* blobs, size_t n_blobs, const uint8_t* commitments,
                                       size_t n_commitments, const uint8_t* proofs, size_t n_proofs, uint8_t* z_out, uint8_t* y_out,
                                       uint64_t* ticket) {
    return submit(p, Job{0, blobs, commitments, proofs, n_blobs, n_commitments, n_proofs, z_out, y_out, false}, ticket);
}

extern "C" int kzgb200_pipeline_submit_device(kzgb200_pipeline* p, const uint8_t* d_blobs, const uint8_t* d_commitments,
                                              const uint8_t* d_proofs, size_t n, uint8_t* d_z_out, uint8_t* d_y_out, uint64_t* ticket) {
    return submit(p, Job{0, d_blobs, d_commitments, d_proofs, n, n, n, d_z_out, d_y_out, true}, ticket);
}

extern "C" int kzgb200_pipeline_wait(kzgb200_pipeline* p, uint64_t ticket, int* ok) {
    if (!p || !ok) return KZGB200_BAD_ARGS;
    std::unique_lock<std::mutex> lk(p->m);
    if (!p->live.erase(ticket)) return KZGB200_BAD_ARGS;               // unknown, already waited for, or another thread is waiting for it
    p->waiters++;
    p->cv_done.wait(lk, [&] { return p->done.count(ticket) != 0; });   // every submitted job completes (destroy drains the queue)
    Done d = p->done[ticket];
    p->done.erase(ticket);
    p->waiters--;
    if (p->stop) p->cv_done.notify_all();                              // destroy waits for the last waiter to leave
    *ok = d.ok;
    return d.rc;
}
