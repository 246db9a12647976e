// Incremental SHA-256 on the host, used for ONE thing: the batch transcript of compute_r_powers (reference
// src/kzg_proof.rs:291-348).  That hash is a single serial chain of 2.5 compressions per blob over data produced by every
// blob of every GPU; a GPU runs a dependent SHA-256 round in ~30 clocks (40 ms at 16384 blobs, measured in round 1), a host
// core with SHA-NI runs the same chain at ~2 GB/s (1.3 ms) and can do it behind the kernels, chunk by chunk, as the (z, y)
// pairs arrive.  The per-blob challenge hashes (128 KiB each, independent) stay on the GPU.
#pragma once
#include <stddef.h>
#include <stdint.h>

namespace kzgb200 {

struct HostSha256 {
    uint32_t h[8];
    uint8_t buf[64];
    uint32_t fill;
    uint64_t total;
};
void host_sha256_init(HostSha256* s);
void host_sha256_update(HostSha256* s, const uint8_t* data, size_t len);
void host_sha256_final(HostSha256* s, uint8_t out[32]);
// 1 if the SHA-NI code path is in use (x86 SHA extensions present and not disabled)
int host_sha256_uses_shani();
// test hook: 1 forces the portable compression function
void host_sha256_force_portable(int on);

}  // namespace kzgb200
