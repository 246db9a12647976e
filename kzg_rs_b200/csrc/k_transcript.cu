// K5: batch challenge r (device-side transcript hashing).
#include "common.cuh"

namespace kzgb200 {

// ------------------------------------------------------------------------------------------------ K5
// Batch challenge r (reference src/kzg_proof.rs:291-348): SHA-256 over
//   "RCKZGBATCH___V1_" | u64be 4096 | u64be n | for each i: C_i[48] | z_i LE[32] | y_i LE[32] | pi_i[48]
// reduced mod q.  The hash is one serial chain over all blobs of the batch (every rank's), so it is a
// single-thread kernel; the message bytes are produced on the fly from the device-resident pieces.
// the same transcript as big-endian 32-bit words (header and entries are word aligned: 8 + 40 n words)
__device__ __forceinline__ uint32_t transcript_word(size_t w, uint64_t n, const uint32_t* C, const ZY* zy, const uint32_t* P) {
    if (w < 8) {
        const uint32_t hdr[8] = {0x52434b5a, 0x47424154, 0x43485f5f, 0x5f56315f, 0, 4096, (uint32_t)(n >> 32), (uint32_t)n};
        return hdr[w];
    }
    size_t q = (w - 8) / 40;
    uint32_t o = (uint32_t)((w - 8) % 40);
    if (o < 12) return sha_bswap(__ldg(C + q * 12 + o));
    if (o < 20) return sha_bswap(zy[q].z.l[o - 12]);
    if (o < 28) return sha_bswap(zy[q].y.l[o - 20]);
    return sha_bswap(__ldg(P + q * 12 + (o - 28)));
}
// K5a (parallel): one thread per 64-byte block of the transcript builds the block from the device-resident
// pieces, expands the SHA-256 message schedule and stores W[t] + K[t], t < 64 -- everything about a block that
// does not depend on the chaining value.
__global__ void __launch_bounds__(128) transcript_schedule_kernel(const uint8_t* __restrict__ commitments, const ZY* __restrict__ zy,
                                                                  const uint8_t* __restrict__ proofs, uint64_t n, uint32_t* __restrict__ wk,
                                                                  uint64_t first_blk, uint64_t blk_count) {
    size_t len = 32 + (size_t)n * 160;
    size_t nblk = (len + 9 + 63) / 64;
    size_t blk = first_blk + (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (blk >= nblk || blk >= first_blk + blk_count) return;
    uint32_t w[16];
    size_t nwords = len / 4;
    const uint32_t* Cw = reinterpret_cast<const uint32_t*>(commitments);
    const uint32_t* Pw = reinterpret_cast<const uint32_t*>(proofs);
    for (int j = 0; j < 16; j++) {
        size_t wi = blk * 16 + j;
        w[j] = wi < nwords ? transcript_word(wi, n, Cw, zy, Pw) : (wi == nwords ? 0x80000000u : 0u);
    }
    if (blk == nblk - 1) { w[14] = (uint32_t)(((uint64_t)len * 8) >> 32); w[15] = (uint32_t)((uint64_t)len * 8); }
    uint4* dst = reinterpret_cast<uint4*>(wk + blk * 64);
#pragma unroll
    for (int i = 0; i < 64; i += 4) {
        uint32_t o[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
            int t = i + u;
            if (t >= 16) {
                uint32_t w15 = w[(t + 1) & 15], w2 = w[(t + 14) & 15];
                uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
                uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
                w[t & 15] = w[t & 15] + s0 + w[(t + 9) & 15] + s1;
            }
            o[u] = w[t & 15] + sha_k(t);
        }
        dst[i / 4] = make_uint4(o[0], o[1], o[2], o[3]);
    }
}
// K5b (serial): the chaining part, 64 rounds per block over the precomputed W+K.  One warp: the lanes stage
// the next blocks into shared memory with coalesced loads, every lane then runs the same rounds on
// broadcast reads (SIMT makes the redundant lanes free); lane 0 publishes r.
constexpr int kTranscriptStage = 8;   // blocks per shared-memory stage
// Processes blocks [first_blk, first_blk + blk_count) and carries the chaining value in `state` (8 words), so the
// chain can advance while later blobs are still being copied / hashed; the call that reaches the last block
// publishes r.
__global__ void __launch_bounds__(32) transcript_chain_kernel(const uint32_t* __restrict__ wk_all, uint64_t n, Fr* __restrict__ r_mont,
                                                              uint32_t* __restrict__ state, uint64_t first_blk, uint64_t blk_count) {
    __shared__ uint4 stage[2][kTranscriptStage * 16];
    size_t len = 32 + (size_t)n * 160;
    size_t total_blk = (len + 9 + 63) / 64;
    size_t nblk = blk_count;
    int lane = threadIdx.x;
    const uint4* src = reinterpret_cast<const uint4*>(wk_all + first_blk * 64);
    uint32_t st[8];
    if (first_blk == 0) sha256_init(st);
    else for (int j = 0; j < 8; j++) st[j] = state[j];
    size_t nstage = (nblk + kTranscriptStage - 1) / kTranscriptStage;
    auto load_stage = [&](size_t sidx, int buf) {
        size_t base = sidx * kTranscriptStage * 16, total = nblk * 16;
#pragma unroll
        for (int k = 0; k < kTranscriptStage * 16 / 32; k++) {
            size_t idx = base + k * 32 + lane;
            stage[buf][k * 32 + lane] = idx < total ? __ldg(src + idx) : make_uint4(0, 0, 0, 0);
        }
    };
    load_stage(0, 0);
    __syncwarp();
    for (size_t sidx = 0; sidx < nstage; sidx++) {
        int buf = sidx & 1;
        if (sidx + 1 < nstage) load_stage(sidx + 1, buf ^ 1);
        size_t blocks_here = nblk - sidx * kTranscriptStage;
        if (blocks_here > kTranscriptStage) blocks_here = kTranscriptStage;
        for (size_t bi = 0; bi < blocks_here; bi++) {
            const uint4* wv = &stage[buf][bi * 16];
            uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
            for (int i = 0; i < 16; i++) {
                uint4 q = wv[i];
                uint32_t kw[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int u = 0; u < 4; u++) {
                    // shortest dependent chain: only sigma1(e)/ch(e) and sigma0(a)/maj(a) sit between e_i -> e_(i+1), a_i -> a_(i+1)
                    uint32_t y = h + kw[u];                 // off the chain (h, kw known early)
                    uint32_t x = y + d;                     // off the chain
                    uint32_t s1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25);
                    uint32_t ch = (e & f) ^ (~e & g);
                    uint32_t s0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22);
                    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
                    uint32_t e2 = x + s1 + ch;
                    uint32_t t1 = y + s1 + ch;
                    uint32_t a2 = t1 + s0 + mj;
                    h = g; g = f; f = e; e = e2; d = c; c = b; b = a; a = a2;
                }
            }
            st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
        }
        __syncwarp();
    }
    if (lane == 0) {
        for (int j = 0; j < 8; j++) state[j] = st[j];
        if (first_blk + blk_count >= total_blk) {
            Fr raw;
            for (int j = 0; j < 8; j++) raw.l[j] = st[7 - j];
            *r_mont = Fr::from_raw(raw);
        }
    }
}

// K5' (opt-in KZGB200_TRANSCRIPT_TREE): the same transcript entries hashed as a two-level tree with domain separation,
//   leaf j = SHA-256("RCKZGBATCH_LEAF_" | entries [16 j, 16 j + 16))          (entry = C_i | z_i LE | y_i LE | pi_i, 160 bytes)
//   root   = SHA-256("RCKZGBATCH___V1_" | u64be 4096 | u64be n | leaf_0 | leaf_1 | ...),      r = root mod q.
// The leaves run here in parallel (41 dependent compressions); the root is 32 bytes per 16 blobs and is hashed by the host
// runtime (host_sha256.cpp) like the exact transcript.  r then differs from kzg-rs's r (the verdict does not: both are
// Fiat-Shamir challenges over the same data), so this mode is NOT the default; the test suite carries an independent restatement of it.
// entry words of the transcript (40 big-endian words per blob), written once in parallel so that the leaf hashes read
// their blocks with plain vector loads
__global__ void __launch_bounds__(256) transcript_words_kernel(const uint8_t* __restrict__ commitments, const ZY* __restrict__ zy,
                                                               const uint8_t* __restrict__ proofs, uint64_t first_entry, uint64_t entry_count,
                                                               uint32_t* __restrict__ words /* [n][40] */) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= entry_count * 40) return;
    size_t w = first_entry * 40 + i;
    words[w] = transcript_word(8 + w, 0, reinterpret_cast<const uint32_t*>(commitments), zy, reinterpret_cast<const uint32_t*>(proofs));
}
__global__ void __launch_bounds__(64) transcript_tree_leaf_words_kernel(const uint32_t* __restrict__ words, uint64_t n, uint32_t* __restrict__ digests,
                                                                        uint64_t first_group, uint64_t group_count) {
    uint64_t g = first_group + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t ngroups = (n + kTreeGroup - 1) / kTreeGroup;
    if (g >= ngroups || g >= first_group + group_count) return;
    uint64_t first = g * kTreeGroup, cnt = n - first < (uint64_t)kTreeGroup ? n - first : (uint64_t)kTreeGroup;
    // message = 4 tag words, then 40 words per entry; 16-byte granules: granule 0 is the tag, granule k > 0 is uint4 k-1 of the entries
    size_t nwords = 4 + (size_t)cnt * 40, nblk = (nwords * 4 + 9 + 63) / 64;
    const uint4* src = reinterpret_cast<const uint4*>(words + first * 40);     // 160-byte entries: 16-byte aligned
    const uint4 tag = make_uint4(0x52434b5a, 0x47424154, 0x43485f4c, 0x4541465f);   // "RCKZGBATCH_LEAF_"
    uint32_t st[8], w[16];
    sha256_init(st);
    for (size_t blk = 0; blk < nblk; blk++) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            size_t gi = blk * 4 + j;                      // granule index; all message granules are whole (nwords % 4 == 0)
            uint4 v = make_uint4(0, 0, 0, 0);
            if (gi == 0) v = tag;
            else if (gi * 4 < nwords) v = __ldg(src + gi - 1);
            else if (gi * 4 == nwords) v.x = 0x80000000u;
            w[4 * j] = v.x; w[4 * j + 1] = v.y; w[4 * j + 2] = v.z; w[4 * j + 3] = v.w;
        }
        if (blk == nblk - 1) { w[14] = (uint32_t)(((uint64_t)nwords * 32) >> 32); w[15] = (uint32_t)((uint64_t)nwords * 32); }
        sha256_compress(st, w);
    }
    // digest bytes (big-endian words) so that the host can hash them as they are
    for (int j = 0; j < 8; j++) digests[g * 8 + j] = sha_bswap(st[j]);
}
// r = (32-byte big-endian digest) mod q in Montgomery form: scalar_from_bytes_unchecked (reference src/kzg_proof.rs:74-91) of the
// transcript digest the host runtime computed (host_sha256.cpp)
__global__ void r_from_digest_kernel(const uint32_t* __restrict__ digest, Fr* __restrict__ r_mont) {
    Fr raw;
    for (int j = 0; j < 8; j++) raw.l[j] = sha_bswap(digest[7 - j]);
    *r_mont = Fr::from_raw(raw);
}

}  // namespace kzgb200
