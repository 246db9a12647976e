// K2: Fiat-Shamir challenge hash, one thread per blob.
#include "common.cuh"

namespace kzgb200 {

// ------------------------------------------------------------------------------------------------ K2
// Fiat-Shamir challenge (reference src/kzg_proof.rs:46-72): z = SHA-256("FSBLOBVERIFY_V1_" | u64be 0 |
// u64be 4096 | blob | commitment) mod q.  One thread per blob: the 2050-block chain is serial per blob, so
// throughput comes from hashing many blobs at once.  The commitment bytes hashed are the caller's: for
// every encoding from_compressed accepts, to_compressed(from_compressed(b)) == b.
// z^(2^k), k = 0..12, for the evaluation tree (K1+K3) are produced here too: 12 squarings per blob after the hash.
// (Measured dead end, tools/microbench/shachain.cu: riding per-element work -- canonicity screen, sum of the elements --
// on this chain costs 12 % of its speed however it is phrased, IADD3 carry chains or IMAD.WIDE column sums: the chain
// is one warp per SM sub-partition issuing ALU-pipe instructions back to back, and ptxas' schedule of it is fragile.)
// Blocks 1..2047 of the challenge hash (99.9 % of the kernel).  The 64 bytes of the blocks ahead are brought in by cp.async
// into a per-thread ring in shared memory (4 stages, three blocks = ~5 us in flight), so the DRAM latency stays hidden
// whatever ptxas does with the loop: with register prefetching the same source ran at 3.3 ms per chain when the loads were
// scheduled at the top of the body and at 4.3 ms when ptxas sank them to the bottom (tools/microbench/shachain.cu).  Each
// thread reads back only what it copied itself, so no barrier is needed, only cp.async.wait_group.
__device__ __forceinline__ void sha_cp_async16(uint4* smem_dst, const uint4* gsrc) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)), "l"(gsrc) : "memory");
}
// Slab-wise arrival (the LAST host chunk of a batch, kzgb200.cu launch_phase1): the chunk is copied as kSlabs strided slabs --
// bytes [16 KiB q, 16 KiB (q+1)) of every blob -- and the hash kernel is launched when slab 0 is there; flags[q] != 0 (set by a
// memset queued behind slab q's copy) tells that slab q has landed.  The chain then finishes ~0.4 ms after the last byte
// instead of a whole chain (2.7 ms) after it.
constexpr int kSlabs = 8, kSlabBlocks = 2048 / kSlabs;
__device__ __forceinline__ void wait_slab(const volatile uint8_t* flags, int q) {
    while (flags[q] == 0) __nanosleep(200);
}
template <int kShaStages, bool kWait>
__device__ __forceinline__ void sha256_blob_body(uint32_t st[8], const uint4* __restrict__ bp, uint4 (*ring)[4][kShaThreads] /* [stage][quarter][thread] */,
                                                 uint32_t one, const volatile uint8_t* flags) {
    const int t = threadIdx.x;
    uint32_t w[16];
    int have = 1;                        // slabs known to have landed (slab 0: before the launch)
#pragma unroll
    for (int k = 1; k < kShaStages; k++) {
#pragma unroll
        for (int q = 0; q < 4; q++) sha_cp_async16(&ring[k % kShaStages][q][t], bp + (4 * k - 2) + q);
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
#pragma unroll 1
    for (int k = 1; k < 2048; k++) {
        asm volatile("cp.async.wait_group %0;" ::"n"(kShaStages - 2) : "memory");    // block k has landed
        uint4 (*sg)[kShaThreads] = ring[k % kShaStages];
        uint4 a = sg[0][t], b = sg[1][t], c = sg[2][t], d = sg[3][t];
        if (k + kShaStages - 1 < 2048) {                                               // refill the stage consumed last time
            if (kWait && (k + kShaStages - 1) / kSlabBlocks >= have) wait_slab(flags, have++);   // uniform over the warp
            const uint4* p = bp + (4 * (k + kShaStages - 1) - 2);
#pragma unroll
            for (int q = 0; q < 4; q++) sha_cp_async16(&ring[(k + kShaStages - 1) % kShaStages][q][t], p + q);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");                          // (possibly empty: keeps the group count uniform)
        w[0] = sha_bswap(a.x); w[1] = sha_bswap(a.y); w[2] = sha_bswap(a.z); w[3] = sha_bswap(a.w);
        w[4] = sha_bswap(b.x); w[5] = sha_bswap(b.y); w[6] = sha_bswap(b.z); w[7] = sha_bswap(b.w);
        w[8] = sha_bswap(c.x); w[9] = sha_bswap(c.y); w[10] = sha_bswap(c.z); w[11] = sha_bswap(c.w);
        w[12] = sha_bswap(d.x); w[13] = sha_bswap(d.y); w[14] = sha_bswap(d.z); w[15] = sha_bswap(d.w);
        sha256_compress_bal(st, w, one);
    }
}
template <int kShaStages, bool kWait>
__global__ void __launch_bounds__(kShaThreads) challenge_kernel(const uint8_t* __restrict__ blobs, const uint8_t* __restrict__ commitments,
                                                       int n, Fr* __restrict__ z_mont, ZY* __restrict__ zy, Fr* __restrict__ zpow,
                                                       uint32_t one /* == 1, opaque to the compiler: see sha256_compress_bal */,
                                                       const volatile uint8_t* flags /* kWait: kSlabs arrival flags */) {
    __shared__ uint4 ring[kShaStages][4][kShaThreads];
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* bp = reinterpret_cast<const uint4*>(blobs + (size_t)i * kBytesPerBlob);
    const uint32_t* cp = reinterpret_cast<const uint32_t*>(commitments + (size_t)i * 48);
    uint32_t st[8], w[16];
    sha256_init(st);
    // block 0: domain | 0 | 4096 | blob[0..32)
    w[0] = 0x4653424c; w[1] = 0x4f425645; w[2] = 0x52494659; w[3] = 0x5f56315f;   // "FSBLOBVERIFY_V1_"
    w[4] = 0; w[5] = 0; w[6] = 0; w[7] = 4096;
    {
        uint4 a = __ldg(bp), b = __ldg(bp + 1);
        w[8] = sha_bswap(a.x); w[9] = sha_bswap(a.y); w[10] = sha_bswap(a.z); w[11] = sha_bswap(a.w);
        w[12] = sha_bswap(b.x); w[13] = sha_bswap(b.y); w[14] = sha_bswap(b.z); w[15] = sha_bswap(b.w);
    }
    sha256_compress(st, w);
    // blocks 1..2047: blob[64k-32 .. 64k+32)
    sha256_blob_body<kShaStages, kWait>(st, bp, ring, one, flags);
    if (kWait) wait_slab(flags, kSlabs - 1);
    // block 2048: blob[131040..131072) | commitment[0..32)
    {
        uint4 a = __ldg(bp + 8190), b = __ldg(bp + 8191);
        w[0] = sha_bswap(a.x); w[1] = sha_bswap(a.y); w[2] = sha_bswap(a.z); w[3] = sha_bswap(a.w);
        w[4] = sha_bswap(b.x); w[5] = sha_bswap(b.y); w[6] = sha_bswap(b.z); w[7] = sha_bswap(b.w);
        for (int j = 0; j < 8; j++) w[8 + j] = sha_bswap(__ldg(cp + j));
    }
    sha256_compress(st, w);
    // block 2049: commitment[32..48) | 0x80 | 0.. | bit length 131152*8
    for (int j = 0; j < 4; j++) w[j] = sha_bswap(__ldg(cp + 8 + j));
    w[4] = 0x80000000u;
    for (int j = 5; j < 15; j++) w[j] = 0;
    w[15] = 131152u * 8u;
    sha256_compress(st, w);
    // scalar_from_bytes_unchecked (kzg_proof.rs:74-91): big-endian 256-bit value reduced mod q
    Fr raw;
    for (int j = 0; j < 8; j++) raw.l[j] = st[7 - j];
    Fr zm = Fr::from_raw(raw);
    z_mont[i] = zm;
    zy[i].z = zm.to_raw();
    Fr s = zm;
#pragma unroll 1
    for (int k = 0; k <= 12; k++) { zpow[(size_t)i * 13 + k] = s; s = s.mul_inl(s); }
}

// ------------------------------------------------------------------------------------------------ K2, warp-specialised form
// The chain above is one warp per SM sub-partition issuing ~1560 instructions per block; only ~830 of them sit on the
// round-to-round dependency (30 clocks per round for a lone warp, tools/microbench/shachain.cu), the rest is loading,
// byte-swapping and the message schedule.  Here a PRODUCER warp does that rest one block ahead -- it leaves W[t] + K[t], t = 0..63,
// in a double-buffered shared-memory tile -- and the CONSUMER warp runs nothing but the 64 rounds.  64-thread CTAs, 32 blobs each;
// two named barriers per buffer (full / empty), the consumer releases a buffer as soon as it has copied it into registers.
constexpr int kWsStride = 68;      // words per thread in a tile: 16-byte aligned rows, conflict-free 128-bit accesses
__device__ __forceinline__ void ws_bar_sync(int id) { asm volatile("barrier.sync %0, 64;" ::"r"(id) : "memory"); }
__device__ __forceinline__ void ws_bar_arrive(int id) { asm volatile("barrier.arrive %0, 64;" ::"r"(id) : "memory"); }
// kConsumers = 2 (one producer warp feeding two consumer warps, 768 instead of 1024 warps for 16384 blobs) was measured: the producer
// becomes the bottleneck, 5.19 against 3.36 ms in the step; only kConsumers = 1 is instantiated.
template <bool kWait, int kConsumers>
__global__ void __launch_bounds__(32 * (kConsumers + 1)) challenge_ws_kernel(const uint8_t* __restrict__ blobs, const uint8_t* __restrict__ commitments, int n,
                                                                             Fr* __restrict__ z_mont, ZY* __restrict__ zy, Fr* __restrict__ zpow,
                                                                             const volatile uint8_t* flags) {
    constexpr int kStages = 8;
    extern __shared__ __align__(16) unsigned char ws_smem[];       // dynamic: two consumers need 66 KB
    uint4 (*ring)[kStages][4][32] = reinterpret_cast<uint4 (*)[kStages][4][32]>(ws_smem);
    uint32_t (*tile)[2][32][kWsStride] = reinterpret_cast<uint32_t (*)[2][32][kWsStride]>(ws_smem + sizeof(uint4) * kConsumers * kStages * 4 * 32);
    const int lane = threadIdx.x & 31, role = threadIdx.x >> 5;
    if (role == kConsumers) {
        // ---- producer: blocks 0 .. 2049 of kConsumers x 32 blobs -> W tiles (the consumers add K[t] as an immediate)
        const uint4* bp[kConsumers];
        const uint32_t* cp[kConsumers];
#pragma unroll
        for (int c = 0; c < kConsumers; c++) {
            int i = (blockIdx.x * kConsumers + c) * 32 + lane;
            if (i >= n) i = n - 1;
            bp[c] = reinterpret_cast<const uint4*>(blobs + (size_t)i * kBytesPerBlob);
            cp[c] = reinterpret_cast<const uint32_t*>(commitments + (size_t)i * 48);
        }
        int have = 1;
        for (int k = 1; k < kStages; k++) {
#pragma unroll
            for (int c = 0; c < kConsumers; c++)
#pragma unroll
                for (int q = 0; q < 4; q++) sha_cp_async16(&ring[c][k % kStages][q][lane], bp[c] + (4 * k - 2) + q);
            asm volatile("cp.async.commit_group;" ::: "memory");
        }
#pragma unroll 1
        for (int b = 0; b < 2050; b++) {
            if (b >= 1 && b < 2048) {
                asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 2) : "memory");      // block b of every consumer has landed
                if (kWait && b + kStages - 1 < 2048 && (b + kStages - 1) / kSlabBlocks >= have) wait_slab(flags, have++);
            }
            if (kWait && b == 2048) wait_slab(flags, kSlabs - 1);
#pragma unroll
            for (int c = 0; c < kConsumers; c++) {
                uint32_t w[16];
                if (b == 0) {
                    w[0] = 0x4653424c; w[1] = 0x4f425645; w[2] = 0x52494659; w[3] = 0x5f56315f;   // "FSBLOBVERIFY_V1_"
                    w[4] = 0; w[5] = 0; w[6] = 0; w[7] = 4096;
                    uint4 a = __ldg(bp[c]), e = __ldg(bp[c] + 1);
                    w[8] = sha_bswap(a.x); w[9] = sha_bswap(a.y); w[10] = sha_bswap(a.z); w[11] = sha_bswap(a.w);
                    w[12] = sha_bswap(e.x); w[13] = sha_bswap(e.y); w[14] = sha_bswap(e.z); w[15] = sha_bswap(e.w);
                } else if (b < 2048) {
                    uint4 (*sg)[32] = ring[c][b % kStages];
                    uint4 a = sg[0][lane], bb = sg[1][lane], e = sg[2][lane], d = sg[3][lane];
                    if (b + kStages - 1 < 2048) {
                        const uint4* p = bp[c] + (4 * (b + kStages - 1) - 2);
#pragma unroll
                        for (int q = 0; q < 4; q++) sha_cp_async16(&ring[c][(b + kStages - 1) % kStages][q][lane], p + q);
                    }
                    w[0] = sha_bswap(a.x); w[1] = sha_bswap(a.y); w[2] = sha_bswap(a.z); w[3] = sha_bswap(a.w);
                    w[4] = sha_bswap(bb.x); w[5] = sha_bswap(bb.y); w[6] = sha_bswap(bb.z); w[7] = sha_bswap(bb.w);
                    w[8] = sha_bswap(e.x); w[9] = sha_bswap(e.y); w[10] = sha_bswap(e.z); w[11] = sha_bswap(e.w);
                    w[12] = sha_bswap(d.x); w[13] = sha_bswap(d.y); w[14] = sha_bswap(d.z); w[15] = sha_bswap(d.w);
                } else if (b == 2048) {
                    uint4 a = __ldg(bp[c] + 8190), e = __ldg(bp[c] + 8191);
                    w[0] = sha_bswap(a.x); w[1] = sha_bswap(a.y); w[2] = sha_bswap(a.z); w[3] = sha_bswap(a.w);
                    w[4] = sha_bswap(e.x); w[5] = sha_bswap(e.y); w[6] = sha_bswap(e.z); w[7] = sha_bswap(e.w);
                    for (int j = 0; j < 8; j++) w[8 + j] = sha_bswap(__ldg(cp[c] + j));
                } else {
                    for (int j = 0; j < 4; j++) w[j] = sha_bswap(__ldg(cp[c] + 8 + j));
                    w[4] = 0x80000000u;
                    for (int j = 5; j < 15; j++) w[j] = 0;
                    w[15] = 131152u * 8u;
                }
                const int buf = b & 1;
                if (b >= 2) ws_bar_sync(3 + 4 * c + buf);         // consumer c has copied tile `buf` (block b - 2) into registers
                uint32_t* out = tile[c][buf][lane];
#pragma unroll
                for (int t = 0; t < 64; t += 4) {
                    uint32_t v[4];
#pragma unroll
                    for (int u = 0; u < 4; u++) {
                        const int x = t + u;
                        if (x >= 16) {
                            uint32_t w15 = w[(x + 1) & 15], w2 = w[(x + 14) & 15];
                            uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
                            uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
                            w[x & 15] = w[x & 15] + s0 + w[(x + 9) & 15] + s1;
                        }
                        v[u] = w[x & 15];
                    }
                    *reinterpret_cast<uint4*>(out + t) = make_uint4(v[0], v[1], v[2], v[3]);
                }
                __threadfence_block();
                ws_bar_arrive(1 + 4 * c + buf);                   // tile `buf` of consumer c is full
            }
            if (b >= 1 && b < 2048) asm volatile("cp.async.commit_group;" ::: "memory");
        }
        return;
    }
    // ---- consumer `role`: the 64 rounds of every block of its 32 blobs
    const int c = role;
    int i = (blockIdx.x * kConsumers + c) * 32 + lane;
    const bool store = i < n;                                 // the tail CTA: surplus lanes redo the last blob (no divergent barriers)
    if (i >= n) i = n - 1;
    uint32_t st[8];
    sha256_init(st);
#pragma unroll 1
    for (int b = 0; b < 2050; b++) {
        const int buf = b & 1;
        ws_bar_sync(1 + 4 * c + buf);
        uint32_t wk[64];
        const uint4* in = reinterpret_cast<const uint4*>(tile[c][buf][lane]);
#pragma unroll
        for (int t = 0; t < 16; t++) { uint4 v = in[t]; wk[4 * t] = v.x; wk[4 * t + 1] = v.y; wk[4 * t + 2] = v.z; wk[4 * t + 3] = v.w; }
        if (b + 2 < 2050) ws_bar_arrive(3 + 4 * c + buf);     // the tile may be refilled (block b + 2)
        uint32_t a = st[0], bq = st[1], cc = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
        for (int t = 0; t < 64; t++) {
            uint32_t t1 = (h + wk[t] + sha_k(t)) + (sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25)) + ((e & f) ^ (~e & g));
            uint32_t t2 = (sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22)) + ((a & bq) ^ (a & cc) ^ (bq & cc));
            h = g; g = f; f = e; e = d + t1; d = cc; cc = bq; bq = a; a = t1 + t2;
        }
        st[0] += a; st[1] += bq; st[2] += cc; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
    }
    if (!store) return;
    Fr raw;
    for (int j = 0; j < 8; j++) raw.l[j] = st[7 - j];
    Fr zm = Fr::from_raw(raw);
    z_mont[i] = zm;
    zy[i].z = zm.to_raw();
    Fr s = zm;
#pragma unroll 1
    for (int k = 0; k <= 12; k++) { zpow[(size_t)i * 13 + k] = s; s = s.mul_inl(s); }
}

// ring depth: 4 stages (three 64-byte blocks, ~5 us, in flight per thread) or 8 (seven blocks); see kzgb200_ctx::sha_stages
void launch_challenge(int stages, cudaStream_t st, const uint8_t* blobs, const uint8_t* commitments, int n, Fr* z_mont, ZY* zy, Fr* zpow,
                      const uint8_t* slab_flags) {
    unsigned grid = (unsigned)((n + kShaThreads - 1) / kShaThreads);
    // The warp-specialised form has the shorter chain per blob (2.18 against 2.71 ms for a 1024-blob launch) but needs two warps per
    // 32 blobs: once chains alone fill the 592 sub-partitions it loses (16384 blobs: 3.49 against 2.91 ms).  It is used where the
    // launch leaves sub-partitions free -- up to kWsMaxBlobs blobs -- and for the slab-wise last host chunk; stages = -1 forces it,
    // stages = 4 keeps the one-warp form everywhere.
    constexpr int kWsMaxBlobs = 8192;
    auto ws_bytes = [](int consumers) { return (size_t)consumers * (sizeof(uint4) * 8 * 4 * 32 + sizeof(uint32_t) * 2 * 32 * kWsStride); };
    if (stages < 0 || (stages != 4 && (slab_flags || n <= kWsMaxBlobs))) {
        unsigned g2 = (unsigned)((n + 31) / 32);
        if (slab_flags) challenge_ws_kernel<true, 1><<<g2, 64, ws_bytes(1), st>>>(blobs, commitments, n, z_mont, zy, zpow, slab_flags);
        else challenge_ws_kernel<false, 1><<<g2, 64, ws_bytes(1), st>>>(blobs, commitments, n, z_mont, zy, zpow, nullptr);
        return;
    }
    if (slab_flags) challenge_kernel<8, true><<<grid, kShaThreads, 0, st>>>(blobs, commitments, n, z_mont, zy, zpow, 1u, slab_flags);
    else if (stages >= 8) challenge_kernel<8, false><<<grid, kShaThreads, 0, st>>>(blobs, commitments, n, z_mont, zy, zpow, 1u, nullptr);
    else challenge_kernel<4, false><<<grid, kShaThreads, 0, st>>>(blobs, commitments, n, z_mont, zy, zpow, 1u, nullptr);
}

}  // namespace kzgb200
