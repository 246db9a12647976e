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

// ring depth: 4 stages (three 64-byte blocks, ~5 us, in flight per thread) or 8 (seven blocks); see kzgb200_ctx::sha_stages
void launch_challenge(int stages, cudaStream_t st, const uint8_t* blobs, const uint8_t* commitments, int n, Fr* z_mont, ZY* zy, Fr* zpow,
                      const uint8_t* slab_flags) {
    unsigned grid = (unsigned)((n + kShaThreads - 1) / kShaThreads);
    if (slab_flags) challenge_kernel<8, true><<<grid, kShaThreads, 0, st>>>(blobs, commitments, n, z_mont, zy, zpow, 1u, slab_flags);
    else if (stages >= 8) challenge_kernel<8, false><<<grid, kShaThreads, 0, st>>>(blobs, commitments, n, z_mont, zy, zpow, 1u, nullptr);
    else challenge_kernel<4, false><<<grid, kShaThreads, 0, st>>>(blobs, commitments, n, z_mont, zy, zpow, 1u, nullptr);
}

}  // namespace kzgb200
