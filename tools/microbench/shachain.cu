// Microbenchmark: the one-thread-per-blob SHA-256 chain on B200 and how much of its ALU-pipe work (SHF / LOP3 / IADD3,
// one warp instruction per two clocks) can be moved to the FMA pipe (IMAD / IMAD.WIDE) that idles beside it.
// F bits: 1 = message-schedule sigmas through 32x32->64 multiplications (x * 2^(32-n): hi = x >> n, lo = x << (32-n)),
//         2 = message-schedule additions as IMAD, 4 = round additions h + K + W as IMAD, 8 = Sigma0 through multiplications,
//         16 = Sigma1 through multiplications, 32 = t2 / a additions as IMAD.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I kzg_rs_b200/csrc -o tools/microbench/shachain tools/microbench/shachain.cu
#include <cstdio>
#include <vector>
#include <cuda_runtime.h>
#include "sha256.cuh"
using namespace kzgb200;
constexpr int kBytesPerBlob = 131072;
__device__ __forceinline__ uint32_t fadd(uint32_t x, uint32_t y, uint32_t one) { return x * one + y; }
__device__ __forceinline__ uint32_t xrot(uint32_t x, uint32_t m) {   // lo ^ hi of x * m  (m = 2^(32-n): rotr(x, n))
    uint64_t p = (uint64_t)x * m;
    return (uint32_t)p ^ (uint32_t)(p >> 32);
}
__device__ __forceinline__ uint32_t xshr(uint32_t x, uint32_t m) { return (uint32_t)(((uint64_t)x * m) >> 32); }
template <int F>
__device__ __forceinline__ void compress_v(uint32_t st[8], uint32_t w[16], uint32_t one) {
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
    for (int i = 0; i < 64; i++) {
        if (i >= 16) {
            uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15], s0, s1;
            if (F & 1) {
                s0 = xrot(w15, one << 25) ^ xrot(w15, one << 14) ^ xshr(w15, one << 29);
                s1 = xrot(w2, one << 15) ^ xrot(w2, one << 13) ^ xshr(w2, one << 22);
            } else {
                s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
                s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
            }
            if (F & 2) w[i & 15] = fadd(fadd(w[i & 15], s0, one), fadd(w[(i + 9) & 15], s1, one), one);
            else w[i & 15] = w[i & 15] + s0 + w[(i + 9) & 15] + s1;
        }
        uint32_t S1 = (F & 16) ? (xrot(e, one << 26) ^ xrot(e, one << 21) ^ xrot(e, one << 7)) : (sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25));
        uint32_t S0 = (F & 8) ? (xrot(a, one << 30) ^ xrot(a, one << 19) ^ xrot(a, one << 10)) : (sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22));
        uint32_t ch = (e & f) ^ (~e & g), mj = (a & b) ^ (a & c) ^ (b & c);
        uint32_t hkw = (F & 4) ? fadd(fadd(w[i & 15], sha_k(i), one), h, one) : h + sha_k(i) + w[i & 15];
        uint32_t t1 = hkw + S1 + ch;
        uint32_t t2 = (F & 32) ? fadd(S0, mj, one) : S0 + mj;
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = (F & 32) ? fadd(t1, t2, one) : t1 + t2;
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}
// F == 64: rounds 16..63 as a rolled loop of 3 x 16 rounds (K from constant memory): ~620 instead of ~1400 instructions per block
__constant__ uint32_t c_k[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5,
    0xd807aa98, 0x12835b01, 0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174,
    0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc, 0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da,
    0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147, 0x06ca6351, 0x14292967,
    0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070,
    0x19a4c116, 0x1e376c08, 0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3,
    0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208, 0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};
#define SHA_ROUND(KV, WV) { \
        uint32_t t1 = h + (sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25)) + ((e & f) ^ (~e & g)) + (KV) + (WV); \
        uint32_t t2 = (sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22)) + ((a & b) ^ (a & c) ^ (b & c)); \
        h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2; }
template <int UNROLL_OUTER>
__device__ __forceinline__ void compress_rolled(uint32_t st[8], uint32_t w[16]) {
    uint32_t a = st[0], b = st[1], c = st[2], d = st[3], e = st[4], f = st[5], g = st[6], h = st[7];
#pragma unroll
    for (int i = 0; i < 16; i++) SHA_ROUND(sha_k(i), w[i])
#pragma unroll UNROLL_OUTER
    for (int r = 16; r < 64; r += 16) {
#pragma unroll
        for (int i = 0; i < 16; i++) {
            uint32_t w15 = w[(i + 1) & 15], w2 = w[(i + 14) & 15];
            uint32_t s0 = sha_rotr(w15, 7) ^ sha_rotr(w15, 18) ^ (w15 >> 3);
            uint32_t s1 = sha_rotr(w2, 17) ^ sha_rotr(w2, 19) ^ (w2 >> 10);
            w[i] = w[i] + s0 + w[(i + 9) & 15] + s1;
            SHA_ROUND(c_k[r + i], w[i])
        }
    }
    st[0] += a; st[1] += b; st[2] += c; st[3] += d; st[4] += e; st[5] += f; st[6] += g; st[7] += h;
}
template <int F>
__global__ void __launch_bounds__(64) chain(const uint8_t* __restrict__ blobs, int n, uint32_t* __restrict__ out, uint32_t one) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint4* bp = reinterpret_cast<const uint4*>(blobs + (size_t)i * kBytesPerBlob);
    uint32_t st[8], w[16];
    sha256_init(st);
    uint4 na = __ldg(bp + 2), nb = __ldg(bp + 3), nc = __ldg(bp + 4), nd = __ldg(bp + 5);
    for (int k = 1; k < 2048; k++) {
        uint4 a = na, b = nb, c = nc, d = nd;
        if (k < 2047) { const uint4* p = bp + (4 * k + 2); na = __ldg(p); nb = __ldg(p + 1); nc = __ldg(p + 2); nd = __ldg(p + 3); }
        w[0] = sha_bswap(a.x); w[1] = sha_bswap(a.y); w[2] = sha_bswap(a.z); w[3] = sha_bswap(a.w);
        w[4] = sha_bswap(b.x); w[5] = sha_bswap(b.y); w[6] = sha_bswap(b.z); w[7] = sha_bswap(b.w);
        w[8] = sha_bswap(c.x); w[9] = sha_bswap(c.y); w[10] = sha_bswap(c.z); w[11] = sha_bswap(c.w);
        w[12] = sha_bswap(d.x); w[13] = sha_bswap(d.y); w[14] = sha_bswap(d.z); w[15] = sha_bswap(d.w);
        if (F == 64) compress_rolled<1>(st, w); else compress_v<F>(st, w, one);
    }
    uint32_t x = 0;
    for (int j = 0; j < 8; j++) x ^= st[j];
    out[i] = x;
}
static std::vector<uint32_t> ref;
template <int F> void run(const uint8_t* d, int n, uint32_t* o) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    chain<F><<<(n + 63) / 64, 64>>>(d, n, o, 1u);
    cudaEventRecord(e0);
    for (int r = 0; r < 3; r++) chain<F><<<(n + 63) / 64, 64>>>(d, n, o, 1u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    std::vector<uint32_t> h(n);
    cudaMemcpy(h.data(), o, n * 4, cudaMemcpyDeviceToHost);
    if (F == 0) ref = h;
    printf("  F=%2d %.3f ms%s", F, ms / 3, h == ref ? "" : " WRONG");
}
int main() {
    uint8_t* d; uint32_t* o;
    int nmax = 16384;
    cudaMalloc(&d, (size_t)nmax * kBytesPerBlob); cudaMalloc(&o, nmax * 4);
    std::vector<uint8_t> hb((size_t)1 << 24);
    for (size_t i = 0; i < hb.size(); i++) hb[i] = (uint8_t)(i * 2654435761u >> 13);
    for (size_t off = 0; off < (size_t)nmax * kBytesPerBlob; off += hb.size()) cudaMemcpy(d + off, hb.data(), hb.size(), cudaMemcpyHostToDevice);
    for (int r = 0; r < 100; r++) chain<0><<<64, 64>>>(d, 4096, o, 1u);   // clocks up
    cudaDeviceSynchronize();
    for (int n : {64, 16384}) {
        printf("n=%5d", n);
        run<0>(d, n, o); run<64>(d, n, o); run<1>(d, n, o); run<2>(d, n, o); run<3>(d, n, o); run<4>(d, n, o); run<7>(d, n, o); run<6>(d, n, o);
        printf("\n       ");
        run<8>(d, n, o); run<11>(d, n, o); run<15>(d, n, o); run<32>(d, n, o); run<38>(d, n, o); run<39>(d, n, o); run<47>(d, n, o); run<63>(d, n, o); run<19>(d, n, o);
        printf("\n");
    }
    return 0;
}
