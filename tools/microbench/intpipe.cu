// Integer-pipe microbenchmarks on the B200 ("measured per-chip peak from a dependency-free
// IMAD microbenchmark").  Prints ops/clk/SM for IMAD, IMAD.WIDE, LOP3/SHF/IADD3 mixes and the field kernels.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -I kzg_rs_b200/csrc tools/microbench/intpipe.cu -o /tmp/intpipe
#include <cstdio>
#include <cuda_runtime.h>
#include "field.cuh"
#include "sha256.cuh"
using namespace kzgb200;

template <int ILP> __global__ void k_imad(uint32_t* out, uint32_t a, uint32_t b, int iters) {
    uint32_t x[ILP];
    for (int j = 0; j < ILP; j++) x[j] = threadIdx.x + j;
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int j = 0; j < ILP; j++) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[j]) : "r"(a), "r"(b));
    uint32_t s = 0; for (int j = 0; j < ILP; j++) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> __global__ void k_imadwide(uint32_t* out, uint32_t a, uint32_t b, int iters) {
    unsigned long long x[ILP];
    for (int j = 0; j < ILP; j++) x[j] = threadIdx.x + j;
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int j = 0; j < ILP; j++) asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[j]) : "r"(a), "r"(b));
    unsigned long long s = 0; for (int j = 0; j < ILP; j++) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = (uint32_t)(s ^ (s >> 32));
}
// carry-chained wide MADs as the field rows use them
template <int ILP> __global__ void k_imadwide_cc(uint32_t* out, uint32_t a, uint32_t b, int iters) {
    uint32_t lo[ILP], hi[ILP];
    for (int j = 0; j < ILP; j++) { lo[j] = threadIdx.x + j; hi[j] = j; }
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int j = 0; j < ILP; j += 2)
            asm volatile("mad.lo.cc.u32 %0, %4, %5, %0;\n\tmadc.hi.cc.u32 %1, %4, %5, %1;\n\tmadc.lo.cc.u32 %2, %4, %5, %2;\n\tmadc.hi.u32 %3, %4, %5, %3;"
                         : "+r"(lo[j]), "+r"(hi[j]), "+r"(lo[j + 1]), "+r"(hi[j + 1]) : "r"(a), "r"(b));
    uint32_t s = 0; for (int j = 0; j < ILP; j++) s += lo[j] ^ hi[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int ILP> __global__ void k_alu(uint32_t* out, uint32_t a, uint32_t b, int iters) {
    uint32_t x[ILP];
    for (int j = 0; j < ILP; j++) x[j] = threadIdx.x + j;
    for (int i = 0; i < iters; i++)
#pragma unroll
        for (int j = 0; j < ILP; j++) { x[j] = __funnelshift_r(x[j], x[j], 7) ^ a; x[j] = x[j] + b + i; }
    uint32_t s = 0; for (int j = 0; j < ILP; j++) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_fpmul(Fp* out, const Fp* in, int iters) {
    Fp a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
    for (int i = 0; i < iters; i++) a = a.mul_inl(b);
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
__global__ void k_fpmul2(Fp* out, const Fp* in, int iters) {   // two independent chains per thread
    Fp a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)], c = b;
    for (int i = 0; i < iters; i++) { a = a.mul_inl(b); c = c.mul_inl(b); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a.add_inl(c);
}
__global__ void k_fpmul_call(Fp* out, const Fp* in, int iters) {
    Fp a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
    for (int i = 0; i < iters; i++) a = a * b;
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
__global__ void k_frdual(Fr* out, const Fr* in, int iters) {
    Fr a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
    for (int i = 0; i < iters; i++) a = Fr::mul_dual_inl(a, b, b, a);
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
__global__ void k_frmul(Fr* out, const Fr* in, int iters) {
    Fr a = in[threadIdx.x & 31], b = in[32 + (threadIdx.x & 31)];
    for (int i = 0; i < iters; i++) a = a.mul_inl(b);
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}
__global__ void k_sha(uint32_t* out, int iters) {
    uint32_t st[8], w[16];
    sha256_init(st);
    for (int j = 0; j < 16; j++) w[j] = threadIdx.x * 16 + j;
    for (int i = 0; i < iters; i++) { uint32_t ww[16]; for (int j = 0; j < 16; j++) ww[j] = w[j] + i; sha256_compress(st, ww); }
    out[blockIdx.x * blockDim.x + threadIdx.x] = st[0] ^ st[7];
}

// SHA-256 rounds only (W+K given), as the transcript chain runs them; active = number of active lanes
__global__ void k_sha_chain(uint32_t* out, int iters, int active) {
    if ((int)threadIdx.x >= active) return;
    uint32_t a = 1, b = 2, c = 3, d = 4, e = 5, f = 6, g = 7, h = 8;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 64; r++) {
            uint32_t kw = i * 64 + r;
            uint32_t y = h + kw, x = y + d;
            uint32_t s1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25), ch = (e & f) ^ (~e & g);
            uint32_t s0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22), mj = (a & b) ^ (a & c) ^ (b & c);
            uint32_t e2 = x + s1 + ch, t1 = y + s1 + ch, a2 = t1 + s0 + mj;
            h = g; g = f; f = e; e = e2; d = c; c = b; b = a; a = a2;
        }
    }
    out[threadIdx.x] = a ^ e;
}
// same with the additions issued as IMAD (FMA pipe) instead of IADD3 (ALU pipe)
__device__ __forceinline__ uint32_t fadd(uint32_t x, uint32_t y) { uint32_t r; asm("mad.lo.u32 %0, %1, 1, %2;" : "=r"(r) : "r"(x), "r"(y)); return r; }
__global__ void k_sha_chain_imad(uint32_t* out, int iters, int active) {
    if ((int)threadIdx.x >= active) return;
    uint32_t a = 1, b = 2, c = 3, d = 4, e = 5, f = 6, g = 7, h = 8;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 64; r++) {
            uint32_t kw = i * 64 + r;
            uint32_t y = fadd(h, kw), x = fadd(y, d);
            uint32_t s1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25), ch = (e & f) ^ (~e & g);
            uint32_t s0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22), mj = (a & b) ^ (a & c) ^ (b & c);
            uint32_t sc = fadd(s1, ch);
            uint32_t e2 = fadd(x, sc), t1 = fadd(y, sc), a2 = fadd(t1, fadd(s0, mj));
            h = g; g = f; f = e; e = e2; d = c; c = b; b = a; a = a2;
        }
    }
    out[threadIdx.x] = a ^ e;
}
// additions as IMAD with a multiplier ptxas cannot see (kernel argument == 1): stays on the FMA pipe
__device__ __forceinline__ uint32_t fadd1(uint32_t x, uint32_t y, uint32_t one) { uint32_t r; asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(x), "r"(one), "r"(y)); return r; }
__global__ void k_sha_chain_imad1(uint32_t* out, int iters, int active, uint32_t one) {
    if ((int)threadIdx.x >= active) return;
    uint32_t a = 1, b = 2, c = 3, d = 4, e = 5, f = 6, g = 7, h = 8;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 64; r++) {
            uint32_t kw = i * 64 + r;
            uint32_t y = fadd1(h, kw, one), x = fadd1(y, d, one);
            uint32_t s1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25), ch = (e & f) ^ (~e & g);
            uint32_t s0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22), mj = (a & b) ^ (a & c) ^ (b & c);
            uint32_t sc = fadd1(s1, ch, one);
            uint32_t e2 = fadd1(x, sc, one), t1 = fadd1(y, sc, one), a2 = fadd1(t1, fadd1(s0, mj, one), one);
            h = g; g = f; f = e; e = e2; d = c; c = b; b = a; a = a2;
        }
    }
    out[threadIdx.x] = a ^ e;
}
// hybrid: only the off-chain additions on the FMA pipe
__global__ void k_sha_chain_hyb(uint32_t* out, int iters, int active, uint32_t one) {
    if ((int)threadIdx.x >= active) return;
    uint32_t a = 1, b = 2, c = 3, d = 4, e = 5, f = 6, g = 7, h = 8;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 64; r++) {
            uint32_t kw = i * 64 + r;
            uint32_t y = fadd1(h, kw, one), x = fadd1(y, d, one);
            uint32_t s1 = sha_rotr(e, 6) ^ sha_rotr(e, 11) ^ sha_rotr(e, 25), ch = (e & f) ^ (~e & g);
            uint32_t s0 = sha_rotr(a, 2) ^ sha_rotr(a, 13) ^ sha_rotr(a, 22), mj = (a & b) ^ (a & c) ^ (b & c);
            uint32_t e2 = x + s1 + ch, t1 = y + s1 + ch, a2 = t1 + fadd1(s0, mj, one);
            h = g; g = f; f = e; e = e2; d = c; c = b; b = a; a = a2;
        }
    }
    out[threadIdx.x] = a ^ e;
}
template <int OP> __global__ void k_lat(uint32_t* out, int iters, uint32_t k) {
    uint32_t x = threadIdx.x, y = k;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int r = 0; r < 64; r++) {
            if (OP == 0) x = __funnelshift_r(x, x, 7) ;
            if (OP == 1) x = (x & y) ^ (~x & k);
            if (OP == 2) x = x + y + k;
            if (OP == 3) x = fadd(x, y);
        }
    }
    out[threadIdx.x] = x;
}

template <class F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
// Do the two pipes overlap?  Even warps run the ALU stream, odd warps the carry-chained wide MADs (mode 3), or every warp one
// of them (modes 1, 2): if the pipes are independent, mode 3 takes max(mode 1, mode 2) of the half-populated runs, not the sum.
template <int ILP> __global__ void k_mix(uint32_t* out, uint32_t a, uint32_t b, int iters, int mode) {
    int warp = threadIdx.x >> 5;
    bool do_alu = mode == 1 || (mode == 3 && !(warp & 1)), do_fma = mode == 2 || (mode == 3 && (warp & 1));
    if (mode != 3 && (warp & 1)) return;             // modes 1, 2: only the even warps work (same warp count per stream as mode 3)
    uint32_t s = 0;
    if (do_alu) {
        uint32_t x[ILP];
        for (int j = 0; j < ILP; j++) x[j] = threadIdx.x + j;
        for (int i = 0; i < iters; i++)
#pragma unroll
            for (int j = 0; j < ILP; j++) { x[j] = __funnelshift_r(x[j], x[j], 7) ^ a; x[j] = x[j] + b + i; }
        for (int j = 0; j < ILP; j++) s += x[j];
    }
    if (do_fma) {
        uint32_t lo[ILP], hi[ILP];
        for (int j = 0; j < ILP; j++) { lo[j] = threadIdx.x + j; hi[j] = j; }
        for (int i = 0; i < iters; i++)
#pragma unroll
            for (int j = 0; j < ILP; j += 2)
                asm volatile("mad.lo.cc.u32 %0, %4, %5, %0;\n\tmadc.hi.cc.u32 %1, %4, %5, %1;\n\tmadc.lo.cc.u32 %2, %4, %5, %2;\n\tmadc.hi.u32 %3, %4, %5, %3;"
                             : "+r"(lo[j]), "+r"(hi[j]), "+r"(lo[j + 1]), "+r"(hi[j + 1]) : "r"(a), "r"(b));
        for (int j = 0; j < ILP; j++) s += lo[j] ^ hi[j];
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    int sms = p.multiProcessorCount; int clk_khz; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    printf("device %s sms %d clock %d kHz\n", p.name, sms, clk_khz);
    void* buf; cudaMalloc(&buf, 1 << 28); cudaMemset(buf, 1, 1 << 28);
    Fp* fin; cudaMalloc(&fin, 64 * sizeof(Fp)); cudaMemset(fin, 3, 64 * sizeof(Fp));
    double hz = clk_khz * 1e3;
    auto rep = [&](const char* name, float ms, double ops_per_thread, int blocks, int threads) {
        double total = ops_per_thread * blocks * threads;
        printf("%-34s %8.3f ms  %8.2f ops/clk/SM  (%.3e ops/s)\n", name, ms, total / (ms * 1e-3) / hz / sms, total / (ms * 1e-3));
    };
    const int T = 256, B = sms * 8, it = 4096;
    rep("IMAD (mad.lo.u32) ilp8", timeit([&] { k_imad<8><<<B, T>>>((uint32_t*)buf, 3, 5, it); }), 8.0 * it, B, T);
    rep("IMAD.WIDE (mad.wide.u32) ilp8", timeit([&] { k_imadwide<8><<<B, T>>>((uint32_t*)buf, 3, 5, it); }), 8.0 * it, B, T);
    rep("IMAD.WIDE.X carry-chained ilp8", timeit([&] { k_imadwide_cc<8><<<B, T>>>((uint32_t*)buf, 3, 5, it); }), 8.0 * it, B, T);
    rep("ALU (SHF+LOP3+IADD3) ilp8, 3 ops", timeit([&] { k_alu<8><<<B, T>>>((uint32_t*)buf, 3, 5, it); }), 8.0 * 3 * it, B, T);
    {   // pipe overlap: 8 warps per SM sub-partition in mode 3 (4 ALU + 4 FMA), 4 in modes 1 / 2
        const int Tm = 1024, Bm = sms;
        float t1 = timeit([&] { k_mix<8><<<Bm, Tm>>>((uint32_t*)buf, 3, 5, it, 1); });
        float t2 = timeit([&] { k_mix<8><<<Bm, Tm>>>((uint32_t*)buf, 3, 5, it, 2); });
        float t3 = timeit([&] { k_mix<8><<<Bm, Tm>>>((uint32_t*)buf, 3, 5, it, 3); });
        printf("pipe overlap: ALU stream alone %.3f ms, IMAD.WIDE.X stream alone %.3f ms, both on the same sub-partitions %.3f ms (sum %.3f, max %.3f)\n",
               t1, t2, t3, t1 + t2, t1 > t2 ? t1 : t2);
    }
    for (int occ : {1, 2, 4, 8}) {
        int Bf = sms * occ, Tf = 128, itf = 512; char nm[64];
        snprintf(nm, 64, "Fp mul inl, %d CTA/SM x128", occ); rep(nm, timeit([&] { k_fpmul<<<Bf, Tf>>>((Fp*)buf, fin, itf); }), itf, Bf, Tf);
        snprintf(nm, 64, "Fp mul inl x2 chains, %d CTA/SM", occ); rep(nm, timeit([&] { k_fpmul2<<<Bf, Tf>>>((Fp*)buf, fin, itf); }), 2.0 * itf, Bf, Tf);
        snprintf(nm, 64, "Fp mul call, %d CTA/SM x128", occ); rep(nm, timeit([&] { k_fpmul_call<<<Bf, Tf>>>((Fp*)buf, fin, itf); }), itf, Bf, Tf);
        snprintf(nm, 64, "Fr mul inl, %d CTA/SM x128", occ); rep(nm, timeit([&] { k_frmul<<<Bf, Tf>>>((Fr*)buf, (Fr*)fin, itf); }), itf, Bf, Tf);
        snprintf(nm, 64, "Fr dual inl, %d CTA/SM x128", occ); rep(nm, timeit([&] { k_frdual<<<Bf, Tf>>>((Fr*)buf, (Fr*)fin, itf); }), itf, Bf, Tf);
        snprintf(nm, 64, "SHA-256 compress, %d CTA/SM x128", occ); rep(nm, timeit([&] { k_sha<<<Bf, Tf>>>((uint32_t*)buf, 256); }), 256, Bf, Tf);
    }
    // single-warp latency of one Fp mul / one SHA compression
    rep("Fp mul inl, 1 warp total (latency)", timeit([&] { k_fpmul<<<1, 32>>>((Fp*)buf, fin, 4096); }), 4096, 1, 32);
    rep("Fp mul call, 1 warp total", timeit([&] { k_fpmul_call<<<1, 32>>>((Fp*)buf, fin, 4096); }), 4096, 1, 32);
    rep("SHA compress, 1 warp total", timeit([&] { k_sha<<<1, 32>>>((uint32_t*)buf, 4096); }), 4096, 1, 32);
    for (int act : {32, 16, 1}) {
        char nm[64]; snprintf(nm, 64, "SHA rounds chain, 1 warp, %d lanes", act);
        float ms = timeit([&] { k_sha_chain<<<1, 32>>>((uint32_t*)buf, 4096, act); });
        printf("%-34s %8.3f ms  %.1f clk/round\n", nm, ms, ms * 1e-3 * hz / (4096.0 * 64));
        snprintf(nm, 64, "SHA rounds chain IMAD-adds, %d lanes", act);
        ms = timeit([&] { k_sha_chain_imad<<<1, 32>>>((uint32_t*)buf, 4096, act); });
        printf("%-34s %8.3f ms  %.1f clk/round\n", nm, ms, ms * 1e-3 * hz / (4096.0 * 64));
    }
    { float ms = timeit([&] { k_sha_chain_imad1<<<1, 32>>>((uint32_t*)buf, 4096, 32, 1); });
      printf("SHA rounds chain IMAD(one) adds      %8.3f ms  %.1f clk/round\n", ms, ms * 1e-3 * hz / (4096.0 * 64));
      ms = timeit([&] { k_sha_chain_hyb<<<1, 32>>>((uint32_t*)buf, 4096, 32, 1); });
      printf("SHA rounds chain hybrid adds         %8.3f ms  %.1f clk/round\n", ms, ms * 1e-3 * hz / (4096.0 * 64)); }
    { float ms;
      ms = timeit([&] { k_lat<0><<<1, 32>>>((uint32_t*)buf, 4096, 5); }); printf("dependent SHF latency   %.2f clk\n", ms * 1e-3 * hz / (4096.0 * 64));
      ms = timeit([&] { k_lat<1><<<1, 32>>>((uint32_t*)buf, 4096, 5); }); printf("dependent LOP3 latency  %.2f clk\n", ms * 1e-3 * hz / (4096.0 * 64));
      ms = timeit([&] { k_lat<2><<<1, 32>>>((uint32_t*)buf, 4096, 5); }); printf("dependent IADD3 latency %.2f clk\n", ms * 1e-3 * hz / (4096.0 * 64));
      ms = timeit([&] { k_lat<3><<<1, 32>>>((uint32_t*)buf, 4096, 5); }); printf("dependent IMAD latency  %.2f clk\n", ms * 1e-3 * hz / (4096.0 * 64)); }
    return 0;
}
