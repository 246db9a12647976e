// Lone-warp microbenchmark for the pairing engine (round 2): what does ONE warp on an SM sub-partition pay per instruction?
//   * streams of independent IMAD.WIDE accumulates (ILP 28), with 32 / 16 / 1 active lanes, accumulate vs RZ addend form
//   * the engine's MUL (fp29.cuh mont_mul29: dual, single) and LIN (vliw29::exec_lin) on a shared-memory register file
//   * 1, 2, 4 warps per sub-partition running the same MUL (is a level issue-bound or latency-bound?)
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I kzg_rs_b200/csrc -o /tmp/lonewarp tools/microbench/lonewarp.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "vliw29.cuh"
using namespace kzgb200;

template <int ILP>
__global__ void k_stream_signed(long long* out, uint32_t* sink, int32_t a, int32_t b) {
    uint64_t x[ILP];
    for (int j = 0; j < ILP; j++) x[j] = threadIdx.x + j;
    long long c0 = clock64();
#pragma unroll 1
    for (int r = 0; r < 256; r++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) f29::madw_s(x[j], a + j, b);
    }
    long long c1 = clock64();
    uint64_t s = 0; for (int j = 0; j < ILP; j++) s ^= x[j];
    sink[threadIdx.x] = (uint32_t)s ^ (uint32_t)(s >> 32);
    if (threadIdx.x == 0) out[0] = c1 - c0;
}
// SHFL stream: ILP independent rotations per iteration
template <int ILP>
__global__ void k_shfl(long long* out, uint32_t* sink) {
    int32_t x[ILP];
    for (int j = 0; j < ILP; j++) x[j] = threadIdx.x * 3 + j;
    long long c0 = clock64();
#pragma unroll 1
    for (int r = 0; r < 256; r++) {
#pragma unroll
        for (int j = 0; j < ILP; j++) x[j] = __shfl_sync(0xffffffffu, x[j], (threadIdx.x - j - 1) & 31);
    }
    long long c1 = clock64();
    int32_t s = 0; for (int j = 0; j < ILP; j++) s ^= x[j];
    sink[threadIdx.x] = (uint32_t)s;
    if (threadIdx.x == 0) out[0] = c1 - c0;
}
template <int ILP, bool ACC>
__global__ void k_stream(long long* out, uint32_t* sink, uint32_t a, uint32_t b, int lanes) {
    uint64_t x[ILP];
    for (int j = 0; j < ILP; j++) x[j] = threadIdx.x + j;
    long long c0 = clock64();
    if ((int)(threadIdx.x & 31) < lanes) {
#pragma unroll 1
        for (int r = 0; r < 256; r++) {
#pragma unroll
            for (int j = 0; j < ILP; j++) {
                if (ACC) f29::madw(x[j], a + j, b);
                else { uint64_t y = 0; f29::madw(y, (uint32_t)x[j], b); x[j] = y; }
            }
        }
    }
    long long c1 = clock64();
    uint64_t s = 0; for (int j = 0; j < ILP; j++) s ^= x[j];
    sink[threadIdx.x] = (uint32_t)s ^ (uint32_t)(s >> 32);
    if (threadIdx.x == 0) out[0] = c1 - c0;
}
// mode 0: dual MUL, 1: single MUL, 2: dual subtracted; every warp of the CTA runs the same thing on its own registers
__global__ void k_mul(long long* out, int mode, int reps) {
    extern __shared__ __align__(16) unsigned char smem[];
    f29::F29* regs = reinterpret_cast<f29::F29*>(smem);
    int t = threadIdx.x;
    for (int i = t; i < 5 * (int)blockDim.x; i += blockDim.x) { for (int j = 0; j < 14; j++) regs[i].l[j] = (0x01234567u * (i + 1) + j * 0x9e3779b9u) & 0x1fffffffu; regs[i].l[13] &= 7u; regs[i].l[14] = regs[i].l[15] = 0; }
    __syncthreads();
    uint32_t ins[4];
    ins[0] = (uint32_t)(5 * t + 4) | ((uint32_t)(5 * t) << 16);
    ins[1] = (uint32_t)(5 * t + 1) | ((uint32_t)(5 * t + 2) << 16);
    ins[2] = (uint32_t)(5 * t + 3) | ((mode == 1 ? 0u : (mode == 2 ? 3u : 1u)) << 16);
    ins[3] = mode == 2 ? 64u : 0u;
    long long c0 = clock64();
    for (int r = 0; r < reps; r++) { vliw29::exec_mul_ref(regs, ins); __syncwarp(); }
    long long c1 = clock64();
    if (t == 0) out[0] = (c1 - c0) / reps;
}
__global__ void k_lin(long long* out, int K, int reps, int reduce) {
    extern __shared__ __align__(16) unsigned char smem[];
    f29::F29* regs = reinterpret_cast<f29::F29*>(smem);
    __shared__ uint32_t terms[64 * 24];
    int t = threadIdx.x;
    for (int i = t; i < 200; i += blockDim.x) { for (int j = 0; j < 14; j++) regs[i].l[j] = (0x01234567u * (i + 1) + j * 0x9e3779b9u) & 0x1fffffffu; regs[i].l[13] &= 7u; regs[i].l[14] = regs[i].l[15] = 0; }
    for (int k = 0; k < 24; k++) terms[t * 24 + k] = (uint32_t)(((t * 7 + k * 13) % 128) * 64) | ((uint32_t)(((k & 1) ? -1 : 1) * ((k % 3 == 0) ? 2 : 1)) << 16);
    __syncthreads();
    uint32_t ins[4] = {(uint32_t)(128 + t), (uint32_t)(t * 24), (uint32_t)K, reduce ? 1u : 0u};
    long long c0 = clock64();
    for (int r = 0; r < reps; r++) { vliw29::exec_lin_ref(regs, ins, terms); __syncwarp(); }
    long long c1 = clock64();
    if (t == 0) out[0] = (c1 - c0) / reps;
}
int main() {
    long long* d; uint32_t* s; cudaMalloc(&d, 8); cudaMalloc(&s, 4096);
    long long h;
    auto get = [&] { cudaDeviceSynchronize(); cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost); return h; };
    for (int lanes : {32, 16, 8, 1}) {
        k_stream<28, true><<<1, 32>>>(d, s, 3, 5, lanes); k_stream<28, true><<<1, 32>>>(d, s, 3, 5, lanes);
        printf("IMAD.WIDE accumulate stream, 1 warp, %2d lanes, ILP 28: %.2f clk/instr\n", lanes, get() / (256.0 * 28));
        k_stream<28, false><<<1, 32>>>(d, s, 3, 5, lanes); k_stream<28, false><<<1, 32>>>(d, s, 3, 5, lanes);
        printf("IMAD.WIDE RZ-addend stream,  1 warp, %2d lanes, ILP 28: %.2f clk/instr\n", lanes, get() / (256.0 * 28));
    }
    for (int warps : {1, 2, 4, 8}) {
        k_stream<28, true><<<1, 32 * warps>>>(d, s, 3, 5, 32); k_stream<28, true><<<1, 32 * warps>>>(d, s, 3, 5, 32);
        printf("IMAD.WIDE accumulate stream, %d warps in the CTA (= %.1f per sub-partition): %.2f clk/instr/warp\n", warps, warps / 4.0, get() / (256.0 * 28));
    }
    for (int warps : {1, 4, 8, 16, 24}) {
        k_stream_signed<28><<<1, 32 * warps>>>(d, s, 3, -5); k_stream_signed<28><<<1, 32 * warps>>>(d, s, 3, -5);
        printf("signed IMAD.WIDE accumulate stream, %2d warps in the CTA (= %.1f per sub-partition): %.2f clk/instr/warp\n", warps, warps / 4.0, get() / (256.0 * 28));
        k_stream<28, true><<<1, 32 * warps>>>(d, s, 3, 5, 32); k_stream<28, true><<<1, 32 * warps>>>(d, s, 3, 5, 32);
        printf("unsigned IMAD.WIDE accumulate stream, %2d warps: %.2f clk/instr/warp\n", warps, get() / (256.0 * 28));
        k_shfl<14><<<1, 32 * warps>>>(d, s); k_shfl<14><<<1, 32 * warps>>>(d, s);
        printf("SHFL.IDX stream (ILP 14), %2d warps: %.2f clk/instr/warp\n", warps, get() / (256.0 * 14));
    }
    cudaFuncSetAttribute(k_mul, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(k_lin, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    for (int mode : {0}) for (int threads : {32}) {
        size_t sm = (size_t)threads * 5 * 64;
        k_mul<<<1, threads, sm>>>(d, mode, 50); k_mul<<<1, threads, sm>>>(d, mode, 50);
        printf("MUL mode %d (0 dual, 1 single, 2 dual subtracted), %3d threads (%4.1f warps per sub-partition): %lld clk per MUL per warp\n", mode, threads, threads / 128.0, get());
    }
    for (int reduce : {0}) for (int K : {4}) for (int threads : {32}) {
        k_lin<<<1, threads, 200 * 64>>>(d, K, 50, reduce); k_lin<<<1, threads, 200 * 64>>>(d, K, 50, reduce);
        printf("LIN %2d terms reduce %d, %2d threads: %lld clk per LIN\n", K, reduce, threads, get());
    }
    return 0;
}
