// Microbenchmark of the pairing engine's LIN instruction (vliw::exec_lin): cycles per term and fixed cost (reduction) for one
// warp whose lanes read different registers, as in a level of the engine.
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I kzg_rs_b200/csrc -o tools/microbench/linlevel tools/microbench/linlevel.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "field.cuh"
#include "vliw_programs.cuh"
#include "pairing.cuh"
#include "vliw.cuh"
using namespace kzgb200;
__global__ void k_lin(int K, int lanes, long long* out, uint32_t* sink) {
    __shared__ Fp regs[160];
    __shared__ uint16_t terms[32 * 24];
    __shared__ uint32_t ins[32][3];
    int t = threadIdx.x;
    for (int i = t; i < 160; i += 32) { for (int j = 0; j < 12; j++) regs[i].l[j] = 0x01234567u * (i + 1) + j; regs[i].l[11] &= 0x0fffffffu; }
    for (int k = 0; k < 24; k++) terms[t * 24 + k] = (uint16_t)(((t * 7 + k * 13) % 128) | ((k & 1) ? 0x4000 : 0) | ((k % 3 == 0) ? 0x8000 : 0));
    ins[t][0] = 128 + t; ins[t][1] = t * 24; ins[t][2] = K;
    __syncwarp();
    long long c0 = clock64();
    for (int r = 0; r < 64; r++) {
        if (t < lanes) vliw::exec_lin(regs, ins[t], terms);
        __syncwarp();
    }
    long long c1 = clock64();
    if (t == 0) out[0] = (c1 - c0) / 64;
    sink[t] = regs[128 + t].l[0];
}
int main() {
    long long* d; uint32_t* s; cudaMalloc(&d, 8); cudaMalloc(&s, 128);
    for (int lanes : {32, 1}) for (int K : {0, 1, 2, 4, 8, 12, 16, 24}) {
        k_lin<<<1, 32>>>(K, lanes, d, s); k_lin<<<1, 32>>>(K, lanes, d, s);
        long long h; cudaMemcpy(&h, d, 8, cudaMemcpyDeviceToHost);
        printf("lanes %2d K %2d: %lld cycles per LIN instruction\n", lanes, K, h);
    }
    return 0;
}
