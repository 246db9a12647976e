// K7 (config 5): m independent verify_kzg_proof tuples.
#include "common.cuh"

namespace kzgb200 {

// m independent verify_kzg_proof tuples, one thread each (reference src/kzg_proof.rs:353-397)
__global__ void __launch_bounds__(64) verify_many_kernel(const uint8_t* __restrict__ c, const uint8_t* __restrict__ z, const uint8_t* __restrict__ y,
                                                         const uint8_t* __restrict__ p, size_t m, const DeviceTables* __restrict__ T,
                                                         uint8_t* __restrict__ verdicts) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= m) return;
    uint8_t cb[48], zb[32], yb[32], pb[48];
    for (int k = 0; k < 48; k++) { cb[k] = c[i * 48 + k]; pb[k] = p[i * 48 + k]; }
    for (int k = 0; k < 32; k++) { zb[k] = z[i * 32 + k]; yb[k] = y[i * 32 + k]; }
    verdicts[i] = verify_kzg_proof_one(cb, zb, yb, pb, &T->pairing);
}

}  // namespace kzgb200
