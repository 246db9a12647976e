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
// per-blob verdicts of a batch whose inputs have been parsed / evaluated by the kernels of the batch path (K1..K4): blob i is
// verify_blob_kzg_proof(blob_i, C_i, pi_i) (reference src/kzg_proof.rs:446-470): 2 = Err(BadArgs) if anything of blob i failed to parse
__global__ void __launch_bounds__(64) verify_parsed_each_kernel(const G1Affine* __restrict__ C, const G1Affine* __restrict__ P, const ZY* __restrict__ zy,
                                                                const uint32_t* __restrict__ status, int n, const DeviceTables* __restrict__ T,
                                                                uint8_t* __restrict__ verdicts) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (status[i]) { verdicts[i] = kBadArgs; return; }
    G1Affine c = C[i], p = P[i];
    ZY s = zy[i];
    verdicts[i] = kzg_pairing_check(kzg_lhs_point(c, s.z, s.y, p), p, &T->pairing) ? kTrue : kFalse;
}

}  // namespace kzgb200
