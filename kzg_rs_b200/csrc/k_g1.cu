// K4: G1 decompression and subgroup checks.
#include "common.cuh"

namespace kzgb200 {

// ------------------------------------------------------------------------------------------------ K4
// G1 decompression + subgroup check for commitments and proofs (reference src/kzg_proof.rs:17-25), as two kernels:
// the decompression (Fp square root) produces the affine points the MSM needs; the subgroup check (two 64-bit scalar
// multiplications, ~2/3 of the work) only feeds the error flags.  The decompression runs on a low-priority stream beside
// the hashing; the subgroup checks are deferred beside the latency-bound tail (window sums, Horner, pairing) on SMs of their
// own -- sharing SMs with the tail's few CTAs costs more than the deferral saves, so the host keeps the two apart with a
// shared-memory reservation (kzgb200.cu, launch_lincomb).
// One thread per point; points [0,n) are commitments, [n,2n) proofs.
__global__ void __launch_bounds__(128) g1_decompress_kernel(const uint8_t* __restrict__ commitments, const uint8_t* __restrict__ proofs, int n,
                                                            G1Affine* __restrict__ C, G1Affine* __restrict__ P, uint32_t* __restrict__ status,
                                                            bool with_subgroup_check) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n) return;
    bool is_proof = i >= n;
    int j = is_proof ? i - n : i;
    const uint8_t* src = (is_proof ? proofs : commitments) + (size_t)j * 48;
    uint8_t b[48];
    for (int k = 0; k < 12; k++) {
        uint32_t v = __ldg(reinterpret_cast<const uint32_t*>(src) + k);
        b[4 * k] = (uint8_t)v; b[4 * k + 1] = (uint8_t)(v >> 8); b[4 * k + 2] = (uint8_t)(v >> 16); b[4 * k + 3] = (uint8_t)(v >> 24);
    }
    G1Affine pt;
    bool ok = g1_from_compressed(pt, b, false);
    (is_proof ? P : C)[j] = pt;
    // fused form (the batch path): every parsing CTA is resident from the start of phase 1, beside the hash chains; as a
    // second kernel the subgroup checks queue behind the evaluation kernel's 16384 CTAs once the hashing is fast
    if (ok && with_subgroup_check) ok = g1_in_subgroup(pt);
    if (!ok) atomicOr(&status[j], is_proof ? kErrProof : kErrCommitment);
}
__global__ void __launch_bounds__(256) g1_subgroup_kernel(const G1Affine* __restrict__ C, const G1Affine* __restrict__ P, int n,
                                                          uint32_t* __restrict__ status) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < 2 * n; i += gridDim.x * blockDim.x) {
        bool is_proof = i >= n;
        int j = is_proof ? i - n : i;
        G1Affine pt = (is_proof ? P : C)[j];
        if (!g1_in_subgroup(pt)) atomicOr(&status[j], is_proof ? kErrProof : kErrCommitment);
    }
}

}  // namespace kzgb200
