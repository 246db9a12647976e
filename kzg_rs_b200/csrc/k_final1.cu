// K7: single-blob final check (cooperative ladder + pairing engine).
#include "common.cuh"
#include "coop.cuh"

namespace kzgb200 {

// Single-blob path (reference src/kzg_proof.rs:446-470 -> verify_kzg_proof_impl :203-223) after z, y and the
// points have been produced by the kernels above:  e(C - [y]G + [z]pi, G2) e(-pi, [tau]G2) == 1.
__global__ void __launch_bounds__(kFinalThreads) single_final_kernel(const G1Affine* __restrict__ C, const G1Affine* __restrict__ P, const ZY* __restrict__ zy,
                                                                     const uint32_t* __restrict__ status, const DeviceTables* __restrict__ T,
                                                                     uint32_t* __restrict__ result, FinalPts* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    FinalSmem& S = *reinterpret_cast<FinalSmem*>(dyn_smem);
    G1* sm = S.sm;
    int t = threadIdx.x;
    if (t == 0) result[2] = 0;
    if (status[0]) { if (t == 0) { result[0] = kBadArgs; result[1] = status[0]; out->go = 0; } return; }
    G1 yg = coop_fixed_base_mul(zy[0].y, T, sm);
    __shared__ CoopPoint ladder;
    if (t < 32) {   // [z]pi: the 255-step double-and-add chain on the warp-cooperative point operations
        if (t == 0) { ladder.v[0] = Fp::one(); ladder.v[1] = Fp::one(); ladder.v[2] = Fp::zero(); }
        __syncwarp();
        G1 pj = G1::from_affine(P[0]);
        Fr z = zy[0].z;
        for (int bit = 254; bit >= 0; bit--) {
            coop_dbl(&ladder, t);
            if ((z.l[bit >> 5] >> (bit & 31)) & 1) coop_add(&ladder, pj, t);
        }
    }
    __syncthreads();
    if (t == 0) {    // hand (-pi, C - [y]G + [z]pi) to pairing_check_kernel: e(pts[1], G2) e(pts[0], [tau]G2) == 1
        G1 zpi = {ladder.v[0], ladder.v[1], ladder.v[2]};
        G1 acc = yg.neg().add_mixed(C[0]).add(zpi);
        Fp zi = vliw::fp_inv_bingcd(acc.z), zi2 = zi.sqr();
        out->pts[1] = acc.is_identity() ? G1Affine{Fp::zero(), Fp::zero(), 1} : G1Affine{acc.x * zi2, acc.y * zi2 * zi, 0};
        G1Affine np = P[0];
        if (!np.inf) np.y = np.y.neg();
        out->pts[0] = np;
        out->go = 1;
    }
}


}  // namespace kzgb200
