// K7: final pairing check of a batch on the cooperative engine.
#include "common.cuh"
#include "coop.cuh"

namespace kzgb200 {

// Final pairing check over the gathered per-rank partials (reference src/kzg_proof.rs:436-441):
//   e(sum_k A_k, [tau]G2) == e(sum_k B_k - [sum_k s_k]G, G2)
// one CTA of kFinalThreads threads: the G1 prelude on a few threads, the pairing on the cooperative engine.
// result: 0 = false, 1 = true, 2 = BadArgs (some rank flagged an unparsable input)
__global__ void __launch_bounds__(kFinalThreads) batch_final_kernel(const Partial* __restrict__ parts, int nparts, const DeviceTables* __restrict__ T,
                                                                    uint32_t* __restrict__ result, long long* __restrict__ ticks) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    FinalSmem& S = *reinterpret_cast<FinalSmem*>(dyn_smem);
    f29::F29* regs = S.regs;
    G1* sm = S.sm;
    __shared__ G1Affine pts[2];
    __shared__ Fr s_sum;
    __shared__ uint32_t s_err;
    int t = threadIdx.x;
    if (t == 0) result[2] = 0;
    if (ticks && t == 0) { ticks[0] = clock64(); for (int i = 8; i < 14; i++) ticks[i] = 0; }
    vliw29::Tables tab = vliw29::load_tables(&S.stab, t, kFinalThreads);
    if (t == 0) {
        Fr s = Fr::zero(); uint32_t err = 0;
        for (int k = 0; k < nparts; k++) { s = s.add_inl(parts[k].ry); err |= parts[k].err; }
        s_sum = s; s_err = err;
    }
    __syncthreads();
    if (s_err) { if (t == 0) { result[0] = kBadArgs; result[1] = s_err; } return; }
    G1 sg = coop_fixed_base_mul(s_sum, T, sm);
    if ((t & 31) == 0) {     // lane 0 of warp 0 and of warp 1: the two data-dependent inversion loops run side by side, not serialised
        int w = t >> 5;
        G1 acc = G1::identity();
        for (int k = 0; k < nparts; k++) acc = acc.add(w == 0 ? parts[k].a : parts[k].b);
        if (w == 1) acc = acc.add(sg.neg());
        Fp zi = vliw::fp_inv_bingcd(acc.z), zi2 = zi.sqr();
        G1Affine a = acc.is_identity() ? G1Affine{Fp::zero(), Fp::zero(), 1} : G1Affine{acc.x * zi2, acc.y * zi2 * zi, 0};
        if (w == 0 && !a.inf) a.y = a.y.neg();     // -A
        pts[w] = a;
    }
    __syncthreads();
    vliw29::Lanes L{t, kFinalThreads, tab, ticks};
    bool ok = vliw29::coop_pairing_product_is_one(regs, pts[1], T->lines29[0], pts[0], T->lines29[1], L);
    L.tick(5);
    if (t == 0) { result[0] = ok ? kTrue : kFalse; result[1] = 0; }
}

}  // namespace kzgb200
