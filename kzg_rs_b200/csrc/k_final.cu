// K7: final pairing check of a batch: G1 prelude kernel + the cooperative pairing engine kernel.
#include "common.cuh"
#include "coop.cuh"

namespace kzgb200 {

// Prelude of the final check over the gathered per-rank partials (reference src/kzg_proof.rs:436-441):
//   e(sum_k A_k, [tau]G2) == e(sum_k B_k - [sum_k s_k]G, G2)
// one CTA of kFinalThreads threads forms the two affine G1 arguments (-A, B - [s]G) and hands them to pairing_check_kernel.
// result: 0 = false, 1 = true, 2 = BadArgs (some rank flagged an unparsable input)
__global__ void __launch_bounds__(kFinalThreads) batch_final_kernel(const Partial* __restrict__ parts, int nparts, const DeviceTables* __restrict__ T,
                                                                    uint32_t* __restrict__ result, FinalPts* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    FinalSmem& S = *reinterpret_cast<FinalSmem*>(dyn_smem);
    G1* sm = S.sm;
    __shared__ Fr s_sum;
    __shared__ uint32_t s_err;
    int t = threadIdx.x;
    if (t == 0) result[2] = 0;
    if (t == 0) {
        Fr s = Fr::zero(); uint32_t err = 0;
        for (int k = 0; k < nparts; k++) { s = s.add_inl(parts[k].ry); err |= parts[k].err; }
        s_sum = s; s_err = err;
    }
    __syncthreads();
    if (s_err) { if (t == 0) { result[0] = kBadArgs; result[1] = s_err; out->go = 0; } return; }
    G1 sg = G1::identity();
    if (!s_sum.is_zero()) sg = coop_fixed_base_mul(s_sum, T, sm);      // zero on the batch path: msm_combine_kernel folded - [s]G into b
    if ((t & 31) == 0) {     // lane 0 of warp 0 and of warp 1: the two data-dependent inversion loops run side by side, not serialised
        int w = t >> 5;
        G1 acc = G1::identity();
        for (int k = 0; k < nparts; k++) acc = acc.add(w == 0 ? parts[k].a : parts[k].b);
        if (w == 1) acc = acc.add(sg.neg());
        Fp zi = vliw::fp_inv_bingcd(acc.z), zi2 = zi.sqr();
        G1Affine a = acc.is_identity() ? G1Affine{Fp::zero(), Fp::zero(), 1} : G1Affine{acc.x * zi2, acc.y * zi2 * zi, 0};
        if (w == 0 && !a.inf) a.y = a.y.neg();     // -A
        out->pts[w] = a;
        if (w == 0) out->go = 1;
    }
}

// e(pts[1], G2) e(pts[0], [tau]G2) == 1 on the cooperative engine (vliw29.cuh): kPairThreads threads = 24 warps.
__global__ void __launch_bounds__(kPairThreads) pairing_check_kernel(const FinalPts* __restrict__ in, const DeviceTables* __restrict__ T,
                                                                     uint32_t* __restrict__ result, long long* __restrict__ ticks) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    PairSmem& S = *reinterpret_cast<PairSmem*>(dyn_smem);
    const int t = threadIdx.x;
    if (in->go == 0) return;            // the prelude already wrote the result
    if (ticks && t == 0) { ticks[0] = clock64(); for (int i = 6; i < 14; i++) ticks[i] = 0; }
    vliw29::Tables tab = vliw29::load_tables(&S.stab, t, kPairThreads);
    vliw29::Lanes L{t, kPairThreads, tab, ticks, S.stab.p29};
    const G1Affine p0 = in->pts[0], p1 = in->pts[1];
    bool ok = vliw29::coop_pairing_product_is_one(S.regs, p1, T->lines29[0], p0, T->lines29[1], L);
    L.tick(5);
    if (t == 0) { result[0] = ok ? kTrue : kFalse; result[1] = 0; }
}

// Self-test of the 16-lane executors against the sequential reference executors (kzgb200_debug_engine_selftest): the same
// programs on two register files seeded alike; after every program all registers are compared as canonical field elements
// (the two forms may leave different representatives of the same residue).
__global__ void __launch_bounds__(kPairThreads) engine_selftest_kernel(uint32_t seed, int rounds, f29::F29* __restrict__ ref, uint32_t* __restrict__ mismatches) {
    extern __shared__ __align__(16) unsigned char dyn_smem[];
    PairSmem& S = *reinterpret_cast<PairSmem*>(dyn_smem);
    const int t = threadIdx.x;
    vliw29::Tables tab = vliw29::load_tables(&S.stab, t, kPairThreads);
    vliw29::Lanes L{t, kPairThreads, tab, nullptr, S.stab.p29};
    for (int i = t; i < vliw29::kTotalRegs; i += kPairThreads) {
        Fp x;
        uint32_t s = seed * 2654435761u + (uint32_t)i * 40503u + 1u;
        for (int k = 0; k < 12; k++) { s = s * 1664525u + 1013904223u; x.l[k] = s; }
        x.l[11] &= 0x0fffffffu;                        // < p
        f29::F29 v = f29::from_fp(x);
        if (i >= vliw29::kRegConst && i < vliw29::kRegConst + 10) v = vliw29::frob_const(i - vliw29::kRegConst);
        S.regs[i] = v; ref[i] = v;
    }
    __syncthreads();
    for (int r = 0; r < rounds; r++)
        for (int prog = 0; prog < vliw29::kNumPrograms; prog++) {
            vliw29::run(prog, S.regs, L);
            if (t == 32) vliw29::run_ref(prog, ref, vliw29::default_tables());
            __syncthreads();
            for (int i = t; i < vliw29::kTotalRegs; i += kPairThreads)
                if (!(vliw29::canonical_signed(S.regs[i], 2048) == vliw29::canonical_signed(ref[i], 2048))) atomicAdd(&mismatches[prog], 1u);
            __syncthreads();
        }
}

}  // namespace kzgb200
