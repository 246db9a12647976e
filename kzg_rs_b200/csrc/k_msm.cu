// K6: random linear combination (Pippenger MSM with GLV).
#include "common.cuh"
#include "coop.cuh"

namespace kzgb200 {

// ------------------------------------------------------------------------------------------------ K6
// Random linear combination (reference src/kzg_proof.rs:399-433), regrouped so that it needs no per-blob
// [y_i]G:   A = sum r_i pi_i ,  B = sum (r_i C_i + (r_i z_i) pi_i) - [sum r_i y_i] G ,  r_i = r^(offset+i).
// v1: one thread per blob does its scalar multiplications (Shamir's trick for the pair), results are then
// tree-summed by pair_sum_kernel.
// Pippenger bucket method with the GLV split: every scalar k = k1 + k2 x^2 (glv.cuh), so a point P contributes
// [k1]P + [k2](-phi(P)) with two 128-bit halves -> 16 windows of 8 bits x 255 buckets.  Three point/scalar sets per rank:
//   set 0: pi_i with r_i  (-> A),   set 1: C_i with r_i,   set 2: pi_i with r_i z_i   (sets 1+2 -> B').
// No sorting network and no atomics on points: a counting sort of the digits per (scalar kind, half, window) gives
// every bucket two contiguous index lists (low halves -> P_i, high halves -> -phi(P_i)); threads own buckets.
__global__ void __launch_bounds__(128) msm_scalars_kernel(const Fr* __restrict__ z_mont, const ZY* __restrict__ zy, const Fr* __restrict__ r_mont,
                                                          uint64_t offset, int n, uint8_t* __restrict__ digits /* [4*16][n] */,
                                                          Fr* __restrict__ ry) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint64_t e = offset + (uint64_t)i;
    uint32_t ee[2] = {(uint32_t)e, (uint32_t)(e >> 32)};
    Fr ri = r_mont->pow(ee, 64);              // r^(offset+i), Montgomery   (compute_powers, kzg_proof.rs:279-289)
    Fr ri_raw = ri.to_raw();
    Fr rz_raw = (ri * z_mont[i]).to_raw();    // r_i z_i   (kzg_proof.rs:425)
    ry[i] = ri * zy[i].y;                     // r_i y_i in normal form
    uint32_t h[2][2][4];
    glv_split(ri_raw.l, h[0][0], h[0][1]);
    glv_split(rz_raw.l, h[1][0], h[1][1]);
    for (int kind = 0; kind < 2; kind++)
        for (int half = 0; half < 2; half++)
            for (int w = 0; w < kWindows; w++)
                digits[((size_t)(kind * 2 + half) * kWindows + w) * n + i] = (uint8_t)(h[kind][half][w >> 2] >> (8 * (w & 3)));
}
// counting sort of one digit row: grid = kDigitRows; order[row][*] = blob indices grouped by digit,
// start[row][b] = first position of digit b (start[row][256] = n)
__global__ void __launch_bounds__(256) msm_sort_kernel(const uint8_t* __restrict__ digits, int n, uint32_t* __restrict__ order,
                                                       uint32_t* __restrict__ start) {
    __shared__ uint32_t hist[kBuckets], cursor[kBuckets];
    int row_id = blockIdx.x, t = threadIdx.x;
    const uint8_t* row = digits + (size_t)row_id * n;
    uint32_t* ord = order + (size_t)row_id * n;
    uint32_t* st = start + (size_t)row_id * (kBuckets + 1);
    hist[t] = 0;
    __syncthreads();
    for (int i = t; i < n; i += blockDim.x) atomicAdd(&hist[row[i]], 1u);
    __syncthreads();
    if (t == 0) {
        uint32_t acc = 0;
        for (int b = 0; b < kBuckets; b++) { cursor[b] = acc; st[b] = acc; acc += hist[b]; }
        st[kBuckets] = acc;
    }
    __syncthreads();
    for (int i = t; i < n; i += blockDim.x) ord[atomicAdd(&cursor[row[i]], 1u)] = (uint32_t)i;
}
// Bucket accumulation, balanced.  A "row" = (point set, GLV half, window): its sorted index list (n entries, grouped by digit) is
// cut into slices of kSlice consecutive entries and every thread sums exactly one slice -- round 1 gave whole bucket lists to
// threads (Poisson-sized: a warp ran at the pace of its longest lane, +40 %).  A slice spans one or a few digit runs: a run that
// lies entirely inside the slice is a finished half-bucket and goes to `halfsum`; the slice's first / last run may continue in
// the neighbouring slices and goes to `part` slot 0 / 1 (a slice that is one single run: slot 0).  msm_bucket_join_kernel adds
// up the few partials per bucket and the two GLV halves.  The next entry's index and point are fetched one addition ahead.
__device__ __forceinline__ G1Affine ldg_point(const G1Affine* p) {
    const uint32_t* w = reinterpret_cast<const uint32_t*>(p);
    G1Affine a;
#pragma unroll
    for (int k = 0; k < 12; k++) { a.x.l[k] = __ldg(w + k); a.y.l[k] = __ldg(w + 12 + k); }
    a.inf = __ldg(w + 24);
    return a;
}
__global__ void __launch_bounds__(128) msm_bucket_kernel(const G1Affine* __restrict__ C, const G1Affine* __restrict__ P, int n,
                                                         const uint32_t* __restrict__ order, const uint32_t* __restrict__ start,
                                                         G1* __restrict__ halfsum /* [96][256] */, G1* __restrict__ part /* [96][slices][2] */) {
    const int slices = (n + kSlice - 1) / kSlice;
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= kMsmRows * slices) return;
    int row = tid / slices, s = tid % slices;                 // row = (set * 2 + half) * kWindows + w
    int w = row % kWindows, half = (row / kWindows) & 1, set = row / (2 * kWindows);
    const G1Affine* pts = set == 1 ? C : P;
    int digit_row = ((set == 2 ? 1 : 0) * 2 + half) * kWindows + w;
    const uint32_t* ord = order + (size_t)digit_row * n;
    const uint32_t* st = start + (size_t)digit_row * (kBuckets + 1);
    const uint32_t pos0 = (uint32_t)s * kSlice, end0 = pos0 + kSlice < (uint32_t)n ? pos0 + kSlice : (uint32_t)n;
    // digit of the first entry: the largest b with st[b] <= pos0
    int b = 0;
#pragma unroll
    for (int step = 128; step >= 1; step >>= 1) if (b + step <= 255 && __ldg(st + b + step) <= pos0) b += step;
    uint32_t next_boundary = __ldg(st + b + 1);
    G1 acc = G1::identity();
    bool first = true;
    auto flush = [&](int bb) {
        uint32_t lo = __ldg(st + bb), hi = __ldg(st + bb + 1);
        if (bb != 0 && hi > lo) {
            if (lo >= pos0 && hi <= end0) halfsum[(size_t)row * kBuckets + bb] = acc;
            else part[((size_t)row * slices + s) * 2 + (first ? 0 : 1)] = acc;
        }
        first = false;
        acc = G1::identity();
    };
    const uint32_t beta_l[12] = KZG_FP_BETA_M;
    const Fp beta = fp_const(beta_l);
    G1Affine nxt = ldg_point(pts + __ldg(ord + pos0));
    for (uint32_t k = pos0; k < end0; k++) {
        G1Affine q = nxt;
        if (k + 1 < end0) nxt = ldg_point(pts + __ldg(ord + k + 1));
        while (k >= next_boundary) { flush(b); b++; next_boundary = __ldg(st + b + 1); }
        if (b != 0) {
            if (half && !q.inf) { q.x = q.x * beta; q.y = q.y.neg(); }     // -phi(P) = (beta x, -y)
            acc = acc.add_mixed(q);
        }
    }
    flush(b);
}
// one thread per (set, window, bucket): the bucket's partial sums of both GLV halves -> buckets[set][window][bucket]
__global__ void __launch_bounds__(128) msm_bucket_join_kernel(int n, const uint32_t* __restrict__ start, const G1* __restrict__ halfsum,
                                                              const G1* __restrict__ part, G1* __restrict__ buckets /* [3][16][256] */) {
    const int slices = (n + kSlice - 1) / kSlice;
    int tid = blockIdx.x * blockDim.x + threadIdx.x;
    if (tid >= kMsmSets * kWindows * kBuckets) return;
    int b = tid % kBuckets, w = (tid / kBuckets) % kWindows, set = tid / (kBuckets * kWindows);
    G1 acc = G1::identity();
    if (b != 0) {
        for (int half = 0; half < 2; half++) {
            int row = (set * 2 + half) * kWindows + w, digit_row = ((set == 2 ? 1 : 0) * 2 + half) * kWindows + w;
            const uint32_t* st = start + (size_t)digit_row * (kBuckets + 1);
            uint32_t lo = __ldg(st + b), hi = __ldg(st + b + 1);
            if (hi <= lo) continue;
            uint32_t s0 = lo / kSlice, s1 = (hi - 1) / kSlice;
            if (s0 == s1) { acc = acc.add(halfsum[(size_t)row * kBuckets + b]); continue; }
            for (uint32_t s = s0; s <= s1; s++) {
                int slot = (s == s0 && lo != s0 * kSlice) ? 1 : 0;     // the bucket opens inside slice s0: that slice's last run
                acc = acc.add(part[((size_t)row * slices + s) * 2 + slot]);
            }
        }
    }
    buckets[tid] = acc;
}
// two warps per (set, window): W = sum_b b * bucket[b].  Lane l owns buckets 4l .. 4l+3 (running-sum trick inside
// the segment, then the segment's offset 8l by a short double-and-add), then a shared-memory tree over the lanes.
__global__ void __launch_bounds__(kWinLanes) msm_window_kernel(const G1* __restrict__ buckets, G1* __restrict__ windows /* [3][16] */) {
    __shared__ G1 sm[kWinLanes];
    int l = threadIdx.x, sw = blockIdx.x;          // sw = set * kWindows + window
    const G1* bk = buckets + (size_t)sw * kBuckets + kWinPer * l;
    G1 run = G1::identity(), acc = G1::identity();
    for (int j = kWinPer - 1; j >= 0; j--) {
        run = run.add(bk[j]);                      // bucket 0 holds the identity
        acc = acc.add(run);                        // after the loop: acc = sum_j (j+1) bk[j], run = sum_j bk[j]
    }
    // sum_j (kWinPer l + j) bk[j] = acc + (kWinPer l - 1) run
    uint32_t k[1] = {(uint32_t)(kWinPer * l)};
    G1 off = scalar_mul(run, k, 8);
    sm[l] = acc.add(off).add(run.neg());
    __syncthreads();
    for (int span = kWinLanes / 2; span >= 1; span >>= 1) {
        if (l < span) sm[l] = sm[l].add(sm[l + span]);
        __syncthreads();
    }
    if (l == 0) windows[sw] = sm[0];
}
// window sums -> A = sum_w 256^w W[0][w], B' = sum_w 256^w (W[1][w] + W[2][w]) (Horner, 8 doublings per window):
// warp 0 does A, warp 1 does B' with the cooperative point operations above; the other warps OR the per-blob error
// flags, tree-sum s = sum r_i y_i and then -- still in the shadow of the Horner chains -- form [s]G from the fixed-base table
// (64 lookups, a 6-level tree); the partial carries B' - [s]G and ry = 0, so the final check has no fixed-base work left
// (round 2: 0.1 ms off the serial tail; with several ranks each folds its own [s_k]G).
// `out` may be a peer-mapped pointer into the group leader's exchange buffer (multi-GPU: the partial-sum gather is this kernel's
// last store, over NVLink); then `flag` (same buffer) receives `epoch` after the partial, with a system-scope fence in between.
__global__ void __launch_bounds__(256) msm_combine_kernel(const G1* __restrict__ windows, const Fr* __restrict__ ry, const uint32_t* __restrict__ status,
                                                          int n, const DeviceTables* __restrict__ T, Partial* __restrict__ out, uint32_t* flag, uint32_t epoch) {
    __shared__ uint32_t s_err;
    __shared__ Fr s_ry[256];
    __shared__ CoopPoint cp[3];
    __shared__ G1 s_sg[64];
    int t = threadIdx.x, warp = t >> 5, lane = t & 31;
    if (t == 0) s_err = 0;
    __syncthreads();
    if (warp < kMsmSets) {
        // warps 0..2: Horner recombination of one point set each (15 x 8 doublings + 16 additions, the serial part of the tail)
        CoopPoint* s = &cp[warp];
        if (lane == 0) { s->v[0] = Fp::one(); s->v[1] = Fp::one(); s->v[2] = Fp::zero(); }
        __syncwarp();
        for (int w = kWindows - 1; w >= 0; w--) {
            if (w != kWindows - 1) for (int k = 0; k < 8; k++) coop_dbl(s, lane);
            coop_add(s, windows[warp * kWindows + w], lane);
        }
    } else {
        // warps 3..7, beside the recombination: OR of the per-blob error flags and sum r_i y_i
        constexpr int kHelpers = 256 - 32 * kMsmSets;
        int h = t - 32 * kMsmSets;
        uint32_t e = 0;
        Fr acc_ry = Fr::zero();
        for (int i = h; i < n; i += kHelpers) { e |= status[i]; acc_ry = acc_ry.add_inl(ry[i]); }
        if (e) atomicOr(&s_err, e);
        s_ry[h] = acc_ry;
        asm volatile("bar.sync 1, %0;" ::"n"(kHelpers) : "memory");
        for (int span = 128; span >= 1; span >>= 1) {
            if (h < span && h + span < kHelpers) s_ry[h] = s_ry[h].add_inl(s_ry[h + span]);
            asm volatile("bar.sync 1, %0;" ::"n"(kHelpers) : "memory");
        }
        // [s]G: helper h < 64 takes the h-th 4-bit digit of s (normal form), then a tree over the 64 table points
        if (h < 64) {
            const Fr sv = s_ry[0];
            const uint32_t d = (sv.l[h / 8] >> (4 * (h % 8))) & 15u;
            s_sg[h] = d ? G1::from_affine(T->gen_table[h][d - 1]) : G1::identity();
        }
        asm volatile("bar.sync 1, %0;" ::"n"(kHelpers) : "memory");
        for (int span = 32; span >= 1; span >>= 1) {
            if (h < span) s_sg[h] = s_sg[h].add(s_sg[h + span]);
            asm volatile("bar.sync 1, %0;" ::"n"(kHelpers) : "memory");
        }
    }
    __syncthreads();
    if (warp == 1) {          // B' = set 1 + set 2 - [s]G
        G1 q = {cp[2].v[0], cp[2].v[1], cp[2].v[2]};
        coop_add(&cp[1], q, lane);
        coop_add(&cp[1], s_sg[0].neg(), lane);
    }
    if (warp < 2 && lane == 0) { G1 r = {cp[warp].v[0], cp[warp].v[1], cp[warp].v[2]}; if (warp == 1) out->b = r; else out->a = r; }
    if (t == 0) { out->ry = Fr::zero(); out->err = s_err; }
    if (flag) {
        __threadfence_system();
        __syncthreads();
        if (t == 0) { *reinterpret_cast<volatile uint32_t*>(flag) = epoch; __threadfence_system(); }
    }
}

}  // namespace kzgb200
