// K1+K3: canonicity + barycentric evaluation; small export / flag kernels.
#include "common.cuh"

namespace kzgb200 {

// ------------------------------------------------------------------------------------------------ K1+K3
// Barycentric evaluation y = p(z) (reference src/kzg_proof.rs:94-133 with batch_inversion :155-201), fused
// with the canonicity check of Blob::as_polynomial (src/dtypes.rs:48-57).
//
// Inversion-free form.  With S = sum_i f_i / (z - w_i) = N / D over the common denominator
// D = prod (z - w_i) = z^4096 - 1, the reference's value is
//     y = (z^n - 1)/n * sum_i f_i w_i / (z - w_i) = (z * N - (z^n - 1) * sum_i f_i) / n        (w/(z-w) = z/(z-w) - 1)
// and N is built by a binary tree over the bit-reversed domain, where the two halves of a node have
// denominators z^(2^k) -+ w:   N = z^(2^k) (Na + Nb) + w (Na - Nb)   -- one fused dual Montgomery product.
// 4095 dual products per blob instead of ~5*4096 products + an inversion, no branch for z in the domain
// (then D = 0 and the formula collapses to f_k exactly), and the result is the same canonical field element.
// Blob elements stay in normal form: MontMul(aR, f) = a f.
__device__ __forceinline__ Fr fr_merge(const Fr& pw, const Fr& a, const Fr& b, const Fr& w) {
    return Fr::mul_dual_inl(pw, a.add_inl(b), w, a.sub_inl(b));
}

// Work split: thread t owns the 32 consecutive leaves [32t, 32t+32) (binary-counter stack of pending left subtrees in shared
// memory, one 16-byte column per thread and half: conflict-free), then the 128 subtree values are merged by a shrinking set
// of threads (64 merges on two warps, then one warp finishes).  z^(2^k) come from K2.  The next leaf and the next twiddle
// (sequential stream twiddle_po) are fetched one merge ahead.  Per leaf beside the merges: the plain sum of the elements is
// kept unreduced in 9 limbs (one carry chain, reduced once per blob), and the canonicity test is one compare of the top
// word -- it decides for every canonical element but a 2^-31 fraction -- with the exact comparison off the fast path.
__global__ void __launch_bounds__(kEvalThreads, 6) eval_kernel(const uint8_t* __restrict__ blobs, int n, const Fr* __restrict__ zpow,
                                                            const DeviceTables* __restrict__ T, ZY* __restrict__ zy,
                                                            uint32_t* __restrict__ status) {
    __shared__ Fr s_pow[13];              // z^(2^k), Montgomery
    __shared__ uint4 s_stack[5][2][kEvalThreads];
    __shared__ Fr s_n[2][kEvalThreads];
    __shared__ uint32_t s_col[kEvalThreads / 32][9][2];   // per warp: sums of the low / high 16-bit halves of each limb of sum f
    int blob = blockIdx.x, t = threadIdx.x;
    if (blob >= n) return;
    if (t < 13) s_pow[t] = ldg_fr(zpow + (size_t)blob * 13 + t);
    const uint4* base = reinterpret_cast<const uint4*>(blobs + (size_t)blob * kBytesPerBlob) + (size_t)t * kLeavesPerThread * 2;
    const Fr* tw = T->twiddle_po[t];
    Fr nxt = load_fe_be(base), wn = ldg_fr(tw), cur;
    int m = 0;
    uint32_t fs[9];
#pragma unroll
    for (int i = 0; i < 9; i++) fs[i] = 0;
    bool bad = false;
    constexpr uint32_t kQ[8] = KZG_FR_Q;
    __syncthreads();
#pragma unroll 1
    for (int j = 0; j < kLeavesPerThread; j++) {
        cur = nxt;
        if (j + 1 < kLeavesPerThread) nxt = load_fe_be(base + 2 * (j + 1));
        if (cur.l[7] >= kQ[7]) bad |= cur.geq_modulus();
        fs[8] += add_n<8>(fs, fs, cur.l);
        int k = 0;
#pragma unroll 1
        for (; (j >> k) & 1; k++) {
            // merge the pending left subtree of level k with cur (right)
            Fr w = wn, left;
            wn = ldg_fr(tw + ++m);                    // m <= 31: the pad entry
            uint4 a = s_stack[k][0][t], b = s_stack[k][1][t];
            left.l[0] = a.x; left.l[1] = a.y; left.l[2] = a.z; left.l[3] = a.w; left.l[4] = b.x; left.l[5] = b.y; left.l[6] = b.z; left.l[7] = b.w;
            cur = fr_merge(s_pow[k], left, cur, w);
        }
        if (k < 5) {                                   // k = trailing ones of j; j == 31 ends with the finished subtree in cur
            s_stack[k][0][t] = make_uint4(cur.l[0], cur.l[1], cur.l[2], cur.l[3]);
            s_stack[k][1][t] = make_uint4(cur.l[4], cur.l[5], cur.l[6], cur.l[7]);
        }
    }
    s_n[0][t] = cur;
    if (bad) atomicOr(&status[blob], kErrBlob);
#pragma unroll
    for (int i = 0; i < 9; i++) {                      // warp sums of 16-bit halves (< 2^21 each) on the integer reduction unit
        uint32_t lo = __reduce_add_sync(0xffffffffu, fs[i] & 0xffffu), hi = __reduce_add_sync(0xffffffffu, fs[i] >> 16);
        if ((t & 31) == 0) { s_col[t >> 5][i][0] = lo; s_col[t >> 5][i][1] = hi; }
    }
    __syncthreads();
    // levels 5..11: node i of level k merges values 2i, 2i+1 of the level below; its twiddle is twiddle[i]
    if (t >= 64) return;
    s_n[1][t] = fr_merge(s_pow[5], s_n[0][2 * t], s_n[0][2 * t + 1], ldg_fr(T->twiddle + t));
    asm volatile("bar.sync 1, 64;" ::: "memory");
    if (t >= 32) return;
    int src = 1;
#pragma unroll 1
    for (int k = 6; k < 12; k++, src ^= 1) {
        if (t < (1 << (11 - k))) s_n[src ^ 1][t] = fr_merge(s_pow[k], s_n[src][2 * t], s_n[src][2 * t + 1], ldg_fr(T->twiddle + t));
        __syncwarp();
    }
    if (t == 0) {
        // sum f as an 8-limb value: carry-propagate the column sums (total < 2^12 q), then subtract q << k where it fits
        uint32_t acc[9];
        uint64_t c = 0;
#pragma unroll
        for (int i = 0; i < 9; i++) {
            for (int wv = 0; wv < kEvalThreads / 32; wv++) c += (uint64_t)s_col[wv][i][0] + ((uint64_t)s_col[wv][i][1] << 16);
            acc[i] = (uint32_t)c; c >>= 32;
        }
#pragma unroll 1
        for (int k = 11; k >= 0; k--) {
            uint32_t qs[9], d[9];
            qs[0] = kQ[0] << k;
#pragma unroll
            for (int i = 1; i < 8; i++) qs[i] = __funnelshift_l(kQ[i - 1], kQ[i], k);
            qs[8] = k ? kQ[7] >> (32 - k) : 0u;
            uint32_t borrow = sub_n<8>(d, acc, qs);
            d[8] = acc[8] - qs[8] - borrow;
            if ((uint64_t)acc[8] >= (uint64_t)qs[8] + borrow) { for (int i = 0; i < 9; i++) acc[i] = d[i]; }   // acc >= q << k
        }
        Fr fsum;
        for (int i = 0; i < 8; i++) fsum.l[i] = acc[i];
        const uint32_t invn[8] = KZG_FR_INV4096_M;
        Fr zn1 = s_pow[12].sub_inl(Fr::one());                              // z^4096 - 1 (Montgomery)
        Fr num = s_pow[0].mul_inl(s_n[src][0]).sub_inl(zn1.mul_inl(fsum));  // z N - (z^n - 1) sum f   (normal form)
        zy[blob].y = fr_const(invn).mul_inl(num);
    }
}
// z / y as 32-byte big-endian strings for the caller (intermediates are part of the parity contract)
__global__ void export_scalars_kernel(const ZY* __restrict__ zy, int n, uint8_t* __restrict__ z_out, uint8_t* __restrict__ y_out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    if (z_out) limbs_to_be32(z_out + (size_t)i * 32, zy[i].z.l);
    if (y_out) limbs_to_be32(y_out + (size_t)i * 32, zy[i].y.l);
}
__global__ void r_to_raw_kernel(const Fr* __restrict__ r_mont, ZY* __restrict__ out) { out->z = r_mont->to_raw(); out->y = Fr::zero(); }
__global__ void status_or_kernel(const uint32_t* __restrict__ status, int n, uint32_t* __restrict__ out) {
    __shared__ uint32_t s;
    if (threadIdx.x == 0) s = 0;
    __syncthreads();
    uint32_t e = 0;
    for (int i = threadIdx.x; i < n; i += blockDim.x) e |= status[i];
    if (e) atomicOr(&s, e);
    __syncthreads();
    if (threadIdx.x == 0) out[0] = s;
}

}  // namespace kzgb200
