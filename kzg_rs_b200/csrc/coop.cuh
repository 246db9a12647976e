// Warp-cooperative Jacobian point operations (lanes select operands and multiply in lockstep); used by the MSM recombination
// (k_msm.cu) and the single-blob ladder (k_final.cu).
#pragma once
#include "common.cuh"

namespace kzgb200 {

// Lanes of a warp run in lockstep, so the lanes do not branch to different products: every lane SELECTS its two
// operands and all of them execute the one multiplication together.
struct CoopPoint { Fp v[20]; };   // [0..2] = X,Y,Z ; the rest scratch
__device__ __forceinline__ Fp sel(int k, const Fp& a, const Fp& b, const Fp& c, const Fp& d) {
    Fp r;
#pragma unroll
    for (int i = 0; i < 12; i++) r.l[i] = k == 0 ? a.l[i] : (k == 1 ? b.l[i] : (k == 2 ? c.l[i] : d.l[i]));
    return r;
}
__device__ __forceinline__ void coop_dbl(CoopPoint* s, int lane) {
    Fp* v = s->v;
    if (v[2].is_zero()) return;                                   // identity (uniform: every lane reads the same value)
    int k = lane & 3;
    Fp X = v[0], Y = v[1], Z = v[2];
    // level 1: X*X, Y*Y, Y*Z
    Fp r1 = sel(k, X, Y, Y, Y).mul_inl(sel(k, X, Y, Z, Z));
    if (lane < 3) v[3 + lane] = r1;
    __syncwarp();
    Fp A = v[3], B = v[4], YZ = v[5];
    Fp E = A.add_inl(A).add_inl(A), XB = X.add_inl(B);
    // level 2: B*B, (X+B)^2, E*E
    Fp o2 = sel(k, B, XB, E, E);
    Fp r2 = o2.mul_inl(o2);
    if (lane < 3) v[6 + lane] = r2;
    __syncwarp();
    Fp C = v[6], D = v[7].sub_inl(A).sub_inl(C); D = D.add_inl(D);
    Fp X3 = v[8].sub_inl(D).sub_inl(D);
    // level 3: E*(D - X3)
    Fp c8 = C.add_inl(C); c8 = c8.add_inl(c8); c8 = c8.add_inl(c8);
    Fp y3 = E.mul_inl(D.sub_inl(X3)).sub_inl(c8);
    __syncwarp();                                                  // everyone has read the old state
    if (lane == 0) { v[0] = X3; v[1] = y3; v[2] = YZ.add_inl(YZ); }
    __syncwarp();
}
// v[0..2] += q (Jacobian)
__device__ __forceinline__ void coop_add(CoopPoint* s, const G1& q, int lane) {
    Fp* v = s->v;
    if (q.is_identity()) return;
    const bool acc_is_identity = v[2].is_zero();                  // uniform; every lane has read Z before lane 0 may overwrite it
    __syncwarp();
    if (acc_is_identity) { if (lane == 0) { v[0] = q.x; v[1] = q.y; v[2] = q.z; } __syncwarp(); return; }
    int k = lane & 3;
    Fp X1 = v[0], Y1 = v[1], Z1 = v[2];
    // level 1: Z1*Z1, Z2*Z2, Z1*Z2
    Fp r = sel(k, Z1, q.z, Z1, Z1).mul_inl(sel(k, Z1, q.z, q.z, q.z));
    if (lane < 3) v[3 + lane] = r;
    __syncwarp();
    Fp Z1Z1 = v[3], Z2Z2 = v[4], Z1Z2 = v[5];
    // level 2: U1 = X1 Z2Z2, U2 = X2 Z1Z1, Y1 Z2Z2, Y2 Z1Z1
    r = sel(k, X1, q.x, Y1, q.y).mul_inl(sel(k, Z2Z2, Z1Z1, Z2Z2, Z1Z1));
    if (lane < 4) v[6 + lane] = r;
    __syncwarp();
    Fp U1 = v[6], H = v[7].sub_inl(U1);
    // level 3: S1 = (Y1 Z2Z2) Z2, S2 = (Y2 Z1Z1) Z1, H*H, Z1Z2*H
    r = sel(k, v[8], v[9], H, Z1Z2).mul_inl(sel(k, q.z, Z1, H, H));
    if (lane < 4) v[10 + lane] = r;
    __syncwarp();
    Fp S1 = v[10], R = v[11].sub_inl(S1), H2 = v[12], Z3 = v[13];
    if (H.is_zero()) {                                             // same x: doubling or cancellation (uniform)
        if (R.is_zero()) { coop_dbl(s, lane); return; }
        if (lane == 0) { v[0] = Fp::one(); v[1] = Fp::one(); v[2] = Fp::zero(); }
        __syncwarp();
        return;
    }
    // level 4: R*R, H2*H, U1*H2
    r = sel(k, R, H2, U1, U1).mul_inl(sel(k, R, H, H2, H2));
    if (lane < 3) v[14 + lane] = r;
    __syncwarp();
    Fp H3 = v[15], UH2 = v[16];
    Fp X3 = v[14].sub_inl(H3).sub_inl(UH2).sub_inl(UH2);
    // level 5: R*(UH2 - X3), S1*H3
    r = sel(k, R, S1, S1, S1).mul_inl(sel(k, UH2.sub_inl(X3), H3, H3, H3));
    if (lane < 2) v[17 + lane] = r;
    __syncwarp();
    if (lane == 0) { v[0] = X3; v[1] = v[17].sub_inl(v[18]); v[2] = Z3; }
    __syncwarp();
}

// [s]G from the fixed-base table: thread t < 64 takes the t-th 4-bit digit of s, then a shared-memory tree sum.
// All kFinalThreads threads call it; the result is returned to every thread.
__device__ __noinline__ inline G1 coop_fixed_base_mul(const Fr& s_raw, const DeviceTables* T, G1* sm /* kFinalThreads */) {
    int t = threadIdx.x;
    uint32_t d = (s_raw.l[t / 8] >> (4 * (t % 8))) & 15u;
    sm[t] = d ? G1::from_affine(T->gen_table[t][d - 1]) : G1::identity();
    __syncthreads();
    for (int span = kFinalThreads / 2; span >= 1; span >>= 1) {
        if (t < span) sm[t] = sm[t].add(sm[t + span]);
        __syncthreads();
    }
    return sm[0];
}

}  // namespace kzgb200
