// Small kernels behind the extra entry points of the C ABI: evaluation at a caller-supplied point
// (evaluate_polynomial_in_evaluation_form, reference src/kzg_proof.rs:94-133) and the pre-parsed batch
// (KzgProof::verify_kzg_proof_batch, reference src/kzg_proof.rs:399-444).
#include "common.cuh"

namespace kzgb200 {

// z given by the caller (32 bytes big-endian per blob) instead of by the Fiat-Shamir hash: canonicity check
// (safe_scalar_affine_from_bytes, kzg_proof.rs:27-43), Montgomery form, and the powers z^(2^k), k = 0..12, eval_kernel consumes.
__global__ void z_setup_kernel(const uint8_t* __restrict__ z_be32, int n, Fr* __restrict__ z_mont, ZY* __restrict__ zy, Fr* __restrict__ zpow,
                               uint32_t* __restrict__ status) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint8_t b[32];
    for (int k = 0; k < 32; k++) b[k] = z_be32[(size_t)i * 32 + k];
    Fr raw;
    if (!scalar_from_be32_checked(raw, b)) atomicOr(&status[i], kErrScalar);
    Fr zm = Fr::from_raw(raw);
    z_mont[i] = zm;
    zy[i].z = zm.to_raw();
    Fr s = zm;
    for (int k = 0; k <= 12; k++) { zpow[(size_t)i * 13 + k] = s; s = s * s; }
}

// Inputs in the reference's in-memory layout (little-endian host, build.rs:185-203): G1Affine = 104 bytes = x (6 x u64 Montgomery
// limbs) | y (6 x u64) | infinity flag byte | 7 bytes padding; Scalar = 4 x u64 Montgomery limbs.  The limb images are
// identical to this library's 32-bit-limb Montgomery forms (same radix).  Produces the parsed points, their compressed encodings
// (the transcript hashes to_compressed of every point, kzg_proof.rs:314-333) and z, y in Montgomery / canonical form.
// Threads [0, 2n): points (commitments then proofs); threads [0, n) also convert the scalars.
__global__ void import_parsed_kernel(const uint8_t* __restrict__ c104, const uint8_t* __restrict__ p104, const uint8_t* __restrict__ z32,
                                     const uint8_t* __restrict__ y32, int n, G1Affine* __restrict__ C, G1Affine* __restrict__ P,
                                     uint8_t* __restrict__ c48, uint8_t* __restrict__ p48, Fr* __restrict__ z_mont, ZY* __restrict__ zy) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= 2 * n) return;
    bool is_proof = i >= n;
    int j = is_proof ? i - n : i;
    const uint32_t* src = reinterpret_cast<const uint32_t*>((is_proof ? p104 : c104) + (size_t)j * 104);
    G1Affine a;
    for (int k = 0; k < 12; k++) { a.x.l[k] = __ldg(src + k); a.y.l[k] = __ldg(src + 12 + k); }
    a.inf = (__ldg(src + 24) & 0xffu) ? 1u : 0u;
    (is_proof ? P : C)[j] = a;
    uint8_t enc[48];
    g1_to_compressed(enc, a);
    uint8_t* dst = (is_proof ? p48 : c48) + (size_t)j * 48;
    for (int k = 0; k < 48; k++) dst[k] = enc[k];
    if (!is_proof) {
        Fr zm, ym;
        const uint32_t* zs = reinterpret_cast<const uint32_t*>(z32 + (size_t)j * 32);
        const uint32_t* ys = reinterpret_cast<const uint32_t*>(y32 + (size_t)j * 32);
        for (int k = 0; k < 8; k++) { zm.l[k] = __ldg(zs + k); ym.l[k] = __ldg(ys + k); }
        z_mont[j] = zm;
        zy[j].z = zm.to_raw();
        zy[j].y = ym.to_raw();
    }
}

// Multi-GPU gather: the group leader waits, on its own stream, until every member's msm_combine_kernel has stored its partial
// (flags[k] == epoch; the stores come over NVLink into this GPU's exchange buffer).  Bounded: ~2 s, then *timed_out = 1.
__global__ void wait_flags_kernel(const uint32_t* flags, int count, uint32_t epoch, uint32_t* __restrict__ timed_out) {
    int k = threadIdx.x;
    if (k >= count) return;
    const volatile uint32_t* f = flags + k;
    long long t0 = clock64();
    while (*f != epoch) {
        if (clock64() - t0 > 4000000000ll) { *timed_out = 1; break; }
        __nanosleep(200);
    }
    __threadfence_system();
}

}  // namespace kzgb200
