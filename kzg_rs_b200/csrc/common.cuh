// Shared declarations of the sm_100a kernels of the kzg-rs verification hot path (K1..K8 of SURVEY.md section 2.1):
// constants, device-resident structures, small load helpers and the prototypes of every kernel.  The kernels live in
// k_*.cu (one translation unit per group so that they compile in parallel); kzgb200.cu is the host runtime.
#pragma once
#include <cuda_runtime.h>
#include "sha256.cuh"
#include "verify.cuh"
#include "vliw.cuh"
#include "vliw29.cuh"
#include "glv.cuh"

namespace kzgb200 {

constexpr int kFieldElementsPerBlob = 4096;       // reference src/consts.rs:7
constexpr int kBytesPerBlob = 4096 * 32;          // src/consts.rs:8
constexpr uint32_t kErrBlob = 1, kErrCommitment = 2, kErrProof = 4, kErrScalar = 8;
// canonical (non-Montgomery) z_i, y_i; the memory image is 2 x 32 little-endian bytes, which is exactly what
// the batch transcript hashes (reference src/kzg_proof.rs:320-328) and what ranks exchange
struct ZY { Fr z, y; };

// Device-resident trusted-setup tables (K8; replaces KzgSettings::load_trusted_setup_file,
// reference src/trusted_setup.rs:94-98 and build.rs:131-170).
struct DeviceTables {
    // twiddle[g] = roots_of_unity[2g] = omega^bitrev12(2g), Montgomery form.  In the bit-reversed domain the
    // group of 2^(k+1) consecutive points starting at index a has prod (z - w_i) = z^(2^(k+1)) - w_a^(2^(k+1)),
    // and w_a^(2^k) = roots_of_unity[2g] for g = a >> (k+1), independent of the level k.
    Fr twiddle[2048];
    // the same twiddles in the order thread t of eval_kernel consumes them (post-order over its 32-leaf subtree): 31 per
    // thread + 1 pad, so the stream is sequential and the next one can be fetched while the current merge runs
    Fr twiddle_po[128][32];
    PairingTables pairing;
    // the same line triples in the representation of the cooperative pairing engine (vliw29.cuh): [0] = G2 generator, [1] = [tau]G2
    vliw29::LineCoeffs29 lines29[2][kMillerSteps];
    // fixed-base table of the G1 generator: gen_table[w][d-1] = [d * 16^w] G, d = 1..15, w < 64
    G1Affine gen_table[64][15];
    uint32_t setup_ok;
};

// ---- launch shapes shared by kernels and host ---------------------------------------------------------------
#ifndef KSHA_THREADS
#define KSHA_THREADS 32
#endif
constexpr int kShaThreads = KSHA_THREADS;     // blobs (threads) per CTA of K2
constexpr int kEvalThreads = 128;
constexpr int kLeavesPerThread = kFieldElementsPerBlob / kEvalThreads;  // 32
constexpr int kTreeGroup = 16;      // transcript entries (160 B each) per leaf hash: 40 compressions
constexpr int kTreeMid = 32;        // leaf digests per middle-level hash: 17 compressions
constexpr int kWindows = 16, kBuckets = 256, kMsmSets = 3, kDigitRows = 4 * kWindows;   // rows: (kind r|rz) x (half lo|hi) x window
constexpr int kMinSlice = 8, kMsmRows = kMsmSets * 2 * kWindows;   // bucket accumulation: fewest entries per thread (msm_slice_len); rows = (set, GLV half, window)
constexpr int kWinLanes = kBuckets;                                 // window sums: one thread per bucket (suffix scan + tree in shared memory)
constexpr int kWinSmemBytes = kWinLanes * 144;                      // one Jacobian point per thread
constexpr int kCombineThreads = 544;    // msm_combine_kernel: 384 engine threads (three Horner chains in lockstep) + 160 helpers
constexpr int kCombineSmemBytes = 48 * 1024;
constexpr int kFinalThreads = 64;       // G1 prelude of the final check (sums, [s]G, affine conversion)
constexpr int kPairThreads = 768;       // pairing engine: 24 warps = 24 products (one per warp) or 48 sums (16 lanes each) at a time
constexpr int kManyThreads = 512, kManyGroups = 28;   // many_pairing_kernel: checks run in lockstep by one CTA (28 x 36 products = 1.97 x 512; 128 registers)
constexpr int kManyCtasPerSm = 1;                     // (two CTAs of 14 checks per SM, not in lockstep with each other: 491 k instead of 555 k checks/s)
constexpr int kManyWarps = 4;         // many_pairing_warp_kernel (identity inputs): one check per warp
constexpr int kHarnessMaxDegree = 16;
constexpr int kLagWindows = 32, kLagEntries = 255;

__device__ __forceinline__ Fr load_fe_be(const uint4* p) {
    uint4 hi = __ldg(p), lo = __ldg(p + 1);   // 32 big-endian bytes: hi holds the most significant 16
    Fr f;
    f.l[7] = sha_bswap(hi.x); f.l[6] = sha_bswap(hi.y); f.l[5] = sha_bswap(hi.z); f.l[4] = sha_bswap(hi.w);
    f.l[3] = sha_bswap(lo.x); f.l[2] = sha_bswap(lo.y); f.l[1] = sha_bswap(lo.z); f.l[0] = sha_bswap(lo.w);
    return f;
}
__device__ __forceinline__ Fr ldg_fr(const Fr* p) {
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 a = __ldg(q), b = __ldg(q + 1);
    Fr f;
    f.l[0] = a.x; f.l[1] = a.y; f.l[2] = a.z; f.l[3] = a.w; f.l[4] = b.x; f.l[5] = b.y; f.l[6] = b.z; f.l[7] = b.w;
    return f;
}

// per-rank partial result exchanged between ranks (the payload of the allgather)
struct Partial {
    G1 a, b;          // sum r_i pi_i ; sum (r_i C_i + r_i z_i pi_i) - [sum r_i y_i]G (the fixed-base term is folded in by msm_combine_kernel)
    Fr ry;            // a scalar s still to be applied as - [s]G by the final check (normal form); zero since the term is folded into b
    uint32_t err;     // OR of the per-blob error flags of this rank
    uint32_t pad[7];
};

// dynamic shared memory of the final kernels: engine register file, program tables, G1 tree scratch (> 48 KB: opt-in)
struct FinalSmem {                       // prelude kernels (padded: see kTailPadSmem in runtime.cuh)
    G1 sm[kFinalThreads];
    unsigned char pad[20 * 1024];
};
struct PairSmem {                        // pairing_check_kernel
    f29::F29 regs[vliw29::kTotalRegs];
    vliw29::SharedTables stab;
};
// hand-over from the prelude kernel to the pairing kernel (device memory, 256 bytes into the context's 512-byte scratch)
struct FinalPts { G1Affine pts[2]; uint32_t go; };

constexpr int kManyStride = vliw::kTotalRegsThr * 12 + 1;     // words between the register files of consecutive groups (odd: bank skew)
constexpr int kManySmemBytes = kManyGroups * kManyStride * 4 + (int)sizeof(vliw::SharedTables) + vliw::kNumSharedRegs * (int)sizeof(Fp) +
                               kManyGroups * (2 * (int)sizeof(G1Affine) + 2) + 64;
// ---- kernels (k_*.cu) ----------------------------------------------------------------------------------------
__global__ void setup_tables_kernel(DeviceTables* T, const uint8_t* g2_points);
__global__ void setup_lines29_kernel(DeviceTables* T);
constexpr int kChallengeSlabs = 8;   // slab-wise arrival of the last host chunk (k_challenge.cu)
void launch_challenge(int stages, cudaStream_t st, const uint8_t* blobs, const uint8_t* commitments, int n, Fr* z_mont, ZY* zy, Fr* zpow,
                      const uint8_t* slab_flags = nullptr);   // K2
__global__ void eval_kernel(const uint8_t* __restrict__ blobs, int n, const Fr* __restrict__ zpow, const DeviceTables* __restrict__ T,
                            ZY* __restrict__ zy, uint32_t* __restrict__ status);
__global__ void g1_decompress_kernel(const uint8_t* __restrict__ commitments, const uint8_t* __restrict__ proofs, int n, G1Affine* __restrict__ C,
                                     G1Affine* __restrict__ P, uint32_t* __restrict__ status, bool with_subgroup_check);
__global__ void g1_subgroup_kernel(const G1Affine* __restrict__ C, const G1Affine* __restrict__ P, int n, uint32_t* __restrict__ status);
__global__ void transcript_schedule_kernel(const uint8_t* __restrict__ commitments, const ZY* __restrict__ zy, const uint8_t* __restrict__ proofs,
                                           uint64_t n, uint32_t* __restrict__ wk, uint64_t first_blk, uint64_t blk_count);
__global__ void transcript_chain_kernel(const uint32_t* __restrict__ wk_all, uint64_t n, Fr* __restrict__ r_mont, uint32_t* __restrict__ state,
                                        uint64_t first_blk, uint64_t blk_count);
__global__ void transcript_words_kernel(const uint8_t* __restrict__ commitments, const ZY* __restrict__ zy, const uint8_t* __restrict__ proofs,
                                        uint64_t first_entry, uint64_t entry_count, uint32_t* __restrict__ words);
__global__ void transcript_tree_leaf_words_kernel(const uint32_t* __restrict__ words, uint64_t n, uint32_t* __restrict__ digests,
                                                  uint64_t first_group, uint64_t group_count);
__global__ void r_from_digest_kernel(const uint32_t* __restrict__ digest, Fr* __restrict__ r_mont);
__global__ void z_setup_kernel(const uint8_t* __restrict__ z_be32, int n, Fr* __restrict__ z_mont, ZY* __restrict__ zy, Fr* __restrict__ zpow,
                               uint32_t* __restrict__ status);
__global__ void import_parsed_kernel(const uint8_t* __restrict__ c104, const uint8_t* __restrict__ p104, const uint8_t* __restrict__ z32,
                                     const uint8_t* __restrict__ y32, int n, G1Affine* __restrict__ C, G1Affine* __restrict__ P,
                                     uint8_t* __restrict__ c48, uint8_t* __restrict__ p48, Fr* __restrict__ z_mont, ZY* __restrict__ zy);
__global__ void msm_scalars_kernel(const Fr* __restrict__ z_mont, const ZY* __restrict__ zy, const Fr* __restrict__ r_mont, uint64_t offset, int n,
                                   uint8_t* __restrict__ digits, Fr* __restrict__ ry);
__global__ void msm_sort_kernel(const uint8_t* __restrict__ digits, int n, uint32_t* __restrict__ order, uint32_t* __restrict__ start);
void launch_msm_bucket(int ctas_per_sm, int n, int slice, cudaStream_t st, const G1Affine* C, const G1Affine* P, const uint32_t* order,
                       const uint32_t* start, G1* halfsum, G1* part);
void launch_msm_bucket_join(int lanes, int n, int slice, cudaStream_t st, const uint32_t* start, const G1* halfsum, const G1* part, G1* buckets);
__global__ void msm_window_kernel(const G1* __restrict__ buckets, G1* __restrict__ windows);
__global__ void msm_combine_kernel(const G1* __restrict__ windows, const Fr* __restrict__ ry, const uint32_t* __restrict__ status, int n,
                                   const DeviceTables* __restrict__ T, Partial* __restrict__ out, uint32_t* flag, uint32_t epoch);
__global__ void wait_flags_kernel(const uint32_t* flags, int count, uint32_t epoch, uint32_t* __restrict__ timed_out);
__global__ void batch_final_kernel(const Partial* __restrict__ parts, int nparts, const DeviceTables* __restrict__ T, uint32_t* __restrict__ result,
                                   FinalPts* __restrict__ out);
__global__ void single_final_kernel(const G1Affine* __restrict__ C, const G1Affine* __restrict__ P, const ZY* __restrict__ zy,
                                    const uint32_t* __restrict__ status, const DeviceTables* __restrict__ T, uint32_t* __restrict__ result,
                                    FinalPts* __restrict__ out);
__global__ void pairing_check_kernel(const FinalPts* __restrict__ in, const DeviceTables* __restrict__ T, uint32_t* __restrict__ result,
                                     long long* __restrict__ ticks);
__global__ void engine_selftest_kernel(uint32_t seed, int rounds, f29::F29* __restrict__ ref, uint32_t* __restrict__ mismatches);
// final check of a batch = G1 prelude + pairing engine, back to back on `st`; scratch512 = the context's 512-byte device scratch
// (ticks at +128 when profiling, the hand-over points at +256)
inline void launch_batch_final(cudaStream_t st, const Partial* parts, int nparts, const DeviceTables* T, uint32_t* result, unsigned char* scratch512,
                               bool ticks) {
    FinalPts* fp = reinterpret_cast<FinalPts*>(scratch512 + 256);
    batch_final_kernel<<<1, kFinalThreads, sizeof(FinalSmem), st>>>(parts, nparts, T, result, fp);
    pairing_check_kernel<<<1, kPairThreads, sizeof(PairSmem), st>>>(fp, T, result, ticks ? reinterpret_cast<long long*>(scratch512 + 128) : nullptr);
}
__global__ void many_lhs_kernel(const uint8_t* __restrict__ z32, const uint8_t* __restrict__ y32, const G1Affine* __restrict__ C,
                                const G1Affine* __restrict__ P, size_t m, const DeviceTables* __restrict__ T, G1Affine* __restrict__ X,
                                uint32_t* __restrict__ status);
__global__ void many_lhs_zy_kernel(const ZY* __restrict__ zy, const G1Affine* __restrict__ C, const G1Affine* __restrict__ P, size_t m,
                                   const DeviceTables* __restrict__ T, G1Affine* __restrict__ X, const uint32_t* __restrict__ status);
__global__ void many_pairing_kernel(const G1Affine* __restrict__ X, const G1Affine* __restrict__ P, const uint32_t* __restrict__ status, size_t m,
                                    const DeviceTables* __restrict__ T, uint8_t* __restrict__ verdicts);
__global__ void many_pairing_warp_kernel(const G1Affine* __restrict__ X, const G1Affine* __restrict__ P, size_t m, const DeviceTables* __restrict__ T,
                                         uint8_t* __restrict__ verdicts);
__global__ void export_scalars_kernel(const ZY* __restrict__ zy, int n, uint8_t* __restrict__ z_out, uint8_t* __restrict__ y_out);
__global__ void r_to_raw_kernel(const Fr* __restrict__ r_mont, ZY* __restrict__ out);
__global__ void status_or_kernel(const uint32_t* __restrict__ status, int n, uint32_t* __restrict__ out);
__global__ void harness_parse_points_kernel(const uint8_t* bytes, int n, G1Affine* out, uint32_t* bad);
__global__ void harness_blob_kernel(uint64_t seed, int n, int D, const DeviceTables* __restrict__ T, uint8_t* __restrict__ blobs);
__global__ void harness_commit_kernel(uint64_t seed, int n, int D, const G1Affine* __restrict__ M, const Fr* __restrict__ z_mont,
                                      uint8_t* __restrict__ out, int want_proof);
__global__ void lag_parse_kernel(const uint8_t* __restrict__ bytes, G1Affine* __restrict__ out, uint32_t* __restrict__ bad);
__global__ void lag_table_kernel(const G1Affine* __restrict__ L, G1* __restrict__ table);
__global__ void blob_scalars_kernel(const uint8_t* __restrict__ blobs, int n, Fr* __restrict__ scalars, uint32_t* __restrict__ status);
__global__ void quotient_kernel(const uint8_t* __restrict__ blobs, int n, const Fr* __restrict__ z_mont, const ZY* __restrict__ zy,
                                const DeviceTables* __restrict__ T, Fr* __restrict__ scalars);
__global__ void lag_msm_kernel(const Fr* __restrict__ scalars, int n, const G1* __restrict__ table, uint8_t* __restrict__ out48);

}  // namespace kzgb200
