// kzg_rs.hpp -- header-only C++17 mirror of kzg-rs's public verification interface over the C ABI of
// kzgb200.h.  Same names, argument meaning and error behaviour as the reference (the north star asks for a
// Rust host; this image has no Rust toolchain, INTEGRATION.md carries the Rust shim as source):
//   kzg_rs::KzgProof::verify_kzg_proof / verify_blob_kzg_proof / verify_blob_kzg_proof_batch   src/kzg_proof.rs:353-525
//   kzg_rs::KzgSettings::load_trusted_setup_file                                                src/trusted_setup.rs:94-98
//   kzg_rs::Blob / Bytes32 / Bytes48 (from_slice length check, as_slice)                        src/dtypes.rs:7-46
//   kzg_rs::KzgError                                                                            src/enums.rs:6-18
// Result<bool, KzgError> is kzg_rs::Result<bool>: either a value or an error.
#pragma once
#include <array>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <string>
#include <utility>
#include <vector>
#include "kzgb200.h"

namespace kzg_rs {

constexpr size_t BYTES_PER_BLOB = KZGB200_BYTES_PER_BLOB;
constexpr size_t BYTES_PER_COMMITMENT = KZGB200_BYTES_PER_COMMITMENT;
constexpr size_t BYTES_PER_PROOF = KZGB200_BYTES_PER_PROOF;
constexpr size_t BYTES_PER_FIELD_ELEMENT = KZGB200_BYTES_PER_FIELD_ELEMENT;

struct KzgError {
    enum Kind { BadArgs, InternalError, InvalidBytesLength, InvalidHexFormat, InvalidTrustedSetup } kind;
    std::string message;
};

template <class T>
class Result {
    bool ok_;
    T value_{};
    KzgError err_{KzgError::InternalError, ""};
public:
    Result(T v) : ok_(true), value_(std::move(v)) {}
    Result(KzgError e) : ok_(false), err_(std::move(e)) {}
    bool is_ok() const { return ok_; }
    bool is_err() const { return !ok_; }
    const T& unwrap() const { return value_; }
    T& unwrap() { return value_; }
    const KzgError& unwrap_err() const { return err_; }
};

template <size_t N>
struct BytesN {
    std::array<uint8_t, N> bytes{};
    static Result<BytesN> from_slice(const uint8_t* p, size_t len) {   // src/dtypes.rs:19-29
        if (len != N) return KzgError{KzgError::InvalidBytesLength, "Invalid slice length"};
        BytesN b;
        std::memcpy(b.bytes.data(), p, N);
        return b;
    }
    const uint8_t* as_slice() const { return bytes.data(); }
};
using Bytes32 = BytesN<32>;
using Bytes48 = BytesN<48>;
using Blob = BytesN<BYTES_PER_BLOB>;
static_assert(sizeof(Blob) == BYTES_PER_BLOB && sizeof(Bytes48) == 48, "vectors of these are contiguous byte arrays");

class KzgSettings {
    struct Deleter { void operator()(kzgb200_ctx* c) const { kzgb200_destroy(c); } };
    std::shared_ptr<kzgb200_ctx> ctx_;
public:
    std::vector<uint8_t> g2_points;   // compressed, 96 bytes each; the verification path reads [0] and [1]
    // setup file: "KZGS" | u32 n1 | u32 n2 | n1*48 G1 Lagrange | n2*96 G2 monomial (kzg_rs_b200/data/mainnet_setup.bin)
    static Result<KzgSettings> load_trusted_setup_file(const std::string& path, int device = 0) {
        std::ifstream f(path, std::ios::binary);
        std::vector<uint8_t> raw((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
        if (raw.size() < 12 || std::memcmp(raw.data(), "KZGS", 4)) return KzgError{KzgError::InvalidTrustedSetup, "Invalid trusted setup"};
        uint32_t n1, n2;
        std::memcpy(&n1, raw.data() + 4, 4); std::memcpy(&n2, raw.data() + 8, 4);
        if (n2 < 2 || raw.size() != 12 + size_t(n1) * 48 + size_t(n2) * 96) return KzgError{KzgError::InvalidTrustedSetup, "Invalid trusted setup"};
        KzgSettings s;
        s.g2_points.assign(raw.begin() + 12 + size_t(n1) * 48, raw.end());
        kzgb200_ctx* c = nullptr;
        int rc = kzgb200_create(&c, device, s.g2_points.data(), 192);
        if (rc) return KzgError{rc == KZGB200_INVALID_SETUP ? KzgError::InvalidTrustedSetup : KzgError::InternalError, "kzgb200_create failed"};
        s.ctx_ = std::shared_ptr<kzgb200_ctx>(c, Deleter());
        return s;
    }
    kzgb200_ctx* ctx() const { return ctx_.get(); }
};

namespace detail {
inline Result<bool> to_result(int rc, int ok, const char* len_msg = "Invalid commitments length") {
    switch (rc) {
        case KZGB200_OK: return ok != 0;
        case KZGB200_BAD_ARGS: return KzgError{KzgError::BadArgs, "Failed to parse G1Affine from bytes"};
        case KZGB200_INVALID_LENGTH: return KzgError{KzgError::InvalidBytesLength, len_msg};
        case KZGB200_INVALID_SETUP: return KzgError{KzgError::InvalidTrustedSetup, "Invalid trusted setup"};
        default: return KzgError{KzgError::InternalError, "Internal error"};
    }
}
}  // namespace detail

struct KzgProof {
    static Result<bool> verify_kzg_proof(const Bytes48& commitment_bytes, const Bytes32& z_bytes, const Bytes32& y_bytes,
                                         const Bytes48& proof_bytes, const KzgSettings& kzg_settings) {
        int ok = 0;
        int rc = kzgb200_verify_kzg_proof(kzg_settings.ctx(), commitment_bytes.as_slice(), z_bytes.as_slice(), y_bytes.as_slice(),
                                          proof_bytes.as_slice(), &ok);
        return detail::to_result(rc, ok);
    }
    static Result<bool> verify_blob_kzg_proof(const Blob& blob, const Bytes48& commitment_bytes, const Bytes48& proof_bytes,
                                              const KzgSettings& kzg_settings) {
        int ok = 0;
        int rc = kzgb200_verify_blob_kzg_proof(kzg_settings.ctx(), blob.as_slice(), commitment_bytes.as_slice(), proof_bytes.as_slice(),
                                               &ok, nullptr, nullptr);
        return detail::to_result(rc, ok);
    }
    static Result<bool> verify_blob_kzg_proof_batch(const std::vector<Blob>& blobs, const std::vector<Bytes48>& commitments_bytes,
                                                    const std::vector<Bytes48>& proofs_bytes, const KzgSettings& kzg_settings) {
        int ok = 0;
        int rc = kzgb200_verify_blob_kzg_proof_batch(kzg_settings.ctx(), reinterpret_cast<const uint8_t*>(blobs.data()), blobs.size(),
                                                     reinterpret_cast<const uint8_t*>(commitments_bytes.data()), commitments_bytes.size(),
                                                     reinterpret_cast<const uint8_t*>(proofs_bytes.data()), proofs_bytes.size(), &ok,
                                                     nullptr, nullptr);
        return detail::to_result(rc, ok, blobs.size() != commitments_bytes.size() ? "Invalid commitments length" : "Invalid proofs length");
    }
};

// Streaming front-end (include/kzgb200.h, kzgb200_pipeline_*): `depth` verify_blob_kzg_proof_batch calls in flight on one GPU.
// The pipeline owns the submitted vectors until the ticket has been waited for (the C ABI reads them asynchronously).
class BatchPipeline {
    struct Deleter { void operator()(kzgb200_pipeline* p) const { kzgb200_pipeline_destroy(p); } };
    struct Held { std::vector<Blob> blobs; std::vector<Bytes48> commitments, proofs; };
    std::shared_ptr<kzgb200_pipeline> p_;
    std::map<uint64_t, Held> held_;
public:
    using Ticket = uint64_t;
    static Result<BatchPipeline> create(const KzgSettings& kzg_settings, int depth = 2, int device = 0) {
        kzgb200_pipeline* raw = nullptr;
        int rc = kzgb200_pipeline_create(&raw, device, kzg_settings.g2_points.data(), 192, depth);
        if (rc) return KzgError{rc == KZGB200_INVALID_SETUP ? KzgError::InvalidTrustedSetup : KzgError::InternalError, "kzgb200_pipeline_create failed"};
        BatchPipeline bp;
        bp.p_ = std::shared_ptr<kzgb200_pipeline>(raw, Deleter());
        return bp;
    }
    // same arguments as KzgProof::verify_blob_kzg_proof_batch (reference src/kzg_proof.rs:472-477), taken by value like there
    Result<Ticket> submit(std::vector<Blob> blobs, std::vector<Bytes48> commitments_bytes, std::vector<Bytes48> proofs_bytes) {
        Held h{std::move(blobs), std::move(commitments_bytes), std::move(proofs_bytes)};
        uint64_t t = 0;
        int rc = kzgb200_pipeline_submit(p_.get(), reinterpret_cast<const uint8_t*>(h.blobs.data()), h.blobs.size(),
                                         reinterpret_cast<const uint8_t*>(h.commitments.data()), h.commitments.size(),
                                         reinterpret_cast<const uint8_t*>(h.proofs.data()), h.proofs.size(), nullptr, nullptr, &t);
        if (rc) return KzgError{KzgError::InternalError, "kzgb200_pipeline_submit failed"};
        held_.emplace(t, std::move(h));          // moving a vector keeps its buffer where it is
        return t;
    }
    Result<bool> wait(Ticket t) {
        int ok = 0;
        int rc = kzgb200_pipeline_wait(p_.get(), t, &ok);
        auto it = held_.find(t);
        bool c_len = it != held_.end() && it->second.blobs.size() != it->second.commitments.size();
        if (it != held_.end()) held_.erase(it);
        return detail::to_result(rc, ok, c_len ? "Invalid commitments length" : "Invalid proofs length");
    }
};

}  // namespace kzg_rs
