"""GPU parity tests (run on the B200 box: pytest -m gpu).  Everything goes through the C ABI
(libkzgb200.so via kzg_rs_b200.api) and is compared with the golden vectors the reference ships and, for the
intermediates z / y, with the CPU oracle on the same inputs -- bit-exact.

Shapes follow the reference's own tests: kzg_proof.rs:604-631, :654-680, :706-737, :739-778.
"""
import pytest
from conftest import unhex

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def K():
    import torch
    assert torch.cuda.is_available(), "gpu tests need a CUDA device"
    import kzg_rs_b200 as K
    return K


@pytest.fixture(scope="module")
def settings(K):
    s = K.KzgSettings.load_trusted_setup_file()
    s.context(0)
    return s


def tri(fn):
    """Ok(b) -> b, Err -> None (the harness rule of kzg_proof.rs:622-629)."""
    import kzg_rs_b200 as K
    try:
        return fn()
    except K.KzgError:
        return None


def test_verify_kzg_proof_vectors(K, settings, vectors):
    for c in vectors["verify_kzg_proof"]:
        def run():
            args = (K.Bytes48.from_hex(c["commitment"]), K.Bytes32.from_hex(c["z"]), K.Bytes32.from_hex(c["y"]),
                    K.Bytes48.from_hex(c["proof"]))
            return K.KzgProof.verify_kzg_proof(*args, settings)
        assert tri(run) == c["output"], c["name"]


def test_verify_kzg_proof_many_matches_vectors(K, settings, vectors):
    cases = [c for c in vectors["verify_kzg_proof"]
             if [len(unhex(c[k])) for k in ("commitment", "z", "y", "proof")] == [48, 32, 32, 48]]
    cat = lambda k: b"".join(unhex(c[k]) for c in cases)
    got = K.KzgProof.verify_kzg_proof_many(cat("commitment"), cat("z"), cat("y"), cat("proof"), len(cases), settings)
    want = bytes({True: 1, False: 0, None: 2}[c["output"]] for c in cases)
    assert got == want


def test_kat_challenge_and_evaluation(K, settings, vectors, oracle):
    k = vectors["kat_compute_challenge"]   # kzg_proof.rs:739-753
    blob = vectors.blobs[k["blob"]]
    case = [c for c in vectors["verify_blob_kzg_proof"] if c["blob"] == k["blob"] and c["commitment"] == k["commitment"]][0]
    ok, z, y = K.KzgProof.verify_blob_kzg_proof(blob, unhex(case["commitment"]), unhex(case["proof"]), settings, want_zy=True)
    assert z == unhex(k["z"])
    assert y == oracle.evaluate_polynomial(blob, z)
    k = vectors["kat_evaluate_polynomial"]  # kzg_proof.rs:755-778: the blob of 19b3f3f8..., single non-zero element
    blob = vectors.blobs[k["blob"]]
    case = [c for c in vectors["verify_blob_kzg_proof"] if c["blob"] == k["blob"] and c["output"] is True][0]
    ok, z, y = K.KzgProof.verify_blob_kzg_proof(blob, unhex(case["commitment"]), unhex(case["proof"]), settings, want_zy=True)
    assert ok is True and y == oracle.evaluate_polynomial(blob, z)


def test_verify_blob_kzg_proof_vectors(K, settings, vectors, oracle):
    for c in vectors["verify_blob_kzg_proof"]:
        def run():
            args = (K.Blob.from_slice(vectors.blobs[c["blob"]]), K.Bytes48.from_hex(c["commitment"]), K.Bytes48.from_hex(c["proof"]))
            return K.KzgProof.verify_blob_kzg_proof(*args, settings, want_zy=True)
        got = tri(run)
        assert (got[0] if got else None) == c["output"], c["name"]
        if got:
            want = oracle.verify_blob_kzg_proof(vectors.blobs[c["blob"]], unhex(c["commitment"]), unhex(c["proof"]), want_zy=True)
            assert (got[1], got[2]) == (want[1], want[2]), c["name"]


def test_verify_blob_kzg_proof_batch_vectors(K, settings, vectors, oracle):
    for c in vectors["verify_blob_kzg_proof_batch"]:
        def run():
            blobs = [K.Blob.from_slice(vectors.blobs[i]) for i in c["blobs"]]
            cs = [K.Bytes48.from_hex(x) for x in c["commitments"]]
            ps = [K.Bytes48.from_hex(x) for x in c["proofs"]]
            return K.KzgProof.verify_blob_kzg_proof_batch(blobs, cs, ps, settings, want_zy=True)
        got = tri(run)
        assert (got[0] if got else None) == c["output"], c["name"]
        if got and len(c["blobs"]) >= 1:
            ok, rc, zs, ys, _ = oracle.verify_blob_kzg_proof_batch([vectors.blobs[i] for i in c["blobs"]],
                                                                   [unhex(x) for x in c["commitments"]],
                                                                   [unhex(x) for x in c["proofs"]], want_trace=True)
            assert (got[1], got[2]) == (zs, ys), c["name"]


def test_reference_batch_harness_shape_n1(K, settings, vectors):
    """kzg_proof.rs:706-737: single-blob vectors through the batch entry with 1-element Vecs."""
    for c in vectors["verify_blob_kzg_proof"]:
        def run():
            return K.KzgProof.verify_blob_kzg_proof_batch([K.Blob.from_slice(vectors.blobs[c["blob"]])],
                                                          [K.Bytes48.from_hex(c["commitment"])], [K.Bytes48.from_hex(c["proof"])], settings)
        assert tri(run) == c["output"], c["name"]


def test_empty_batch_is_true(K, settings):
    assert K.KzgProof.verify_blob_kzg_proof_batch([], [], [], settings) is True   # kzg_proof.rs:478-480
