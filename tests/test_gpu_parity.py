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


def _batch_case(vectors, n):
    return [c for c in vectors["verify_blob_kzg_proof_batch"] if c["output"] is True and len(c["blobs"]) == n][0]


def test_batch_intermediates_r_and_msm_sums_bit_exact(K, settings, vectors, oracle):
    """Exact transcript mode: r, sum r_i pi_i and rhs_g1 equal the oracle's (kzg_proof.rs:413-433), for n = 2..6."""
    from kzg_rs_b200 import api
    from oracle import pyref as R
    for n in (2, 3, 5, 6):
        c = _batch_case(vectors, n)
        blobs = [vectors.blobs[i] for i in c["blobs"]]
        cs, ps = [unhex(x) for x in c["commitments"]], [unhex(x) for x in c["proofs"]]
        assert K.KzgProof.verify_blob_kzg_proof_batch(blobs, cs, ps, settings) is True
        got = api.last_batch_intermediates(settings)
        ok, rc, zs, ys, tr = oracle.verify_blob_kzg_proof_batch(blobs, cs, ps, want_trace=True)
        assert got["r"] == tr["r"], n
        assert R.g1_to_compressed(got["A"]) == tr["proof_lincomb"], n
        rhs = R.g1_add(got["B_prime"], R.g1_neg(R.g1_mul(R.G1_GEN, got["sum_r_y"])))
        assert R.g1_to_compressed(rhs) == tr["rhs_g1"], n


def test_tree_transcript_mode_same_verdicts(K, settings, vectors, oracle):
    """Opt-in tree transcript: r differs from kzg-rs's, verdicts and z / y do not."""
    from kzg_rs_b200 import api
    api.set_transcript_mode(settings, api.TRANSCRIPT_TREE)
    try:
        for c in vectors["verify_blob_kzg_proof_batch"]:
            def run():
                return K.KzgProof.verify_blob_kzg_proof_batch([K.Blob.from_slice(vectors.blobs[i]) for i in c["blobs"]],
                                                              [K.Bytes48.from_hex(x) for x in c["commitments"]],
                                                              [K.Bytes48.from_hex(x) for x in c["proofs"]], settings)
            assert tri(run) == c["output"], c["name"]
        c = _batch_case(vectors, 4)
        args = ([vectors.blobs[i] for i in c["blobs"]], [unhex(x) for x in c["commitments"]], [unhex(x) for x in c["proofs"]])
        assert K.KzgProof.verify_blob_kzg_proof_batch(*args, settings) is True
        r_tree = api.last_batch_intermediates(settings)["r"]
    finally:
        api.set_transcript_mode(settings, api.TRANSCRIPT_EXACT)
    assert K.KzgProof.verify_blob_kzg_proof_batch(*args, settings) is True
    assert api.last_batch_intermediates(settings)["r"] != r_tree


def test_synthetic_batch_64_against_oracle(K, settings, oracle):
    """BASELINE configs[2]: 64 synthetic blobs (harness generator): verdict, every z_i, y_i, r and both sums vs the
    oracle; then a corrupted-proof batch (-> false) and a non-canonical element (-> Err)."""
    import ctypes as C
    import os
    import torch
    from kzg_rs_b200 import api
    from oracle import pyref as R
    lib, ctx, n = K.Library.get().dll, settings.context(0), 64
    tau = open(os.path.join(os.path.dirname(K.__file__), "data", "tau_powers_g1.bin"), "rb").read()
    blobs = torch.empty(n * 131072, dtype=torch.uint8, device="cuda")
    cs = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
    ps = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
    assert lib.kzgb200_harness_generate(ctx, 0x4B5A47, n, 8, tau, blobs.data_ptr(), cs.data_ptr(), ps.data_ptr()) == 0
    hb, hc, hp = (t.cpu().numpy().tobytes() for t in (blobs, cs, ps))
    # the generator itself: commitments and proofs equal the oracle's commit/prove on the same blobs
    for i in (0, 17, 63):
        b = hb[i * 131072:(i + 1) * 131072]
        assert oracle.blob_to_kzg_commitment(b) == hc[i * 48:(i + 1) * 48]
        assert oracle.compute_blob_kzg_proof(b, hc[i * 48:(i + 1) * 48]) == hp[i * 48:(i + 1) * 48]
    ok, zs, ys = K.KzgProof.verify_blob_kzg_proof_batch_raw(hb, n, hc, n, hp, n, settings, want_zy=True)
    split = lambda raw, k: [raw[k * i:k * i + k] for i in range(n)]
    ok_ref, rc, z_ref, y_ref, tr = oracle.verify_blob_kzg_proof_batch(split(hb, 131072), split(hc, 48), split(hp, 48), nthreads=8, want_trace=True)
    assert ok is True and ok_ref is True and zs == z_ref and ys == y_ref
    got = api.last_batch_intermediates(settings)
    assert got["r"] == tr["r"] and R.g1_to_compressed(got["A"]) == tr["proof_lincomb"]
    rhs = R.g1_add(got["B_prime"], R.g1_neg(R.g1_mul(R.G1_GEN, got["sum_r_y"])))
    assert R.g1_to_compressed(rhs) == tr["rhs_g1"]
    # proof 17 replaced by proof 18 -> false
    bad = bytearray(hp); bad[17 * 48:18 * 48] = hp[18 * 48:19 * 48]
    assert K.KzgProof.verify_blob_kzg_proof_batch_raw(hb, n, hc, n, bytes(bad), n, settings) is False
    # element 9 of blob 5 set to the modulus -> Err(BadArgs)
    q = bytes.fromhex("73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001")
    badb = bytearray(hb); badb[5 * 131072 + 9 * 32:5 * 131072 + 10 * 32] = q
    with pytest.raises(K.KzgError) as e:
        K.KzgProof.verify_blob_kzg_proof_batch_raw(bytes(badb), n, hc, n, hp, n, settings)
    assert e.value.kind == "BadArgs"
    # length mismatch -> InvalidBytesLength (kzg_proof.rs:491-501)
    with pytest.raises(K.KzgError) as e:
        K.KzgProof.verify_blob_kzg_proof_batch_raw(hb, n, hc, n - 1, hp, n, settings)
    assert e.value.kind == "InvalidBytesLength"
    # a commitment / a proof on the curve but outside the subgroup (found by the deferred subgroup checks) -> Err(BadArgs)
    from kzg_rs_b200.sharded import NOT_IN_G1
    for which in (0, 1):
        for idx in (0, 40, 63):
            arrs = [bytearray(hc), bytearray(hp)]
            arrs[which][idx * 48:idx * 48 + 48] = NOT_IN_G1
            with pytest.raises(K.KzgError) as e:
                K.KzgProof.verify_blob_kzg_proof_batch_raw(hb, n, bytes(arrs[0]), n, bytes(arrs[1]), n, settings)
            assert e.value.kind == "BadArgs"
    assert K.KzgProof.verify_blob_kzg_proof_batch_raw(hb, n, hc, n, hp, n, settings) is True      # and the flags do not leak into the next call


def test_canonicity_boundary_elements(K, settings, vectors, oracle):
    """Blob::as_polynomial (src/dtypes.rs:48-57): q - 1 and values sharing q's top word are canonical (Ok, here false
    because the commitment no longer matches), q and q + 1 are not (Err(BadArgs)) -- the kernel decides on the top word
    first and compares exactly only then.  z and y of the accepted blobs against the oracle."""
    q = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    case = next(c for c in vectors["verify_blob_kzg_proof"] if c["output"] is True)
    blob, c, p = vectors.blobs[case["blob"]], unhex(case["commitment"]), unhex(case["proof"])
    for value, want_err in ((q - 1, False), (q - (1 << 200), False), ((0x73eda753 << 224) | 5, False), (q, True), (q + 1, True),
                            ((0x73eda754 << 224), True), (2**256 - 1, True)):
        for idx in (0, 1, 2047, 4095):
            b = bytearray(blob)
            b[32 * idx:32 * idx + 32] = value.to_bytes(32, "big")
            got = tri(lambda: K.KzgProof.verify_blob_kzg_proof(K.Blob.from_slice(bytes(b)), K.Bytes48.from_slice(c), K.Bytes48.from_slice(p),
                                                               settings, want_zy=True))
            if want_err:
                assert got is None, (hex(value), idx, got)
            else:
                ok, z, y = got
                assert ok is False
                assert z == oracle.compute_challenge(bytes(b), c) and y == oracle.evaluate_polynomial(bytes(b), z)


def test_pipeline_matches_blocking_calls(K, settings, oracle):
    """Streaming front-end: batches in flight on two contexts give the verdicts, errors and z / y of the
    blocking entry, whatever the completion order; host and device submissions mixed."""
    import ctypes as C
    import os
    import torch
    lib, ctx = K.Library.get().dll, settings.context(0)
    tau = open(os.path.join(os.path.dirname(K.__file__), "data", "tau_powers_g1.bin"), "rb").read()
    sizes, data = (64, 700, 1, 2), []
    for k, n in enumerate(sizes):
        b = torch.empty(n * 131072, dtype=torch.uint8, device="cuda")
        c = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
        p = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
        assert lib.kzgb200_harness_generate(ctx, 0x1000 + k, n, 8, tau, b.data_ptr(), c.data_ptr(), p.data_ptr()) == 0
        data.append((b, c, p))
    torch.cuda.synchronize()
    host = [tuple(t.cpu().numpy().tobytes() for t in d) for d in data]
    q = bytes.fromhex("73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001")
    hb, hc, hp = host[0]
    bad_proof = bytearray(hp); bad_proof[17 * 48:18 * 48] = hp[18 * 48:19 * 48]
    bad_blob = bytearray(hb); bad_blob[5 * 131072 + 9 * 32:5 * 131072 + 10 * 32] = q
    z0, y0 = bytearray(64 * 32), bytearray(64 * 32)
    want_ok, want_z, want_y = K.KzgProof.verify_blob_kzg_proof_batch_raw(hb, 64, hc, 64, hp, 64, settings, want_zy=True)
    assert want_ok is True
    with K.BatchPipeline(settings, depth=2) as pipe:
        jobs = []      # (ticket or exception kind, expectation)
        for rep in range(2):
            jobs.append((pipe.submit(hb, 64, hc, 64, hp, 64, z_out=z0, y_out=y0), True))
            jobs.append((pipe.submit_device(*data[1], 700), True))
            jobs.append((pipe.submit(hb, 64, hc, 64, bytes(bad_proof), 64), False))
            jobs.append((pipe.submit(bytes(bad_blob), 64, hc, 64, hp, 64), "BadArgs"))
            jobs.append((pipe.submit(host[2][0], 1, host[2][1], 1, host[2][2], 1), True))
            jobs.append((pipe.submit(hb, 64, hc, 63, hp, 64), "InvalidBytesLength"))
            jobs.append((pipe.submit(b"", 0, b"", 0, b"", 0), True))
            jobs.append((pipe.submit_device(*data[3], 2), True))
        for ticket, want in reversed(jobs):          # wait in the opposite order
            if isinstance(want, str):
                with pytest.raises(K.KzgError) as e:
                    pipe.wait(ticket)
                assert e.value.kind == want
            else:
                assert pipe.wait(ticket) is want
        assert bytes(z0) == b"".join(want_z) and bytes(y0) == b"".join(want_y)
        with pytest.raises(K.KzgError):
            pipe.wait(jobs[0][0])                    # a ticket can be waited for once


def test_gpu_commit_and_prove_reproduce_reference_vector_bytes(K, settings, vectors, oracle):
    """blob_to_kzg_commitment / compute_blob_kzg_proof on the GPU regenerate the commitment and proof
    bytes of every valid verify_blob_kzg_proof vector, and agree with the oracle on random blobs."""
    import random
    import torch
    lib, ctx = K.Library.get().dll, settings.context(0)
    assert lib.kzgb200_load_g1_lagrange(ctx, settings.g1_lagrange_bytes, 4096) == 0
    cases = [c for c in vectors["verify_blob_kzg_proof"] if c["output"] is True]
    rnd = random.Random(99)
    q = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    extra = [b"".join(rnd.randrange(q).to_bytes(32, "big") for _ in range(4096)) for _ in range(2)]
    blobs = [vectors.blobs[c["blob"]] for c in cases] + extra
    n = len(blobs)
    d_b = torch.frombuffer(bytearray(b"".join(blobs)), dtype=torch.uint8).cuda()
    d_c = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
    d_p = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
    assert lib.kzgb200_blob_to_kzg_commitment_batch(ctx, d_b.data_ptr(), n, d_c.data_ptr()) == 0
    assert lib.kzgb200_compute_blob_kzg_proof_batch(ctx, d_b.data_ptr(), d_c.data_ptr(), n, d_p.data_ptr()) == 0
    hc, hp = d_c.cpu().numpy().tobytes(), d_p.cpu().numpy().tobytes()
    for i, c in enumerate(cases):
        assert hc[48 * i:48 * i + 48] == unhex(c["commitment"]), c["name"]
        assert hp[48 * i:48 * i + 48] == unhex(c["proof"]), c["name"]
    for k, blob in enumerate(extra):
        i = len(cases) + k
        assert hc[48 * i:48 * i + 48] == oracle.blob_to_kzg_commitment(blob)
        assert hp[48 * i:48 * i + 48] == oracle.compute_blob_kzg_proof(blob, hc[48 * i:48 * i + 48])
    # and the verifier accepts what the prover produced (random blobs, batch path)
    assert K.KzgProof.verify_blob_kzg_proof_batch_raw(b"".join(blobs), n, hc, n, hp, n, settings) is True
    # a non-canonical element is rejected by the commit path as well
    bad = bytearray(blobs[0]); bad[0:32] = q.to_bytes(32, "big")
    d_bad = torch.frombuffer(bad, dtype=torch.uint8).cuda()
    assert lib.kzgb200_blob_to_kzg_commitment_batch(ctx, d_bad.data_ptr(), 1, d_c.data_ptr()) == 1


def _random_workload(K, settings, n, seed):
    """n uniformly random blobs with commitments / proofs from the GPU commit/prove path (device tensors)."""
    import torch
    lib, ctx = K.Library.get().dll, settings.context(0)
    assert lib.kzgb200_load_g1_lagrange(ctx, settings.g1_lagrange_bytes, 4096) == 0
    g = torch.Generator(device="cuda").manual_seed(seed)
    blobs = torch.randint(0, 256, (n * 131072,), dtype=torch.uint8, device="cuda", generator=g)
    blobs.view(n * 4096, 32)[:, 0] &= 0x3f
    cs = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
    ps = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
    torch.cuda.synchronize()        # the library works on its own stream: torch's kernels must have finished
    assert lib.kzgb200_blob_to_kzg_commitment_batch(ctx, blobs.data_ptr(), n, cs.data_ptr()) == 0
    assert lib.kzgb200_compute_blob_kzg_proof_batch(ctx, blobs.data_ptr(), cs.data_ptr(), n, ps.data_ptr()) == 0
    return blobs, cs, ps


@pytest.mark.parametrize("n", [2500, 4133])
def test_chunked_paths_keep_r_and_sums_bit_exact(K, settings, oracle, n):
    """Multi-chunk execution: the host path cuts the batch into >= 1024-blob chunks and advances the exact transcript
    chain chunk by chunk; the resident path does the same above 4096 blobs.  r, z, y and both MSM sums must still equal
    the oracle's (ragged sizes: the last chunk is partial)."""
    import ctypes as C
    import os
    from kzg_rs_b200 import api
    from oracle import pyref as R
    lib, ctx = K.Library.get().dll, settings.context(0)
    blobs, cs, ps = _random_workload(K, settings, n, 1234 + n)
    hb, hc, hp = (t.cpu().numpy().tobytes() for t in (blobs, cs, ps))
    rc, ok_ref, z_ref, y_ref = oracle.verify_batch_raw(hb, hc, hp, n, nthreads=os.cpu_count())
    assert rc == 0 and ok_ref is True
    split = lambda raw, k: [raw[k * i:k * i + k] for i in range(n)]
    r_ref, _ = oracle.compute_r_powers(split(hc, 48), split(z_ref, 32), split(y_ref, 32), split(hp, 48))
    # host path (chunked copies + incremental chain)
    ok, zs, ys = K.KzgProof.verify_blob_kzg_proof_batch_raw(hb, n, hc, n, hp, n, settings, want_zy=True)
    assert ok is True and b"".join(zs) == z_ref and b"".join(ys) == y_ref
    host = api.last_batch_intermediates(settings)
    assert host["r"] == r_ref
    # resident path
    okc = C.c_int(-1)
    assert lib.kzgb200_verify_blob_kzg_proof_batch_device(ctx, blobs.data_ptr(), cs.data_ptr(), ps.data_ptr(), n, C.byref(okc), None, None) == 0
    dev = api.last_batch_intermediates(settings)
    assert okc.value == 1 and dev["r"] == r_ref and dev["A"] == host["A"] and dev["B_prime"] == host["B_prime"]
    # the sums against an independent evaluation on a 3-blob prefix is covered by the vector tests; here: pairing-level check
    rhs = R.g1_add(dev["B_prime"], R.g1_neg(R.g1_mul(R.G1_GEN, dev["sum_r_y"])))
    assert oracle.pairings_verify(R.g1_to_compressed(dev["A"]), 1, R.g1_to_compressed(rhs), 0) is True
    # tree mode on the same data: same verdict, different r
    api.set_transcript_mode(settings, api.TRANSCRIPT_TREE)
    try:
        assert K.KzgProof.verify_blob_kzg_proof_batch_raw(hb, n, hc, n, hp, n, settings) is True
        assert api.last_batch_intermediates(settings)["r"] != r_ref
        bad = bytearray(hp); bad[48 * (n - 1):48 * n] = hp[:48]
        assert K.KzgProof.verify_blob_kzg_proof_batch_raw(hb, n, hc, n, bytes(bad), n, settings) is False
    finally:
        api.set_transcript_mode(settings, api.TRANSCRIPT_EXACT)


def test_cpp_mirror_runs_reference_vectors(K, settings, vectors, tmp_path):
    """The C++ mirror of the reference interface (include/kzg_rs.hpp) on the GPU: every well-formed verify_kzg_proof
    vector and every verify_blob_kzg_proof vector through kzg_rs::KzgProof, tri-state results as in kzg_proof.rs:604-680."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "run_vectors.cpp"
    src.write_text(r'''
#include "kzg_rs.hpp"
#include <cstdio>
#include <vector>
using namespace kzg_rs;
static int tri(const Result<bool>& r) { return r.is_err() ? 2 : (r.unwrap() ? 1 : 0); }
int main(int argc, char** argv) {
    auto s = KzgSettings::load_trusted_setup_file(argv[1]);
    if (s.is_err()) return 3;
    const KzgSettings& ks = s.unwrap();
    std::vector<uint8_t> rec(160);
    uint32_t n = 0;
    if (fread(&n, 4, 1, stdin) != 1) return 4;
    for (uint32_t i = 0; i < n; i++) {
        if (fread(rec.data(), 1, 160, stdin) != 160) return 4;
        auto c = Bytes48::from_slice(rec.data(), 48).unwrap(); auto z = Bytes32::from_slice(rec.data() + 48, 32).unwrap();
        auto y = Bytes32::from_slice(rec.data() + 80, 32).unwrap(); auto p = Bytes48::from_slice(rec.data() + 112, 48).unwrap();
        putchar('0' + tri(KzgProof::verify_kzg_proof(c, z, y, p, ks)));
    }
    putchar('\n');
    if (fread(&n, 4, 1, stdin) != 1) return 4;
    std::vector<uint8_t> blob(BYTES_PER_BLOB + 96);
    for (uint32_t i = 0; i < n; i++) {
        if (fread(blob.data(), 1, blob.size(), stdin) != blob.size()) return 4;
        auto b = Blob::from_slice(blob.data(), BYTES_PER_BLOB).unwrap();
        auto c = Bytes48::from_slice(blob.data() + BYTES_PER_BLOB, 48).unwrap(); auto p = Bytes48::from_slice(blob.data() + BYTES_PER_BLOB + 48, 48).unwrap();
        putchar('0' + tri(KzgProof::verify_blob_kzg_proof(b, c, p, ks)));
        std::vector<Blob> bs{b}; std::vector<Bytes48> cs{c}, ps{p};
        putchar('0' + tri(KzgProof::verify_blob_kzg_proof_batch(bs, cs, ps, ks)));    // kzg_proof.rs:706-737 shape
    }
    putchar('\n');
    std::vector<Blob> b2; std::vector<Bytes48> c2(1), p2;
    putchar('0' + tri(KzgProof::verify_blob_kzg_proof_batch(b2, c2, p2, ks)));          // empty -> Ok(true)
    b2.resize(2); c2.resize(1); p2.resize(2);
    auto r = KzgProof::verify_blob_kzg_proof_batch(b2, c2, p2, ks);                      // length mismatch
    putchar(r.is_err() && r.unwrap_err().kind == KzgError::InvalidBytesLength ? 'L' : '?');
    // streaming front-end: three tickets in flight on two contexts, waited for in reverse order
    auto pipe = BatchPipeline::create(ks, 2).unwrap();
    std::vector<Blob> e0; std::vector<Bytes48> e1, e2;
    auto t1 = pipe.submit(e0, e1, e2).unwrap();                                          // empty -> Ok(true)
    auto t2 = pipe.submit(std::vector<Blob>(2), std::vector<Bytes48>(1), std::vector<Bytes48>(2)).unwrap();   // length mismatch
    auto t3 = pipe.submit(std::vector<Blob>(2), std::vector<Bytes48>(2), std::vector<Bytes48>(2)).unwrap();   // zero blobs, zero points: Err(BadArgs)
    auto r3 = pipe.wait(t3); auto r2 = pipe.wait(t2); auto r1 = pipe.wait(t1);
    putchar(r1.is_ok() && r1.unwrap() ? '1' : '?');
    putchar(r2.is_err() && r2.unwrap_err().kind == KzgError::InvalidBytesLength ? 'L' : '?');
    putchar(r3.is_err() && r3.unwrap_err().kind == KzgError::BadArgs ? 'B' : '?');
    putchar('\n');
    return 0;
}
''')
    exe = tmp_path / "run_vectors"
    lib_dir = os.path.join(root, "kzg_rs_b200")
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"), str(src), "-o", str(exe),
                           "-L", lib_dir, "-lkzgb200", "-Wl,-rpath," + lib_dir])
    import struct
    k_cases = [c for c in vectors["verify_kzg_proof"] if [len(unhex(c[k])) for k in ("commitment", "z", "y", "proof")] == [48, 32, 32, 48]]
    b_cases = [c for c in vectors["verify_blob_kzg_proof"]
               if len(vectors.blobs[c["blob"]]) == 131072 and len(unhex(c["commitment"])) == 48 and len(unhex(c["proof"])) == 48]
    payload = struct.pack("<I", len(k_cases)) + b"".join(unhex(c["commitment"]) + unhex(c["z"]) + unhex(c["y"]) + unhex(c["proof"]) for c in k_cases)
    payload += struct.pack("<I", len(b_cases)) + b"".join(vectors.blobs[c["blob"]] + unhex(c["commitment"]) + unhex(c["proof"]) for c in b_cases)
    out = subprocess.run([str(exe), os.path.join(lib_dir, "data", "mainnet_setup.bin")], input=payload, capture_output=True)
    assert out.returncode == 0, out.stderr.decode()
    l1, l2, l3 = out.stdout.decode().split("\n")[:3]
    code = {True: "1", False: "0", None: "2"}
    assert l1 == "".join(code[c["output"]] for c in k_cases)
    assert l2 == "".join(code[c["output"]] * 2 for c in b_cases)
    assert l3 == "1L1LB"
