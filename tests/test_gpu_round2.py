"""GPU parity tests, part 2 (pytest -m gpu): the exact transcript at full size, the reference's helper KATs on the CUDA kernels,
the pre-parsed batch, per-blob verdicts, custom trusted setups, pageable host memory and the multi-GPU group (two contexts on
one GPU, in one process and in two processes) -- all through the C ABI, bit-exact against the CPU oracle.
"""
import ctypes as C
import os
import random
import subprocess
import sys

import pytest
from conftest import GOLDEN, ROOT, unhex

pytestmark = pytest.mark.gpu
Q = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
P = 0x1a0111ea397fe69a4b1ba7b6434bacd764774b84f38512bf6730d2a0f6b0f6241eabfffeb153ffffb9feffffffffaaab


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
    import kzg_rs_b200 as K
    try:
        return fn()
    except K.KzgError as e:
        assert e.kind == "BadArgs", e.kind
        return None


def harness(K, settings, n, seed, degree=8):
    """n synthetic blobs (harness generator) with valid commitments / proofs: (device tensors), (host bytes)"""
    import torch
    lib, ctx = K.Library.get().dll, settings.context(0)
    tau = open(os.path.join(os.path.dirname(K.__file__), "data", "tau_powers_g1.bin"), "rb").read()
    b = torch.empty(n * 131072, dtype=torch.uint8, device="cuda")
    c = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
    p = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
    assert lib.kzgb200_harness_generate(ctx, seed, n, degree, tau, b.data_ptr(), c.data_ptr(), p.data_ptr()) == 0
    torch.cuda.synchronize()
    return (b, c, p), tuple(t.cpu().numpy().tobytes() for t in (b, c, p))


def check_sums(got, tr):
    """A = sum r_i pi_i and rhs = B' - [sum r_i y_i] G against the oracle's trace (compressed points)"""
    from oracle import pyref as R
    assert R.g1_to_compressed(got["A"]) == tr["proof_lincomb"]
    rhs = R.g1_add(got["B_prime"], R.g1_neg(R.g1_mul(R.G1_GEN, got["sum_r_y"])))
    assert R.g1_to_compressed(rhs) == tr["rhs_g1"]


def test_full_size_batch_r_and_msm_sums_bit_exact(K, settings, oracle):
    """BASELINE configs[3] size: 16384 uniformly random blobs.  Default (exact, host-hashed) transcript: every z_i, y_i, r,
    sum r_i pi_i and rhs_g1 equal the oracle's; the device-chain mode gives the same r; the tree mode gives the r of its
    restatement in oracle/pyref.py; resident and host (pageable) paths agree."""
    import numpy as np
    from kzg_rs_b200 import api
    from oracle import pyref as R
    from test_gpu_parity import _random_workload
    n = 16384
    lib, ctx = K.Library.get().dll, settings.context(0)
    blobs, cs, ps = _random_workload(K, settings, n, 20261017)
    hb, hc, hp = blobs.cpu().numpy(), cs.cpu().numpy().tobytes(), ps.cpu().numpy().tobytes()
    rc, ok_ref, z_ref, y_ref, tr = oracle.verify_batch_raw(hb.tobytes(), hc, hp, n, nthreads=os.cpu_count(), want_trace=True)
    assert rc == 0 and ok_ref is True
    import torch
    zo = torch.empty(n * 32, dtype=torch.uint8, device="cuda"); yo = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    okc = C.c_int(-1)
    assert lib.kzgb200_verify_blob_kzg_proof_batch_device(ctx, blobs.data_ptr(), cs.data_ptr(), ps.data_ptr(), n, C.byref(okc), zo.data_ptr(), yo.data_ptr()) == 0
    assert okc.value == 1
    assert zo.cpu().numpy().tobytes() == z_ref and yo.cpu().numpy().tobytes() == y_ref
    got = api.last_batch_intermediates(settings)
    assert got["r"] == tr["r"]
    check_sums(got, tr)
    # host path from ordinary (pageable) memory: the numpy array behind hb
    assert lib.kzgb200_verify_blob_kzg_proof_batch(ctx, hb.ctypes.data, n, hc, n, hp, n, C.byref(okc), None, None) == 0
    host = api.last_batch_intermediates(settings)
    assert okc.value == 1 and host["r"] == tr["r"] and host["A"] == got["A"] and host["B_prime"] == got["B_prime"]
    # the GPU chain: same r
    api.set_transcript_mode(settings, api.TRANSCRIPT_EXACT_DEVICE)
    try:
        assert lib.kzgb200_verify_blob_kzg_proof_batch_device(ctx, blobs.data_ptr(), cs.data_ptr(), ps.data_ptr(), n, C.byref(okc), None, None) == 0
        assert okc.value == 1 and api.last_batch_intermediates(settings)["r"] == tr["r"]
        api.set_transcript_mode(settings, api.TRANSCRIPT_TREE)
        assert lib.kzgb200_verify_blob_kzg_proof_batch_device(ctx, blobs.data_ptr(), cs.data_ptr(), ps.data_ptr(), n, C.byref(okc), None, None) == 0
        zs = [int.from_bytes(z_ref[32 * i:32 * i + 32], "big") for i in range(n)]
        ys = [int.from_bytes(y_ref[32 * i:32 * i + 32], "big") for i in range(n)]
        r_tree = R.tree_transcript_r([hc[48 * i:48 * i + 48] for i in range(n)], zs, ys, [hp[48 * i:48 * i + 48] for i in range(n)])
        assert okc.value == 1 and api.last_batch_intermediates(settings)["r"] == r_tree.to_bytes(32, "big")
        # a corrupted proof is still rejected in tree mode, host path
        bad = bytearray(hp); bad[48 * (n - 1):48 * n] = hp[:48]
        assert lib.kzgb200_verify_blob_kzg_proof_batch(ctx, hb.ctypes.data, n, hc, n, bytes(bad), n, C.byref(okc), None, None) == 0 and okc.value == 0
    finally:
        api.set_transcript_mode(settings, api.TRANSCRIPT_EXACT)


@pytest.mark.parametrize("n", [2, 3, 17, 64, 333])
def test_tree_transcript_r_matches_its_restatement(K, settings, oracle, n):
    """KZGB200_TRANSCRIPT_TREE is pinned: r equals oracle/pyref.py's tree_transcript_r for ragged leaf counts, and the
    verdict follows the data (valid -> true, one proof swapped -> false)."""
    from kzg_rs_b200 import api
    from oracle import pyref as R
    _, (hb, hc, hp) = harness(K, settings, n, 0x7ee + n)
    api.set_transcript_mode(settings, api.TRANSCRIPT_TREE)
    try:
        ok, zs, ys = K.KzgProof.verify_blob_kzg_proof_batch_raw(hb, n, hc, n, hp, n, settings, want_zy=True)
        got = api.last_batch_intermediates(settings)
        r = R.tree_transcript_r([hc[48 * i:48 * i + 48] for i in range(n)], [int.from_bytes(z, "big") for z in zs],
                                [int.from_bytes(y, "big") for y in ys], [hp[48 * i:48 * i + 48] for i in range(n)])
        assert ok is True and got["r"] == r.to_bytes(32, "big")
        bad = hp[48:96] + hp[:48] + hp[96:]
        assert K.KzgProof.verify_blob_kzg_proof_batch_raw(hb, n, hc, n, bad, n, settings) is False
    finally:
        api.set_transcript_mode(settings, api.TRANSCRIPT_EXACT)
    assert K.KzgProof.verify_blob_kzg_proof_batch_raw(hb, n, hc, n, hp, n, settings) is True
    assert api.last_batch_intermediates(settings)["r"] != r.to_bytes(32, "big")


def test_reference_helper_kats_on_the_kernels(K, settings, vectors, oracle):
    """compute_challenge / evaluate_polynomial_in_evaluation_form (src/lib.rs:8) as GPU entry points: the reference's own KATs
    (src/kzg_proof.rs:739-778), z inside the evaluation domain (:109-111 returns polynomial[i]; the kernel's inversion-free
    formula must collapse to exactly that), random z against the oracle, and the error cases."""
    from oracle import pyref as R
    k = vectors["kat_compute_challenge"]
    assert K.compute_challenge(vectors.blobs[k["blob"]], unhex(k["commitment"]), settings) == unhex(k["z"])
    k = vectors["kat_evaluate_polynomial"]
    assert K.evaluate_polynomial_in_evaluation_form(vectors.blobs[k["blob"]], unhex(k["z"]), settings) == unhex(k["y"])
    rnd = random.Random(11)
    blob = b"".join(rnd.randrange(Q).to_bytes(32, "big") for _ in range(4096))
    roots = R.roots_of_unity()
    for i in (0, 1, 2, 5, 2047, 2048, 4095):
        z = roots[i].to_bytes(32, "big")
        y = K.evaluate_polynomial_in_evaluation_form(blob, z, settings)
        assert y == blob[32 * i:32 * i + 32], "z = roots_of_unity[%d]" % i
        assert y == oracle.evaluate_polynomial(blob, z)
    for z in (0, 1, Q - 1, rnd.randrange(Q), rnd.randrange(Q)):
        zb = z.to_bytes(32, "big")
        assert K.evaluate_polynomial_in_evaluation_form(blob, zb, settings) == oracle.evaluate_polynomial(blob, zb)
    for c in vectors["verify_blob_kzg_proof"][:6]:
        b = vectors.blobs[c["blob"]]
        if len(b) == 131072 and len(unhex(c["commitment"])) == 48 and oracle.compute_challenge(b, unhex(c["commitment"])):
            assert K.compute_challenge(b, unhex(c["commitment"]), settings) == oracle.compute_challenge(b, unhex(c["commitment"]))
    with pytest.raises(K.KzgError) as e:
        K.evaluate_polynomial_in_evaluation_form(blob, Q.to_bytes(32, "big"), settings)              # z not canonical
    assert e.value.kind == "BadArgs"
    bad = bytearray(blob); bad[32 * 77:32 * 78] = (Q + 5).to_bytes(32, "big")
    with pytest.raises(K.KzgError) as e:
        K.evaluate_polynomial_in_evaluation_form(bytes(bad), (5).to_bytes(32, "big"), settings)      # Blob::as_polynomial fails
    assert e.value.kind == "BadArgs"


def _affine104(pt):
    """pyref affine point (or None) -> the reference's in-memory G1Affine: x, y Montgomery (6 x u64 LE), infinity byte, padding"""
    if pt is None:
        return bytes(48) + ((1 << 384) % P).to_bytes(48, "little") + bytes([1]) + bytes(7)
    return (pt[0] * (1 << 384) % P).to_bytes(48, "little") + (pt[1] * (1 << 384) % P).to_bytes(48, "little") + bytes(8)


def _scalar32(v):
    return (v * (1 << 256) % Q).to_bytes(32, "little")


def test_preparsed_verify_kzg_proof_batch(K, settings, vectors, oracle):
    """KzgProof::verify_kzg_proof_batch (src/kzg_proof.rs:399-444) on typed inputs in the reference's in-memory layout."""
    from kzg_rs_b200 import api
    from oracle import pyref as R
    for n in (2, 5):
        c = [x for x in vectors["verify_blob_kzg_proof_batch"] if x["output"] is True and len(x["blobs"]) == n][0]
        blobs = [vectors.blobs[i] for i in c["blobs"]]
        cs, ps = [unhex(x) for x in c["commitments"]], [unhex(x) for x in c["proofs"]]
        ok, rc, zs, ys, tr = oracle.verify_blob_kzg_proof_batch(blobs, cs, ps, want_trace=True)
        Cpts = [R.g1_from_compressed(x)[1] for x in cs]
        Ppts = [R.g1_from_compressed(x)[1] for x in ps]
        zi, yi = [int.from_bytes(z, "big") for z in zs], [int.from_bytes(y, "big") for y in ys]
        args = lambda C_, z_, y_, P_: (b"".join(_affine104(p) for p in C_), b"".join(_scalar32(v) for v in z_), b"".join(_scalar32(v) for v in y_),
                                       b"".join(_affine104(p) for p in P_), len(C_), settings)
        assert K.KzgProof.verify_kzg_proof_batch(*args(Cpts, zi, yi, Ppts)) is True
        got = api.last_batch_intermediates(settings)
        assert got["r"] == tr["r"]
        check_sums(got, tr)
        assert K.KzgProof.verify_kzg_proof_batch(*args(Cpts, zi, [yi[0] + 1] + yi[1:], Ppts)) is False
        assert K.KzgProof.verify_kzg_proof_batch(*args(Cpts, zi, yi, [R.G1_GEN] + Ppts[1:])) is False      # (the vectors' own proofs may all be the identity)
    # n = 1 and n = 0 (empty sums: both pairing arguments are the identity -> Ok(true))
    assert K.KzgProof.verify_kzg_proof_batch(*args(Cpts[:1], zi[:1], yi[:1], Ppts[:1])) is True
    assert K.KzgProof.verify_kzg_proof_batch(*args(Cpts[:1], zi[:1], [yi[0] + 1], Ppts[:1])) is False
    assert K.KzgProof.verify_kzg_proof_batch(b"", b"", b"", b"", 0, settings) is True
    # identity commitment and proof with y = 0: C - [0]G = O, pi = O -> both sums are the identity -> true (pairs are skipped)
    assert K.KzgProof.verify_kzg_proof_batch(*args([None, None], [3, 4], [0, 0], [None, None])) is True


def test_per_blob_verdicts(K, settings, oracle):
    """kzgb200_verify_blob_kzg_proof_batch_each: verdict i = verify_blob_kzg_proof(blob_i, C_i, pi_i) (src/kzg_proof.rs:446-470)."""
    from kzg_rs_b200.sharded import NOT_IN_G1
    n = 24
    _, (hb, hc, hp) = harness(K, settings, n, 0xeac4)
    V = K.KzgProof.verify_blob_kzg_proof_batch_each
    assert V(hb, hc, hp, n, settings) == [True] * n
    b, c, p = bytearray(hb), bytearray(hc), bytearray(hp)
    p[3 * 48:4 * 48] = hp[4 * 48:5 * 48]                                    # wrong proof -> Ok(false)
    b[7 * 131072 + 32 * 100:7 * 131072 + 32 * 101] = Q.to_bytes(32, "big")   # non-canonical element -> Err
    c[11 * 48:12 * 48] = NOT_IN_G1                                          # commitment outside the subgroup -> Err
    p[13 * 48:14 * 48] = bytes(48)                                          # not a compressed encoding -> Err
    b[20 * 131072 + 32 * 5 + 31] ^= 1                                       # blob no longer matches its commitment -> Ok(false)
    got, z, y = V(bytes(b), bytes(c), bytes(p), n, settings, want_zy=True)
    want = [True] * n
    want[3], want[7], want[11], want[13], want[20] = False, None, None, None, False
    assert got == want
    for i in (0, 3, 20, 23):       # z, y of parsable blobs against the oracle
        blob = bytes(b[131072 * i:131072 * (i + 1)])
        zi = oracle.compute_challenge(blob, bytes(c[48 * i:48 * i + 48]))
        assert z[32 * i:32 * i + 32] == zi and y[32 * i:32 * i + 32] == oracle.evaluate_polynomial(blob, zi)
    # each verdict against the oracle's single-blob function
    for i in range(n):
        assert got[i] == oracle.verify_blob_kzg_proof(bytes(b[131072 * i:131072 * (i + 1)]), bytes(c[48 * i:48 * i + 48]), bytes(p[48 * i:48 * i + 48])), i
    assert V(hb[:131072], hc[:48], hp[:48], 1, settings) == [True]
    assert V(hb[:131072], hc[:48], hp[48:96], 1, settings) == [False]
    assert V(b"", b"", b"", 0, settings) == []


def test_custom_trusted_setup_end_to_end(K, settings, oracle):
    """EnvKzgSettings::Custom (reference src/trusted_setup.rs:52-57): a context built from a NON-mainnet setup (toy tau,
    tests/golden/make_custom_setup.py) through KzgSettings.load_trusted_setup_file(path) -> kzgb200_create.  Commitments /
    proofs made under that setup (GPU commit/prove) verify under it, not under the mainnet setup, and agree with the oracle
    initialised from the same file."""
    import torch
    path = os.path.join(GOLDEN, "custom_setup.bin")
    custom = K.KzgSettings.load_trusted_setup_file(path)
    assert custom is not settings and custom.g2_monomial_bytes[96:192] != settings.g2_monomial_bytes[96:192]
    lib, ctx = K.Library.get().dll, custom.context(0)
    try:
        assert lib.kzgb200_load_g1_lagrange(ctx, custom.g1_lagrange_bytes, 4096) == 0
        rnd = random.Random(5)
        n = 3
        blobs = [b"".join(rnd.randrange(Q).to_bytes(32, "big") for _ in range(4096)) for _ in range(n)]
        d_b = torch.frombuffer(bytearray(b"".join(blobs)), dtype=torch.uint8).cuda()
        d_c = torch.empty(n * 48, dtype=torch.uint8, device="cuda"); d_p = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
        assert lib.kzgb200_blob_to_kzg_commitment_batch(ctx, d_b.data_ptr(), n, d_c.data_ptr()) == 0
        assert lib.kzgb200_compute_blob_kzg_proof_batch(ctx, d_b.data_ptr(), d_c.data_ptr(), n, d_p.data_ptr()) == 0
        hc, hp = d_c.cpu().numpy().tobytes(), d_p.cpu().numpy().tobytes()
        cs, ps = [hc[48 * i:48 * i + 48] for i in range(n)], [hp[48 * i:48 * i + 48] for i in range(n)]
        oracle.use_setup(path)
        try:
            for i in range(n):
                assert oracle.blob_to_kzg_commitment(blobs[i]) == cs[i] and oracle.compute_blob_kzg_proof(blobs[i], cs[i]) == ps[i]
            assert oracle.verify_blob_kzg_proof_batch(blobs, cs, ps) is True
            assert oracle.verify_blob_kzg_proof(blobs[0], cs[0], ps[0]) is True
        finally:
            oracle.use_setup(None)
        assert oracle.verify_blob_kzg_proof_batch(blobs, cs, ps) is False                       # mainnet oracle: other tau
        assert K.KzgProof.verify_blob_kzg_proof_batch(blobs, cs, ps, custom) is True
        assert K.KzgProof.verify_blob_kzg_proof(blobs[1], cs[1], ps[1], custom) is True
        assert K.KzgProof.verify_blob_kzg_proof_batch(blobs, cs, ps, settings) is False         # the mainnet context says no
        assert K.KzgProof.verify_blob_kzg_proof(blobs[1], cs[1], ps[1], settings) is False
        y = oracle.evaluate_polynomial(blobs[0], (9).to_bytes(32, "big"))
        # verify_kzg_proof under the custom setup: (C, z, y, pi) from the oracle's compute_kzg_proof on that setup
        oracle.use_setup(path)
        try:
            proof, y2 = oracle.compute_kzg_proof(blobs[0], (9).to_bytes(32, "big"))
        finally:
            oracle.use_setup(None)
        assert y2 == y
        assert K.KzgProof.verify_kzg_proof(cs[0], (9).to_bytes(32, "big"), y, proof, custom) is True
        assert K.KzgProof.verify_kzg_proof(cs[0], (9).to_bytes(32, "big"), y, proof, settings) is False
    finally:
        custom.close()
    # a setup whose G2 points do not decode is refused: Err(InvalidTrustedSetup)
    bad = K.KzgSettings(custom.g1_lagrange_bytes, bytes(192))
    with pytest.raises(K.KzgError) as e:
        bad.context(0)
    assert e.value.kind == "InvalidTrustedSetup"


@pytest.mark.parametrize("mode", ["stage", "direct", "register"])
def test_pageable_host_memory_paths(K, settings, oracle, mode):
    """An ordinary (pageable) Vec<Blob>: the pinned staging ring (default), the driver's staging and in-place pinning all give the
    same z, y, r as the oracle; 1100 blobs = two copy chunks, five staging buffers' worth."""
    import numpy as np
    from kzg_rs_b200 import api
    n = 1100
    _, (hb, hc, hp) = harness(K, settings, n, 0x9a9e)
    rc, ok_ref, z_ref, y_ref, tr = oracle.verify_batch_raw(hb, hc, hp, n, nthreads=os.cpu_count(), want_trace=True)
    arr = np.frombuffer(hb, dtype=np.uint8).copy()         # malloc'd, unpinned
    old = os.environ.get("KZGB200_PAGEABLE")
    os.environ["KZGB200_PAGEABLE"] = mode
    try:
        s = K.KzgSettings(settings.g1_lagrange_bytes, settings.g2_monomial_bytes)      # fresh context: the mode is read at creation
        ok, zs, ys = K.KzgProof.verify_blob_kzg_proof_batch_raw(arr.ctypes.data, n, hc, n, hp, n, s, want_zy=True)
        assert ok is True and b"".join(zs) == z_ref and b"".join(ys) == y_ref
        assert api.last_batch_intermediates(s)["r"] == tr["r"]
        s.close()
    finally:
        if old is None:
            os.environ.pop("KZGB200_PAGEABLE")
        else:
            os.environ["KZGB200_PAGEABLE"] = old


def _sum_partials(api, parts):
    """gathered per-rank partials -> what the final kernel forms: A, B', sum r_i y_i"""
    from oracle import pyref as R
    A = Bp = None
    s = 0
    for raw in parts:
        d = api.decode_partial(raw)
        A, Bp, s = R.g1_add(A, d["A"]), R.g1_add(Bp, d["B_prime"]), (s + d["sum_r_y"]) % Q
    return {"A": A, "B_prime": Bp, "sum_r_y": s}


def test_pairing_engine_16_lane_instructions_match_their_reference(K, settings):
    """csrc/vliw29.cuh: every generated engine program (Fp12 products and squarings, cyclotomic squaring, line products,
    Frobenius maps, inversion halves, G1 formulas) on the 16-lane cooperative executors versus the sequential reference
    executors -- the ones the CPU suite runs the reference's 114 verify_kzg_proof vectors on (tools/hosttest/vliw29_host.cu) --
    from the same seeded register files, programs chained so that the loose signed limb forms feed later programs; every
    register compared as a canonical field element after every program."""
    lib, ctx = K.Library.get().dll, settings.context(0)
    lib.kzgb200_debug_engine_selftest.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
    for seed in (1, 2, 3, 0xB200):
        mis = (C.c_uint32 * 32)()
        n = C.c_int(0)
        assert lib.kzgb200_debug_engine_selftest(ctx, seed, 2, mis, C.byref(n)) == 0
        assert n.value >= 12 and not any(mis[i] for i in range(32)), (seed, list(mis))


def test_group_across_all_visible_gpus(K, settings, oracle):
    """One process driving every visible GPU (kzgb200_group_create with device ids 0..N-1): the partial sums travel by peer
    stores over NVLink into the leader GPU.  Skipped on a single-GPU box (the two-contexts-on-one-GPU tests cover the protocol)."""
    import torch
    from kzg_rs_b200 import api
    ng = torch.cuda.device_count()
    if ng < 2:
        pytest.skip("needs >= 2 GPUs")
    n = 64 * ng + 24
    _, (hb, hc, hp) = harness(K, settings, n, 0x9d0 + ng)
    rc, ok_ref, z_ref, y_ref, tr = oracle.verify_batch_raw(hb, hc, hp, n, nthreads=os.cpu_count(), want_trace=True)
    with K.DeviceGroup.create(settings, list(range(ng)), 4096) as g:
        assert [g.uses_peer_stores(i) for i in range(ng)] == [True] * ng
        ok, z, y = g.verify_blob_kzg_proof_batch_raw(hb, n, hc, n, hp, n, want_zy=True)
        assert ok is True and z == z_ref and y == y_ref
        lib = K.Library.get().dll
        from kzg_rs_b200.sharded import shard_ranges
        active = sum(1 for lo, hi in shard_ranges(n, ng) if hi > lo)      # shards are multiples of 16 blobs: the last GPUs may stay idle
        for i in range(active):
            r = C.create_string_buffer(32)
            assert lib.kzgb200_last_r(g.context(i), r) == 0 and r.raw == tr["r"]
        check_sums(_sum_partials(api, g.last_partials(active)), tr)
        bad = bytearray(hp); bad[48 * (n - 1):48 * n] = hp[:48]
        assert g.verify_blob_kzg_proof_batch_raw(hb, n, hc, n, bytes(bad), n) is False
        badb = bytearray(hb); badb[(n - 1) * 131072 + 64:(n - 1) * 131072 + 96] = Q.to_bytes(32, "big")
        assert tri(lambda: g.verify_blob_kzg_proof_batch_raw(bytes(badb), n, hc, n, hp, n)) is None
        assert g.verify_blob_kzg_proof_batch_raw(hb, n, hc, n, hp, n) is True


@pytest.mark.parametrize("no_p2p", [0, 1])
def test_group_of_two_contexts_on_one_gpu(K, settings, oracle, no_p2p):
    """The multi-GPU path behind the ABI with world = 2 on cuda:0 (one process, two contexts): shards, transcript exchange, r,
    the gathered partials and the verdict against the oracle on the unsplit batch; deferred subgroup flags of the LAST shard;
    exact and tree transcripts; peer stores and the host path."""
    import ctypes as C
    from kzg_rs_b200 import api
    from kzg_rs_b200.sharded import NOT_IN_G1
    from oracle import pyref as R
    n = 80
    _, (hb, hc, hp) = harness(K, settings, n, 0x6209)
    rc, ok_ref, z_ref, y_ref, tr = oracle.verify_batch_raw(hb, hc, hp, n, nthreads=os.cpu_count(), want_trace=True)
    assert rc == 0 and ok_ref
    old = os.environ.get("KZGB200_GROUP_NO_P2P")
    os.environ["KZGB200_GROUP_NO_P2P"] = str(no_p2p)
    try:
        with K.DeviceGroup.create(settings, [0, 0], 4096) as g:
            assert g.world == 2 and g.local == 2 and g.uses_peer_stores(1) == (not no_p2p)
            ok, z, y = g.verify_blob_kzg_proof_batch_raw(hb, n, hc, n, hp, n, want_zy=True)
            assert ok is True and z == z_ref and y == y_ref
            lib = K.Library.get().dll
            for i in range(2):      # both ranks derived the oracle's r
                r = C.create_string_buffer(32)
                assert lib.kzgb200_last_r(g.context(i), r) == 0 and r.raw == tr["r"]
            parts = g.last_partials(2)
            assert all(api.decode_partial(p)["A"] is not None for p in parts)          # both shards contributed
            check_sums(_sum_partials(api, parts), tr)
            # negatives: a swapped proof in either shard -> false; bad field element / non-subgroup point in the LAST shard -> Err
            for i in (0, n - 2):
                bad = bytearray(hp); bad[48 * i:48 * i + 48], bad[48 * i + 48:48 * i + 96] = hp[48 * i + 48:48 * i + 96], hp[48 * i:48 * i + 48]
                assert g.verify_blob_kzg_proof_batch_raw(hb, n, hc, n, bytes(bad), n) is False
            badb = bytearray(hb); badb[(n - 1) * 131072 + 64:(n - 1) * 131072 + 96] = Q.to_bytes(32, "big")
            assert tri(lambda: g.verify_blob_kzg_proof_batch_raw(bytes(badb), n, hc, n, hp, n)) is None
            for which, idx in ((0, n - 1), (1, n - 3), (0, 2)):
                arrs = [bytearray(hc), bytearray(hp)]
                arrs[which][48 * idx:48 * idx + 48] = NOT_IN_G1
                assert tri(lambda: g.verify_blob_kzg_proof_batch_raw(hb, n, bytes(arrs[0]), n, bytes(arrs[1]), n)) is None
            assert g.verify_blob_kzg_proof_batch_raw(hb, n, hc, n, hp, n) is True          # flags do not leak into the next call
            with pytest.raises(K.KzgError) as e:
                g.verify_blob_kzg_proof_batch_raw(hb, n, hc, n - 1, hp, n)
            assert e.value.kind == "InvalidBytesLength"
            assert g.verify_blob_kzg_proof_batch_raw(b"", 0, b"", 0, b"", 0) is True
            assert g.verify_blob_kzg_proof_batch_raw(hb[:131072 * 3], 3, hc[:144], 3, hp[:144], 3) is True      # below 32 blobs: first GPU
            # tree transcript across the group: r of the restatement over the unsplit batch
            g.set_transcript_mode(api.TRANSCRIPT_TREE)
            assert g.verify_blob_kzg_proof_batch_raw(hb, n, hc, n, hp, n) is True
            r = C.create_string_buffer(32)
            assert lib.kzgb200_last_r(g.context(1), r) == 0
            r_tree = R.tree_transcript_r([hc[48 * i:48 * i + 48] for i in range(n)], [int.from_bytes(z_ref[32 * i:32 * i + 32], "big") for i in range(n)],
                                         [int.from_bytes(y_ref[32 * i:32 * i + 32], "big") for i in range(n)], [hp[48 * i:48 * i + 48] for i in range(n)])
            assert r.raw == r_tree.to_bytes(32, "big")
    finally:
        if old is None:
            os.environ.pop("KZGB200_GROUP_NO_P2P")
        else:
            os.environ["KZGB200_GROUP_NO_P2P"] = old


_RANK_SCRIPT = r'''
import ctypes as C, json, os, sys
sys.path.insert(0, sys.argv[1])
rank, world, session, n_total = int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], int(sys.argv[5])
import torch
import kzg_rs_b200 as K
from kzg_rs_b200 import api
from kzg_rs_b200.sharded import ShardedBatch, shard_ranges
S = K.KzgSettings.load_trusted_setup_file()
lib = K.Library.get().dll
data = open(sys.argv[6], "rb").read()
hb, hc, hp = data[:n_total * 131072], data[n_total * 131072:n_total * 131120], data[n_total * 131120:]
lo, hi = shard_ranges(n_total, world)[rank]
n = hi - lo
plan = ShardedBatch(lib, S, n, rank, world, 0, session)
dev = lambda b: torch.frombuffer(bytearray(b), dtype=torch.uint8).cuda()
d_b, d_c, d_p = dev(hb[lo * 131072:hi * 131072]), dev(hc[lo * 48:hi * 48]), dev(hp[lo * 48:hi * 48])
torch.cuda.synchronize()
out = {"peer": plan.group.uses_peer_stores(0)}
out["device"] = plan.verify_device(d_b, d_c, d_p)
if rank == 0:
    out["partials"] = [p.hex() for p in plan.group.last_partials(world)]
z, y = plan.last_zy_host(n)
out["z"], out["y"] = z.hex(), y.hex()
r = C.create_string_buffer(32); lib.kzgb200_last_r(plan.ctx, r); out["r"] = r.raw.hex()
h = lambda b: torch.frombuffer(bytearray(b), dtype=torch.uint8)
out["host"] = plan.verify_host(h(hb[lo * 131072:hi * 131072]), h(hc[lo * 48:hi * 48]), h(hp[lo * 48:hi * 48]))
out["neg"] = plan.check_negatives(d_b, d_c, d_p)
plan.close()
print("RESULT " + json.dumps(out))
'''


def test_group_of_two_processes_on_one_gpu(K, settings, oracle, tmp_path):
    """kzgb200_group_join: one process per rank (as under torchrun), both on cuda:0: rendezvous in POSIX shared memory, the
    leader's exchange buffer mapped through CUDA IPC, collective calls.  z, y, r, partials and verdicts against the oracle."""
    import json
    from kzg_rs_b200 import api
    n = 96
    _, (hb, hc, hp) = harness(K, settings, n, 0x2b0c)
    rc, ok_ref, z_ref, y_ref, tr = oracle.verify_batch_raw(hb, hc, hp, n, nthreads=os.cpu_count(), want_trace=True)
    data = tmp_path / "batch.bin"
    data.write_bytes(hb + hc + hp)
    script = tmp_path / "rank.py"
    script.write_text(_RANK_SCRIPT)
    session = "pytest%d" % os.getpid()
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, str(k), "2", session, str(n), str(data)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
             for k in range(2)]
    outs = []
    for p in procs:
        so, se = p.communicate(timeout=600)
        assert p.returncode == 0, se[-3000:]
        outs.append(json.loads([l for l in so.splitlines() if l.startswith("RESULT ")][0][7:]))
    assert "".join(o["z"] for o in outs) == z_ref.hex() and "".join(o["y"] for o in outs) == y_ref.hex()
    for o in outs:
        assert o["device"] is True and o["host"] is True and o["r"] == tr["r"].hex()
        assert o["neg"] == {"swapped_proofs_verdict": False, "element_equal_to_modulus": "Err(BadArgs)", "commitment_outside_subgroup": "Err(BadArgs)"}
    check_sums(_sum_partials(api, [bytes.fromhex(x) for x in outs[0]["partials"]]), tr)
