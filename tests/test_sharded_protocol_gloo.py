"""The N>1 path on CPU: world_size 2 under gloo.  The multi-GPU exchanges live inside libkzgb200.so (csrc/group.cu): transcript
entries of every rank flow through a POSIX shared-memory block and the leader hashes them in global order into r.  That
host-side protocol needs no GPU, so it runs here for real: two processes (rendezvous and session name over gloo), each passing
the entries of its shard -- z, y from the CPU oracle -- through kzgb200_group_host_protocol_test; both must receive the digest
whose reduction mod q is the oracle's r for the UNSPLIT batch (reference src/kzg_proof.rs:291-348).  Also: ragged shards,
many chunks, and the shard-range helper.
"""
import ctypes as C
import hashlib
import os
import random
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Q = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001


def _worker(rank, world, port, case_index, results):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from conftest import Vectors, unhex
        from oracle import oracle as O
        from kzg_rs_b200.api import Library
        from kzg_rs_b200.sharded import shard_ranges
        lib = Library.get().dll
        name = [None]
        if rank == 0:
            name[0] = "t%d_%d" % (os.getpid(), port)
        dist.broadcast_object_list(name, src=0)
        out = {}
        # (1) a golden batch vector split over the ranks: z, y of the rank's own blobs from the oracle
        V = Vectors()
        case = [c for c in V["verify_blob_kzg_proof_batch"] if c["output"] is True and len(c["blobs"]) >= 4][case_index]
        n = len(case["blobs"])
        lo, hi = shard_ranges(n, world, align=1)[rank]
        cs = [unhex(x) for x in case["commitments"]][lo:hi]
        ps = [unhex(x) for x in case["proofs"]][lo:hi]
        zy = b""
        for i, c in zip(case["blobs"][lo:hi], cs):
            z = O.compute_challenge(V.blobs[i], c)
            y = O.evaluate_polynomial(V.blobs[i], z)
            zy += z[::-1] + y[::-1]                      # little-endian, as the kernels leave them (kzg_proof.rs:320-328)
        digest = C.create_string_buffer(32)
        rc = lib.kzgb200_group_host_protocol_test((name[0] + "a").encode(), rank, world, b"".join(cs), zy, b"".join(ps), hi - lo, 1, digest)
        out["vector"] = (rc, digest.raw, n)
        # (2) ragged synthetic shards, many chunks: the digest is SHA-256 of the reference's byte layout over ALL entries
        rnd = random.Random(1234)
        sizes = [1000, 777]
        allc, allzy, allp = [], [], []
        for k in range(world):
            allc.append(rnd.randbytes(48 * sizes[k])); allzy.append(rnd.randbytes(64 * sizes[k])); allp.append(rnd.randbytes(48 * sizes[k]))
        rc = lib.kzgb200_group_host_protocol_test((name[0] + "b").encode(), rank, world, allc[rank], allzy[rank], allp[rank], sizes[rank], 96, digest)
        total = sum(sizes)
        msg = b"RCKZGBATCH___V1_" + (4096).to_bytes(8, "big") + total.to_bytes(8, "big")
        for k in range(world):
            for i in range(sizes[k]):
                msg += allc[k][48 * i:48 * i + 48] + allzy[k][64 * i:64 * i + 64] + allp[k][48 * i:48 * i + 48]
        out["synthetic"] = (rc, digest.raw == hashlib.sha256(msg).digest())
        results[rank] = out
    finally:
        dist.destroy_process_group()


def _run(case_index):
    mgr = mp.Manager()
    results = mgr.dict()
    port = 29650 + (os.getpid() + case_index) % 200
    mp.spawn(_worker, args=(2, port, case_index, results), nprocs=2, join=True)
    return dict(results)


@pytest.mark.parametrize("case_index", [0, 1])
def test_two_rank_transcript_exchange_gives_the_unsplit_r(case_index, vectors, oracle):
    from conftest import unhex
    res = _run(case_index)
    case = [c for c in vectors["verify_blob_kzg_proof_batch"] if c["output"] is True and len(c["blobs"]) >= 4][case_index]
    ok, rc, zs, ys, tr = oracle.verify_blob_kzg_proof_batch([vectors.blobs[i] for i in case["blobs"]], [unhex(x) for x in case["commitments"]],
                                                           [unhex(x) for x in case["proofs"]], want_trace=True)
    assert ok is True
    for rank in (0, 1):
        rc, digest, n = res[rank]["vector"]
        assert rc == 0 and n == len(case["blobs"])
        assert (int.from_bytes(digest, "big") % Q).to_bytes(32, "big") == tr["r"], "rank %d: r differs from the unsplit batch's" % rank
        assert res[rank]["synthetic"] == (0, True)


def test_shard_ranges():
    sys.path.insert(0, ROOT)
    from kzg_rs_b200.sharded import shard_ranges
    assert shard_ranges(16384, 8) == [(2048 * k, 2048 * (k + 1)) for k in range(8)]
    assert shard_ranges(100, 4) == [(0, 32), (32, 64), (64, 96), (96, 100)]
    assert shard_ranges(40, 4) == [(0, 16), (16, 32), (32, 40), (40, 40)]
    assert shard_ranges(0, 2) == [(0, 0), (0, 0)]
    for n, w in ((16384, 3), (5000, 7), (33, 2)):
        r = shard_ranges(n, w)
        assert r[0][0] == 0 and r[-1][1] == n and all(a[1] == b[0] for a, b in zip(r, r[1:]))
        assert all(lo % 16 == 0 for lo, _ in r if lo < n)
