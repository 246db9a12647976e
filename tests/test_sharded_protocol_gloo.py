"""The N>1 path on CPU: world_size 2 under gloo.  The orchestration code is the product's own
(kzg_rs_b200/sharded.py: ShardedBatch -- shard ranges, the two allgathers, offsets of the r powers); the four
per-rank phases are supplied by an oracle-backed stand-in with the same interface as the GPU backend, so the
test pins the protocol: a batch split over 2 ranks gives the verdict of the unsplit batch, for a valid batch, a
corrupted proof on either rank, and an unparsable input on one rank.
"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
Q = 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001


class OracleBackend:
    """CPU stand-in for GpuBackend: same methods, same payload sizes (64 B per blob of zy, 352 B partials)."""

    def __init__(self, late_flags=False):
        # late_flags: the rank's own error flags are known only at finalize (deferred subgroup checks in the CUDA backend)
        self.late_flags = late_flags
        sys.path.insert(0, ROOT)
        from oracle import oracle as O
        from oracle import pyref as R
        self.O, self.R = O, R

    @staticmethod
    def _b(t):
        return t.numpy().tobytes()

    def evaluate(self, blobs, cs, ps, n, zy_out):
        O = self.O
        hb, hc, hp = self._b(blobs), self._b(cs), self._b(ps)
        self.c = [hc[48 * i:48 * i + 48] for i in range(n)]
        self.p = [hp[48 * i:48 * i + 48] for i in range(n)]
        self.err, out, self.z, self.y = 0, b"", [], []
        for i in range(n):
            blob = hb[131072 * i:131072 * (i + 1)]
            okc, okp = O.g1_check(self.c[i]), O.g1_check(self.p[i])
            z = O.compute_challenge(blob, self.c[i]) if okc and okc[0] else None
            y = O.evaluate_polynomial(blob, z) if z else None
            if not (okc and okc[0] and okp and okp[0] and y):
                self.err, z, y = 1, z or bytes(32), y or bytes(32)
            self.z.append(int.from_bytes(z, "big")); self.y.append(int.from_bytes(y, "big"))
            out += self.z[-1].to_bytes(32, "little") + self.y[-1].to_bytes(32, "little")
        zy_out.copy_(torch.frombuffer(bytearray(out), dtype=torch.uint8))

    def challenge(self, all_c, all_zy, all_p, n_total):
        c, zy, p = self._b(all_c), self._b(all_zy), self._b(all_p)
        try:
            r = self.O.compute_r_powers([c[48 * i:48 * i + 48] for i in range(n_total)],
                                        [zy[64 * i:64 * i + 32][::-1] for i in range(n_total)],
                                        [zy[64 * i + 32:64 * i + 64][::-1] for i in range(n_total)],
                                        [p[48 * i:48 * i + 48] for i in range(n_total)])
            self.r = int.from_bytes(r[0], "big")
        except Exception:
            self.r = 1
        if r is None:
            self.r = 1

    def lincomb(self, offset, partial_out):
        O, n = self.O, len(self.c)
        be = lambda v: v.to_bytes(32, "big")
        if self.err:
            A = Bp = bytes([0xc0]) + bytes(47); s = 0
        else:
            ri = [pow(self.r, offset + i, Q) for i in range(n)]
            A = O.g1_lincomb(self.p, [be(x) for x in ri])
            Bp = O.g1_lincomb(self.c + self.p, [be(x) for x in ri] + [be(x * z % Q) for x, z in zip(ri, self.z)])
            s = sum(x * y for x, y in zip(ri, self.y)) % Q
        raw = A + Bp + be(s) + (0 if self.late_flags else self.err).to_bytes(4, "little")
        partial_out.copy_(torch.frombuffer(bytearray(raw + bytes(352 - len(raw))), dtype=torch.uint8))

    def finalize(self, partials, world):
        O, R = self.O, self.R
        if self.late_flags and self.err:
            return None
        raw = self._b(partials)
        parts = [raw[352 * k:352 * (k + 1)] for k in range(world)]
        if any(int.from_bytes(p[128:132], "little") for p in parts):
            return None
        one = (1).to_bytes(32, "big")
        A = O.g1_lincomb([p[:48] for p in parts], [one] * world)
        Bp = O.g1_lincomb([p[48:96] for p in parts], [one] * world)
        s = sum(int.from_bytes(p[96:128], "big") for p in parts) % Q
        gen = R.g1_to_compressed(R.G1_GEN)
        rhs = O.g1_lincomb([Bp, gen], [one, ((Q - s) % Q).to_bytes(32, "big")])
        return O.pairings_verify(A, 1, rhs, 0)     # e(-A, [tau]G2) e(rhs, G2) == 1

    def sync_collectives(self):
        pass


def _worker(rank, world, port, cases, results):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    sys.path.insert(0, ROOT)
    from kzg_rs_b200.sharded import ShardedBatch
    out = []
    for late in (False, True):
        for blobs, cs, ps in cases:
            n_local = len(cs) // 48 // world
            sl = lambda raw, k: torch.frombuffer(bytearray(raw[k * n_local * rank:k * n_local * (rank + 1)]), dtype=torch.uint8)
            plan = ShardedBatch(None, None, n_local, rank, world, dist, device=torch.device("cpu"), backend=OracleBackend(late))
            out.append(plan.verify_device(sl(blobs, 131072), sl(cs, 48), sl(ps, 48)))
    results[rank] = out
    dist.destroy_process_group()


def test_two_rank_sharded_batch_matches_unsharded(vectors, oracle):
    from conftest import unhex
    c = [c for c in vectors["verify_blob_kzg_proof_batch"] if c["output"] is True and len(c["blobs"]) == 4][0]
    blobs = b"".join(vectors.blobs[i] for i in c["blobs"])
    cs = b"".join(unhex(x) for x in c["commitments"])
    ps = b"".join(unhex(x) for x in c["proofs"])
    # a valid G1 point that is not the proof: the blob's own commitment
    wrong = lambda i: ps[:48 * i] + cs[48 * i:48 * i + 48] + ps[48 * i + 48:]
    bad_c = cs[:48 * 3] + bytes([0x81]) + bytes(range(1, 48))          # unparsable commitment on rank 1
    cases = [(blobs, cs, ps),                       # valid
             (blobs, cs, wrong(1)),                 # wrong proof in rank 0's shard
             (blobs, cs, wrong(3)),                 # wrong proof in rank 1's shard
             (blobs, bad_c, ps)]
    want = [oracle.verify_blob_kzg_proof_batch([blobs[131072 * i:131072 * (i + 1)] for i in range(4)],
                                               [k[48 * i:48 * i + 48] for i in range(4)], [p[48 * i:48 * i + 48] for i in range(4)])
            for _, k, p in cases]
    assert want == [True, False, False, None]
    ctx = mp.get_context("spawn")
    results = ctx.Manager().dict()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, cases, results)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert results[0] == want + want and results[1] == want + want      # flags in the partial / flags known only at finalize
