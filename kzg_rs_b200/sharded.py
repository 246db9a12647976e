"""Blob-sharded verify_blob_kzg_proof_batch: one rank (process) per GPU, contiguous blob ranges per rank.

The reference's batch verification (src/kzg_proof.rs:472-525 -> :399-444) has two places where all blobs
meet: the transcript hash that yields r (compute_r_powers, :291-348) and the final sums that feed the single
pairing check (:419-441).  Everything else is per blob.  So a rank
    1. evaluates its shard (parse C/pi, canonicity, z_i, y_i)                        kzgb200_shard_evaluate
    2. allgathers (C_i, z_i, y_i, pi_i) -- 160 B per blob -- and derives the same r   kzgb200_shard_challenge
    3. forms its partial sums with r^(offset+i)                                       kzgb200_shard_lincomb
    4. allgathers the 352-byte partials and runs the final pairing check              kzgb200_shard_finalize
torch.distributed (NCCL over NVLink) only moves the two small payloads; the 128 KiB blobs never leave their GPU.
"""
import ctypes as C
import math

import torch

BLOB = 131072
PARTIAL_BYTES = 352
FR_MODULUS_BE = bytes.fromhex("73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001")
NOT_IN_G1 = bytes.fromhex("8123456789abcdef" + "0123456789abcdef" * 5)      # decompresses to a curve point of the wrong order


class GpuBackend:
    """The four phases on libkzgb200.so (device pointers)."""

    def __init__(self, lib, ctx):
        self.lib, self.ctx = lib, ctx

    def _check(self, rc):
        if rc == 1:
            return None           # Err(BadArgs)
        if rc:
            raise RuntimeError("kzgb200 rc=%d: %s" % (rc, self.lib.kzgb200_last_error(self.ctx).decode()))
        return True

    def batch(self, blobs, cs, ps, n, z_out, y_out):
        ok = C.c_int(-1)
        rc = self.lib.kzgb200_verify_blob_kzg_proof_batch_device(self.ctx, blobs.data_ptr(), cs.data_ptr(), ps.data_ptr(), n, C.byref(ok),
                                                                 z_out.data_ptr(), y_out.data_ptr())
        return self._check(rc) and bool(ok.value)

    def evaluate(self, blobs, cs, ps, n, zy_out):
        self._check(self.lib.kzgb200_shard_evaluate(self.ctx, blobs.data_ptr(), cs.data_ptr(), ps.data_ptr(), n, zy_out.data_ptr()))

    def evaluate_host(self, h_blobs, h_cs, h_ps, n, c_out, p_out, zy_out):
        self._check(self.lib.kzgb200_shard_evaluate_host(self.ctx, h_blobs.data_ptr(), h_cs.data_ptr(), h_ps.data_ptr(), n,
                                                         c_out.data_ptr(), p_out.data_ptr(), zy_out.data_ptr()))

    def challenge(self, all_c, all_zy, all_p, n_total):
        self._check(self.lib.kzgb200_shard_challenge(self.ctx, all_c.data_ptr(), all_zy.data_ptr(), all_p.data_ptr(), n_total))

    def lincomb(self, offset, partial_out):
        self._check(self.lib.kzgb200_shard_lincomb(self.ctx, offset, partial_out.data_ptr()))

    def finalize(self, partials, world):
        ok = C.c_int(-1)
        rc = self.lib.kzgb200_shard_finalize(self.ctx, partials.data_ptr(), world, C.byref(ok))
        return self._check(rc) and bool(ok.value)

    def sync_collectives(self):
        torch.cuda.current_stream().synchronize()


class ShardedBatch:
    """Orchestration of one sharded batch; `backend` supplies the four phases (GpuBackend in production; the CPU
    tests drive the same code with an oracle-backed stand-in under gloo)."""

    def __init__(self, lib, ctx, n_local, rank=0, world=1, dist=None, device=None, backend=None):
        self.lib, self.ctx, self.n, self.rank, self.world, self.dist = lib, ctx, n_local, rank, world, dist
        self.backend = backend or GpuBackend(lib, ctx)
        dev = device or torch.device("cuda", torch.cuda.current_device())
        u8 = dict(dtype=torch.uint8, device=dev)
        self.z_out = torch.empty(n_local * 32, **u8)
        self.y_out = torch.empty(n_local * 32, **u8)
        if world > 1:
            self.zy = torch.empty(n_local * 64, **u8)
            self.all_c = torch.empty(world * n_local * 48, **u8)
            self.all_p = torch.empty(world * n_local * 48, **u8)
            self.all_zy = torch.empty(world * n_local * 64, **u8)
            self.partial = torch.empty(PARTIAL_BYTES, **u8)
            self.partials = torch.empty(world * PARTIAL_BYTES, **u8)
            self.stage = None
        self.launches_per_step = self.count_launches(n_local, resident=True, tree=True)

    @staticmethod
    def count_launches(n, resident, tree):
        """Kernels of libkzgb200.so launched by one batch on one rank (mirrors launch_phase1 / advance_transcript /
        launch_lincomb in csrc/kzgb200.cu): G1 decompress + subgroup, per chunk challenge + evaluation, export of z/y,
        transcript (tree: words + leaf per chunk, then root; exact: schedule + chain per chunk), 5 MSM kernels, pairing, flag merge."""
        if resident:
            chunks = 1 if (tree or n < 4096) else math.ceil(n / max(1024, math.ceil(n / 8)))
        else:
            chunks = math.ceil(n / max(1024, math.ceil(n / 64)))
        transcript = (2 * chunks + 1) if tree else 2 * chunks
        return 2 + 2 * chunks + 1 + transcript + 5 + 2

    def _check(self, rc):
        if rc == 1:
            return None           # Err(BadArgs)
        if rc:
            raise RuntimeError("kzgb200 rc=%d: %s" % (rc, self.lib.kzgb200_last_error(self.ctx).decode()))
        return True

    def verify_device(self, d_blobs, d_cs, d_ps):
        """Inputs resident in HBM.  Returns True / False / None (= Err(BadArgs))."""
        be = self.backend
        be.sync_collectives()     # the library runs on its own stream: the caller's pending writes to the inputs must be done
        if self.world == 1:
            return be.batch(d_blobs, d_cs, d_ps, self.n, self.z_out, self.y_out)
        be.evaluate(d_blobs, d_cs, d_ps, self.n, self.zy)
        return self._exchange_and_finish(d_cs, d_ps)

    def _exchange_and_finish(self, d_cs, d_ps):
        be, dist = self.backend, self.dist
        dist.all_gather_into_tensor(self.all_c, d_cs)
        dist.all_gather_into_tensor(self.all_p, d_ps)
        dist.all_gather_into_tensor(self.all_zy, self.zy)
        be.sync_collectives()
        be.challenge(self.all_c, self.all_zy, self.all_p, self.n * self.world)
        be.lincomb(self.rank * self.n, self.partial)
        dist.all_gather_into_tensor(self.partials, self.partial)
        be.sync_collectives()
        res = be.finalize(self.partials, self.world)
        # a rank's own subgroup checks may finish after its partial was exported (they run beside the tail): agree on the
        # outcome -- Err(BadArgs) on any rank is Err(BadArgs) for the batch (reference: first failure aborts, src/kzg_proof.rs:503-516)
        code = torch.tensor([2 if res is None else int(res)], dtype=torch.int32, device=self.partials.device)
        dist.all_reduce(code, op=dist.ReduceOp.MAX)
        worst = int(code.item())
        return None if worst == 2 else (res if worst == int(bool(res)) else bool(worst))

    def verify_host(self, h_blobs, h_cs, h_ps):
        """Inputs in (pinned) host memory; host->device copies are part of the call."""
        if self.world == 1:
            ok = C.c_int(-1)
            rc = self.lib.kzgb200_verify_blob_kzg_proof_batch(self.ctx, h_blobs.data_ptr(), self.n, h_cs.data_ptr(), self.n,
                                                              h_ps.data_ptr(), self.n, C.byref(ok), None, None)
            return self._check(rc) and bool(ok.value)
        if self.stage is None:
            dev = self.z_out.device
            self.stage = (torch.empty(self.n * 48, dtype=torch.uint8, device=dev), torch.empty(self.n * 48, dtype=torch.uint8, device=dev))
        self.backend.evaluate_host(h_blobs, h_cs, h_ps, self.n, self.stage[0], self.stage[1], self.zy)
        return self._exchange_and_finish(*self.stage)

    def last_zy_host(self, m):
        """(z bytes, y bytes), 32-byte big-endian each, of the first m blobs of the last single-GPU call."""
        return self.z_out[:m * 32].cpu().numpy().tobytes(), self.y_out[:m * 32].cpu().numpy().tobytes()

    def check_negatives(self, d_blobs, d_cs, d_ps):
        """Corrupted-proof batch -> false; non-canonical field element -> Err(BadArgs).  Restores the inputs."""
        res = {}
        if self.n >= 2:
            a, b = d_ps[:48].clone(), d_ps[48:96].clone()
            if self.rank == 0:
                d_ps[:48], d_ps[48:96] = b, a
            res["swapped_proofs_verdict"] = self.verify_device(d_blobs, d_cs, d_ps)
            d_ps[:48], d_ps[48:96] = a, b
        pos = 5 * 32 if self.n == 1 else BLOB + 7 * 32
        saved = d_blobs[pos:pos + 32].clone()
        if self.rank == 0:
            d_blobs[pos:pos + 32] = torch.tensor(list(FR_MODULUS_BE), dtype=torch.uint8, device=d_blobs.device)
        r = self.verify_device(d_blobs, d_cs, d_ps)
        res["element_equal_to_modulus"] = "Err(BadArgs)" if r is None else r
        d_blobs[pos:pos + 32] = saved
        # a commitment that is on the curve but outside the subgroup (c-kzg vector invalid_commitment_32afa9561a4b3b91), in the
        # LAST rank's shard: found by the deferred subgroup checks, must be Err(BadArgs) on every rank
        k = (self.n - 1) * 48
        saved = d_cs[k:k + 48].clone()
        if self.rank == self.world - 1:
            d_cs[k:k + 48] = torch.tensor(list(NOT_IN_G1), dtype=torch.uint8, device=d_cs.device)
        r = self.verify_device(d_blobs, d_cs, d_ps)
        res["commitment_outside_subgroup"] = "Err(BadArgs)" if r is None else r
        d_cs[k:k + 48] = saved
        return res
