"""Blob-sharded verify_blob_kzg_proof_batch, one rank (process) per GPU: thin driver over the library's group API.

The reference's batch verification (src/kzg_proof.rs:472-525 -> :399-444) has two places where all blobs meet: the transcript
hash that yields r (compute_r_powers, :291-348) and the final sums that feed the single pairing check (:419-441).  Both
exchanges live INSIDE libkzgb200.so (csrc/group.cu): transcript entries through a shared host block hashed by the leader,
partial sums stored by the reduction kernels straight into the leader GPU's memory over NVLink.  This module only picks the
session name, calls kzgb200_group_join once and then the collective verify call; torch.distributed is not on the data path.
"""
import ctypes as C
import os

import torch

from .api import DeviceGroup, KzgError

BLOB = 131072
PARTIAL_BYTES = 352
FR_MODULUS_BE = bytes.fromhex("73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001")
NOT_IN_G1 = bytes.fromhex("8123456789abcdef" + "0123456789abcdef" * 5)      # decompresses to a curve point of the wrong order


def shard_ranges(n_total, world, align=16):
    """Contiguous blob ranges per rank (the split kzgb200_group_verify_blob_kzg_proof_batch uses): equal shards rounded up to
    `align` blobs, the last rank takes the remainder; ranks beyond the data get empty ranges."""
    per = -(-(-(-n_total // world)) // align) * align if n_total else 0
    out = []
    for k in range(world):
        lo = min(k * per, n_total)
        out.append((lo, min(lo + per, n_total)))
    return out


def session_name():
    """One name per job: the rendezvous port torchrun already hands to every rank (or the parent pid when run by hand)."""
    return "%s_%s" % (os.environ.get("MASTER_PORT", "0"), os.environ.get("TORCHELASTIC_RUN_ID", str(os.getppid())))


class ShardedBatch:
    """One rank's view of a sharded batch: this rank's n_local blobs on its GPU.  world == 1 runs the plain single-GPU entry."""

    def __init__(self, lib, settings, n_local, rank=0, world=1, device=0, session=None, cap=None):
        self.lib, self.n, self.rank, self.world, self.device = lib, n_local, rank, world, device
        self.group = None
        if world > 1:
            # cap = blobs per rank the shared block is sized for: must be the same on every rank
            self.group = DeviceGroup.join(settings, session or session_name(), rank, world, device, cap or n_local)
            self.ctx = self.group.context(0)
        else:
            self.ctx = settings.context(device)
        dev = torch.device("cuda", device)
        self.z_out = torch.empty(n_local * 32, dtype=torch.uint8, device=dev)
        self.y_out = torch.empty(n_local * 32, dtype=torch.uint8, device=dev)

    def set_transcript_mode(self, mode):
        self.lib.kzgb200_set_transcript_mode(self.ctx, mode)

    def _check(self, rc):
        if rc == 1:
            return None           # Err(BadArgs)
        if rc:
            raise RuntimeError("kzgb200 rc=%d: %s" % (rc, self.lib.kzgb200_last_error(self.ctx).decode()))
        return True

    def verify_device(self, d_blobs, d_cs, d_ps):
        """Inputs resident in HBM (complete: the library works on its own streams).  True / False / None (= Err(BadArgs))."""
        if self.world == 1:
            ok = C.c_int(-1)
            rc = self.lib.kzgb200_verify_blob_kzg_proof_batch_device(self.ctx, d_blobs.data_ptr(), d_cs.data_ptr(), d_ps.data_ptr(), self.n,
                                                                     C.byref(ok), self.z_out.data_ptr(), self.y_out.data_ptr())
            return self._check(rc) and bool(ok.value)
        try:
            return self.group.verify_shards([d_blobs], [d_cs], [d_ps], [self.n], True, [self.z_out], [self.y_out])
        except KzgError as e:
            if e.kind == "BadArgs":
                return None
            raise

    def verify_host(self, h_blobs, h_cs, h_ps):
        """Inputs in host memory (pinned or pageable); host->device copies are part of the call."""
        if self.world == 1:
            ok = C.c_int(-1)
            rc = self.lib.kzgb200_verify_blob_kzg_proof_batch(self.ctx, h_blobs.data_ptr(), self.n, h_cs.data_ptr(), self.n,
                                                              h_ps.data_ptr(), self.n, C.byref(ok), None, None)
            return self._check(rc) and bool(ok.value)
        try:
            return self.group.verify_shards([h_blobs], [h_cs], [h_ps], [self.n], False)
        except KzgError as e:
            if e.kind == "BadArgs":
                return None
            raise

    def last_zy_host(self, m):
        """(z bytes, y bytes), 32-byte big-endian each, of the first m blobs of this rank's last device-input call."""
        return self.z_out[:m * 32].cpu().numpy().tobytes(), self.y_out[:m * 32].cpu().numpy().tobytes()

    def check_negatives(self, d_blobs, d_cs, d_ps):
        """Corrupted-proof batch -> false; non-canonical field element -> Err(BadArgs).  Restores the inputs."""
        res = {}
        if self.n >= 2:
            a, b = d_ps[:48].clone(), d_ps[48:96].clone()
            if self.rank == 0:
                d_ps[:48], d_ps[48:96] = b, a
            torch.cuda.synchronize()
            res["swapped_proofs_verdict"] = self.verify_device(d_blobs, d_cs, d_ps)
            d_ps[:48], d_ps[48:96] = a, b
        pos = 5 * 32 if self.n == 1 else BLOB + 7 * 32
        saved = d_blobs[pos:pos + 32].clone()
        if self.rank == 0:
            d_blobs[pos:pos + 32] = torch.tensor(list(FR_MODULUS_BE), dtype=torch.uint8, device=d_blobs.device)
        torch.cuda.synchronize()
        r = self.verify_device(d_blobs, d_cs, d_ps)
        res["element_equal_to_modulus"] = "Err(BadArgs)" if r is None else r
        d_blobs[pos:pos + 32] = saved
        # a commitment that is on the curve but outside the subgroup (c-kzg vector invalid_commitment_32afa9561a4b3b91), in the
        # LAST rank's shard: found by the deferred subgroup checks, must be Err(BadArgs) on every rank
        k = (self.n - 1) * 48
        saved = d_cs[k:k + 48].clone()
        if self.rank == self.world - 1:
            d_cs[k:k + 48] = torch.tensor(list(NOT_IN_G1), dtype=torch.uint8, device=d_cs.device)
        torch.cuda.synchronize()
        r = self.verify_device(d_blobs, d_cs, d_ps)
        res["commitment_outside_subgroup"] = "Err(BadArgs)" if r is None else r
        d_cs[k:k + 48] = saved
        torch.cuda.synchronize()
        return res

    def close(self):
        if self.group is not None:
            self.group.close()
            self.group = None
