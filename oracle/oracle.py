"""ctypes binding of the C oracle (oracle/kzg_oracle.c).  TEST INFRASTRUCTURE ONLY.

May be imported by tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs; never by the
product package.  Mirrors the reference API (kzg_proof.rs:353-525): tri-state results are returned
as True / False / None (None = the reference would return Err).
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
RC_OK, RC_BADARGS, RC_INTERNAL, RC_BADLEN = 0, 1, 2, 3
SETUP_BIN = os.path.join(os.path.dirname(_HERE), "kzg_rs_b200", "data", "mainnet_setup.bin")


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "libkzg_oracle.so")
        if not os.path.exists(path):
            build()
        _LIB = C.CDLL(path)
        _LIB.kzgo_init.argtypes = [C.c_char_p, C.c_size_t]
        with open(SETUP_BIN, "rb") as fh:
            raw = fh.read()
        rc = _LIB.kzgo_init(raw, len(raw))
        if rc:
            raise RuntimeError("oracle setup load failed rc=%d" % rc)
        _LIB.kzgo_verify_blob_kzg_proof_batch.argtypes = [
            C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.c_int,
            C.POINTER(C.c_int), C.c_char_p, C.c_char_p, C.c_char_p]
        _LIB.kzgo_compute_r_powers.argtypes = [C.c_char_p] * 4 + [C.c_size_t, C.c_char_p, C.c_char_p]
        _LIB.kzgo_g1_lincomb.argtypes = [C.c_char_p, C.c_char_p, C.c_size_t, C.c_int, C.c_char_p]
        _LIB.kzgo_sha256.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p]
    return _LIB


def _tri(rc, ok):
    return None if rc else bool(ok.value)


def verify_kzg_proof(commitment, z, y, proof):
    if len(commitment) != 48 or len(z) != 32 or len(y) != 32 or len(proof) != 48:
        return None  # Bytes48/Bytes32::from_slice length error (dtypes.rs:19-29)
    ok = C.c_int(0)
    rc = lib().kzgo_verify_kzg_proof(bytes(commitment), bytes(z), bytes(y), bytes(proof), C.byref(ok))
    return _tri(rc, ok)


def verify_blob_kzg_proof(blob, commitment, proof, want_zy=False):
    if len(blob) != 131072 or len(commitment) != 48 or len(proof) != 48:
        return (None, None, None) if want_zy else None
    ok = C.c_int(0)
    z, y = C.create_string_buffer(32), C.create_string_buffer(32)
    rc = lib().kzgo_verify_blob_kzg_proof(bytes(blob), bytes(commitment), bytes(proof), C.byref(ok), z, y)
    if want_zy:
        return (_tri(rc, ok), None if rc else z.raw, None if rc else y.raw)
    return _tri(rc, ok)


def verify_blob_kzg_proof_batch(blobs, commitments, proofs, nthreads=1, want_trace=False):
    """blobs/commitments/proofs: sequences of bytes objects (Vec<Blob>, Vec<Bytes48>, Vec<Bytes48>).
    Returns verdict, or (verdict, rc, z_list, y_list, trace) with want_trace."""
    if any(len(b) != 131072 for b in blobs) or any(len(c) != 48 for c in commitments) or any(len(p) != 48 for p in proofs):
        return (None, RC_BADLEN, None, None, None) if want_trace else None
    n = len(blobs)
    ok = C.c_int(0)
    z, y = C.create_string_buffer(32 * max(n, 1)), C.create_string_buffer(32 * max(n, 1))
    tr = C.create_string_buffer(128)
    rc = lib().kzgo_verify_blob_kzg_proof_batch(b"".join(blobs), n, b"".join(commitments), len(commitments),
                                                b"".join(proofs), len(proofs), nthreads, C.byref(ok), z, y, tr)
    if not want_trace:
        return _tri(rc, ok)
    if rc:
        return (None, rc, None, None, None)
    zs = [z.raw[32 * i:32 * i + 32] for i in range(n)]
    ys = [y.raw[32 * i:32 * i + 32] for i in range(n)]
    trace = None if n < 2 else {"r": tr.raw[:32], "proof_lincomb": tr.raw[32:80], "rhs_g1": tr.raw[80:128]}
    return (bool(ok.value), rc, zs, ys, trace)


def verify_batch_raw(blobs, commitments, proofs, n, nthreads=1, want_trace=False):
    """Contiguous-buffer variant used for timing: returns (rc, ok, z_bytes, y_bytes) [+ trace dict r / proof_lincomb / rhs_g1]."""
    ok = C.c_int(0)
    z, y = C.create_string_buffer(32 * n), C.create_string_buffer(32 * n)
    tr = C.create_string_buffer(128) if want_trace else None
    rc = lib().kzgo_verify_blob_kzg_proof_batch(blobs, n, commitments, n, proofs, n, nthreads, C.byref(ok), z, y, tr)
    if want_trace:
        return rc, bool(ok.value), z.raw, y.raw, {"r": tr.raw[:32], "proof_lincomb": tr.raw[32:80], "rhs_g1": tr.raw[80:128]}
    return rc, bool(ok.value), z.raw, y.raw


def use_setup(path=None):
    """Re-initialise the oracle with another trusted setup ("KZGS" container); None = back to the mainnet setup."""
    with open(path or SETUP_BIN, "rb") as fh:
        raw = fh.read()
    rc = lib().kzgo_init(raw, len(raw))
    if rc:
        raise RuntimeError("oracle setup load failed rc=%d" % rc)


def compute_challenge(blob, commitment):
    z = C.create_string_buffer(32)
    rc = lib().kzgo_compute_challenge(bytes(blob), bytes(commitment), z)
    return None if rc else z.raw


def evaluate_polynomial(blob, z):
    y = C.create_string_buffer(32)
    rc = lib().kzgo_evaluate_polynomial(bytes(blob), bytes(z), y)
    return None if rc else y.raw


def compute_r_powers(commitments, zs, ys, proofs):
    n = len(commitments)
    r, pw = C.create_string_buffer(32), C.create_string_buffer(32 * max(n, 1))
    rc = lib().kzgo_compute_r_powers(b"".join(commitments), b"".join(zs), b"".join(ys), b"".join(proofs), n, r, pw)
    if rc:
        return None
    return r.raw, [pw.raw[32 * i:32 * i + 32] for i in range(n)]


def g1_check(b48):
    f, s = C.c_int(0), C.c_int(0)
    rc = lib().kzgo_g1_check(bytes(b48), C.byref(f), C.byref(s))
    return None if rc else (bool(f.value), bool(s.value))


def g1_lincomb(points, scalars, use_msm=True):
    out = C.create_string_buffer(48)
    rc = lib().kzgo_g1_lincomb(b"".join(points), b"".join(scalars), len(points), int(use_msm), out)
    return None if rc else out.raw


def pairings_verify(a1, a2_idx, b1, b2_idx):
    ok = C.c_int(0)
    rc = lib().kzgo_pairings_verify(bytes(a1), a2_idx, bytes(b1), b2_idx, C.byref(ok))
    return None if rc else bool(ok.value)


def blob_to_kzg_commitment(blob):
    out = C.create_string_buffer(48)
    rc = lib().kzgo_blob_to_kzg_commitment(bytes(blob), out)
    return None if rc else out.raw


def compute_blob_kzg_proof(blob, commitment):
    out = C.create_string_buffer(48)
    rc = lib().kzgo_compute_blob_kzg_proof(bytes(blob), bytes(commitment), out)
    return None if rc else out.raw


def compute_kzg_proof(blob, z):
    out, y = C.create_string_buffer(48), C.create_string_buffer(32)
    rc = lib().kzgo_compute_kzg_proof(bytes(blob), bytes(z), out, y)
    return None if rc else (out.raw, y.raw)


def tau_power_g1(j):
    out = C.create_string_buffer(48)
    rc = lib().kzgo_tau_power_g1(j, out)
    return None if rc else out.raw


def sha256(msg, portable=False):
    out = C.create_string_buffer(32)
    lib().kzgo_sha256_force_portable(int(portable))
    lib().kzgo_sha256(bytes(msg), len(msg), out)
    lib().kzgo_sha256_force_portable(0)
    return out.raw
