"""The CUDA sources are __host__ __device__ down to the carry-chain primitives, so the product's own field / curve /
pairing / cooperative-engine code is unit-tested here on the CPU (nvcc host compile of the same headers):
  * Montgomery products (single, fused dual, dedicated squaring), add, sub vs Python integers, edge values included;
  * per-proof verification logic (decompression, subgroup check, G1-side pairing equation with precomputed lines)
    and the cooperative pairing engines + binary-GCD inversion vs the reference's 114 well-formed verify_kzg_proof vectors
    (vliw_host: the 12 x 32-limb engine of the many-tuple path; vliw29_host: the 29-bit signed-limb engine of the batch path on
    its sequential reference executors, plus its Montgomery products and conversions against the 12 x 32 field layer);
  * the generated engine programs vs the Python oracle (tools/gen_vliw.py self-test).
"""
import os
import random
import shutil
import subprocess
import sys

import pytest
from conftest import unhex

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "kzg_rs_b200", "csrc")
pytestmark = pytest.mark.skipif(shutil.which("nvcc") is None, reason="nvcc not found")


def build(tmp, name):
    exe = os.path.join(tmp, name)
    subprocess.check_call(["nvcc", "-std=c++17", "-O1", "-x", "cu", "-I", CSRC, os.path.join(ROOT, "tools", "hosttest", name + ".cu"),
                           "-o", exe], stderr=subprocess.DEVNULL)
    return exe


def test_montgomery_arithmetic_matches_python(tmp_path):
    from oracle.pyref import P, Q
    exe = build(str(tmp_path), "field_host")
    rnd = random.Random(5)
    lines, exp = [], []
    for name, mod, nb in (("r", Q, 256), ("p", P, 384)):
        rinv = pow(1 << nb, -1, mod)
        special = [0, 1, 2, mod - 1, mod - 2, (1 << nb) % mod, mod >> 1, (1 << (nb - 32)) - 1]
        for _ in range(1500):
            v = [rnd.choice(special) if rnd.random() < 0.2 else rnd.randrange(mod) for _ in range(4)]
            a, b, c, d = v
            w = nb // 4
            lines.append("%s %0*x %0*x %0*x %0*x" % (name, w, a, w, b, w, c, w, d))
            exp.append((a * b * rinv % mod, (a * b + c * d) * rinv % mod, (a + b) % mod, (a - b) % mod, a * a * rinv % mod, c * c * rinv % mod))
    out = subprocess.run([exe], input="\n".join(lines) + "\n", capture_output=True, text=True).stdout.split("\n")
    for l, e, o in zip(lines, exp, out):
        assert tuple(int(x, 16) for x in o.split()) == e, l[:40]


def _vector_records(vectors):
    recs, want = [], []
    for c in vectors["verify_kzg_proof"]:
        a = [unhex(c[k]) for k in ("commitment", "z", "y", "proof")]
        if [len(x) for x in a] != [48, 32, 32, 48]:
            continue
        recs.append(b"".join(a))
        want.append({True: "1", False: "0", None: "2"}[c["output"]])
    return b"".join(recs), "".join(want)


@pytest.mark.parametrize("prog", ["verify_host", "vliw_host", "vliw29_host"])
def test_verification_logic_on_reference_vectors(tmp_path, vectors, prog):
    exe = build(str(tmp_path), prog)
    recs, want = _vector_records(vectors)
    out = subprocess.run([exe, os.path.join(ROOT, "kzg_rs_b200", "data", "mainnet_setup.bin")], input=recs, capture_output=True)
    assert out.returncode == 0 and out.stdout.decode().strip() == want and len(want) == 114


def test_engine_programs_match_python_oracle():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_vliw.py"), "--check"], capture_output=True, text=True)
    assert out.returncode == 0 and "self-test of all programs: ok" in out.stdout
