#!/usr/bin/env python3
"""A NON-mainnet trusted setup for the custom-setup tests (EnvKzgSettings::Custom, reference src/trusted_setup.rs:52-57): the
same shape as src/trusted_setup.txt (4096 Lagrange G1 points, 65 monomial G2 points) for the toy secret tau below, written in the
"KZGS" container tests/golden/make_fixtures.py uses (kzg_rs_b200.KzgSettings and the oracle both load it).
    g1_lagrange[i] = [L_i(tau)] G1,   L_i(X) = (X^n - 1) / n * w^i / (X - w^i)   (natural order; loaders bit-reverse)
    g2_monomial[j] = [tau^j] G2
Output: tests/golden/custom_setup.bin (203 KB).  Pure-Python group arithmetic (oracle/pyref.py): takes about a minute.
    python tests/golden/make_custom_setup.py
"""
import os
import struct
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import pyref as R

TAU = int.from_bytes(b"kzg-rs_b200 custom setup: NOT a secret", "big") % R.Q
N = 4096


def g2_to_compressed(pt):
    if pt is None:
        return bytes([0xc0]) + bytes(95)
    (x0, x1), (y0, y1) = pt
    largest = (y1 > (R.P - 1) // 2) if y1 else (y0 > (R.P - 1) // 2)
    out = bytearray(x1.to_bytes(48, "big") + x0.to_bytes(48, "big"))
    out[0] |= 0x80 | (0x20 if largest else 0)
    return bytes(out)


def main():
    w = R.ROOT_OF_UNITY_4096
    zn1 = (pow(TAU, N, R.Q) - 1) % R.Q
    inv_n = pow(N, -1, R.Q)
    g1 = []
    wi = 1
    for i in range(N):
        li = zn1 * inv_n % R.Q * wi % R.Q * pow((TAU - wi) % R.Q, -1, R.Q) % R.Q
        g1.append(R.g1_to_compressed(R.g1_mul(R.G1_GEN, li)))
        wi = wi * w % R.Q
    g2, t = [], 1
    for j in range(65):
        g2.append(g2_to_compressed(R.g2_mul(R.G2_GEN, t)))
        t = t * TAU % R.Q
    # sanity: sum of the Lagrange points is the generator (sum_i L_i = 1), and the G2 points round-trip
    acc = None
    for b in g1:
        acc = R.g1_add(acc, R.g1_from_compressed(b, check_subgroup=False)[1])
    assert R.g1_to_compressed(acc) == R.g1_to_compressed(R.G1_GEN)
    assert R.g2_from_compressed(g2[1]) == (True, R.g2_mul(R.G2_GEN, TAU))
    out = os.path.join(ROOT, "tests", "golden", "custom_setup.bin")
    with open(out, "wb") as fh:
        fh.write(b"KZGS" + struct.pack("<II", N, 65) + b"".join(g1) + b"".join(g2))
    print("wrote", out)


if __name__ == "__main__":
    main()
