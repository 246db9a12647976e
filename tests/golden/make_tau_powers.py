#!/usr/bin/env python3
"""[tau^j]G1 for j < 16, derived from the mainnet Lagrange setup with the CPU oracle:
    [tau^j]G1 = sum_i w_i^j * L_i(tau)G      (the commitment to X^j)
Output: kzg_rs_b200/data/tau_powers_g1.bin (16 x 48 bytes, compressed).  The workload generator uses them to
commit to / open low-degree polynomials with a handful of scalar multiplications per blob.
Checks: j = 0 gives the generator; e([tau]G1, G2) == e(G1, [tau]G2).
    python tests/golden/make_tau_powers.py
"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import oracle as O
from oracle import pyref as R

pts = [O.tau_power_g1(j) for j in range(16)]
assert pts[0] == R.g1_to_compressed(R.G1_GEN)
assert O.pairings_verify(pts[1], 0, pts[0], 1) is True
# [tau^2]G via the pairing too: e(T2, G2) == e(T1, tauG2)
assert O.pairings_verify(pts[2], 0, pts[1], 1) is True
with open(os.path.join(ROOT, "kzg_rs_b200", "data", "tau_powers_g1.bin"), "wb") as fh:
    fh.write(b"".join(pts))
print("wrote", len(pts), "points")
