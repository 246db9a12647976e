"""Secondary BASELINE.json configs on one B200 (not the bench.py headline):
  config 2: verify_blob_kzg_proof latency on one blob          config 3: 64-blob batch
  config 5: m independent verify_kzg_proof tuples (valid tuples C = aG + bT, y = a + bz, pi = bG built with the oracle)
Prints one JSON line."""
import ctypes as C, json, os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kzg_rs_b200 as K
from oracle import oracle as O
from oracle import pyref as R

lib = K.Library.get().dll
S = K.KzgSettings.load_trusted_setup_file()
ctx = S.context(0)
tau = open(os.path.join(os.path.dirname(K.__file__), "data", "tau_powers_g1.bin"), "rb").read()
out = {}


def gen(n, seed=0x4B5A47):
    b = torch.empty(n * 131072, dtype=torch.uint8, device="cuda"); c = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
    p = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
    assert lib.kzgb200_harness_generate(ctx, seed, n, 8, tau, b.data_ptr(), c.data_ptr(), p.data_ptr()) == 0
    return [t.cpu().numpy().tobytes() for t in (b, c, p)]


def med(fn, reps):
    ts = []
    for _ in range(reps):
        t = time.perf_counter(); fn(); ts.append(time.perf_counter() - t)
    ts.sort()
    return ts[len(ts) // 2]


hb, hc, hp = gen(1)
assert K.KzgProof.verify_blob_kzg_proof(hb, hc, hp, S) is True
out["config2_single_blob_latency_ms"] = 1e3 * med(lambda: K.KzgProof.verify_blob_kzg_proof(hb, hc, hp, S), 30)
t = time.perf_counter(); assert O.verify_blob_kzg_proof(hb, hc, hp) is True
out["config2_cpu_oracle_ms"] = 1e3 * (time.perf_counter() - t)
hb, hc, hp = gen(64)
assert K.KzgProof.verify_blob_kzg_proof_batch_raw(hb, 64, hc, 64, hp, 64, S) is True
out["config3_batch64_ms"] = 1e3 * med(lambda: K.KzgProof.verify_blob_kzg_proof_batch_raw(hb, 64, hc, 64, hp, 64, S), 20)
t = time.perf_counter(); O.verify_batch_raw(hb, hc, hp, 64, nthreads=os.cpu_count())
out["config3_cpu_oracle_ms_all_cores"] = 1e3 * (time.perf_counter() - t)
# config 5: tuples from a small pool of valid (C, z, y, pi), ~1 % wrong y, ~0.1 % malformed z
rnd = random.Random(5)
G, T = R.g1_to_compressed(R.G1_GEN), tau[48:96]
pool = []
for _ in range(32):
    a, b, z = rnd.randrange(R.Q), rnd.randrange(R.Q), rnd.randrange(R.Q)
    be = lambda v: v.to_bytes(32, "big")
    Cc = O.g1_lincomb([G, T], [be(a), be(b)]); pi = O.g1_lincomb([G], [be(b)])
    pool.append((Cc, be(z), be((a + b * z) % R.Q), pi))
m = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
idx = [rnd.randrange(32) for _ in range(m)]
cs = bytearray(b"".join(pool[i][0] for i in idx)); zs = bytearray(b"".join(pool[i][1] for i in idx))
ys = bytearray(b"".join(pool[i][2] for i in idx)); ps = bytearray(b"".join(pool[i][3] for i in idx))
want = bytearray([1] * m)
for k in range(0, m, 97):
    ys[32 * k + 31] ^= 1; want[k] = 0
for k in range(50, m, 997):
    zs[32 * k:32 * k + 32] = b"\xff" * 32; want[k] = 2
got = K.KzgProof.verify_kzg_proof_many(bytes(cs), bytes(zs), bytes(ys), bytes(ps), m, S)
assert got == bytes(want), "config 5 verdict mismatch"
dt = med(lambda: K.KzgProof.verify_kzg_proof_many(bytes(cs), bytes(zs), bytes(ys), bytes(ps), m, S), 3)
out["config5_tuples"] = m
out["config5_checks_per_s"] = m / dt
t = time.perf_counter()
for i in range(16):
    O.verify_kzg_proof(*pool[i])
out["config5_cpu_oracle_checks_per_s_1thread"] = 16 / (time.perf_counter() - t)
print(json.dumps(out))
