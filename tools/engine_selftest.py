"""GPU self-test of the pairing engine's 16-lane instructions against the sequential reference executors (kzgb200_debug_engine_selftest)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import kzg_rs_b200 as K
from kzg_rs_b200.api import Library
lib = Library.get().dll
S = K.KzgSettings.load_trusted_setup_file()
ctx = S.context(0)
lib.kzgb200_debug_engine_selftest.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.POINTER(C.c_uint32), C.POINTER(C.c_int)]
bad = 0
for seed in range(int(sys.argv[1]) if len(sys.argv) > 1 else 4):
    mis = (C.c_uint32 * 32)(); n = C.c_int(0)
    print("seed", seed, "...", flush=True)
    rc = lib.kzgb200_debug_engine_selftest(ctx, seed, int(sys.argv[2]) if len(sys.argv) > 2 else 2, mis, C.byref(n))
    print("seed", seed, "rc", rc, "mismatches per program", list(mis)[:n.value])
    bad += rc != 0 or any(mis[i] for i in range(n.value))
print("engine selftest:", "FAIL" if bad else "ok")
sys.exit(1 if bad else 0)
