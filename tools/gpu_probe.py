"""Quick device-side probe: generate n synthetic blobs with the harness, check a few against the oracle,
time the phases of the single-GPU batch path."""
import ctypes as C, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import kzg_rs_b200 as K
from kzg_rs_b200.api import Library
from oracle import oracle as O

lib = Library.get().dll
S = K.KzgSettings.load_trusted_setup_file()
ctx = S.context(0)
tau = open(os.path.join(os.path.dirname(K.__file__), "data", "tau_powers_g1.bin"), "rb").read()
names = ["parse", "challenge", "eval", "transcript", "lincomb", "reduce", "final"]
lib.kzgb200_set_profiling(ctx, 1)
for n in [int(x) for x in (sys.argv[1:] or ["64", "1024", "4096"])]:
    blobs = torch.empty(n * 131072, dtype=torch.uint8, device="cuda")
    cs = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
    ps = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
    t = time.time()
    rc = lib.kzgb200_harness_generate(ctx, 0x4B5A47, n, 8, tau, blobs.data_ptr(), cs.data_ptr(), ps.data_ptr())
    torch.cuda.synchronize()
    print("n", n, "harness rc", rc, "gen s", round(time.time() - t, 3))
    if n <= 64:
        hb, hc, hp = blobs.cpu().numpy().tobytes(), cs.cpu().numpy().tobytes(), ps.cpu().numpy().tobytes()
        for i in (0, n - 1):
            b = hb[i * 131072:(i + 1) * 131072]
            assert O.blob_to_kzg_commitment(b) == hc[i * 48:(i + 1) * 48], "commitment mismatch"
            assert O.compute_blob_kzg_proof(b, hc[i * 48:(i + 1) * 48]) == hp[i * 48:(i + 1) * 48], "proof mismatch"
        print("harness commitments/proofs match oracle")
        rc_, ok_, z_, y_ = O.verify_batch_raw(hb, hc, hp, n, nthreads=8)
        print("oracle verdict", rc_, ok_)
    ok = C.c_int(-1)
    zo = torch.empty(n * 32, dtype=torch.uint8, device="cuda"); yo = torch.empty(n * 32, dtype=torch.uint8, device="cuda")
    for it in range(2):
        t = time.time()
        rc = lib.kzgb200_verify_blob_kzg_proof_batch_device(ctx, blobs.data_ptr(), cs.data_ptr(), ps.data_ptr(), n, C.byref(ok), zo.data_ptr(), yo.data_ptr())
        dt = time.time() - t
        ph = (C.c_float * 7)(); lib.kzgb200_get_phase_ms(ctx, ph)
        print("  rc", rc, "ok", ok.value, "wall ms", round(dt * 1e3, 2), {k: round(v, 3) for k, v in zip(names, ph)})
    if n <= 64:
        assert zo.cpu().numpy().tobytes() == z_ and yo.cpu().numpy().tobytes() == y_, "z/y mismatch vs oracle"
        print("z,y match oracle for all", n)
    # negative: swap two proofs
    ps2 = ps.clone(); ps2[:48] = ps[48:96]; ps2[48:96] = ps[:48]
    rc = lib.kzgb200_verify_blob_kzg_proof_batch_device(ctx, blobs.data_ptr(), cs.data_ptr(), ps2.data_ptr(), n, C.byref(ok), None, None)
    print("  swapped proofs: rc", rc, "ok", ok.value)
    try:
        lib.kzgb200_debug_final_ticks.argtypes = [C.c_void_p, C.POINTER(C.c_longlong)]
        tk = (C.c_longlong * 14)()
        lib.kzgb200_verify_blob_kzg_proof_batch_device(ctx, blobs.data_ptr(), cs.data_ptr(), ps.data_ptr(), n, C.byref(ok), None, None)
        lib.kzgb200_debug_final_ticks(ctx, tk)
        names2 = ["prelude(G1 sums, [s]G, affine)", "miller loop", "final exp easy", "final exp hard", "compare"]
        print("  final kernel sections (us @1.965GHz):", {k: round((tk[i + 1] - tk[i]) / 1965.0, 1) for i, k in enumerate(names2)})
        print("  engine levels (lane 0): LIN body %.0f us, LIN barrier wait %.0f us over %d levels; MUL body %.0f us, MUL wait %.0f us over %d levels"
              % (tk[8] / 1965.0, tk[9] / 1965.0, tk[12], tk[10] / 1965.0, tk[11] / 1965.0, tk[13]))
        print("  slot copies %.0f us, line loads %.0f us (thread 0, barrier included)" % (tk[6] / 1965.0, tk[7] / 1965.0))
    except Exception as e:
        print("ticks unavailable", e)
