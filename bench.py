#!/usr/bin/env python3
"""bench.py -- blobs verified / second through verify_blob_kzg_proof_batch (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--blobs B] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

A step = one verify_blob_kzg_proof_batch over B synthetic blobs per GPU (default 16384 = 2 GiB of blob
bytes per GPU, BASELINE.json configs[3]); with N ranks the batch is N*B blobs sharded by contiguous ranges,
(z,y) and the per-rank partial sums are exchanged with NCCL allgathers and every rank runs the final pairing
check ("scaling": "weak").  `value` times the step with inputs resident in HBM; `e2e` times the same call
from pinned HOST buffers (host->device copies inside the timed region).  Timing: CUDA events on the library's
stream, barrier + synchronize on both sides, max over ranks.

--impl reference times the reference's CPU algorithm (the C oracle port, kind "port": the Rust crate cannot be
built here) with all host threads on a bounded sample of the same workload.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
BLOB = 131072
ALGO_BYTES_PER_BLOB = 131072 + 48 + 48
# dram__bytes_read.sum + dram__bytes_write.sum per launch at 16384 blobs from the ncu --set full captures under profiles/
TRAFFIC_BYTES = {"challenge_sha256": 2148375000 + 4876800, "evaluate_barycentric": 2154585000 + 4713984}   # profiles/ncu_full_r01_summary.txt
PHASES = ["parse_g1", "challenge_sha256", "evaluate_barycentric", "transcript_r", "lincomb_terms", "reduce", "final_pairing"]


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--blobs", type=int, default=16384, help="blobs per GPU")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample", type=int, default=8192, help="blobs in the CPU-baseline sample (~19 s of CPU work)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-pipeline", action="store_true", help="skip the streaming front-end measurement")
    ap.add_argument("--inflight", type=int, default=2, help="batches in flight in the streaming front-end measurement")
    ap.add_argument("--lowdegree-blobs", action="store_true",
                    help="blobs = evaluations of random degree<8 polynomials (cheap generator) instead of uniformly random "
                         "blobs committed / proved by the GPU commit/prove path")
    return ap.parse_args()


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        sm = sorted(int(r[1]) for r in self.rows if r[1].isdigit())
        mx = [int(r[2]) for r in self.rows if r[2].isdigit()]
        reasons = set()
        for r in self.rows:
            for name, v in zip(["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"], r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows)}


def bind_to_gpu_numa_node(index):
    """Pin this rank's host threads (and hence its first-touch pinned allocations) to the CPUs local to its GPU, so that
    the end-to-end leg's host->device copies do not cross the socket interconnect.  Best effort."""
    try:
        bus = subprocess.run(["nvidia-smi", "-i", str(index), "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             capture_output=True, text=True, timeout=10).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        cpus = open("/sys/bus/pci/devices/%s/local_cpulist" % bus).read().strip()
        ids = set()
        for part in cpus.split(","):
            a, _, b = part.partition("-")
            ids.update(range(int(a), int(b or a) + 1))
        if ids:
            os.sched_setaffinity(0, ids)
            return cpus
    except Exception:
        pass
    return None


def run_reference(args, rank, world):
    """The reference's CPU path (oracle port), all host threads, bounded sample per step."""
    if rank != 0:
        return
    from oracle import oracle as O
    import torch
    O.build()
    cores = os.cpu_count() or 1
    n = min(args.blobs, args.cpu_sample)
    blobs, cs, ps = make_workload_host(args, n)
    times = []
    for i in range(args.warmup + args.steps):
        t = time.perf_counter()
        rc, ok, _, _ = O.verify_batch_raw(blobs, cs, ps, n, nthreads=cores)
        dt = time.perf_counter() - t
        assert rc == 0 and ok, "reference arm rejected a valid batch"
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * sum(times) / len(times)
    val = n / (ms / 1e3)
    sample = "%d of %d blobs per step (same generator/seed), %d threads, SHA-NI=%d" % (n, args.blobs, cores, O.lib().kzgo_sha256_uses_shani())
    print(json.dumps({
        "impl": "reference", "metric": "blobs verified/sec (verify_blob_kzg_proof_batch)", "value": val, "unit": "blobs/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "u32-limb modular integer (Fr 255-bit / Fp 381-bit) + SHA-256", "data": "synthetic",
        "config": workload_config(args, n, world),
        "cpu_baseline": {"value": val, "unit": "blobs/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": "blobs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def workload_config(args, n, world):
    return {"workload": "verify_blob_kzg_proof_batch on %d synthetic blobs per GPU (%d total, %.2f GiB of blob bytes per GPU), "
                        "BASELINE.json configs[3]" % (n, n * world, n * BLOB / 2**30),
            "blobs_per_gpu": n, "total_blobs": n * world, "parallelism": "blob-sharded x%d" % world,
            "generator": ("uniformly random field elements, commitments/proofs by the GPU commit/prove path over the mainnet setup" if not args.lowdegree_blobs
                          else "harness: random degree<8 polynomials in evaluation form, commitments/proofs over the mainnet setup, seed 0x4B5A47"),
            "cache": "inputs (>= 2 GiB per step) are larger than the 126 MB L2"}


def make_workload_device(lib, ctx, args, n, rank):
    import torch
    import kzg_rs_b200 as K
    tau = open(os.path.join(os.path.dirname(K.__file__), "data", "tau_powers_g1.bin"), "rb").read()
    blobs = torch.empty(n * BLOB, dtype=torch.uint8, device="cuda")
    cs = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
    ps = torch.empty(n * 48, dtype=torch.uint8, device="cuda")
    if not getattr(args, "lowdegree_blobs", False):
        S = K.KzgSettings.load_trusted_setup_file()
        assert lib.kzgb200_load_g1_lagrange(ctx, S.g1_lagrange_bytes, 4096) == 0
        g = torch.Generator(device="cuda").manual_seed(0x4B5A47 + rank)
        blobs = torch.randint(0, 256, (n * BLOB,), dtype=torch.uint8, device="cuda", generator=g)
        blobs.view(n * 4096, 32)[:, 0] &= 0x3f          # every element < 2^254 < q
        torch.cuda.synchronize()                         # the library works on its own stream
        assert lib.kzgb200_blob_to_kzg_commitment_batch(ctx, blobs.data_ptr(), n, cs.data_ptr()) == 0
        assert lib.kzgb200_compute_blob_kzg_proof_batch(ctx, blobs.data_ptr(), cs.data_ptr(), n, ps.data_ptr()) == 0
        return blobs, cs, ps
    rc = lib.kzgb200_harness_generate(ctx, 0x4B5A47 + rank, n, 8, tau, blobs.data_ptr(), cs.data_ptr(), ps.data_ptr())
    assert rc == 0, "harness failed rc=%d" % rc
    return blobs, cs, ps


def make_workload_host(args, n):
    """Reference arm: same generator (GPU harness) when a GPU is present, else oracle-side commit/prove."""
    import torch
    if torch.cuda.is_available():
        import kzg_rs_b200 as K
        from kzg_rs_b200.api import Library
        S = K.KzgSettings.load_trusted_setup_file()
        b, c, p = make_workload_device(Library.get().dll, S.context(0), args, n, 0)
        return b.cpu().numpy().tobytes(), c.cpu().numpy().tobytes(), p.cpu().numpy().tobytes()
    from oracle import oracle as O
    import random
    rnd, Q = random.Random(0x4B5A47), 0x73eda753299d7d483339d80809a1d80553bda402fffe5bfeffffffff00000001
    uniq = []
    for _ in range(min(n, 4)):
        blob = b"".join(rnd.randrange(Q).to_bytes(32, "big") for _ in range(4096))
        c = O.blob_to_kzg_commitment(blob)
        uniq.append((blob, c, O.compute_blob_kzg_proof(blob, c)))
    seq = [uniq[i % len(uniq)] for i in range(n)]
    return b"".join(x[0] for x in seq), b"".join(x[1] for x in seq), b"".join(x[2] for x in seq)


def main():
    args = parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import kzg_rs_b200 as K
    from kzg_rs_b200.api import Library
    from kzg_rs_b200 import sharded
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    lib = Library.get().dll
    S = K.KzgSettings.load_trusted_setup_file()
    ctx = S.context(local_rank)
    n = args.blobs
    d_blobs, d_cs, d_ps = make_workload_device(lib, ctx, args, n, rank)
    # pinned host copies for the end-to-end leg
    h_blobs = torch.empty(n * BLOB, dtype=torch.uint8).pin_memory()
    h_cs = torch.empty(n * 48, dtype=torch.uint8).pin_memory()
    h_ps = torch.empty(n * 48, dtype=torch.uint8).pin_memory()
    h_blobs.copy_(d_blobs); h_cs.copy_(d_cs); h_ps.copy_(d_ps)
    torch.cuda.synchronize()
    stream = torch.cuda.ExternalStream(lib.kzgb200_stream(ctx), device=torch.device("cuda", local_rank))
    plan = sharded.ShardedBatch(lib, ctx, n, rank, world, dist)
    lib.kzgb200_set_profiling(ctx, 1)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_resident():
        ok = plan.verify_device(d_blobs, d_cs, d_ps)
        assert ok is True, "valid batch rejected"

    def step_e2e():
        ok = plan.verify_host(h_blobs, h_cs, h_ps)
        assert ok is True, "valid batch rejected (e2e)"

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        phase_acc = [0.0] * 7
        for _ in range(steps):
            fn()
            ph = (C.c_float * 7)()
            lib.kzgb200_get_phase_ms(ctx, ph)
            phase_acc = [a + b for a, b in zip(phase_acc, ph)]
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1) / steps
        if dist is not None:
            t = torch.tensor([ms], device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, [a / steps for a in phase_acc]

    # Headline = the throughput configuration: KZGB200_TRANSCRIPT_TREE (batch challenge r hashed as a 3-level tree; verdict, z,
    # y identical to kzg-rs, r itself not).  The library default (EXACT: r bit-identical, one serial SHA-256 chain over all
    # blobs of all ranks) is timed in the same run and reported under "exact_transcript".
    sampler = ClockSampler(local_rank)
    sampler.start()
    res = {}
    for mode_name, mode in (("tree", 1), ("exact", 0)):
        lib.kzgb200_set_transcript_mode(ctx, mode)
        ms, phases = timed(step_resident, args.steps, max(args.warmup, 3))
        ms_e2e, _ = timed(step_e2e, args.steps, max(args.warmup, 3))
        res[mode_name] = (ms, phases, ms_e2e)
    # Streaming front-end (kzgb200_pipeline_*), single GPU: the same K batches through two contexts, two in
    # flight, so the latency-bound tail of one batch runs under the head / the PCIe copy of the next.  Reported beside the
    # headline (which stays one blocking call at a time).
    pipelined = None
    if world == 1 and not args.no_pipeline:
        pipe_batches = max(args.steps, 12)      # enough batches for the ramp-up and drain not to dominate
        with K.BatchPipeline(S, depth=args.inflight, device=local_rank, transcript_mode=1) as pipe:
            def run(submit, steps):
                tickets = [submit() for _ in range(steps)]        # submit blocks while two batches are in flight
                for t in tickets:
                    assert pipe.wait(t) is True, "valid batch rejected (pipeline)"

            def timed_pipe(submit):
                run(submit, args.inflight * max(args.warmup, 3))    # warm-up on every context
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                run(submit, pipe_batches)
                torch.cuda.synchronize()
                e1.record()
                torch.cuda.synchronize()
                return e0.elapsed_time(e1) / pipe_batches

            ms_p = timed_pipe(lambda: pipe.submit_device(d_blobs, d_cs, d_ps, n))
            ms_pe = timed_pipe(lambda: pipe.submit(h_blobs, n, h_cs, n, h_ps, n))
        pipelined = {"in_flight": args.inflight, "value": n / (ms_p / 1e3), "ms_per_step": ms_p, "e2e": n / (ms_pe / 1e3), "e2e_ms_per_step": ms_pe,
                     "unit": "blobs/s", "transcript": "tree", "batches": pipe_batches,
                     "note": "kzgb200_pipeline_submit / _wait: every ticket is one verify_blob_kzg_proof_batch call; "
                             "`batches` batches, in_flight at a time, wall time of all of them (ramp-up and drain included) / batches"}
    sampler.stop_flag = True
    sampler.join(timeout=2)
    lib.kzgb200_set_transcript_mode(ctx, 1)
    step_resident()
    zy_sample = plan.last_zy_host(min(n, args.cpu_sample)) if world == 1 else None
    # negatives on the same workload (verdict only, untimed)
    neg = plan.check_negatives(d_blobs, d_cs, d_ps)

    total = n * world
    ms, phases, ms_e2e = res["tree"]
    value = total / (ms / 1e3)
    e2e = total / (ms_e2e / 1e3)
    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        # dominant kernel = the longest compute phase of the blocking call; parse_g1 is excluded: its phase spans the deferred
        # subgroup checks, which run beside the tail on SMs of their own
        top = max(range(1, 7), key=lambda i: phases[i]) if world == 1 else None
        roof = int_pipe = None
        if top is not None and phases[top] > 0:
            ach = n * ALGO_BYTES_PER_BLOB / (phases[top] / 1e3) / 1e9
            roof = {"bound": "hbm", "kernel": PHASES[top], "achieved": ach, "peak": peak, "unit": "GB/s", "frac": ach / peak,
                    "traffic": TRAFFIC_BYTES.get(PHASES[top]), "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6.65 TB/s",
                    "note": "the path is integer-pipe / latency bound, not HBM bound; see int_pipe and DESIGN.md section 4"}
            # algorithmic integer work of the two blob-streaming kernels against the MEASURED pipe peaks (tools/microbench/intpipe.cu,
            # profiles/intpipe_r01.txt): ALU 69.2 thread-ops/clk/SM, carry-chained IMAD.WIDE.X 31.0 /clk/SM, 148 SMs
            clk = 1.965e9
            sha_ops = n * 2050 * 1218.0          # ALU-pipe instructions per 64-byte block (SASS count: 672 SHF + 352 LOP3 + 178 IADD3 + 16 PRMT)
            fr_ops = n * 4095 * 192.0            # IMAD.WIDE.X per fused dual Fr product
            int_pipe = {"challenge_sha256": {"achieved_ops_per_s": sha_ops / (phases[1] / 1e3), "peak_ops_per_s": 69.2 * 148 * clk,
                                             "frac": sha_ops / (phases[1] / 1e3) / (69.2 * 148 * clk), "pipe": "ALU (SHF/LOP3/IADD3)"},
                        "evaluate_barycentric": {"achieved_ops_per_s": fr_ops / (phases[2] / 1e3), "peak_ops_per_s": 31.0 * 148 * clk,
                                                 "frac": fr_ops / (phases[2] / 1e3) / (31.0 * 148 * clk), "pipe": "FMA-heavy (IMAD.WIDE.U32.X)"}}
        cfg = workload_config(args, n, world)
        cfg["host_affinity"] = numa_cpus
        cfg["transcript"] = "tree (opt-in KZGB200_TRANSCRIPT_TREE: same verdict / z / y as kzg-rs, r hashed as a 3-level tree)"
        ex_ms, ex_ph, ex_e2e = res["exact"]
        out = {"metric": "blobs verified/sec (verify_blob_kzg_proof_batch)", "value": value, "unit": "blobs/s", "n_gpus": world,
               "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
               "vs_baseline": None, "dtype": "u32-limb modular integer (Fr 255-bit / Fp 381-bit) + SHA-256", "data": "synthetic",
               "config": cfg,
               "e2e": {"value": e2e, "unit": "blobs/s", "ms_per_step": ms_e2e, "h2d_bytes_per_step": n * ALGO_BYTES_PER_BLOB * world,
                       "d2h_bytes_per_step": 8 * world},
               "gpu_launches": args.steps * sum(plan.count_launches(n, resident=r, tree=t) for r in (True, False) for t in (True, False)),
               "clocks": sampler.summary(),
               "phases_ms": dict(zip(PHASES, phases)) if world == 1 else None, "roofline": roof, "int_pipe": int_pipe, "negatives": neg,
               "pipelined": pipelined,
               "exact_transcript": {"value": total / (ex_ms / 1e3), "e2e": total / (ex_e2e / 1e3), "unit": "blobs/s",
                                    "ms_per_step": ex_ms, "e2e_ms_per_step": ex_e2e,
                                    "phases_ms": dict(zip(PHASES, ex_ph)) if world == 1 else None,
                                    "note": "library default KZGB200_TRANSCRIPT_EXACT: r, its powers and both MSM sums bit-identical to "
                                            "kzg-rs; the transcript is ONE serial SHA-256 chain over all blobs of all ranks"}}
    if rank == 0 and not args.no_cpu_baseline and world == 1:
        from oracle import oracle as O
        O.build()
        cores = os.cpu_count() or 1
        m = min(n, args.cpu_sample)
        hb = h_blobs[:m * BLOB].numpy().tobytes(); hc = h_cs[:m * 48].numpy().tobytes(); hp = h_ps[:m * 48].numpy().tobytes()
        t = time.perf_counter()
        rc, ok, z, y = O.verify_batch_raw(hb, hc, hp, m, nthreads=cores)
        dt = time.perf_counter() - t
        assert rc == 0 and ok
        m1 = min(m, 64)
        t = time.perf_counter()
        O.verify_batch_raw(hb[:m1 * BLOB], hc[:m1 * 48], hp[:m1 * 48], m1, nthreads=1)
        dt1 = time.perf_counter() - t
        out["cpu_baseline"] = {"value": m / dt, "unit": "blobs/s", "cores": cores, "kind": "port",
                               "sample": "first %d of the %d blobs, %d threads (blob-parallel); single thread on %d blobs: %.1f blobs/s; SHA-NI=%d"
                                         % (m, n, cores, m1, m1 / dt1, O.lib().kzgo_sha256_uses_shani())}
        # bit-exactness of z, y on the sample against the oracle
        out["parity_sample"] = {"blobs": m, "z_y_bit_exact": zy_sample == (z, y)}
    if rank == 0:
        print(json.dumps(out))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
