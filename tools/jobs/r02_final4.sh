#!/bin/bash
# refresh of the single-GPU records on the final code of round 2: tests, smoke, bench lines, launch list, captures of the changed kernels, sanitizers
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 python tools/bench_configs.py > gpurun_out/bench_configs.json 2> gpurun_out/bench_configs.err; echo "configs rc=$?" >> gpurun_out/bench_configs.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench.log 2>&1
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras --blobs 16384"
for k in msm_bucket_kernel msm_bucket_join_kernel msm_window_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/full_$k $B > gpurun_out/ncu_$k.log 2>&1
  ncu -i gpurun_out/full_$k.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$k.csv 2>/dev/null
  rm -f gpurun_out/full_$k.ncu-rep
done
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest "tests/test_gpu_parity.py::test_synthetic_batch_64_against_oracle" -x -q > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest "tests/test_gpu_parity.py::test_synthetic_batch_64_against_oracle" -x -q > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log
tail -2 gpurun_out/bench.err gpurun_out/bench_configs.err; tail -3 gpurun_out/sanitizer_memcheck.log gpurun_out/sanitizer_racecheck.log
python - <<'PY'
import json
o=json.loads(open('gpurun_out/bench.json').read().strip().split('\n')[-1])
print(o['value'], o['ms_per_step'], json.dumps({k:round(v,3) for k,v in o['phases_ms'].items()}), o['e2e']['value'], o['e2e']['ms_per_step'], o['e2e'].get('pageable'), o.get('pipelined'), o['roofline'], o['gpu_launches'])
PY
