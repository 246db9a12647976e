#!/bin/bash
# last refresh of the single-GPU bench line, launch list and the changed kernels' captures (final code of round 2)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench.log 2>&1
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras --blobs 16384"
for k in msm_bucket_kernel pairing_check_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/full_$k $B > gpurun_out/ncu_$k.log 2>&1
  ncu -i gpurun_out/full_$k.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$k.csv 2>/dev/null
  rm -f gpurun_out/full_$k.ncu-rep
done
timeout 300 python tools/gpu_probe.py 16384 > gpurun_out/probe.log 2>&1; tail -4 gpurun_out/probe.log
python - <<'PY'
import json
o=json.loads(open('gpurun_out/bench.json').read().strip().split('\n')[-1])
print(o['value'], o['ms_per_step'], json.dumps(o['phases_ms']), o['e2e']['value'], o['e2e']['ms_per_step'], o['e2e'].get('pageable'), o.get('pipelined'), o['roofline'], o['gpu_launches'])
PY
