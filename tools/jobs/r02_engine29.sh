#!/bin/bash
# first GPU run of the 29-bit-limb pairing engine: GPU tests, section probe, short bench
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 300 python tools/gpu_probe.py 16384 > gpurun_out/probe.log 2>&1
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "bench rc=$?" >> gpurun_out/bench_short.err
tail -15 gpurun_out/pytest_gpu.log; tail -6 gpurun_out/probe.log; tail -3 gpurun_out/bench_short.err
python - <<'PY'
import json
o=json.loads(open('gpurun_out/bench_short.json').read().strip().split('\n')[-1])
print(o['value'], o['ms_per_step'], json.dumps(o['phases_ms']), o['e2e']['value'])
PY
