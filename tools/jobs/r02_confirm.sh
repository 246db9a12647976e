#!/bin/bash
# last confirmation of the shipped build: GPU tests, smoke, one short bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -n 3 gpurun_out/pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -n 1 gpurun_out/smoke.log
timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_confirm.json 2> gpurun_out/bench_confirm.err; echo "bench rc=$?"
python - <<'PY'
import json
o=json.loads(open('gpurun_out/bench_confirm.json').read().strip().split('\n')[-1])
print(o['value'], o['ms_per_step'], o['e2e']['value'], o['e2e']['ms_per_step'], o['roofline']['traffic'], o['gpu_launches'])
PY
