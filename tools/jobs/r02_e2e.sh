#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -u tools/engine_selftest.py 2 2 > gpurun_out/selftest.log 2>&1; tail -2 gpurun_out/selftest.log
timeout 900 python -m pytest tests -m gpu -x -q -k "batch or pageable or pipeline or chunked or full_size" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_quick.log
tail -4 gpurun_out/pytest_quick.log
for st in 1 0; do
KZGB200_SLAB_TAIL=$st timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_slab$st.json 2> gpurun_out/bench_slab$st.err
python - <<PY
import json
o=json.loads(open('gpurun_out/bench_slab$st.json').read().strip().split('\n')[-1])
print('slab_tail=$st', round(o['value']), o['ms_per_step'], json.dumps(o['phases_ms']), round(o['e2e']['value']), o['e2e']['ms_per_step'])
PY
done
timeout 300 python tools/gpu_probe.py 16384 > gpurun_out/probe.log 2>&1; tail -3 gpurun_out/probe.log
