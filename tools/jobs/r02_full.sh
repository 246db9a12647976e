#!/bin/bash
# full GPU test-suite + default bench + engine self-test
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -14 gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
tail -3 gpurun_out/bench.err
python - <<'PY'
import json
o=json.loads(open('gpurun_out/bench.json').read().strip().split('\n')[-1])
print(o['value'], o['ms_per_step'], json.dumps(o['phases_ms']), o['e2e']['value'], o['e2e']['ms_per_step'], o.get('pipelined'))
PY
timeout 100 python -u tools/engine_selftest.py 1 1 > gpurun_out/selftest.log 2>&1; tail -5 gpurun_out/selftest.log
