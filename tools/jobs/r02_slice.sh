#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "intermediates or full_size or chunked or batch_vectors or synthetic" > gpurun_out/pytest_quick.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_quick.log
tail -3 gpurun_out/pytest_quick.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err
python - <<'PY'
import json
o=json.loads(open('gpurun_out/bench_short.json').read().strip().split('\n')[-1])
print(round(o['value']), o['ms_per_step'], json.dumps(o['phases_ms']), round(o['e2e']['value']), o['e2e']['ms_per_step'])
PY
