#!/bin/bash
# MSM tail experiment: parity tests that pin r / the MSM sums bit-exactly, then a short bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "intermediates or full_size or chunked or synthetic_batch or batch_vectors or per_blob or group" > gpurun_out/pytest_tail.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tail.log
tail -3 gpurun_out/pytest_tail.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_tail.json 2> gpurun_out/bench_tail.err; echo "bench rc=$?"
python - <<'PY'
import json
o=json.loads(open('gpurun_out/bench_tail.json').read().strip().split('\n')[-1])
print(o['value'], o['ms_per_step'], json.dumps({k:round(v,3) for k,v in o['phases_ms'].items()}), o['e2e']['value'], o['e2e']['ms_per_step'])
PY
