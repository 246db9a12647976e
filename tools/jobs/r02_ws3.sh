#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
for nb in 8192 2048; do for st in 8 4; do
KZGB200_SHA_STAGES=$st timeout 600 python bench.py --blobs $nb --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_ws.json 2> gpurun_out/bench_ws.err
python - <<PY
import json
o=json.loads(open('gpurun_out/bench_ws.json').read().strip().split('\n')[-1])
print('blobs=$nb sha_stages=$st', round(o['value']), o['ms_per_step'], json.dumps(o['phases_ms']), round(o['e2e']['value']), o['e2e']['ms_per_step'])
PY
done; done
timeout 600 python tools/bench_configs.py 2>/dev/null | head -3
