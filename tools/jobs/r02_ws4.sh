#!/bin/bash
mkdir -p gpurun_out
KZGB200_SHA_STAGES=-2 timeout 900 python -m pytest tests -m gpu -x -q -k "kat or blob_kzg_proof_vectors or batch_vectors or synthetic_batch_64 or full_size" > gpurun_out/pytest_ws.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ws.log
tail -3 gpurun_out/pytest_ws.log
for st in -2 8; do
KZGB200_SHA_STAGES=$st timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_ws.json 2> gpurun_out/bench_ws.err
python - <<PY
import json
o=json.loads(open('gpurun_out/bench_ws.json').read().strip().split('\n')[-1])
print('sha_stages=$st', round(o['value']), o['ms_per_step'], json.dumps(o['phases_ms']), round(o['e2e']['value']), o['e2e']['ms_per_step'])
PY
done
