#!/bin/bash
mkdir -p gpurun_out
for st in -2 8 -2 8; do
KZGB200_SHA_STAGES=$st timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_ws.json 2> gpurun_out/bench_ws.err
python - <<PY
import json
o=json.loads(open('gpurun_out/bench_ws.json').read().strip().split('\n')[-1])
print('sha_stages=$st', round(o['value']), o['ms_per_step'], o['phases_ms']['challenge_sha256'], o['phases_ms']['g1_decompress'])
PY
done
