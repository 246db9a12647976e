#!/bin/bash
mkdir -p gpurun_out
for st in 8 4 8; do
KZGB200_SHA_STAGES=$st timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_ws.json 2> gpurun_out/bench_ws.err
python - <<PY
import json
o=json.loads(open('gpurun_out/bench_ws.json').read().strip().split('\n')[-1])
print('sha_stages=$st', round(o['value']), o['ms_per_step'], json.dumps(o['phases_ms']), round(o['e2e']['value']), o['e2e']['ms_per_step'], o['clocks'])
PY
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ws.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_ws.log 2>&1
python tools/ncu_summary.py gpurun_out/launches_ws.csv 2>/dev/null | head -6
