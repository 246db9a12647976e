#!/bin/bash
# short bench x2 + ncu launch list
mkdir -p gpurun_out
for i in 1 2; do
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_short$i.json 2> gpurun_out/bench_short$i.err
python - <<PY
import json
o=json.loads(open('gpurun_out/bench_short$i.json').read().strip().split('\n')[-1])
print(o['value'], o['ms_per_step'], json.dumps(o['phases_ms']), o['e2e']['value'])
PY
done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench.log 2>&1
python tools/ncu_summary.py gpurun_out/launches.csv 2>/dev/null | head -40
