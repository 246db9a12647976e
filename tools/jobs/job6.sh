#!/bin/bash
mkdir -p gpurun_out
for m in 262144; do
timeout 900 python bench.py --config tuples --tuples $m --steps 3 --warmup 2 > gpurun_out/bench_tuples_$m.json 2> gpurun_out/bench_tuples.err; echo "rc=$?"
python -c "
import json
o=json.loads(open('gpurun_out/bench_tuples_$m.json').read().strip().split('\n')[-1]); print($m, round(o['value']), o['verdicts_match_construction'], o['phases_ms_last_chunk'], o['roofline']['frac'])"
done
tail -3 gpurun_out/bench_tuples.err
