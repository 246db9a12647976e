#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_short.json 2> gpurun_out/bench_short.err; echo "bench rc=$?"
python - <<'PY'
import json
o=json.loads(open('gpurun_out/bench_short.json').read().strip().split('\n')[-1])
print(o['value'], o['ms_per_step'], json.dumps(o['phases_ms']), o['e2e']['value'])
PY
for m in 262144; do
timeout 900 python bench.py --config tuples --tuples $m --steps 3 --warmup 2 > gpurun_out/bench_tuples_$m.json 2> gpurun_out/bench_tuples.err; echo "rc=$?"
python -c "
import json
o=json.loads(open('gpurun_out/bench_tuples_$m.json').read().strip().split('\n')[-1]); print($m, round(o['value']), o['verdicts_match_construction'], o['phases_ms_last_chunk'], o['roofline']['frac'])"
done
