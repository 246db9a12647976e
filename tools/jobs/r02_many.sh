#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x -k "verify_kzg_proof or per_blob or cpp_mirror or many" > gpurun_out/pytest_many.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_many.log
tail -3 gpurun_out/pytest_many.log
for m in 262144 1000000; do
timeout 900 python bench.py --config tuples --tuples $m --steps 3 --warmup 2 > gpurun_out/bench_tuples_$m.json 2> gpurun_out/bench_tuples.err; echo "rc=$?"
python -c "
import json
o=json.loads(open('gpurun_out/bench_tuples_$m.json').read().strip().split('\n')[-1]); print($m, round(o['value']), o['verdicts_match_construction'], o['roofline']['frac'])"
done
