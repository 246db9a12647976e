#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "pageable or full_size or chunked or batch_vectors or pipeline" > gpurun_out/pytest_ws.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ws.log
tail -4 gpurun_out/pytest_ws.log
for st in 8 4; do
KZGB200_SHA_STAGES=$st timeout 600 python bench.py --steps 5 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/bench_ws$st.json 2> gpurun_out/bench_ws$st.err
python - <<PY
import json
o=json.loads(open('gpurun_out/bench_ws$st.json').read().strip().split('\n')[-1])
print('sha_stages=$st', round(o['value']), o['ms_per_step'], round(o['e2e']['value']), o['e2e']['ms_per_step'])
PY
done
