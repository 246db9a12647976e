#!/bin/bash
# engine prefetch experiment: selftest of the 16-lane executors, parity subset, short bench, probe
mkdir -p gpurun_out
timeout 300 python tools/engine_selftest.py > gpurun_out/selftest.log 2>&1; tail -1 gpurun_out/selftest.log
timeout 900 python -m pytest tests -m gpu -q -x -k "intermediates or full_size or chunked or synthetic_batch or batch_vectors or per_blob or group or verify_kzg_proof_vectors" > gpurun_out/pytest_tail.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_tail.log
tail -3 gpurun_out/pytest_tail.log
timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_tail.json 2> gpurun_out/bench_tail.err; echo "bench rc=$?"
python - <<'PY'
import json
o=json.loads(open('gpurun_out/bench_tail.json').read().strip().split('\n')[-1])
print(o['value'], o['ms_per_step'], json.dumps({k:round(v,3) for k,v in o['phases_ms'].items()}), o['e2e']['value'], o['e2e']['ms_per_step'])
PY
timeout 300 python tools/gpu_probe.py 16384 > gpurun_out/probe.log 2>&1; tail -4 gpurun_out/probe.log
