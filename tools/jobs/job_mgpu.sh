#!/bin/bash
# multi-GPU job: N = number of visible GPUs
N=$(nvidia-smi -L | wc -l)
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_$N.txt 2>&1
timeout 600 python -m pytest tests/test_gpu_round2.py -m gpu -q -x -k "group" > gpurun_out/pytest_group_$N.log 2>&1; tail -3 gpurun_out/pytest_group_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err; echo "bench rc=$?"
tail -5 gpurun_out/bench_n$N.err
python - <<PY
import json
o=json.loads(open('gpurun_out/bench_n$N.json').read().strip().split('\n')[-1])
print('weak', o['value'], o['ms_per_step'], 'e2e', o['e2e']['value'], o['e2e']['ms_per_step'])
print('phases', json.dumps(o['phases_ms']))
print('strong', json.dumps(o['strong']))
print('group', o['group'], 'neg', o['negatives'], 'parity', o.get('parity_sample'))
print('modes', json.dumps(o.get('other_transcript_modes'))[:600])
PY
