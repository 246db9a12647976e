#!/bin/bash
# hash-phase bimodality: runs with and without the shared-memory cap of the chain kernel
mkdir -p gpurun_out
for i in 1 2 3 4 5 6 7 8; do
  for pad in 32768 0; do
    KZGB200_SHA_PAD=$pad timeout 300 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_pad.json 2> gpurun_out/bench_pad.err
    python - $pad <<'PY'
import json,sys
o=json.loads(open('gpurun_out/bench_pad.json').read().strip().split('\n')[-1])
p=o['phases_ms']
print('pad',sys.argv[1],'ms',round(o['ms_per_step'],3),'hash',round(p['challenge_sha256'],3),'decomp',round(p['g1_decompress'],3),'lincomb',round(p['lincomb_terms'],3),'e2e',round(o['e2e']['ms_per_step'],2))
PY
  done
done
