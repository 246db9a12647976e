#!/bin/bash
# join kernel: lanes per bucket
mkdir -p gpurun_out
for rep in 1 2; do
for j in 8 4 2; do
  KZGB200_MSM_JOIN=$j timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_join.json 2> gpurun_out/bench_join.err
  python - $j <<'PY'
import json,sys
o=json.loads(open('gpurun_out/bench_join.json').read().strip().split('\n')[-1])
p=o['phases_ms']
print('join lanes',sys.argv[1],'ms',round(o['ms_per_step'],3),'hash',round(p['challenge_sha256'],2),'lincomb',round(p['lincomb_terms'],3),'reduce',round(p['reduce'],3),'ok',o.get('parity_sample'))
PY
done
done
