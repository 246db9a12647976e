#!/bin/bash
# bucket kernel: occupancy variant x slice length
mkdir -p gpurun_out
for occ in 2 3 4; do
  for sl in 0 12 14 16 21; do
    KZGB200_MSM_OCC=$occ KZGB200_MSM_SLICE=$sl timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/bench_occ${occ}_sl${sl}.json 2> gpurun_out/bench_occ.err
    python - $occ $sl <<'PY'
import json,sys
o=json.loads(open('gpurun_out/bench_occ%s_sl%s.json'%(sys.argv[1],sys.argv[2])).read().strip().split('\n')[-1])
p=o['phases_ms']
print('occ',sys.argv[1],'slice',sys.argv[2],'ms',round(o['ms_per_step'],3),'lincomb',round(p['lincomb_terms'],3),'reduce',round(p['reduce'],3),'final',round(p['final_pairing'],3),'e2e',round(o['e2e']['ms_per_step'],2))
PY
  done
done
