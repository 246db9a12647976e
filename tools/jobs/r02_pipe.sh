#!/bin/bash
# streaming front-end (two batches in flight) against the MSM tail variants
mkdir -p gpurun_out
run() {
  timeout 400 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_pipe.json 2> gpurun_out/bench_pipe.err
  python - "$1" <<'PY'
import json,sys
o=json.loads(open('gpurun_out/bench_pipe.json').read().strip().split('\n')[-1])
p=o['pipelined']
print(sys.argv[1],'| blocking ms',round(o['ms_per_step'],3),'| pipelined ms',round(p['ms_per_step'],3),'M/s',round(p['value']/1e6,3),'e2e ms',round(p['e2e_ms_per_step'],2))
PY
}
run "default (occ3 join2)"
KZGB200_MSM_OCC=2 run "occ2"
KZGB200_MSM_OCC=2 KZGB200_MSM_JOIN=8 run "occ2 join8"
run "default again"
