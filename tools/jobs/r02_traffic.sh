#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,launch__grid_size --clock-control none -k regex:"challenge_kernel|eval_kernel" -c 60 --csv --log-file gpurun_out/traffic_r02.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_traffic.log 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/traffic_r02.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]; ki=H.index('Kernel Name'); mi=H.index('Metric Name'); vi=H.index('Metric Value'); ui=H.index('Metric Unit'); idi=H.index('ID')
d={}
for r in rows[hdr+1:]:
    if len(r)<=vi: continue
    d.setdefault((r[idi], r[ki][:40]),{})[r[mi]]=(r[vi],r[ui])
for k,v in d.items():
    if v.get('launch__grid_size',('0',''))[0].replace(',','') in ('512','16384'): print(k, v)
PY
