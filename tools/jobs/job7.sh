#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:many_pairing_kernel -c 1 -o gpurun_out/many_pairing python bench.py --config tuples --tuples 12432 --steps 1 --warmup 0 > gpurun_out/ncu_many.log 2>&1
ncu -i gpurun_out/many_pairing.ncu-rep --page raw --csv > gpurun_out/many_pairing_raw.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/many_pairing_raw.csv 2>/dev/null | head -70
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/many_pairing_raw.csv')))
hdr,units,vals=rows[0],rows[1],rows[2]
for i,h in enumerate(hdr):
    if any(k in h for k in ('bank_conflict','l1tex__data_pipe','smsp__inst_executed_pipe','sm__inst_executed_pipe','lsu','shared','smsp__average_warps_issue_stalled')) and vals[i] not in ('0','0.000000','n/a'):
        print(h, units[i], vals[i])
PY
