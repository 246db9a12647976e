#!/bin/bash
mkdir -p gpurun_out
for c in 1 2 3; do
KZGB200_MANY_CTAS=$c timeout 600 python bench.py --config tuples --tuples 65536 --steps 2 --warmup 1 2>/dev/null | python -c "
import json,sys
o=json.loads(sys.stdin.read().strip().split('\n')[-1]); print('ctas/SM=$c', round(o['value']), o['phases_ms_last_chunk'])"
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:many_pairing -c 1 -o gpurun_out/many_pairing python bench.py --config tuples --tuples 16384 --steps 1 --warmup 0 > gpurun_out/ncu_many.log 2>&1
ncu -i gpurun_out/many_pairing.ncu-rep --page raw --csv > gpurun_out/many_pairing_raw.csv 2>/dev/null
python tools/ncu_raw_summary.py gpurun_out/many_pairing_raw.csv 2>/dev/null | head -70
