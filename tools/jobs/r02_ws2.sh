#!/bin/bash
mkdir -p gpurun_out
KZGB200_SHA_STAGES=-1 timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_ws.csv python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras > gpurun_out/ncu_ws.log 2>&1
python tools/ncu_summary.py gpurun_out/launches_ws.csv 2>/dev/null | head -8
