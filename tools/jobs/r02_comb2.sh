#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -k regex:msm_combine -c 2 python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-extras 2>&1 | grep -E "msm_combine|duration|inst_executed" | head -8
