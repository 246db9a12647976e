#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_probe.py 16384 > gpurun_out/probe.log 2>&1
tail -4 gpurun_out/probe.log
