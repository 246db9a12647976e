#!/bin/bash
mkdir -p gpurun_out
timeout 200 python -u tools/engine_selftest.py 3 2 > gpurun_out/selftest.log 2>&1; tail -6 gpurun_out/selftest.log
timeout 300 python tools/gpu_probe.py 16384 > gpurun_out/probe.log 2>&1
tail -3 gpurun_out/probe.log
