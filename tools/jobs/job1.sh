#!/bin/bash
# GPU job 1 of round 2: full GPU test-suite, bench (default flags), launch list, sanitizer on the 64-blob test
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
lscpu | head -25 >> gpurun_out/gpu.txt; grep -o -m1 "sha_ni" /proc/cpuinfo >> gpurun_out/gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
timeout 1500 python -m pytest tests -m gpu -x -q --durations=25 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?" >> gpurun_out/bench.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r02.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extras > gpurun_out/ncu_bench.log 2>&1
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest "tests/test_gpu_parity.py::test_synthetic_batch_64_against_oracle" -x -q > gpurun_out/sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
tail -5 gpurun_out/pytest_gpu.log; tail -3 gpurun_out/bench.err; head -c 1500 gpurun_out/bench.json; tail -3 gpurun_out/sanitizer_memcheck.log
