#!/bin/bash
# round-2 profile job: ncu --set full of the kernels VERDICT names, racecheck on the 64-blob test, engine section probe
mkdir -p gpurun_out
B="python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-extras --blobs 16384"
for k in g1_decompress_kernel g1_subgroup_kernel batch_final_kernel msm_bucket_kernel msm_combine_kernel msm_window_kernel; do
  timeout 400 ncu --set full --clock-control none --import-source on -k regex:$k -c 1 -f -o gpurun_out/full_$k $B > gpurun_out/ncu_$k.log 2>&1
  ncu -i gpurun_out/full_$k.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$k.csv 2>/dev/null
  rm -f gpurun_out/full_$k.ncu-rep
done
python tools/ncu_raw_summary.py gpurun_out/ncu_raw_*.csv > gpurun_out/ncu_full_r02_summary.txt 2>&1
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest "tests/test_gpu_parity.py::test_synthetic_batch_64_against_oracle" -x -q > gpurun_out/sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log
timeout 300 python tools/gpu_probe.py 16384 > gpurun_out/probe.log 2>&1
tail -4 gpurun_out/sanitizer_racecheck.log; tail -6 gpurun_out/probe.log; head -60 gpurun_out/ncu_full_r02_summary.txt
