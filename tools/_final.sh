set -x
timeout 500 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; tail -c 300 gpurun_out/bench_final.err
for k in challenge_kernel:17 eval_kernel:17 batch_final_kernel:1 msm_bucket_kernel:1; do
  name=${k%%:*}; skip=${k##*:}
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$name -s $skip -c 1 -o /tmp/cap_$name -f python bench.py --steps 1 --no-cpu-baseline --no-pipeline > /dev/null 2>&1
  ncu -i /tmp/cap_$name.ncu-rep --page raw --csv > gpurun_out/ncu_raw_$name.csv 2>/dev/null
done
ls -la gpurun_out | tail -8
