#!/bin/bash
# launch list of the bench command + ncu --set full of the two-stage collision kernels (after a clean run)
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-c1 --no-c4 --no-c5"
$CMD > gpurun_out/plain_final.log 2>&1 || { tail -5 gpurun_out/plain_final.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/bench_launches.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:"pq_collect_kernel|pq_test_kernel|pq_slow_kernel" -s 8 -c 8 -o gpurun_out/prof_pq -f $CMD > gpurun_out/ncu_pq.log 2>&1
tail -2 gpurun_out/ncu_pq.log | cut -c1-200
