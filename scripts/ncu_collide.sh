#!/bin/bash
# launch list (and optionally --set full) of the two-stage collision kernels, after a clean run of the same command
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-c1 --no-c5 --no-c4"
$CMD > gpurun_out/plain_cq.log 2>&1 || { tail -5 gpurun_out/plain_cq.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"pq_|cover_|sphere_grid|sweep_table|sphere_table|compact_flags|add_sweep|edge_check" -c 200 --csv --log-file gpurun_out/cq_launches.csv $CMD > /dev/null 2>&1
if [ "$1" = full ]; then
ncu --set full --clock-control none --import-source on -k regex:"pq_collect_kernel|pq_test_kernel" -s 4 -c 4 -o gpurun_out/prof_cq -f $CMD > gpurun_out/ncu_cq.log 2>&1
tail -2 gpurun_out/ncu_cq.log | cut -c1-200
fi
