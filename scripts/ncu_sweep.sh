#!/bin/bash
# ncu --set full of the collision kernels (edge-centric add sweep + edge batch), after a clean run of the same command
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-c1 --no-c4 --no-c5"
$CMD > gpurun_out/plain_sw.log 2>&1 || { tail -5 gpurun_out/plain_sw.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:"add_sweep_edge_kernel|edge_check_grid_kernel" -s 6 -c 2 -o gpurun_out/prof_sweep_v2 -f $CMD > gpurun_out/ncu_sw.log 2>&1
tail -3 gpurun_out/ncu_sw.log | cut -c1-200
