#!/bin/bash
mkdir -p gpurun_out
python scripts/exp_c1.py 2>&1 | tail -1
ncu --set full --clock-control none --import-source on -k regex:extend_query_kernel -s 4000 -c 1 -o gpurun_out/prof_extend -f python scripts/exp_c1.py > gpurun_out/ncu_ext.log 2>&1
tail -n 2 gpurun_out/ncu_ext.log | cut -c1-200
