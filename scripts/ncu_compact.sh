#!/bin/bash
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-c1 --no-c4 --no-c5"
ncu --set full --clock-control none --import-source on -k regex:flag_compact_fused -s 8 -c 1 -f -o gpurun_out/r02_compact $CMD > gpurun_out/r02_ncu_compact.log 2>&1
tail -3 gpurun_out/r02_ncu_compact.log
