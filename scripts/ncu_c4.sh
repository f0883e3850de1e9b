#!/bin/bash
# ncu --set full of the final Dubins trajectory check and of the flag compaction (after a clean run of the same commands)
CMD4="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-sweep --no-c1 --no-c5"
CMD3="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-c1 --no-c4 --no-c5"
$CMD4 > gpurun_out/plain_c4.json 2> gpurun_out/plain_c4.err || { tail -5 gpurun_out/plain_c4.err; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:dubins_check_kernel -s 3 -c 1 -f -o gpurun_out/r02_dubins_check_final $CMD4 > gpurun_out/r02_ncu_dubins.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:flag_compact_kernel -s 8 -c 1 -f -o gpurun_out/r02_flag_compact $CMD3 > gpurun_out/r02_ncu_fc.log 2>&1
tail -2 gpurun_out/r02_ncu_dubins.log gpurun_out/r02_ncu_fc.log | cut -c1-160
