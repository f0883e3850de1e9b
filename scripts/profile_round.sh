#!/bin/bash
# GPU side of the round's evidence (run under gpurun, after the same commands exited 0 without a profiler):
#   gpurun_out/<tag>_bench_launches.csv   launch list of the default bench command (short)
#   gpurun_out/<tag>_range_v5.ncu-rep     ncu --set full of the dominant kernel
#   gpurun_out/<tag>_ig_test.ncu-rep      ncu --set full of the item-grid test kernel (C3 sweep)
TAG=${1:-r02}
CMD="python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-c1 --no-c4 --no-c5"
$CMD > gpurun_out/${TAG}_plain.json 2> gpurun_out/${TAG}_plain.err || { tail -5 gpurun_out/${TAG}_plain.err; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/${TAG}_bench_launches.csv $CMD > /dev/null 2>&1
ncu --set full --clock-control none --import-source on -k regex:range_v5_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_range_v5 $CMD > gpurun_out/${TAG}_ncu_range.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:'ig_test_kernel.*Sweep' -s 3 -c 1 -f -o gpurun_out/${TAG}_ig_test $CMD > gpurun_out/${TAG}_ncu_ig.log 2>&1
ls -la gpurun_out/${TAG}_*.ncu-rep gpurun_out/${TAG}_bench_launches.csv
