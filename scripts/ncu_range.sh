#!/bin/bash
# one ncu --set full capture of the range kernel (after the same command ran clean without ncu)
# usage: scripts/ncu_range.sh <tag> [kernel regex]
TAG=${1:-v5}; KRE=${2:-range_v5_kernel}
mkdir -p gpurun_out
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-sweep --no-c1 --no-c4 --no-c5 > gpurun_out/plain_$TAG.log 2>&1 || { tail -5 gpurun_out/plain_$TAG.log; exit 1; }
tail -c 600 gpurun_out/plain_$TAG.log
ncu --set full --clock-control none --import-source on -k regex:$KRE -s 3 -c 1 -o gpurun_out/prof_$TAG -f python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e --no-sweep --no-c1 --no-c4 --no-c5 > gpurun_out/ncu_$TAG.log 2>&1
tail -3 gpurun_out/ncu_$TAG.log
