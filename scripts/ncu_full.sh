#!/bin/bash
# one `ncu --set full` capture of a kernel selected by a regex on the demangled name:
#   bash scripts/ncu_full.sh '<regex>' <out-name> [launch-skip] -- bench flags...
re="$1"; out="$2"; skip="${3:-0}"; shift 3 || true
[ "$1" = "--" ] && shift
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$re" --launch-skip "$skip" -c 1 \
    -f -o gpurun_out/"$out" python bench.py "$@" > gpurun_out/"$out".log 2>&1
ls -la gpurun_out/"$out".ncu-rep
