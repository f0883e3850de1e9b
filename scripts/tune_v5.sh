#!/bin/bash
# grid / query-sort sweep for the v5 range kernel on C2: "occupancy aspect S F" per config
IFS=";" read -ra ARR <<< "${CFGS:-4 0 3 1;4 2 3 1;4 4 3 1;3 0 3 1;6 0 3 1;6 2 3 1;4 0 2 1;4 0 4 1;4 2 4 1;8 4 3 1}"
for cfg in "${ARR[@]}"; do
  set -- $cfg
  RRTQX_GRID_OCCUPANCY=$1 RRTQX_GRID_ASPECT=$2 RRTQX_QSORT_S=$3 RRTQX_QSORT_F=$4 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-sweep --no-c1 --no-c4 --no-c5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('occ=$1 asp=$2 S=$3 F=$4', 'ms', round(d['ms_per_step'],3), 'fill', round(d['roofline']['kernel_ms']['range_fill'],3), 'sort', round(d['roofline']['kernel_ms']['range_sort'],3), 'frac', round(d['roofline']['frac'],4), 'sparse', round(d['sparse_variant']['ms'],3), 'nearest', round(d['nearest']['ms'],3))
"
done
