#!/bin/bash
# A/B of the range kernels on the GPU box: parity tests first, then bench lines per configuration.
# usage: CFGS="gen occ aspect;..." scripts/exp_v5.sh
mkdir -p gpurun_out
python -m pytest tests/test_gpu_kdtree.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -5
IFS=";" read -ra ARR <<< "${CFGS:-5 4 1;4 4 1}"
for cfg in "${ARR[@]}"; do
  set -- $cfg
  RRTQX_RANGE_KERNEL=$1 RRTQX_GRID_OCCUPANCY=$2 RRTQX_GRID_ASPECT=$3 RRTQX_V5_NW=${4:-24} python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-sweep --no-c1 --no-c4 --no-c5 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('gen=$1 occ=$2 asp=$3 nw=${4:-24}', 'ms', round(d['ms_per_step'],3), 'fill', round(d['roofline']['kernel_ms']['range_fill'],3), 'sort', round(d['roofline']['kernel_ms']['range_sort'],3), 'frac', round(d['roofline']['frac'],3), 'K', d['config']['neighbours_per_step_rank0'])
    elif 'rror' in l: print(l.strip()[:300])
"
done
