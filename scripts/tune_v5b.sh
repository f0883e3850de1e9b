#!/bin/bash
# warps per block / work-unit size of the v5 kernel on C2: "NW UNIT" per config
IFS=";" read -ra ARR <<< "${CFGS:-24 0;20 0;28 0;24 48;24 64;24 128;24 192}"
for cfg in "${ARR[@]}"; do
  set -- $cfg
  RRTQX_V5_NW=$1 RRTQX_V5_UNIT=$2 python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-sweep --no-c1 --no-c4 --no-c5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('NW=$1 UNIT=$2', 'ms', round(d['ms_per_step'],3), 'fill', round(d['roofline']['kernel_ms']['range_fill'],3), 'frac', round(d['roofline']['frac'],4), 'sparse', round(d['sparse_variant']['ms'],3))
"
done
