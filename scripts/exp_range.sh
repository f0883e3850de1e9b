#!/bin/bash
# A/B of range-kernel builds: for each "lib:nw" pair run the C2 leg only and print ms / frac
for spec in "$@"; do
  lib=${spec%%:*}; nw=${spec##*:}
  RRTQX_B200_LIB=$PWD/rrtqx_3d_b200/$lib RRTQX_V5_NW=$nw python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --no-sweep --no-c1 --no-c4 --no-c5 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$spec', 'ms/step %.4f' % d['ms_per_step'], 'kernel %.4f' % d['roofline']['kernel_ms']['range_fill'], 'frac %.4f' % d['roofline']['frac'], 'sparse', d.get('sparse_variant',{}).get('ms'))
"
done
