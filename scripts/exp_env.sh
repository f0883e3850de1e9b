#!/bin/bash
# usage: ENVS="A=1 B=2;C=3" scripts/exp_env.sh  -> bench line summary per environment setting
IFS=";" read -ra ARR <<< "${ENVS:-X=0}"
for cfg in "${ARR[@]}"; do
  env $cfg python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-sweep --no-c1 --no-c4 --no-c5 2>&1 | python -c "
import sys,json
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('$cfg', 'ms', round(d['ms_per_step'],3), 'fill', round(d['roofline']['kernel_ms']['range_fill'],3), 'sort', round(d['roofline']['kernel_ms']['range_sort'],3), 'frac', round(d['roofline']['frac'],3))
    elif 'rror' in l: print(l.strip()[:300])
"
done
