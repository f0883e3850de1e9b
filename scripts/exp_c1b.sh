#!/bin/bash
for env in "RRTQX_GRID_ASPECT=2" "RRTQX_GRID_ASPECT=1" "RRTQX_GRID_ASPECT=2" "RRTQX_GRID_ASPECT=1"; do
  env $env python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-sweep --no-c4 --no-c5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('$env', d['planner_iteration']['gpu_us_per_iteration'], d['planner_iteration']['nodes_final'])
"
done
