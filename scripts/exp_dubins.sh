#!/bin/bash
# parity + A/B of the Dubins trajectory check kernels (RRTQX_DUBINS_CHECK_V1=1: first form)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dubins.py tests/test_gpu_collision.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -3
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-c1 --no-c5 --no-sweep"
for mode in v2 v1; do
  if [ $mode = v1 ]; then export RRTQX_DUBINS_CHECK_V1=1; else unset RRTQX_DUBINS_CHECK_V1; fi
  timeout 600 $CMD > gpurun_out/dub_$mode.json 2> gpurun_out/dub_$mode.err || tail -5 gpurun_out/dub_$mode.err
  python - <<PY
import json
l = json.loads(open("gpurun_out/dub_$mode.json").read().strip().splitlines()[-1])
print("$mode", json.dumps(l.get("c4_dubins"))[:900])
PY
done
