#!/bin/bash
# A/B of the warp-queue collision kernels vs the thread-per-edge grid kernels (RRTQX_EDGE_NO_QUEUE=1)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_collision.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -5
CMD="python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-c1 --no-c5 --no-c4"
for mode in queue noqueue; do
  if [ $mode = noqueue ]; then export RRTQX_EDGE_NO_QUEUE=1; else unset RRTQX_EDGE_NO_QUEUE; fi
  timeout 600 $CMD > gpurun_out/collide_$mode.json 2> gpurun_out/collide_$mode.err || tail -5 gpurun_out/collide_$mode.err
  python - <<PY
import json
l = json.loads(open("gpurun_out/collide_$mode.json").read().strip().splitlines()[-1])
s, b = l.get("edge_sweep", {}), l.get("edge_batch", {})
print("$mode", "sweep ms", s.get("ms"), "blocked", s.get("blocked_edges"), s.get("orphans"), "err", s.get("error"),
      "| batch ms", b.get("ms"), "colliding", b.get("colliding_edges"))
PY
done
