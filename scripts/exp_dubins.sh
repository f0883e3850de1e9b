#!/bin/bash
# A/B of builds of the Dubins trajectory check and the Otte / Dubins sweep: for each library name run the C4 leg only
for lib in "$@"; do
  RRTQX_B200_LIB=$PWD/rrtqx_3d_b200/$lib python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-sweep --no-c1 --no-c5 2>gpurun_out/exp_dubins.err | python -c '
import json, sys
d = json.loads(sys.stdin.read().strip().splitlines()[-1])["dubins"]
print(sys.argv[1], "check_ms %.4f" % d["check_ms"], "solve_ms %.4f" % d["solve_ms"], d["colliding_edges"],
      {k: v for k, v in d.get("sweep_2d", {}).items() if k != "note"}, d.get("error"))
' $lib
done
