#!/bin/bash
# one point of the scaling curve, launched as the driver launches it
N=$1; TAG=${2:-r2_scale3}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/${TAG}_n$N.json 2> gpurun_out/${TAG}_n$N.err
tail -2 gpurun_out/${TAG}_n$N.err
python - <<P
import json
d = json.loads(open("gpurun_out/${TAG}_n$N.json").read().strip().splitlines()[-1])
print($N, d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], {k: d["c5"][k] for k in ("edge_batch_ms", "edge_weak_ms", "planner_iterations_per_s")})
P
