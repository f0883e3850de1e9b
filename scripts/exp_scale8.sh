#!/bin/bash
# C2 weak scaling at N GPUs: NCCL CTA caps for the per-step count gather, and no gather at all (diagnosis)
N=${1:-8}
run() {
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --no-e2e --no-sweep --no-c1 --no-c4 $2 2>gpurun_out/exp_scale8.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=$N $1 $2', 'value %.4g' % d['value'], 'ms/step %.4f' % d['ms_per_step'], d['roofline']['kernel_ms'], 'c5 edge batch %.4f weak %.4f' % (d['c5']['edge_batch_ms'], d['c5']['edge_weak_ms']))
"
}
RRTQX_NCCL_MAX_CTAS=0 run ctas=default ""
RRTQX_NCCL_MAX_CTAS=2 run ctas=2 ""
RRTQX_NCCL_MAX_CTAS=1 run ctas=1 ""
RRTQX_NCCL_MAX_CTAS=4 run ctas=4 ""
