#!/bin/bash
# C2 weak scaling + C5 legs at N GPUs: gathers through the IPC windows (default) against NCCL (RRTQX_COMM_NO_IPC=1)
N=${1:-8}
run() {
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --no-cpu --no-e2e --no-sweep --no-c1 --no-c4 $2 2>gpurun_out/exp_scale8.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('N=$N $1 $2', 'value %.4g' % d['value'], 'ms/step %.4f' % d['ms_per_step'], d['roofline']['kernel_ms'], 'c5 edge batch %.4f weak %.4f' % (d['c5']['edge_batch_ms'], d['c5']['edge_weak_ms']), 'peer_stores', d['c5']['comm']['peer_stores'], 'colliding', d['c5']['colliding_edges'])
" || tail -5 gpurun_out/exp_scale8.err
}
MODE=${2:-both}
if [ "$MODE" != nccl ]; then run ipc ""; fi
if [ "$MODE" != ipc ]; then RRTQX_COMM_NO_IPC=1 run nccl ""; fi
