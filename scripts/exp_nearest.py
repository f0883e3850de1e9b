import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from rrtqx_3d_b200 import workloads as W
from rrtqx_3d_b200.device import Context, DeviceTree
stream = torch.cuda.current_stream()
ctx = Context(0, stream.cuda_stream)
pts, qs, r = W.c2_workload(1000000, 1000000)
t = DeviceTree(ctx, 3); t.insert_batch(pts)
dq = torch.from_numpy(qs).cuda()
idx = torch.empty(len(qs), dtype=torch.int32, device='cuda'); dist = torch.empty(len(qs), dtype=torch.float64, device='cuda')
flush = torch.empty(256<<20, dtype=torch.uint8, device='cuda')
ts=[]
for i in range(8):
    flush.zero_()
    t.nearest(dq.data_ptr(), n_queries=len(qs), idx_out=idx.data_ptr(), dist_out=dist.data_ptr())
    if i>=3: ts.append(ctx.last_phase_ms("nearest"))
print("nearest 1M queries on 1M nodes: ms", round(float(np.mean(ts)),3), "-> queries/s", len(qs)/np.mean(ts)*1e3, "hbm frac", 60e6/(np.mean(ts)*1e-3)/6545e9)
print("tree build ms", ctx.last_phase_ms("tree_build"))
