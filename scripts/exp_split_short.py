import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from rrtqx_3d_b200 import workloads as W
from rrtqx_3d_b200.device import Context, DeviceTree, RangeResult
stream = torch.cuda.current_stream()
ctx = Context(0, stream.cuda_stream)
pts, qs, r = W.c2_workload(1000000, 1000000)
t = DeviceTree(ctx, 3); t.insert_batch(pts)
dq = torch.from_numpy(qs).cuda()
res = RangeResult(ctx)
flush = torch.empty(256<<20, dtype=torch.uint8, device='cuda')
def run(name, **kw):
    ts=[]
    for i in range(6):
        flush.zero_()
        t.range_query(dq, r, result=res, n_queries=len(qs), **kw)
        if i>=2: ts.append(ctx.last_phase_ms("range_fill"))
    print(name, "fill ms", round(float(np.mean(ts)),3), end='; ')
run("idx+dist", want_dist=True); run("idx only", want_dist=False); run("count only", want_dist=False, count_only=True); print()
