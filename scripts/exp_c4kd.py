import sys, math, numpy as np, torch
sys.path.insert(0,'/root/repo')
from rrtqx_3d_b200 import workloads as W
from rrtqx_3d_b200.device import Context, DeviceTree, RangeResult
ctx = Context(0, torch.cuda.current_stream().cuda_stream)
lo, hi = [-50.0, -50.0, 0.0, 0.0], [50.0, 50.0, 0.0, 2.0 * math.pi]
nodes = W.uniform_points(4, 200000, lo, hi)
qs = W.uniform_points(5, 200000, lo, hi)
for wrap in (True, False):
    t = DeviceTree(ctx, 4, wraps=[3], wrap_points=[2*math.pi]) if wrap else DeviceTree(ctx, 4)
    t.insert_batch(nodes)
    dq = torch.from_numpy(qs).cuda()
    res = RangeResult(ctx)
    for r in (1.06, 2.0):
        ts=[]
        for i in range(5):
            _, tot = t.range_query(dq, r, result=res, n_queries=len(qs))
            if i>=2: ts.append(ctx.last_phase_ms("range_query"))
        print("wrap" if wrap else "flat", "r", r, "K/q", tot/len(qs), "ms", round(float(np.mean(ts)),3))
