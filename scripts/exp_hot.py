"""Upper bound of what staging candidate regions on-chip could give: all 1M queries drawn from ONE supercell,
so every record the kernel touches is L1-resident (same neighbours per query, same output volume)."""
import sys, numpy as np, torch
sys.path.insert(0,'/root/repo')
from rrtqx_3d_b200 import workloads as W
from rrtqx_3d_b200.device import Context, DeviceTree, RangeResult
stream = torch.cuda.current_stream()
ctx = Context(0, stream.cuda_stream)
pts, qs, r = W.c2_workload(1000000, 1000000)
t = DeviceTree(ctx, 3); t.insert_batch(pts)
res = RangeResult(ctx)
flush = torch.empty(256<<20, dtype=torch.uint8, device='cuda')
def run(name, q, **kw):
    dq = torch.from_numpy(np.ascontiguousarray(q)).cuda()
    ts=[]
    for i in range(6):
        flush.zero_()
        _, tot = t.range_query(dq, r, result=res, n_queries=len(q), **kw)
        if i>=2: ts.append(ctx.last_phase_ms("range_fill"))
    print(name, "K/q", tot/len(q), "fill ms", round(float(np.mean(ts)),3))
run("uniform", qs, want_dist=True)
box = qs[(np.abs(qs[:,0])<0.95)&(np.abs(qs[:,1])<0.95)&(np.abs(qs[:,2])<0.95)]   # one 1.9-unit cube ~ 3 cells
hot = np.tile(box, (len(qs)//len(box)+1, 1))[:len(qs)]
run("hot cube (%d distinct)" % len(box), hot, want_dist=True)
run("hot cube count only", hot, want_dist=False, count_only=True)
inner = qs[(np.abs(qs[:,0])<17)&(np.abs(qs[:,1])<17)&(np.abs(qs[:,2])<17)]
inner = np.tile(inner, (2,1))[:len(qs)]
run("interior only", inner, want_dist=True)
