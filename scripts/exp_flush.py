#!/usr/bin/env python
"""How the L2 flush regime changes the sub-millisecond C3 legs (torch events AND library phase events)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from rrtqx_3d_b200 import workloads as W
from rrtqx_3d_b200.device import Context, DeviceTree, EdgeSet, SphereSet, SweepResult, edge_check_batch

stream = torch.cuda.Stream()
torch.cuda.set_stream(stream)   # a non-default stream: its handle is non-zero, so the library really uses it
ctx = Context(0, stream.cuda_stream)
pts, _, r = W.c2_workload(1_000_000, 1)
tree = DeviceTree(ctx, 3); tree.insert_batch(pts)
src, dst, parent = bench.build_c3_edges(tree, pts, 0.5346)
E = EdgeSet(tree); E.upload(src, dst, parent)
centers, radii = W.c3_obstacles(256)
S = SphereSet(ctx, centers, radii)
ids = np.arange(256, dtype=np.int32)
sres = SweepResult(ctx)
dsrc, ddst = torch.from_numpy(src).cuda(), torch.from_numpy(dst).cuda()
dflag = torch.empty(len(src), dtype=torch.uint8, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
f64 = flush.view(torch.int64)
big = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")

def regime(name):
    if name == "none": return
    if name == "write": flush.zero_()
    if name == "write+read": flush.zero_(); f64.sum()
    if name == "copy": big[:256 << 20].copy_(big[256 << 20:])
    if name == "write+sync": flush.zero_(); torch.cuda.synchronize()

calls = {
  "add_sweep": (lambda: E.add_sweep(S, ids, W.ROBOT_RADIUS, W.DELTA, result=sres), "add_sweep"),
  "edges_check": (lambda: E.check_all(S, W.ROBOT_RADIUS, out=dflag.data_ptr()), "edges_check"),
  "edge_check_batch": (lambda: edge_check_batch(tree, S, dsrc, ddst, W.ROBOT_RADIUS, n_edges=len(src), out=dflag), "edge_check"),
}
for cname, (fn, phase) in calls.items():
    for reg in ("none", "write", "write+read", "copy", "write+sync"):
        ev, ph = [], []
        for it in range(8):
            regime(reg)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(stream); fn(); b.record(stream); b.synchronize()
            if it >= 3: ev.append(a.elapsed_time(b)); ph.append(ctx.last_phase_ms(phase))
        print(f"{cname:18s} {reg:12s} torch-events {np.mean(ev):.4f} ms   phase {np.mean(ph):.4f} ms", flush=True)
