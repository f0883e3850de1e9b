#!/usr/bin/env python
"""A/B of library builds on the C3 legs (RRTQX_B200_LIB selects the build): add sweep (256 obstacles / one), resident
edge check, per-call edge check; library phase events, L2 flushed by a 256 MiB memset on the same stream."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import bench
from rrtqx_3d_b200 import workloads as W
from rrtqx_3d_b200.device import Context, DeviceTree, EdgeSet, SphereSet, SweepResult, edge_check_batch

stream = torch.cuda.Stream(); torch.cuda.set_stream(stream)
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
calls = {
  "add_sweep_256": (lambda: E.add_sweep(S, ids, W.ROBOT_RADIUS, W.DELTA, result=sres), "add_sweep"),
  "add_sweep_1": (lambda: E.add_sweep(S, ids[7:8], W.ROBOT_RADIUS, W.DELTA, result=sres), "add_sweep"),
  "edges_check": (lambda: E.check_all(S, W.ROBOT_RADIUS, out=dflag.data_ptr()), "edges_check"),
  "edge_check_batch": (lambda: edge_check_batch(tree, S, dsrc, ddst, W.ROBOT_RADIUS, n_edges=len(src), out=dflag), "edge_check"),
}
out = []
for cname, (fn, phase) in calls.items():
    ph = []
    for it in range(13):
        flush.zero_()
        fn()
        if it >= 3: ph.append(ctx.last_phase_ms(phase))
    out.append(f"{cname} {np.mean(ph):.4f}")
print(os.environ.get("RRTQX_B200_LIB", "default").split("/")[-1], " | ".join(out), "| hits", sres.sizes()[:2], int(dflag.sum().item()), flush=True)
