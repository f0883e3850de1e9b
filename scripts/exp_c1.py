"""Where a C1 planner iteration spends its time: kernel (events) vs host call overhead."""
import sys, time, numpy as np, torch
sys.path.insert(0,'/root/repo')
from rrtqx_3d_b200 import workloads as W, _abi as A
from rrtqx_3d_b200.device import Context, DeviceTree, SphereSet, extend_query
ctx = Context(0)
centers, radii, _ = W.building2_spheres()
S = SphereSet(ctx, centers, radii)
n_iter = 6000
samples = W.uniform_points(1, n_iter, [-W.ENV_RAD] * 3, [W.ENV_RAD] * 3)
t = DeviceTree(ctx, 3); t.insert(np.array([4.0, 16.5, -7.5]))
bufs = (np.empty(8192, np.int32), np.empty(8192, np.float64), np.empty(8192, np.uint8), np.empty(8192, np.uint8))
n = 1; tq = ti = 0.0; kms = []
for it in range(n_iter):
    r = W.shrinking_ball_radius(n, 3, W.DELTA, W.BALL_CONSTANT)
    t0 = time.perf_counter()
    res = extend_query(t, S, samples[it], r, W.ROBOT_RADIUS, A.CHECK_QUICK_PASS, capacity=8192, bufs=bufs)
    t1 = time.perf_counter()
    t2 = time.perf_counter()
    if not res.point_collides:
        t.insert(samples[it]); n += 1
    t3 = time.perf_counter()
    tq += t1 - t0; ti += t3 - t2
print(f"extend_query call {1e6*tq/n_iter:.1f} us/iter insert call {1e6*ti/n_iter:.1f} us/iter, nodes {n}")
