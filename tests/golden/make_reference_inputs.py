"""Writes tests/golden/reference_inputs.txt: the seeded inputs that julia/make_reference_vectors.jl replays through the
REAL reference (no RNG on the Julia side) and that tests/test_reference_vectors.py replays through the oracle.

    python tests/golden/make_reference_inputs.py

Deterministic (counter-based generator of rrtqx_3d_b200/workloads.py); re-running reproduces the committed file."""
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, HERE)
from refvec_io import write_arrays  # noqa: E402
from rrtqx_3d_b200 import workloads as W  # noqa: E402
from rrtqx_3d_b200.formats import read_polygon_obstacles  # noqa: E402


def main():
    A = {}
    # ---- kd trees: d = 2, 3, 4 and the Dubins tree (d = 4, theta wraps at 2 pi)
    for d in (2, 3, 4):
        A[f"kd{d}_pts"] = W.uniform_points(100 + d, 300, [-10.0] * d, [10.0] * d)
        A[f"kd{d}_qs"] = W.uniform_points(200 + d, 30, [-11.0] * d, [11.0] * d)
    A["kd_radii"] = np.array([0.0, 1.5, 4.0, 9.0])
    u = W.uniform01(301, 0, 3 * 300).reshape(300, 3)
    A["kdw_pts"] = np.stack([-10 + 20 * u[:, 0], -10 + 20 * u[:, 1], np.zeros(300), 2 * math.pi * u[:, 2]], axis=1)
    u = W.uniform01(302, 0, 3 * 30).reshape(30, 3)
    A["kdw_qs"] = np.stack([-10 + 20 * u[:, 0], -10 + 20 * u[:, 1], np.zeros(30), 2 * math.pi * u[:, 2]], axis=1)
    A["kdw_radii"] = np.array([1.0, 3.0, 5.0])          # 5 > period / 2: the dedup of ghost identities fires
    # ---- sphere world: building2 + segments near the surfaces, zero-length edges, inactive obstacles
    centers, radii, _ = W.building2_spheres()
    A["sph"] = np.column_stack([centers, radii])
    unused = np.zeros(len(radii), dtype=np.int64)
    unused[::7] = 1
    A["sph_unused"] = unused
    n = 400
    base = W.uniform_points(31, n, [-1.0] * 3, [1.0] * 3)
    base /= np.linalg.norm(base, axis=1, keepdims=True)
    k = (W.splitmix64(32, 0, n) % np.uint64(len(radii))).astype(int)
    scale = 0.97 + 0.06 * W.uniform01(33, 0, n)
    s = centers[k] + base * ((radii[k] + 0.5) * scale)[:, None]
    e = s + W.uniform_points(34, n, [-1.5] * 3, [1.5] * 3)
    e[::13] = s[::13]                                   # zero-length: collides with every active obstacle
    A["seg_s"], A["seg_e"] = s, e
    A["robot_radius"] = np.array([0.5])
    A["points"] = np.vstack([W.uniform_points(35, 150, [-20.0] * 3, [20.0] * 3), centers[:10] + 0.01])
    # ---- obstacle add / remove sweep on a 2000-node graph (out-edges within 2.4, chain parents)
    pts, _, _ = W.c2_workload(2000, 1)
    A["sw_pts"] = pts
    d2 = ((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1)
    src, dst = np.nonzero((d2 < 2.4 ** 2) & ~np.eye(len(pts), dtype=bool))
    A["sw_edges"] = np.column_stack([src, dst])
    A["sw_parent"] = np.arange(len(pts), dtype=np.int64) - 1
    A["sw_obstacles"] = np.array([[2.0, -3.0, 1.0, 3.0], [-8.0, 6.0, -4.0, 2.0], [pts[0, 0], pts[0, 1], pts[0, 2], 1.0]])
    A["sw_delta"] = np.array([W.DELTA])
    # ---- 2-D polygon world: the reference's own rand_Static.txt polygons + two balls
    polys = read_polygon_obstacles(os.path.join(HERE, "ref_rand_Static.txt"))[:12]
    A["poly_ptr"] = np.concatenate([[0], np.cumsum([len(p["polygon"]) for p in polys])])
    A["poly_xy"] = np.vstack([p["polygon"] for p in polys])
    A["balls2d"] = np.array([[5.0, 5.0, 2.0], [-20.0, 31.0, 4.0]])
    m = 300
    s2 = W.uniform_points(61, m, [-50.0, -50.0], [50.0, 50.0])
    e2 = s2 + W.uniform_points(62, m, [-6.0, -6.0], [6.0, 6.0])
    e2[::9, 0] = s2[::9, 0]                             # exactly vertical (the 1e-6 branch of segmentDistSqrd)
    e2[::11] = s2[::11]
    A["seg2_s"], A["seg2_e"] = s2, e2
    # ---- Dubins: pose pairs for calculateTrajectory / saturate, edges for the trajectory check
    u = W.uniform01(71, 0, 6 * 200).reshape(200, 6)
    st = np.stack([-50 + 100 * u[:, 0], -50 + 100 * u[:, 1], np.zeros(200), 2 * math.pi * u[:, 2]], axis=1)
    ang, rad = 2 * math.pi * u[:, 3], 8.0 * np.sqrt(u[:, 4])
    gl = np.stack([st[:, 0] + rad * np.cos(ang), st[:, 1] + rad * np.sin(ang), np.zeros(200), 2 * math.pi * u[:, 5]], axis=1)
    st[:10, 3] = 0.0
    gl[:10, 3] = 0.0                                    # aligned headings (degenerate tangents)
    A["dub_start"], A["dub_goal"] = st, gl
    A["dub_rmin"] = np.array([1.0])
    A["sat_delta"] = np.array([5.0])
    write_arrays(os.path.join(HERE, "reference_inputs.txt"), A)
    print("wrote", os.path.join(HERE, "reference_inputs.txt"), {k: np.asarray(v).shape for k, v in A.items()})


if __name__ == "__main__":
    main()
