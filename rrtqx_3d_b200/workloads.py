"""Seeded synthetic workloads for the BASELINE.json configs.

RNG: counter-based splitmix64.  Sample i of stream `seed` is
    x = mix(seed * 0x9E3779B97F4A7C15 + (i + 1) * 0x9E3779B97F4A7C15)
    u = (x >> 11) * 2**-53                       in [0, 1)
and a point is built as `lo + u .* width` in that operation order, the order of
the reference sampler (DRRT_Q.jl:600  S.lowerBounds + rand(1,d) .* S.width).
Pure numpy; no dependency on the CUDA library or the oracle.
"""
from __future__ import annotations

import math
import os

import numpy as np

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)
_M1 = np.uint64(0xBF58476D1CE4E5B9)
_M2 = np.uint64(0x94D049BB133111EB)


def splitmix64(seed: int, start: int, count: int) -> np.ndarray:
    """uint64 outputs for counters start .. start+count-1 of stream `seed`."""
    with np.errstate(over="ignore"):
        ctr = np.arange(start + 1, start + 1 + count, dtype=np.uint64)
        z = (np.uint64(seed) * _GOLDEN) + ctr * _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * _M1
        z = (z ^ (z >> np.uint64(27))) * _M2
        z = z ^ (z >> np.uint64(31))
    return z


def uniform01(seed: int, start: int, count: int) -> np.ndarray:
    """float64 in [0,1): (x >> 11) * 2^-53."""
    return (splitmix64(seed, start, count) >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def uniform_points(seed: int, n: int, lo, hi, start: int = 0) -> np.ndarray:
    """n x d points, `lo + u .* width` (DRRT_Q.jl:600)."""
    lo = np.asarray(lo, dtype=np.float64).reshape(1, -1)
    hi = np.asarray(hi, dtype=np.float64).reshape(1, -1)
    d = lo.shape[1]
    width = hi - lo  # CSpace.width = U - L (DRRT_data_structures.jl:389)
    u = uniform01(seed, start * d, n * d).reshape(n, d)
    return lo + u * width


def shrinking_ball_radius(n: int, d: int, delta: float, ball_constant: float) -> float:
    """rrtqx.jl:382  min(delta, ballConstant*((log(1+n)/n)^(1/d)))."""
    return min(delta, ball_constant * ((math.log(1 + n) / n) ** (1 / d)))


# ---- reference constants (experimentsForRRTQX.jl:37-41,76-77,104,132) ----
ENV_RAD = 20.0
ROBOT_RADIUS = 0.5
DELTA = 8.0
BALL_CONSTANT = 80.0
C2_RADIUS = 1.9196069361095438  # r(n=1e6, d=3); float.fromhex('0x1.eb6b5c33c3e79p+0')
C2_SPARSE_RADIUS = 0.7795


def c2_workload(n_nodes: int = 1_000_000, n_queries: int = 1_000_000, radius: float | None = None,
                seed_nodes: int = 1, seed_queries: int = 2):
    """Config C2: uniform tree + uniform queries in [-20,20]^3 at the RRTx ball radius."""
    lo, hi = [-ENV_RAD] * 3, [ENV_RAD] * 3
    pts = uniform_points(seed_nodes, n_nodes, lo, hi)
    qs = uniform_points(seed_queries, n_queries, lo, hi)
    if radius is None:
        radius = shrinking_ball_radius(n_nodes, 3, DELTA, BALL_CONSTANT)
    return pts, qs, float(radius)


def c3_obstacles(n_obs: int = 256, seed: int = 3, rmin: float = 1.0, rmax: float = 3.5):
    """Config C3: sphere obstacles, centres uniform in the box, radii U[rmin,rmax]."""
    lo, hi = [-ENV_RAD] * 3, [ENV_RAD] * 3
    centres = uniform_points(seed, n_obs, lo, hi)
    u = uniform01(seed + 1000, 0, n_obs)
    radii = rmin + u * (rmax - rmin)
    return centres, radii


def read_sphere_obstacle_file(path: str):
    """Sphere-obstacle text format of readDiscoverable3DObstaclesFromfile
    (DRRT_Q.jl:901-947): count; then per obstacle `x, y, z` / radius / behaviour.
    Returns centres (n x 3), radii (n), behaviour (n; 0 normal, -1 vanishing, 1 appearing)."""
    with open(path, "r") as f:
        lines = [ln.strip() for ln in f.readlines()]
    p = int(lines[0])
    centres = np.zeros((p, 3))
    radii = np.zeros(p)
    beh = np.zeros(p, dtype=np.int32)
    k = 1
    for i in range(p):
        centres[i] = [float(t) for t in lines[k].split(",")[:3]]
        radii[i] = float(lines[k + 1])
        beh[i] = int(lines[k + 2])
        k += 3
    return centres, radii, beh


def building2_spheres():
    """Centres (31 x 3), radii, behaviour of the reference fixture building2.txt,
    from the committed CSV tests/golden/building2_spheres.csv."""
    path = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden",
                        "building2_spheres.csv")
    rows = np.loadtxt(path, delimiter=",", comments="#")
    return np.ascontiguousarray(rows[:, 0:3]), np.ascontiguousarray(rows[:, 3]), rows[:, 4].astype(np.int32)


# ------------------------------------------------------------ config C4 (Dubins, 2-D polygon world)
def c4_city_blocks(n_side: int = 10, half: float = 3.0, pitch: float = 10.0, first: float = -45.0):
    """10 x 10 grid of 6 x 6 square city blocks centred at (first + pitch*i, first + pitch*j), as kind-3 polygons."""
    obs = []
    for j in range(n_side):
        for i in range(n_side):
            cx, cy = first + pitch * i, first + pitch * j
            obs.append(("polygon", np.array([[cx - half, cy - half], [cx + half, cy - half], [cx + half, cy + half],
                                             [cx - half, cy + half]])))
    return obs


def arc_line_arc_trajectory(start, end, r_turn: float, step: float = 0.1):
    """A Dubins-shaped polyline (left arc, straight, left arc) from pose `start` (x, y, theta) towards `end`,
    arcs sampled every `step` rad like the reference's collect(phi_start:0.1:phi_end)
    (DRRT_DubinsEdge_functions.jl:506-655).  Synthetic workload generator -- NOT the reference's solver; the
    GPU check is bit-exact given the trajectory points, whatever produced them."""
    x0, y0, th0 = start
    x1, y1, th1 = end
    pts = [(x0, y0)]
    c0 = (x0 - r_turn * math.sin(th0), y0 + r_turn * math.cos(th0))          # left-turn circle at start
    c1 = (x1 - r_turn * math.sin(th1), y1 + r_turn * math.cos(th1))          # left-turn circle at end
    hdg = math.atan2(c1[1] - c0[1], c1[0] - c0[0])                            # LSL: tangent heading
    a0 = (hdg - th0) % (2 * math.pi)
    k = 1
    while k * step < a0:
        a = th0 + k * step
        pts.append((c0[0] + r_turn * math.sin(a), c0[1] - r_turn * math.cos(a)))
        k += 1
    pts.append((c0[0] + r_turn * math.sin(hdg), c0[1] - r_turn * math.cos(hdg)))
    pts.append((c1[0] + r_turn * math.sin(hdg), c1[1] - r_turn * math.cos(hdg)))
    a1 = (th1 - hdg) % (2 * math.pi)
    k = 1
    while k * step < a1:
        a = hdg + k * step
        pts.append((c1[0] + r_turn * math.sin(a), c1[1] - r_turn * math.cos(a)))
        k += 1
    pts.append((x1, y1))
    return np.asarray(pts, dtype=np.float64)


def c4_workload(n_nodes: int = 200_000, n_edges: int = 100_000, r_edge: float = 4.0, r_turn: float = 1.0, seed: int = 4):
    """Config C4 pieces: nodes uniform in [-50,50]^2 x {0} x [0,2pi) (dubinsExperimentsForPaper.jl:77-78), random
    directed edges of length <= r_edge with arc-line-arc trajectories (CSR)."""
    lo, hi = [-50.0, -50.0, 0.0, 0.0], [50.0, 50.0, 0.0, 2.0 * math.pi]
    nodes = uniform_points(seed, n_nodes, lo, hi)
    u = uniform01(seed + 1, 0, 4 * n_edges).reshape(n_edges, 4)
    src = (u[:, 0] * n_nodes).astype(np.int64)
    ang, rad = 2 * math.pi * u[:, 1], r_edge * np.sqrt(u[:, 2])
    starts = nodes[src][:, [0, 1, 3]]
    ends = np.stack([starts[:, 0] + rad * np.cos(ang), starts[:, 1] + rad * np.sin(ang), 2 * math.pi * u[:, 3]], axis=1)
    ptr = np.zeros(n_edges + 1, dtype=np.int64)
    chunks = []
    for e in range(n_edges):
        t = arc_line_arc_trajectory(starts[e], ends[e], r_turn)
        chunks.append(t)
        ptr[e + 1] = ptr[e] + len(t)
    return nodes, starts, ends, ptr, np.concatenate(chunks, axis=0)
