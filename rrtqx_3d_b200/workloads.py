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
