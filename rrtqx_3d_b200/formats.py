"""On-disk formats of the reference, either side of the hot path.

Readers: the obstacle files the experiments load (`readDiscoverable3DObstaclesFromfile` DRRT_Q.jl:901-947,
`readDiscoverablecObstaclesFromfile` :853-899, both through `str2array` :171-210).
Writers: the per-slice dumps MATLAB consumes (`saveRRTNodes` / `saveRRTTree` / `saveRRTGraph` DRRT_Q.jl:252-337,
`saveObstacleLocations` :562-588), produced from bulk arrays (positions + kd visit order fetched from the
device in one copy each) instead of a recursive walk over a million heap objects.  Rows are in the order of the
reference's recursive kd traversal (node, kdChildL subtree, kdChildR subtree: `DeviceTree.preorder()`), numbers are
printed the way Julia 1.0 `writedlm` prints a Float64 (shortest round-trip digits, Grisu `_show` layout).
"""
from __future__ import annotations

import io
import math
from decimal import Decimal

import numpy as np


# ------------------------------------------------------------------ number formatting
def julia_float(x: float) -> str:
    """print(::Float64) of Julia 1.0 (base/grisu/grisu.jl `_show`, SHORTEST mode): shortest digits that round-trip;
    exponential iff the decimal point position pt (x = 0.DIGITS * 10^pt) satisfies pt <= -4 or pt > 6."""
    x = float(x)
    if math.isnan(x):
        return "NaN"
    if math.isinf(x):
        return "Inf" if x > 0 else "-Inf"
    sign = "-" if math.copysign(1.0, x) < 0 else ""
    if x == 0.0:
        return sign + "0.0"
    t = Decimal(repr(abs(x))).as_tuple()          # repr(): shortest round-trip digits, like Grisu SHORTEST
    raw = "".join(map(str, t.digits))
    digits = raw.rstrip("0") or "0"
    exp = t.exponent + (len(raw) - len(digits))   # value = int(digits) * 10^exp
    n = len(digits)
    pt = n + exp                                   # x = 0.DIGITS * 10^pt
    if pt <= -4 or pt > 6:                         # => #.#######e###
        return f"{sign}{digits[0]}.{digits[1:] or '0'}e{pt - 1}"
    if pt <= 0:                                    # => 0.00########
        return f"{sign}0.{'0' * (-pt)}{digits}"
    if n <= pt:                                    # => ########00.0
        return f"{sign}{digits}{'0' * (pt - n)}.0"
    return f"{sign}{digits[:pt]}.{digits[pt:]}"    # => ####.####


def writedlm_rows(fp, rows) -> None:
    """writedlm(fptr, A, ',') for a 2-D float array: one line per row, comma separated."""
    rows = np.asarray(rows, dtype=np.float64)
    if rows.ndim == 1:
        rows = rows.reshape(1, -1)
    for r in rows:
        fp.write(",".join(julia_float(v) for v in r))
        fp.write("\n")


# ------------------------------------------------------------------ readers
def str2array(s: str) -> np.ndarray:
    """str2array (DRRT_Q.jl:171-210): comma-separated numbers up to the first newline / NUL; no comma after the
    last number."""
    for stop in ("\n", "\0"):
        k = s.find(stop)
        if k >= 0:
            s = s[:k]
    return np.array([float(t) for t in s.split(",")], dtype=np.float64).reshape(1, -1)


_BEHAVIOUR = {0: (False, False, False), -1: (True, True, False), 1: (True, False, True)}


def _behaviour(code: int):
    """obsBehavourType -> (senseableObstacle, obstacleUnusedAfterSense, obstacleUnused), DRRT_Q.jl:873-889/924-940."""
    if code not in _BEHAVIOUR:
        raise ValueError("unknown behavoiur type")     # the reference's error text
    return _BEHAVIOUR[code]


def read_sphere_obstacles(path_or_file, obs_mult: int = 1):
    """readDiscoverable3DObstaclesFromfile: count; per obstacle `x, y, z` / radius / behaviour.  Returns a list of
    dicts (position 1x3, radius, lifeSpan = 300.0 as the reader sets it, the three behaviour flags), each obstacle
    repeated obs_mult times in file order (the caller pushes them to the FRONT of S.obstacles, like addObsToCSpace)."""
    f = open(path_or_file, "r") if isinstance(path_or_file, str) else path_or_file
    try:
        n = int(f.readline())
        out = []
        for _ in range(n):
            center = str2array(f.readline())[:, :3]
            radius = float(f.readline())
            sense, unused_after, unused = _behaviour(int(f.readline()))
            for _ in range(obs_mult):
                out.append({"position": center.copy(), "radius": radius, "lifeSpan": 300.0, "senseableObstacle": sense,
                            "obstacleUnusedAfterSense": unused_after, "obstacleUnused": unused})
        return out
    finally:
        if isinstance(path_or_file, str):
            f.close()


def read_polygon_obstacles(path_or_file, obs_mult: int = 1):
    """readDiscoverablecObstaclesFromfile: count; per polygon: vertex count, that many `x, y` lines, behaviour.
    Returns a list of dicts (kind 3, polygon Nx2, bounding circle as the Obstacle constructor computes it
    DRRT_data_structures.jl:229-241, behaviour flags)."""
    from .device import PolygonSet
    f = open(path_or_file, "r") if isinstance(path_or_file, str) else path_or_file
    try:
        n = int(f.readline())
        out = []
        for _ in range(n):
            nv = int(f.readline())
            poly = np.vstack([str2array(f.readline())[:, :2] for _ in range(nv)])
            sense, unused_after, unused = _behaviour(int(f.readline()))
            cx, cy, rad = PolygonSet.polygon_bound(poly)
            for _ in range(obs_mult):
                out.append({"kind": 3, "polygon": poly.copy(), "position": np.array([[cx, cy]]), "radius": rad,
                            "senseableObstacle": sense, "obstacleUnusedAfterSense": unused_after, "obstacleUnused": unused})
        return out
    finally:
        if isinstance(path_or_file, str):
            f.close()


# ------------------------------------------------------------------ writers (visualisation dumps)
def save_rrt_nodes(fp, positions, order, tree_cost, lmc) -> None:
    """saveRRTNodes (DRRT_Q.jl:316-337): one row [position rrtTreeCost rrtLMC] per node in kd visit order."""
    positions = np.asarray(positions, dtype=np.float64)
    rows = np.column_stack([positions[order], np.asarray(tree_cost)[order], np.asarray(lmc)[order]])
    writedlm_rows(fp, rows)


def save_rrt_tree(fp, positions, order, tree_cost, parent, parent_used=None) -> None:
    """saveRRTTree (DRRT_Q.jl:252-275): for every node with rrtParentUsed, in kd visit order, the row
    [node.position node.rrtTreeCost] followed by the same row of rrtParentEdge.endNode."""
    positions = np.asarray(positions, dtype=np.float64)
    tree_cost, parent = np.asarray(tree_cost, dtype=np.float64), np.asarray(parent)
    used = (parent >= 0) if parent_used is None else np.asarray(parent_used, dtype=bool)
    sel = order[used[order]]
    rows = np.empty((2 * len(sel), positions.shape[1] + 1))
    rows[0::2, :-1], rows[0::2, -1] = positions[sel], tree_cost[sel]
    rows[1::2, :-1], rows[1::2, -1] = positions[parent[sel]], tree_cost[parent[sel]]
    writedlm_rows(fp, rows)


def save_rrt_graph(fp, positions, order, row_ptr, col) -> None:
    """saveRRTGraph (DRRT_Q.jl:278-307): for every node in kd visit order and every edge of its rrtNeighborsOut
    (CSR row in list order, front first), the start position row then the end position row."""
    positions = np.asarray(positions, dtype=np.float64)
    row_ptr, col = np.asarray(row_ptr), np.asarray(col)
    deg = (row_ptr[1:] - row_ptr[:-1])[order]
    src = np.repeat(order, deg)
    starts = np.repeat(row_ptr[:-1][order], deg)
    within = np.arange(len(src)) - np.repeat(np.cumsum(deg) - deg, deg)
    dst = col[starts + within]
    rows = np.empty((2 * len(src), positions.shape[1]))
    rows[0::2], rows[1::2] = positions[src], positions[dst]
    writedlm_rows(fp, rows)


def save_obstacle_locations(fp, centers, radii, unused, expired=None) -> None:
    """saveObstacleLocations for the sphere list (DRRT_Q.jl:562-588): `x,y,z,radius` per obstacle that is neither
    unused nor expired, in list order."""
    centers, radii = np.asarray(centers, dtype=np.float64).reshape(-1, 3), np.asarray(radii, dtype=np.float64)
    skip = np.asarray(unused, dtype=bool)
    if expired is not None:
        skip = skip | np.asarray(expired, dtype=bool)
    keep = ~skip
    writedlm_rows(fp, np.column_stack([centers[keep], radii[keep]]))


def save_rrt_nodes_collision(fp, positions, order, tree_cost, lmc) -> None:
    """saveRRTNodesCollision (DRRT_Q.jl:340-362): [position min(rrtTreeCost, rrtLMC)] per node in kd visit order.
    Julia's `min` propagates NaN (either operand), unlike np.minimum's ordering of signed zeros only."""
    positions = np.asarray(positions, dtype=np.float64)
    a, b = np.asarray(tree_cost, dtype=np.float64)[order], np.asarray(lmc, dtype=np.float64)[order]
    writedlm_rows(fp, np.column_stack([positions[order], np.minimum(a, b)]))   # np.minimum propagates NaN too


def save_data(fp, data) -> None:
    """saveData (DRRT_Q.jl:41-48): every row of a matrix as one comma-separated line (robotMovePath dumps,
    rrtqx.jl:879)."""
    data = np.asarray(data, dtype=np.float64)
    if data.size:
        writedlm_rows(fp, data.reshape(data.shape[0], -1))


def save_rrt_path_q(fp, positions, node, root, parent, parent_used=None, trajectories=None, max_hops=1000) -> None:
    """saveRRTPath_Q (DRRT_Q.jl:3880-3891): follow parent edges from `node` while the node is not the root, its
    parent edge is in use and fewer than 1000 hops were taken; each hop writes saveEdgeTrajectory of the parent edge
    (SimpleEdge: start row then end row, DRRT_SimpleEdge_functions.jl:193-196; DubinsEdge: the trajectory rows,
    DRRT_DubinsEdge_functions.jl:734-736 - pass `trajectories`, a callable node -> rows of its parent edge);
    finally the first three coordinates of the node reached."""
    positions = np.asarray(positions, dtype=np.float64)
    parent = np.asarray(parent)
    used = (parent >= 0) if parent_used is None else np.asarray(parent_used, dtype=bool)
    this, hops = int(node), 0
    while this != root and used[this] and hops < max_hops:
        if trajectories is None:
            writedlm_rows(fp, positions[[this, int(parent[this])]])
        else:
            writedlm_rows(fp, trajectories(this))
        this = int(parent[this])
        hops += 1
    writedlm_rows(fp, positions[this, :3])


def save_kds_q(fp, kino_dist, aug_dist) -> None:
    """saveKds_Q (DRRT_Q.jl:3873-3878): the two scalars, one per line (writedlm of a scalar)."""
    fp.write(julia_float(kino_dist) + "\n")
    fp.write(julia_float(aug_dist) + "\n")


def dump_to_string(writer, *args, **kw) -> str:
    buf = io.StringIO()
    writer(buf, *args, **kw)
    return buf.getvalue()
