"""Collision checks with the reference's names (SimpleEdge / sphere world of DRRT_Q.jl)
running on the GPU: explicitEdgeCheck, explicitPointCheck(3D), explicitNodeCheck(3D),
plus batched twins.  Obstacles are mirrored from `S.obstacles` to a device sphere set;
the mirror is refreshed whenever the list, a radius or an active flag changed.
"""
from __future__ import annotations

import math

import numpy as np

from . import _abi as A
from .device import Context, SphereSet, edge_check_batch, node_check_batch, segment_check_batch
from .structures import CSpace, SimpleEdge, SphereObstacle


class ObstacleMirror:
    """Device copy of CSpace.obstacles (front-to-back order = device order)."""

    def __init__(self, ctx: Context):
        self.ctx = ctx
        self.set = SphereSet(ctx)
        self._sig = None

    def sync(self, S: CSpace) -> SphereSet:
        obs = list(S.obstacles)
        sig = tuple((id(o), o.radius, o.active(), o.position.tobytes()) for o in obs)
        if sig != self._sig:
            n = len(obs)
            centers = np.zeros((n, 3))
            radii = np.zeros(n)
            active = np.zeros(n, dtype=np.uint8)
            for i, o in enumerate(obs):
                centers[i] = o.position.reshape(-1)[:3]
                radii[i] = o.radius
                active[i] = 1 if o.active() else 0
                o.deviceId = i
            self.set.upload(centers, radii, active)
            self._sig = sig
        return self.set


def _mirror(S: CSpace, ctx: Context) -> ObstacleMirror:
    m = getattr(S, "_gpu_mirror", None)
    if m is None or m.ctx is not ctx:
        m = ObstacleMirror(ctx)
        S._gpu_mirror = m
    return m


def calculateTrajectory(S: CSpace, edge: SimpleEdge):
    """calculateTrajectory, SimpleEdge (DRRT_SimpleEdge_functions.jl:177-181)."""
    from .kdtree import euclidianDist
    a, b = edge.startNode.position, edge.endNode.position
    edge.dist = euclidianDist(a, b)                       # dist
    edge.distOriginal = edge.dist
    edge.Wdist = euclidianDist(a[0, :3], b[0, :3])        # Wdist on [1:3]


def saturate(newPoint, closestPoint, delta):
    """saturate, SimpleEdge (DRRT_SimpleEdge_functions.jl:69-74): the reference rebinds its local
    `newPoint`, so the caller's array is never modified -- a no-op, reproduced as such."""
    return None


def explicitEdgeCheck(ctx: Context, S: CSpace, edge: SimpleEdge, obstacle: SphereObstacle | None = None,
                      flags: int = 0) -> bool:
    """explicitEdgeCheck(S, edge) (DRRT_Q.jl:1802-1826) or, with `obstacle`, the per-obstacle form
    explicitEdgeCheck(S, edge, ob) (DRRT_SimpleEdge_functions.jl:210-212)."""
    if obstacle is None and S.inWarmupTime:
        return False
    s = edge.startNode.position.reshape(1, -1)[:, :3]
    e = edge.endNode.position.reshape(1, -1)[:, :3]
    if obstacle is not None:
        one = SphereSet(ctx, obstacle.position, [obstacle.radius], [1 if obstacle.active() else 0])
        return bool(segment_check_batch(ctx, one, s, e, S.robotRadius, flags)[0])
    spheres = _mirror(S, ctx).sync(S)
    return bool(segment_check_batch(ctx, spheres, s, e, S.robotRadius, flags)[0])


def explicitEdgeCheckBatch(ctx: Context, S: CSpace, tree, src_idx, dst_idx, flags: int = 0) -> np.ndarray:
    """Batched explicitEdgeCheck over edges between nodes of a device tree (indices)."""
    if S.inWarmupTime:
        return np.zeros(len(src_idx), dtype=np.uint8)
    spheres = _mirror(S, ctx).sync(S)
    return edge_check_batch(tree.dev if hasattr(tree, "dev") else tree, spheres, src_idx, dst_idx, S.robotRadius, flags)


def explicitPointCheck(ctx: Context, S: CSpace, point):
    """explicitPointCheck (DRRT_Q.jl:1520-1556) -> (Bool, certificate)."""
    if S.inWarmupTime:
        return False, math.inf
    spheres = _mirror(S, ctx).sync(S)
    p = np.asarray(point, dtype=np.float64).reshape(1, -1)[:, :3]
    hit, cert = node_check_batch(ctx, spheres, p, S.robotRadius, A.CHECK_QUICK_PASS)
    return bool(hit[0]), float(cert[0])


def explicitPointCheck3D(ctx: Context, S: CSpace, point):
    """explicitPointCheck3D (DRRT_Q.jl:1558-1590): no quickCheck pass."""
    if S.inWarmupTime:
        return False, math.inf
    spheres = _mirror(S, ctx).sync(S)
    p = np.asarray(point, dtype=np.float64).reshape(1, -1)[:, :3]
    hit, cert = node_check_batch(ctx, spheres, p, S.robotRadius, 0)
    return bool(hit[0]), float(cert[0])


def explicitNodeCheck(ctx, S, node):
    """explicitNodeCheck (DRRT_Q.jl:1594)."""
    return explicitPointCheck(ctx, S, node.position)


def explicitNodeCheck3D(ctx, S, node):
    """explicitNodeCheck3D (DRRT_Q.jl:1595)."""
    return explicitPointCheck3D(ctx, S, node.position)


def explicitPointCheckBatch(ctx: Context, S: CSpace, points, quick_pass=True):
    if S.inWarmupTime:
        n = len(points)
        return np.zeros(n, dtype=np.uint8), np.full(n, math.inf)
    spheres = _mirror(S, ctx).sync(S)
    return node_check_batch(ctx, spheres, np.asarray(points, dtype=np.float64).reshape(-1, 3), S.robotRadius,
                            A.CHECK_QUICK_PASS if quick_pass else 0)
