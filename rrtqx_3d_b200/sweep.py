"""Obstacle add / remove sweeps with the reference's names and side effects
(DRRT_Q.jl:3195-3362).  The GPU decides WHICH edges are blocked / restored and which
nodes are orphaned; this module applies the reference's mutations to the host
structures (edge.dist = Inf, rrtParentUsed = false, successor-list surgery, queue
callbacks), so the planner above it is unchanged.
"""
from __future__ import annotations

import math

import numpy as np

from . import _abi as A
from .collision import _mirror, explicitEdgeCheck
from .device import EdgeSet, SweepResult
from .kdtree import KDTree, emptyRangeList, kdFindWithinRange
from .structures import CSpace, JList, RRTNode, SphereObstacle


def findPointsInConflictWithObstacle(S: CSpace, KD: KDTree, ob: SphereObstacle, root=None) -> JList:
    """findPointsInConflictWithObstacle (DRRT_Q.jl:3195-3215).  The returned range list must be
    destroyed with emptyRangeList, as in the reference."""
    if not S.spaceHasTime and not S.spaceHasTheta:
        searchRange = S.robotRadius + S.delta + ob.radius
        return kdFindWithinRange(KD, searchRange, ob.position)
    if not S.spaceHasTime and S.spaceHasTheta:
        searchRange = S.robotRadius + S.delta + ob.radius + math.pi
        p = ob.position.reshape(-1)
        return kdFindWithinRange(KD, searchRange, np.array([[p[0], p[1], p[2], 0.0, math.pi]])[:, :KD.d])
    raise RuntimeError("this type of obstacle not coded for this type of space")


class EdgeMirror:
    """Device mirror of every node's out-edge lists (InitialNeighborListOut then rrtNeighborsOut,
    the order of RRTNodeNeighborIterator, DRRT_Q.jl:2408-2431) and parent edges."""

    def __init__(self, KD: KDTree):
        self.KD = KD
        self.set = EdgeSet(KD.dev)
        self.edges = []        # edge id -> (JListNode holding the edge)
        self.result = SweepResult(KD.ctx)

    def rebuild(self):
        nodes = self.KD.nodes
        src, dst, items = [], [], []
        parent = np.full(len(nodes), -1, dtype=np.int32)
        for n in nodes:
            for lst in (n.InitialNeighborListOut, n.rrtNeighborsOut):
                for item in lst:
                    e = item.data
                    src.append(n.kdIndex)
                    dst.append(e.endNode.kdIndex)
                    items.append(item)
            if n.rrtParentUsed and n.rrtParentEdge is not None:
                parent[n.kdIndex] = n.rrtParentEdge.endNode.kdIndex
        self.edges = items
        self.set.upload(np.asarray(src, dtype=np.int32), np.asarray(dst, dtype=np.int32), parent)
        return self


def addNewObstacle(S: CSpace, KD: KDTree, Q, ob: SphereObstacle, root, fileCounter, R, edges: EdgeMirror | None = None,
                   flags: int = 0):
    """addNewObstacle (DRRT_Q.jl:3220-3290).  Q may provide verifyInOSQueue(node)."""
    ctx = KD.ctx
    ob.obstacleUnused = False                                    # :3222
    if edges is None:
        edges = EdgeMirror(KD).rebuild()
    spheres = _mirror(S, ctx).sync(S)
    if ob.deviceId < 0 or ob not in list(S.obstacles):
        raise ValueError("obstacle must be in S.obstacles (addObsToCSpace) before addNewObstacle")
    if ob.lifeSpan > 0:                                          # explicitEdgeCheck3D early-out (:1777)
        res = edges.set.add_sweep(spheres, [ob.deviceId], S.robotRadius, S.delta, flags, edges.result)
        blocked, orphans = res.fetch()
    else:
        blocked, orphans = np.zeros(0, np.int32), np.zeros(0, np.int32)
    for e in blocked:                                            # :3248-3249
        edges.edges[int(e)].data.dist = math.inf
    for v in orphans:                                            # :3257-3270
        thisNode = KD.nodes[int(v)]
        pe = thisNode.rrtParentEdge
        if thisNode.successorListItemInParent is not None:
            pe.endNode.SuccessorList.remove(thisNode.successorListItemInParent)
        pe.endNode = thisNode
        pe.dist = math.inf
        thisNode.rrtParentUsed = False
        if Q is not None and hasattr(Q, "verifyInOSQueue"):
            Q.verifyInOSQueue(thisNode)
    if R is not None and R.robotEdgeUsed and explicitEdgeCheck(ctx, S, R.robotEdge, ob, flags):   # :3287-3289
        R.currentMoveInvalid = True
    return blocked, orphans


def removeObstacle(S: CSpace, KD: KDTree, Q, ob: SphereObstacle, root, hyberBallRad, timeElapsed, moveGoal,
                   edges: EdgeMirror | None = None, qx_semantics: bool = True, flags: int = 0):
    """removeObstacle (DRRT_Q.jl:3295-3362).  qx_semantics=True reproduces the QX fork, which sets
    ob.obstacleUnused = true BEFORE the loop so nothing is ever restored;
    False gives the Otte generation (DRRT.jl:3202-3268).  Q may provide
    recalculateLMCMineVTwo(node, root, r), verifyInQueue(node) and lessQ(a, b) (the key comparison of
    DRRT_Q.jl:3355; when Q has no lessQ the comparison counts as true)."""
    ctx = KD.ctx
    if edges is None:
        edges = EdgeMirror(KD).rebuild()
    spheres = _mirror(S, ctx).sync(S)                            # before the flags change (indices)
    ob_id = ob.deviceId
    ob.expired = True
    ob.obstacleUnused = True                                     # :3301-3302
    others = [o.deviceId for o in S.obstacles
              if o is not ob and not o.obstacleUnused and o.startTime <= timeElapsed <= (o.startTime + o.lifeSpan)
              and o.lifeSpan > 0]                                # :3330 + explicitEdgeCheck3D early-out
    inf = np.array([1 if item.data.dist == math.inf else 0 for item in edges.edges], dtype=np.uint8)
    f = flags | (A.SWEEP_REMOVED_INACTIVE if qx_semantics else 0)
    res = edges.set.remove_sweep(spheres, ob_id, others, inf, S.robotRadius, S.delta, f, edges.result)
    restored, requeue = res.fetch()
    for e in restored:                                           # :3340-3346
        edge = edges.edges[int(e)].data
        edge.dist = edge.distOriginal
    for v in requeue:                                            # :3352-3357
        n = KD.nodes[int(v)]
        if Q is not None and hasattr(Q, "recalculateLMCMineVTwo"):
            Q.recalculateLMCMineVTwo(n, root, hyberBallRad)
        # :3354-3356  only nodes that became inconsistent and are relevant to the robot are queued
        less = getattr(Q, "lessQ", None) if Q is not None else None
        if (Q is not None and hasattr(Q, "verifyInQueue") and n.rrtTreeCost != n.rrtLMC
                and (less is None or less(n, moveGoal))):
            Q.verifyInQueue(n)
    ob.obstacleUnused = True                                     # :3361
    return restored, requeue
