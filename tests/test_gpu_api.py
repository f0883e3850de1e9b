"""The reference-shaped API (kdInsert / kdFindNearest / kdFindWithinRange / explicitEdgeCheck /
addNewObstacle ...) on the GPU, checked the way the reference's own (commented) tests check the kd
tree: against the naive whole-tree search (kdTree_general.jl:1039-1087), here the CPU oracle."""
import math

import numpy as np
import pytest

import oracle
from rrtqx_3d_b200 import workloads as W
from rrtqx_3d_b200.collision import (explicitEdgeCheck, explicitNodeCheck, explicitPointCheck3D, calculateTrajectory)
from rrtqx_3d_b200.kdtree import (KDTree, KDdist, emptyRangeList, kdFindMoreWithinRange, kdFindNearest, kdFindWithinRange,
                                  kdInsert, kdInsertBatch, popFromRangeList)
from rrtqx_3d_b200.structures import (CSpace, RRTNode, RobotData, SphereObstacle, addObsToCSpace, newEdge)
from rrtqx_3d_b200.sweep import EdgeMirror, addNewObstacle, findPointsInConflictWithObstacle, removeObstacle

pytestmark = pytest.mark.gpu


def _space(building2):
    centers, radii, _ = building2
    S = CSpace(3, -1.0, [-20.0] * 3, [20.0] * 3, [-14.9, -13.5, -7.5], [4.0, 16.5, -7.5])
    S.robotRadius, S.delta = W.ROBOT_RADIUS, W.DELTA
    for c, r in zip(centers, radii):
        addObsToCSpace(S, SphereObstacle(c, r))
    return S


def test_kd_api_matches_naive(ctx):
    pts, qs, _ = W.c2_workload(3000, 60)
    KD = KDTree(ctx, 3, KDdist)
    nodes = [RRTNode(p) for p in pts]
    for n in nodes[:50]:
        kdInsert(KD, n)                       # planner-style single inserts
    kdInsert(KD, nodes[3])                    # idempotent via kdInTree (kdTree_general.jl:122-124)
    kdInsertBatch(KD, nodes[50:])
    assert KD.treeSize == 3000 and KD.root is nodes[0]
    orc = oracle.KDTree(3)
    orc.insert_batch(pts)
    par, cl, cr, sp = orc.fields()
    for i in (0, 1, 17, 49, 50, 2999):        # kd fields of the host nodes stay populated
        n = nodes[i]
        assert n.kdSplit == sp[i] + 1
        assert (n.kdParent.kdIndex if n.kdParentExist else -1) == par[i]
        assert (n.kdChildL.kdIndex if n.kdChildLExist else -1) == cl[i]
        assert (n.kdChildR.kdIndex if n.kdChildRExist else -1) == cr[i]
    for q in qs:
        node, d = kdFindNearest(KD, q)
        oi, od = orc.find_nearest(q, naive=True)
        assert node.kdIndex == oi and d == od
        L = kdFindWithinRange(KD, 5.0, q)
        oidx, okey = orc.find_within_range(5.0, q)       # marks stay set, as the reference's inHeap flags do
        got = {n.data.kdIndex: n.key for n in L}
        assert got == dict(zip(oidx.tolist(), okey.tolist()))
        assert all(n.data.inHeap for n in L)                  # members are marked until the list is emptied
        # kdFindMoreWithinRange from a second point: union, no duplicates (addToRangeList dedup)
        before = L.length
        kdFindMoreWithinRange(KD, 5.0, q + 1.0, L)
        o2, _ = orc.find_within_range(5.0, q + 1.0, prev=(oidx, okey))
        orc.empty(o2)
        assert L.length == len(o2) >= before
        assert sorted(n.data.kdIndex for n in L) == sorted(o2.tolist())
        node0, key0 = popFromRangeList(L)
        assert not node0.inHeap
        emptyRangeList(L)
        assert L.length == 0 and not any(n.inHeap for n in nodes)


def test_explicit_checks_match_reference_semantics(ctx, building2):
    S = _space(building2)
    sph, ns = oracle.make_spheres(building2[0], building2[1])
    L = oracle.lib()
    P = lambda a: oracle._p(np.ascontiguousarray(a, dtype=np.float64), oracle.c_f64p)
    import ctypes as C
    pts = W.uniform_points(51, 200, [-20.0] * 3, [20.0] * 3)
    for i in range(0, 198, 2):
        e = newEdge(RRTNode(pts[i]), RRTNode(pts[i] + (pts[i + 1] - pts[i]) * 0.1))
        calculateTrajectory(S, e)
        assert e.dist == e.distOriginal == L.orc_euclid(P(e.startNode.position), P(e.endNode.position), 3)
        # obstacle order of the oracle = S.obstacles order (newest first); OR is order independent
        assert explicitEdgeCheck(ctx, S, e) == bool(L.orc_edge_check_all(sph, ns, 0, P(e.startNode.position),
                                                                        P(e.endNode.position), S.robotRadius, 0))
        ob = list(S.obstacles)[i % 31]
        one, _ = oracle.make_spheres(ob.position, [ob.radius])
        assert explicitEdgeCheck(ctx, S, e, ob) == bool(L.orc_edge_check_sphere(one, P(e.startNode.position),
                                                                             P(e.endNode.position), S.robotRadius, 0))
        c = C.c_double()
        hit = L.orc_point_check(sph, ns, 0, P(pts[i]), S.robotRadius, C.byref(c))
        assert explicitNodeCheck(ctx, S, RRTNode(pts[i])) == (bool(hit), c.value)
        hit3 = L.orc_point_check_3d(sph, ns, 0, P(pts[i]), S.robotRadius, C.byref(c))
        assert explicitPointCheck3D(ctx, S, pts[i]) == (bool(hit3), c.value)
    S.inWarmupTime = True                      # warm-up: obstacles are ignored (DRRT_Q.jl:1805-1807,1523-1525)
    assert explicitEdgeCheck(ctx, S, e) is False and explicitNodeCheck(ctx, S, RRTNode(pts[0])) == (False, math.inf)


class _Queue:
    def __init__(self):
        self.os, self.q = [], []

    def verifyInOSQueue(self, n):
        self.os.append(n)

    def verifyInQueue(self, n):
        self.q.append(n)

    # removeObstacle queues a node only if it became inconsistent AND lessQ(node, moveGoal) (DRRT_Q.jl:3352-3357)
    def recalculateLMCMineVTwo(self, n, root, r):
        if n.kdIndex % 2 == 0:
            n.rrtLMC = 1.0          # rrtTreeCost stays Inf: inconsistent

    def lessQ(self, a, b):
        return a.kdIndex % 3 != 0


def test_add_and_remove_obstacle_mutations(ctx):
    pts, _, _ = W.c2_workload(4000, 1)
    S = CSpace(3, -1.0, [-20.0] * 3, [20.0] * 3, pts[0], pts[1])
    S.robotRadius, S.delta = W.ROBOT_RADIUS, W.DELTA
    KD = KDTree(ctx, 3, KDdist)
    nodes = [RRTNode(p) for p in pts]
    kdInsertBatch(KD, nodes)
    # RRTx-style graph: every node linked to its neighbours within 3.0 (initial out lists), chain parents
    from rrtqx_3d_b200.kdtree import kdFindWithinRangeBatch
    counts, offsets, idx, dist = kdFindWithinRangeBatch(KD, 3.0, pts)
    for i, n in enumerate(nodes):
        for j, dd in zip(idx[offsets[i]:offsets[i] + counts[i]], dist[offsets[i]:offsets[i] + counts[i]]):
            if j != i:
                e = newEdge(n, nodes[int(j)])
                e.dist = e.distOriginal = float(dd)
                n.InitialNeighborListOut.push(e)
        if i > 0:
            pe = newEdge(n, nodes[i - 1])
            n.rrtParentEdge, n.rrtParentUsed = pe, True
            n.successorListItemInParent = nodes[i - 1].SuccessorList.push(pe)
    ob = SphereObstacle([2.0, -3.0, 1.0], 3.0)
    ob.obstacleUnused = True                      # "appearing" obstacle (DRRT_Q.jl:923-927)
    other = SphereObstacle([4.0, -3.0, 1.0], 2.5)
    addObsToCSpace(S, other)
    addObsToCSpace(S, ob)
    Q, R = _Queue(), RobotData()
    # unit-length robot edge through the obstacle centre.  (A LONG edge through it would NOT be reported:
    # the reference projects with dot/L instead of dot/L^2, DRRT_Q.jl:1208 -- reproduced bit for bit.)
    R.robotEdgeUsed, R.robotEdge = True, newEdge(RRTNode([2.0, -3.0, 1.5]), RRTNode([2.0, -3.0, 0.5]))
    long_edge = newEdge(RRTNode([2.0, -3.0, 5.0]), RRTNode([2.0, -3.0, -5.0]))
    assert explicitEdgeCheck(ctx, S, long_edge, ob) is False
    mirror = EdgeMirror(KD).rebuild()
    Lc = findPointsInConflictWithObstacle(S, KD, ob, nodes[0])
    cand = {n.data.kdIndex for n in Lc}
    emptyRangeList(Lc)
    blocked, orphans = addNewObstacle(S, KD, Q, ob, nodes[0], 0, R, edges=mirror)
    assert not ob.obstacleUnused and R.currentMoveInvalid
    # oracle decision per edge
    Lo = oracle.lib()
    P = lambda a: oracle._p(np.ascontiguousarray(a, dtype=np.float64), oracle.c_f64p)
    one, _ = oracle.make_spheres(ob.position, [ob.radius])
    n_blocked = 0
    for n in nodes:
        for item in n.InitialNeighborListOut:
            e = item.data
            want = n.kdIndex in cand and bool(Lo.orc_edge_check_sphere(one, P(n.position), P(e.endNode.position), 0.5, 0))
            assert (e.dist == math.inf) == want
            n_blocked += want
    assert n_blocked == len(blocked) > 0
    orphan_set = {n.kdIndex for n in Q.os}
    assert orphan_set == set(orphans.tolist()) and len(orphan_set) > 0
    for n in Q.os:
        assert not n.rrtParentUsed and n.rrtParentEdge.endNode is n and n.rrtParentEdge.dist == math.inf
    # QX removeObstacle: obstacle disabled before the loop -> nothing restored
    mirror.rebuild()
    restored, requeue = removeObstacle(S, KD, Q, ob, nodes[0], 3.0, 0.0, nodes[0], edges=mirror, qx_semantics=True)
    assert len(restored) == 0 and ob.obstacleUnused and ob.expired
    # Otte semantics: edges blocked only by `ob` come back; those also hit by `other` stay blocked
    ob.obstacleUnused = False
    restored, requeue = removeObstacle(S, KD, Q, ob, nodes[0], 3.0, 0.0, nodes[0], edges=mirror, qx_semantics=False)
    oth, _ = oracle.make_spheres(other.position, [other.radius])
    for n in nodes:
        for item in n.InitialNeighborListOut:
            e = item.data
            hit_other = bool(Lo.orc_edge_check_sphere(oth, P(n.position), P(e.endNode.position), 0.5, 0))
            hit_ob = n.kdIndex in cand and bool(Lo.orc_edge_check_sphere(one, P(n.position), P(e.endNode.position), 0.5, 0))
            assert (e.dist == math.inf) == (hit_ob and hit_other)
    want_q = {v for v in requeue.tolist() if v % 2 == 0 and v % 3 != 0}
    assert len(restored) > 0 and {n.kdIndex for n in Q.q} == want_q and 0 < len(want_q) < len(requeue)
