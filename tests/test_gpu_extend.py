"""Config C1 replay: the planner's per-iteration geometric work (nearest -> node check -> range(r(n)) ->
2k edge checks -> insert) through the fused rrtqx_extend_query, every iteration compared with the oracle."""
import ctypes as C

import numpy as np
import pytest

import oracle
from rrtqx_3d_b200 import _abi as A
from rrtqx_3d_b200 import workloads as W
from rrtqx_3d_b200.device import DeviceTree, SphereSet, extend_query

pytestmark = pytest.mark.gpu


def test_c1_replay_matches_oracle(ctx, building2):
    centers, radii, _ = building2
    sph, ns = oracle.make_spheres(centers, radii)          # all 31 spheres forced active
    S = SphereSet(ctx, centers, radii)
    L = oracle.lib()
    P = lambda a: oracle._p(np.ascontiguousarray(a, dtype=np.float64), oracle.c_f64p)
    n_iter = 20000                                          # the full C1 configuration
    samples = W.uniform_points(1, n_iter, [-W.ENV_RAD] * 3, [W.ENV_RAD] * 3)
    t = DeviceTree(ctx, 3)
    orc = oracle.KDTree(3)
    goal = np.array([4.0, 16.5, -7.5])                      # root of the search tree (experimentsForRRTQX.jl:41)
    t.insert(goal)
    orc.insert(goal)
    bufs = None
    accepted = 0
    for it in range(n_iter):
        p = samples[it]
        n = len(orc)
        r = W.shrinking_ball_radius(n, 3, W.DELTA, W.BALL_CONSTANT)      # rrtqx.jl:382, n = treeSize before the insert
        res = extend_query(t, S, p, r, W.ROBOT_RADIUS, A.CHECK_QUICK_PASS, capacity=4096)
        oi, od = orc.find_nearest(p)
        assert (res.nearest_idx, res.nearest_dist) == (oi, od)
        c = C.c_double()
        hit = L.orc_point_check(sph, ns, 0, P(p), W.ROBOT_RADIUS, C.byref(c))
        assert (res.point_collides, res.point_cert) == (bool(hit), c.value)
        idx, key = orc.find_within_range(r, p)
        orc.empty(idx)
        assert res.count == len(idx)
        go, oo = np.argsort(res.idx), np.argsort(idx)
        assert np.array_equal(res.idx[go], idx[oo])
        assert np.array_equal(res.dist[go].view(np.uint64), key[oo].view(np.uint64))
        if it % 25 == 0 or it < 50:        # edge flags: forward new->n and reverse n->new (not symmetric)
            pts = orc_positions(orc)
            for k in range(len(res.idx)):
                nb = pts[res.idx[k]]
                assert res.fwd[k] == L.orc_edge_check_all(sph, ns, 0, P(p), P(nb), W.ROBOT_RADIUS, 0)
                assert res.rev[k] == L.orc_edge_check_all(sph, ns, 0, P(nb), P(p), W.ROBOT_RADIUS, 0)
        if not res.point_collides:         # the planner only inserts collision-free samples (rrtqx.jl:940-950)
            assert t.insert(p) == orc.insert(p)
            accepted += 1
    assert accepted > 0.6 * n_iter and len(t) == accepted + 1
    # kd topology after thousands of single inserts + re-indexing is still the sequential one
    for a, b in zip(t.kd_fields(), orc.fields()):
        assert np.array_equal(a, b)


def orc_positions(orc):
    n = len(orc)
    ptr = oracle.lib().orc_kd_positions
    ptr.restype = C.POINTER(C.c_double)
    ptr.argtypes = [C.c_void_p]
    return np.ctypeslib.as_array(ptr(orc.h), shape=(n, orc.d))


def test_extend_query_capacity_and_empty_ball(ctx, building2):
    centers, radii, _ = building2
    S = SphereSet(ctx, centers, radii)
    pts, _, _ = W.c2_workload(5000, 1)
    t = DeviceTree(ctx, 3)
    t.insert_batch(pts)
    orc = oracle.KDTree(3)
    orc.insert_batch(pts)
    q = np.array([1.0, 2.0, 3.0])
    res = extend_query(t, S, q, 6.0, 0.5, 0, capacity=16)            # more neighbours than capacity
    idx, _ = orc.find_within_range(6.0, q)
    orc.empty(idx)
    assert res.count == len(idx) > 16 and len(res.idx) == 16 and set(res.idx.tolist()) <= set(idx.tolist())
    res = extend_query(t, S, np.array([500.0, 0.0, 0.0]), 1.0, 0.5, 0)   # empty ball: nearest by brute force
    oi, od = orc.find_nearest(np.array([500.0, 0.0, 0.0]))
    assert res.count == 0 and (res.nearest_idx, res.nearest_dist) == (oi, od)


def test_extend_query_ignores_inactive_entries_of_a_long_obstacle_list(ctx, building2):
    """The reference's obstacle list only grows (expired agent obstacles stay in it, rrtqx.jl:462-530): 5000 entries
    of which 31 are active must work (the kernel's table holds the ACTIVE ones), and the answer must follow an
    in-place update of the flags."""
    centers, radii, _ = building2
    rng = np.random.default_rng(3)
    n_dead = 5000
    allc = np.vstack([rng.uniform(-20, 20, (n_dead, 3)), centers])
    allr = np.concatenate([rng.uniform(1.0, 3.0, n_dead), radii])
    active = np.concatenate([np.zeros(n_dead, np.uint8), np.ones(len(radii), np.uint8)])
    S_long = SphereSet(ctx, allc, allr, active)
    S_ref = SphereSet(ctx, centers, radii)
    pts, _, _ = W.c2_workload(5000, 1)
    t = DeviceTree(ctx, 3)
    t.insert_batch(pts)
    for q in (np.array([1.0, 2.0, 3.0]), np.array([-5.0, 5.0, 0.5]), centers[4] + 0.1):
        a = extend_query(t, S_long, q, 4.0, 0.5, A.CHECK_QUICK_PASS)
        b = extend_query(t, S_ref, q, 4.0, 0.5, A.CHECK_QUICK_PASS)
        assert (a.point_collides, a.point_cert, a.count) == (b.point_collides, b.point_cert, b.count)
        oa, ob = np.argsort(a.idx), np.argsort(b.idx)
        assert np.array_equal(a.idx[oa], b.idx[ob])
        assert np.array_equal(a.fwd[oa], b.fwd[ob]) and np.array_equal(a.rev[oa], b.rev[ob])
    # switch every active obstacle off: nothing collides any more
    S_long.update(n_dead, active=np.zeros(len(radii), np.uint8))
    a = extend_query(t, S_long, centers[4] + 0.1, 4.0, 0.5, A.CHECK_QUICK_PASS)
    assert not a.point_collides and not a.fwd.any() and not a.rev.any()
    # more than 768 ACTIVE obstacles is refused loudly, not silently truncated
    S_big = SphereSet(ctx, allc[:1000], allr[:1000])
    with pytest.raises(A.RRTQXError):
        extend_query(t, S_big, np.array([1.0, 2.0, 3.0]), 4.0, 0.5, 0)
