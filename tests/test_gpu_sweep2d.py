"""GPU parity: obstacle add / remove sweeps of the Otte generation with DubinsEdge (DRRT.jl:3048-3268) against the
oracle twins orc_obstacle_add_sweep_2d / orc_obstacle_remove_sweep_2d.  The obstacles are the reference's own
polygon fixtures (environments/rand_Static.txt, rand_Disc2.txt, rand_DiscForest_3.txt, copied unchanged into
tests/golden/ref_*.txt and parsed by the reader of DRRT_Q.jl:853-899) plus a few balls.  Booleans are bit-exact
GIVEN the trajectory rows: the oracle is fed the rows the device holds (solved on the device, or uploaded)."""
import ctypes as C
import math
import os

import numpy as np
import pytest

import oracle
from rrtqx_3d_b200 import workloads as W
from rrtqx_3d_b200.device import DeviceTree, EdgeSet, PolygonSet
from rrtqx_3d_b200.formats import read_polygon_obstacles

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
RHO, DELTA, R_TURN = 0.5, 5.0, 1.0


def _reference_polygons(name, limit=None):
    obs = read_polygon_obstacles(os.path.join(GOLDEN, name))
    return [("polygon", o["polygon"]) for o in obs[:limit]]


def _dubins_graph(ctx, n, r, seed=7):
    """C4-style nodes [x y 0 theta] in [-50,50]^2 x [0,2pi) with the theta wrap, out-edges to all neighbours within
    r (4-D euclid incl. ghosts, what kdFindWithinRange returns), parent = the lowest-index neighbour."""
    u = W.uniform01(seed, 0, 3 * n).reshape(n, 3)
    pts = np.zeros((n, 4))
    pts[:, 0], pts[:, 1], pts[:, 3] = -50 + 100 * u[:, 0], -50 + 100 * u[:, 1], 2 * math.pi * u[:, 2]
    t = DeviceTree(ctx, 4, wraps=(3,), wrap_points=(2 * math.pi,))
    t.insert_batch(pts)
    res, total = t.range_query(pts, r, want_dist=False)
    counts, offsets = res.layout()
    idx, _ = res.fetch(want_dist=False)
    src = np.repeat(np.arange(n, dtype=np.int32), counts)
    dst = np.concatenate([idx[offsets[q]:offsets[q] + counts[q]] for q in range(n)]).astype(np.int32)
    keep = src != dst
    src, dst = np.ascontiguousarray(src[keep]), np.ascontiguousarray(dst[keep])
    parent = np.full(n, -1, dtype=np.int32)
    first = np.r_[True, src[1:] != src[:-1]]
    for s_, d_ in zip(src[first], dst[first]):          # out-edges are grouped by start node
        parent[s_] = d_
    parent[0] = -1                                      # the root has no parent
    return pts, t, src, dst, parent


def _csr(src, n):
    row_ptr = np.zeros(n + 1, dtype=np.int64)
    np.add.at(row_ptr, src + 1, 1)
    return np.cumsum(row_ptr)


def _orc_obstacles(P):
    from test_gpu_collision import _orc_obstacles_2d
    return _orc_obstacles_2d(P)


def _oracle_add(orc, ob, pts, row_ptr, col, parent, tptr, txy):
    L = oracle.lib()
    n_e = len(col)
    be = np.zeros(n_e + 8, dtype=np.int32)
    on = np.zeros(len(pts) + 8, dtype=np.int32)
    nb, no, nc, nt = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
    rc = L.orc_obstacle_add_sweep_2d(orc.h, C.byref(ob), 1, RHO, DELTA, R_TURN, oracle._p(row_ptr, oracle.c_i64p),
                                     oracle._p(col, oracle.c_i32p), oracle._p(parent, oracle.c_i32p),
                                     oracle._p(tptr, oracle.c_i64p), oracle._p(txy, oracle.c_f64p),
                                     oracle._p(be, oracle.c_i32p), C.byref(nb), len(be), oracle._p(on, oracle.c_i32p),
                                     C.byref(no), len(on), C.byref(nc), C.byref(nt))
    assert rc == 0
    return set(be[:nb.value].tolist()), set(on[:no.value].tolist()), nc.value


@pytest.mark.parametrize("fixture,limit,extra_balls", [("ref_rand_Static.txt", None, True),
                                                        ("ref_rand_Disc2.txt", None, False),
                                                        ("ref_rand_DiscForest_3.txt", 60, False)])
def test_add_and_remove_sweep_2d_reference_polygons(ctx, fixture, limit, extra_balls):
    n = 20000
    pts, t, src, dst, parent = _dubins_graph(ctx, n, 2.2)
    assert 4 < len(src) / n < 40
    obstacles = _reference_polygons(fixture, limit)
    if extra_balls:
        obstacles += [("ball", (5.0, 5.0), 2.0), ("ball", (-20.0, 31.0), 4.0), ("ball", (pts[0, 0], pts[0, 1]), 1.0)]
    P = PolygonSet(ctx)
    P.upload(obstacles)
    orc_obs = _orc_obstacles(P)
    E = EdgeSet(t)
    E.upload(src, dst, parent)
    rows = E.solve_trajectories(R_TURN)                 # device solver for every item (out-edges + parent edges)
    tptr, txy = E.trajectories()
    assert len(tptr) == len(src) + n + 1 and tptr[-1] == rows == len(txy) and rows > 10 * len(src)
    orc = oracle.KDTree(4, (3,), (2 * math.pi,))
    orc.insert_batch(pts)
    row_ptr = _csr(src, n)                              # src is sorted: upload order == CSR order
    assert np.all(src[1:] >= src[:-1])
    # single obstacles: spread over the set, the first and the last included
    pick = sorted(set(np.linspace(0, len(obstacles) - 1, 9).astype(int).tolist()))
    union_b, union_o, n_hits = set(), set(), 0
    for o in pick:
        res = E.add_sweep_2d(P, [o], RHO, DELTA, R_TURN)
        ge, gn = res.fetch()
        wb, wo, nc = _oracle_add(orc, orc_obs[o], pts, row_ptr, dst, parent, tptr, txy)
        assert set(ge.tolist()) == wb and len(ge) == len(wb), (fixture, o)
        assert set(gn.tolist()) == wo and len(gn) == len(wo), (fixture, o)
        assert nc > 0
        union_b |= wb
        union_o |= wo
        n_hits += len(wb)
    assert n_hits > 50 and len(union_o) > 0
    # a batch of obstacles = the union of the single sweeps
    res = E.add_sweep_2d(P, pick, RHO, DELTA, R_TURN)
    ge, gn = res.fetch()
    assert set(ge.tolist()) == union_b and set(gn.tolist()) == union_o
    ef, nf = res.flags(len(src), n)
    assert ef.sum() == len(union_b) and nf.sum() == len(union_o)

    # removeObstacle (Otte semantics): flagged edges that hit the removed obstacle and none of the others come back
    L = oracle.lib()
    inf = np.zeros(len(src), dtype=np.uint8)
    inf[sorted(union_b)] = 1
    inf[::7] = 1                                        # plus flags the obstacle set does not explain
    removed = pick[len(pick) // 2]
    others = [o for o in pick if o != removed]
    rr = E.remove_sweep_2d(P, removed, others, inf, RHO, DELTA, R_TURN)
    re_, rn_ = rr.fetch()
    oth = (oracle.Obstacle2D * len(others))(*[orc_obs[o] for o in others])
    be = np.zeros(len(src) + 8, dtype=np.int32)
    on = np.zeros(n + 8, dtype=np.int32)
    nr, nq = C.c_int64(0), C.c_int64(0)
    rc = L.orc_obstacle_remove_sweep_2d(orc.h, C.byref(orc_obs[removed]), 1, oth, len(others), RHO, DELTA, R_TURN,
                                        oracle._p(row_ptr, oracle.c_i64p), oracle._p(dst, oracle.c_i32p),
                                        oracle._p(tptr, oracle.c_i64p), oracle._p(txy, oracle.c_f64p),
                                        oracle._p(inf, oracle.c_u8p), oracle._p(be, oracle.c_i32p), C.byref(nr), len(be),
                                        oracle._p(on, oracle.c_i32p), C.byref(nq), len(on))
    assert rc == 0
    assert set(re_.tolist()) == set(be[:nr.value].tolist()) and len(re_) == nr.value
    assert set(rn_.tolist()) == set(on[:nq.value].tolist()) and len(rn_) == nq.value
    # with no other obstacle every flagged candidate edge that hits the removed one is restored
    r0 = E.remove_sweep_2d(P, removed, [], inf, RHO, DELTA, R_TURN)
    assert r0.sizes()[0] >= nr.value


def test_uploaded_trajectories_and_stale_state(ctx):
    """Trajectories uploaded by the caller (the planner's own edge.trajectory rows, here the oracle's solver) give
    the oracle's sets; changing the edge set without renewing the trajectories is refused."""
    from rrtqx_3d_b200 import _abi as A
    n = 3000
    pts, t, src, dst, parent = _dubins_graph(ctx, n, 5.0, seed=11)
    P = PolygonSet(ctx)
    obstacles = _reference_polygons("ref_rand_Static.txt")
    P.upload(obstacles)
    orc_obs = _orc_obstacles(P)
    E = EdgeSet(t)
    E.upload(src, dst, parent)
    with pytest.raises(A.RRTQXError):                   # no trajectories resident yet
        E.add_sweep_2d(P, [0], RHO, DELTA, R_TURN)
    items = [(s, d) for s, d in zip(src, dst)] + [(v, parent[v] if parent[v] >= 0 else v) for v in range(n)]
    trajs = [oracle.dubins_trajectory(pts[a], pts[b], R_TURN)[2] for a, b in items]
    tptr = np.zeros(len(items) + 1, dtype=np.int64)
    tptr[1:] = np.cumsum([len(x) for x in trajs])
    txy = np.ascontiguousarray(np.vstack(trajs))
    E.set_trajectories(tptr, txy)
    orc = oracle.KDTree(4, (3,), (2 * math.pi,))
    orc.insert_batch(pts)
    row_ptr = _csr(src, n)
    total = 0
    for o in range(0, len(obstacles), 5):
        ge, gn = E.add_sweep_2d(P, [o], RHO, DELTA, R_TURN).fetch()
        wb, wo, _ = _oracle_add(orc, orc_obs[o], pts, row_ptr, dst, parent, tptr, txy)
        assert set(ge.tolist()) == wb and set(gn.tolist()) == wo
        total += len(wb)
    assert total > 20
    E.append(np.array([0], dtype=np.int32), np.array([1], dtype=np.int32))
    with pytest.raises(A.RRTQXError):                   # the edge set changed: trajectories are stale
        E.add_sweep_2d(P, [0], RHO, DELTA, R_TURN)


def test_candidate_rule_is_the_wrap_range_query(ctx):
    """The start-node filter of the sweeps equals kdFindWithinRange(KD, ((rho+delta)+R)+pi, [x y 0 pi]) on the
    theta-wrapped tree: an obstacle placed on every edge (a huge ball) blocks exactly the out-edges of the nodes
    the device range query returns."""
    n = 6000
    pts, t, src, dst, parent = _dubins_graph(ctx, n, 3.0, seed=13)
    P = PolygonSet(ctx)
    P.upload([("ball", (10.0, -12.0), 9.0), ("ball", (0.0, 0.0), 500.0)])
    E = EdgeSet(t)
    E.upload(src, dst, parent)
    E.solve_trajectories(R_TURN)
    for o, (c, rad) in enumerate((((10.0, -12.0), 9.0), ((0.0, 0.0), 500.0))):
        q = np.array([[c[0], c[1], 0.0, math.pi]])
        res, _ = t.range_query(q, ((RHO + DELTA) + rad) + math.pi, want_dist=False)
        cand = set(res.fetch(want_dist=False)[0].tolist())
        ge, gn = E.add_sweep_2d(P, [o], RHO, DELTA, R_TURN).fetch()
        starts = set(src[ge].tolist())
        assert starts <= cand and set(gn.tolist()) <= cand
        if rad > 100:                                   # everything within reach collides: the filter alone decides
            assert starts == {v for v in cand if (src == v).any()}
            assert set(gn.tolist()) == {v for v in cand if parent[v] >= 0}
