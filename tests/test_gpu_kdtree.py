"""GPU parity: kd build / range / nearest through the C ABI vs the CPU oracle.

Bar: neighbour index SETS bit-exact, distances bit-exact (they are the same
IEEE operations), kd topology identical to sequential kdInsert.
"""
import numpy as np
import pytest

import oracle
from rrtqx_3d_b200 import workloads as W
from rrtqx_3d_b200.device import DeviceTree

pytestmark = pytest.mark.gpu

TWO_PI = 2.0 * np.pi


def _points(seed, n, d):
    lo = [-20.0] * 3 + [0.0]
    hi = [20.0] * 3 + [TWO_PI]
    return W.uniform_points(seed, n, lo[:d] if d < 4 else lo, hi[:d] if d < 4 else hi)


def _compare_range(gpu_lists, orc, r, queries, per_query_r=None):
    for qi, (gi, gd) in enumerate(gpu_lists):
        rr = r if per_query_r is None else per_query_r[qi]
        oi, ok = orc.find_within_range(rr, queries[qi])
        orc.empty(oi)
        go, oo = np.argsort(gi, kind="stable"), np.argsort(oi, kind="stable")
        assert np.array_equal(gi[go], oi[oo]), f"query {qi}: index sets differ ({len(gi)} vs {len(oi)})"
        assert len(set(gi.tolist())) == len(gi), f"query {qi}: duplicates"
        if gd is not None:
            assert np.array_equal(gd[go].view(np.uint64), ok[oo].view(np.uint64)), f"query {qi}: keys differ"


@pytest.mark.parametrize("d", [2, 3, 4])
def test_kd_topology_equals_sequential_insert(ctx, d):
    pts = _points(21, 30000, d)
    orc = oracle.KDTree(d)
    orc.insert_batch(pts)
    t = DeviceTree(ctx, d)
    # batch, then single inserts, then another batch: same tree as one-by-one insertion
    assert t.insert_batch(pts[:20000]) == 0
    for i in range(20000, 20010):
        assert t.insert(pts[i]) == i
    assert t.insert_batch(pts[20010:]) == 20010
    assert len(t) == len(pts)
    for a, b, name in zip(t.kd_fields(), orc.fields(), ["parent", "childL", "childR", "split"]):
        assert np.array_equal(a, b), name
    assert np.array_equal(t.positions().view(np.uint64), pts.view(np.uint64))


def test_kd_topology_sorted_input_is_a_chain(ctx):
    # adversarial insertion order: depth == n (many build rounds)
    n = 300
    pts = np.stack([np.arange(n, dtype=np.float64)] * 3, axis=1)
    orc = oracle.KDTree(3)
    orc.insert_batch(pts)
    t = DeviceTree(ctx, 3)
    t.insert_batch(pts)
    for a, b in zip(t.kd_fields(), orc.fields()):
        assert np.array_equal(a, b)


@pytest.mark.parametrize("d,n,nq", [(3, 20000, 300), (2, 5000, 200)])
def test_range_sets_and_keys_bitexact(ctx, d, n, nq):
    pts, qs = _points(1, n, d), _points(2, nq, d)
    r = W.shrinking_ball_radius(n, d, W.DELTA, W.BALL_CONSTANT)
    orc = oracle.KDTree(d)
    orc.insert_batch(pts)
    t = DeviceTree(ctx, d)
    t.insert_batch(pts)
    res, total = t.range_query(qs, r)
    lists = res.lists()
    assert total == sum(len(i) for i, _ in lists)
    _compare_range(lists, orc, r, qs)
    # counts-only and no-dist variants agree
    res2, total2 = t.range_query(qs, r, want_dist=False, count_only=True)
    assert total2 == total
    assert np.array_equal(res2.layout()[0], res.layout()[0])


def test_range_large_batch_sorted_queries(ctx):
    # > 2048 queries takes the cell-sorted query path
    pts, qs, _ = W.c2_workload(50000, 4000)
    r = 1.7
    orc = oracle.KDTree(3)
    orc.insert_batch(pts)
    t = DeviceTree(ctx, 3)
    t.insert_batch(pts)
    res, total = t.range_query(qs, r)
    counts, offsets = res.layout()
    oc, ooff, oidx, okey = orc.range_batch(r, qs, nthreads=8)
    assert np.array_equal(counts, oc)
    assert total == int(oc.sum())
    idx, dist = res.fetch()
    for q in range(0, len(qs), 37):
        g = idx[offsets[q]:offsets[q] + counts[q]]
        o = oidx[ooff[q]:ooff[q + 1]]
        assert np.array_equal(np.sort(g), np.sort(o))
        gd = dist[offsets[q]:offsets[q] + counts[q]][np.argsort(g)]
        od = okey[ooff[q]:ooff[q + 1]][np.argsort(o)]
        assert np.array_equal(gd.view(np.uint64), od.view(np.uint64))


def test_range_wrap_dimension_ghost_identities(ctx):
    # Dubins-style space [x y t theta]: theta wraps with period 2*pi (DRRT.jl:3312)
    lo, hi = [-5.0, -5.0, 0.0, 0.0], [5.0, 5.0, 0.0, TWO_PI]
    pts = W.uniform_points(5, 6000, lo, hi)
    qs = W.uniform_points(6, 300, lo, hi)
    qs[:10, 3] = np.linspace(0.0, 0.05, 10)           # near the seam
    qs[10:20, 3] = TWO_PI - np.linspace(0.0, 0.05, 10)
    orc = oracle.KDTree(4, wraps=[3], wrap_points=[TWO_PI])
    orc.insert_batch(pts)
    t = DeviceTree(ctx, 4, wraps=[3], wrap_points=[TWO_PI])
    t.insert_batch(pts)
    for r in (0.9, 2.5):
        res, _ = t.range_query(qs, r)
        _compare_range(res.lists(), orc, r, qs)


def test_range_root_is_admitted_with_le(ctx):
    # kdTree_general.jl:896-898: root uses <=, every other node strict <
    pts = np.array([[0.0, 0.0, 0.0], [6.0, 0.0, 0.0], [3.0, 3.0, 0.0], [3.0, -1.0, 0.0]])
    q = np.array([[3.0, 0.0, 0.0]])
    orc = oracle.KDTree(3)
    orc.insert_batch(pts)
    t = DeviceTree(ctx, 3)
    t.insert_batch(pts)
    res, total = t.range_query(q, 3.0)
    (gi, gd), = res.lists()
    oi, ok = orc.find_within_range(3.0, q[0])
    orc.empty(oi)
    assert sorted(gi.tolist()) == sorted(oi.tolist()) == [0, 3]   # root at exactly r kept, node 1/2 at exactly r dropped


def test_range_degenerate_radii_and_per_query_radii(ctx):
    pts, qs, _ = W.c2_workload(3000, 64)
    orc = oracle.KDTree(3)
    orc.insert_batch(pts)
    t = DeviceTree(ctx, 3)
    t.insert_batch(pts)
    radii = np.array([0.0, -1.0, np.nan, np.inf, 1e-300, 5.0, 100.0, 3.3] * 8)
    qs[0] = pts[0]   # query on the root with r = 0: root admitted (dist 0 <= 0)
    res, total = t.range_query(qs, 0.0, ranges=radii)
    _compare_range(res.lists(), orc, None, qs, per_query_r=radii)


def test_range_sees_unsorted_tail_after_inserts(ctx):
    pts, qs, _ = W.c2_workload(8000, 100)
    r = 5.0
    orc = oracle.KDTree(3)
    t = DeviceTree(ctx, 3)
    orc.insert_batch(pts[:6000])
    t.insert_batch(pts[:6000])
    res, _ = t.range_query(qs, r)             # builds the grid index over 6000 nodes
    _compare_range(res.lists(), orc, r, qs)
    for i in range(6000, 6100):               # planner-style single inserts -> tail
        orc.insert(pts[i])
        t.insert(pts[i])
    res, _ = t.range_query(qs, r, result=res)
    _compare_range(res.lists(), orc, r, qs)
    orc.insert_batch(pts[6100:])
    t.insert_batch(pts[6100:])
    res, _ = t.range_query(qs, r, result=res)  # tail > limit -> re-index
    _compare_range(res.lists(), orc, r, qs)
    gi, gd = t.nearest(qs)
    oi, od = orc.nearest_batch(qs)
    assert np.array_equal(gd.view(np.uint64), od.view(np.uint64))
    assert np.array_equal(gi, oi)


@pytest.mark.parametrize("d,wrap,nq", [(3, False, 3000), (2, False, 3000), (4, True, 3000),
                                       (2, False, 6000), (4, False, 6000), (4, True, 6000)])   # >= 4096: thread per query
def test_nearest_bitexact(ctx, d, wrap, nq):
    pts, qs = _points(3, 20000, d), _points(4, nq, d)
    kw = dict(wraps=[3], wrap_points=[TWO_PI]) if wrap else {}
    orc = oracle.KDTree(d, **kw)
    orc.insert_batch(pts)
    t = DeviceTree(ctx, d, **kw)
    t.insert_batch(pts)
    # include queries far outside the bounding box of the tree
    qs[:5] *= 10.0
    gi, gd = t.nearest(qs)
    oi, od = orc.nearest_batch(qs, nthreads=8)
    assert np.array_equal(gd.view(np.uint64), od.view(np.uint64))
    assert np.array_equal(gi, oi)


def test_empty_tree_is_an_error(ctx):
    from rrtqx_3d_b200._abi import RRTQXError, ERR_EMPTY_TREE
    t = DeviceTree(ctx, 3)
    with pytest.raises(RRTQXError) as e:
        t.range_query(np.zeros((1, 3)), 1.0)
    assert e.value.status == ERR_EMPTY_TREE
    with pytest.raises(RRTQXError):
        t.nearest(np.zeros((1, 3)))


def test_preorder_is_the_reference_traversal_order(ctx):
    """rrtqx_tree_preorder: node, kdChildL subtree, kdChildR subtree (saveRRTSubNodes, DRRT_Q.jl:316-327),
    checked against a recursive walk over the ORACLE's kd links."""
    import sys
    pts = _points(9, 30000, 3)
    orc = oracle.KDTree(3)
    orc.insert_batch(pts)
    t = DeviceTree(ctx, 3)
    t.insert_batch(pts[:20000])
    for p in pts[20000:20050]:
        t.insert(p)                       # single inserts take the one-thread walk
    t.insert_batch(pts[20050:])
    parent, left, right, split = orc.fields()
    want = []
    sys.setrecursionlimit(10000)

    def walk(v):
        want.append(v)
        if left[v] >= 0:
            walk(left[v])
        if right[v] >= 0:
            walk(right[v])
    walk(0)
    got = t.preorder()
    assert np.array_equal(got, np.asarray(want, dtype=np.int32))
    assert sorted(got.tolist()) == list(range(len(pts)))


def test_nearest_large_batch_thread_per_query_and_fallback(ctx):
    """Batches >= 4096 take the thread-per-query kernel; queries far outside the grid or in empty regions are
    finished by the warp-per-query kernel; a pending unsorted tail is indexed first.  Bit-exact vs the oracle."""
    rng = np.random.default_rng(13)
    clustered = np.concatenate([rng.normal(size=(20000, 3)) * 0.5 + c for c in ([-15, -15, -15], [10, 12, -3], [0, 0, 18])])
    pts = np.ascontiguousarray(np.concatenate([clustered, _points(5, 5000, 3)]))
    orc = oracle.KDTree(3)
    orc.insert_batch(pts)
    t = DeviceTree(ctx, 3)
    t.insert_batch(pts[:60000])
    t.range_query(pts[:10], 1.0)                      # builds the index over the first 60000 nodes
    for p in pts[60000:60200]:
        t.insert(p)                                   # unsorted tail
    t.insert_batch(pts[60200:])
    qs = _points(6, 6000, 3)                          # mostly empty space between the clusters
    qs[:200] = qs[:200] * 50.0                        # far outside the bounding box
    qs[200:400] = pts[:200]                           # exactly on nodes (distance 0)
    qs[400:600] = clustered[:200] + 1e-9
    gi, gd = t.nearest(qs)
    oi, od = orc.nearest_batch(qs, nthreads=8)
    assert np.array_equal(gd.view(np.uint64), od.view(np.uint64))
    assert np.array_equal(gi, oi)
