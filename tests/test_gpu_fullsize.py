"""Full-size (BASELINE.json C2 / C3) checks through size-independent properties + oracle samples."""
import ctypes as C

import numpy as np
import pytest

import oracle
from rrtqx_3d_b200 import workloads as W
from rrtqx_3d_b200.device import DeviceTree, EdgeSet, RangeResult, SphereSet, edge_check_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2(ctx):
    pts, qs, r = W.c2_workload(1_000_000, 1_000_000)
    t = DeviceTree(ctx, 3)
    t.insert_batch(pts)
    return pts, qs, r, t


def _per_query_checksums(res):
    counts, offsets = res.layout()
    idx, dist = res.fetch(want_dist=res_has_dist(res))
    order = np.argsort(offsets, kind="stable")
    # lists are tightly packed in offset order: segment sums via cumulative sums
    cs = np.concatenate([[0], np.cumsum(idx.astype(np.int64))])
    seg = np.zeros(len(counts), dtype=np.int64)
    seg[order] = cs[offsets[order] + counts[order]] - cs[offsets[order]]
    return counts, seg, idx, dist, offsets


def res_has_dist(res):
    return res.device_pointers()[3] is not None


def test_c2_full_size_properties(ctx, c2):
    pts, qs, r, t = c2
    res, total = t.range_query(qs, r, want_dist=True)
    counts, seg, idx, dist, offsets = _per_query_checksums(res)
    assert total == int(counts.sum()) and 4.3e8 < total < 4.5e8
    assert np.sort(offsets)[0] == 0 and np.array_equal(np.sort(offsets)[1:], np.cumsum(counts[np.argsort(offsets, kind="stable")])[:-1])
    assert dist.min() >= 0.0 and dist.max() < r                       # every key is a distance below the radius
    # idempotence + the idx-only path returns the same sets
    res2, total2 = t.range_query(qs, r, want_dist=False, result=RangeResult(ctx))
    counts2, seg2, _, _, _ = _per_query_checksums(res2)
    assert total2 == total and np.array_equal(counts, counts2) and np.array_equal(seg, seg2)
    # counts-only mode
    res3, total3 = t.range_query(qs, r, want_dist=False, count_only=True, result=RangeResult(ctx))
    assert total3 == total and np.array_equal(res3.layout()[0], counts)
    # oracle on a stride sample of the million queries (sets and keys bit-exact)
    orc = oracle.KDTree(3)
    orc.insert_batch(pts)
    sample = np.arange(0, len(qs), 20011)
    oc, ooff, oidx, okey = orc.range_batch(r, qs[sample], nthreads=16)
    for k, q in enumerate(sample):
        g = idx[offsets[q]:offsets[q] + counts[q]]
        o = oidx[ooff[k]:ooff[k + 1]]
        go, oo = np.argsort(g), np.argsort(o)
        assert np.array_equal(g[go], o[oo])
        assert np.array_equal(dist[offsets[q]:offsets[q] + counts[q]][go].view(np.uint64), okey[ooff[k]:ooff[k + 1]][oo].view(np.uint64))
    res2.close(); res3.close()
    # symmetry: sum_q |N(q)| over tree P  ==  sum_p |N(p)| over tree Q   (up to the two roots' <= rule)
    t2 = DeviceTree(ctx, 3)
    t2.insert_batch(qs)
    res4, total4 = t2.range_query(pts, r, want_dist=False, count_only=True, result=RangeResult(ctx))
    assert abs(total4 - total) <= 2
    res4.close(); res.close()


def test_c2_nearest_full_size(ctx, c2):
    pts, qs, r, t = c2
    gi, gd = t.nearest(qs)
    assert gi.min() >= 0 and gi.max() < len(pts)
    # the reported distance is the distance to the reported node, bit for bit
    d = qs - pts[gi]
    s = d[:, 0] * d[:, 0] + d[:, 1] * d[:, 1]
    s = s + d[:, 2] * d[:, 2]
    assert np.array_equal(np.sqrt(s).view(np.uint64), gd.view(np.uint64))
    orc = oracle.KDTree(3)
    orc.insert_batch(pts)
    sample = np.arange(0, len(qs), 997)
    oi, od = orc.nearest_batch(qs[sample], nthreads=16)
    assert np.array_equal(gi[sample], oi) and np.array_equal(gd[sample].view(np.uint64), od.view(np.uint64))


def test_c3_full_size_sweep_is_union_of_single_obstacle_sweeps(ctx, c2):
    pts, qs, r, t = c2
    # C3 edge set: all ordered pairs within 0.5346 + parents (same construction as bench.py)
    import bench
    src, dst, parent = bench.build_c3_edges(t, pts, 0.5346)
    assert 9.5e6 < len(src) < 1.05e7
    centers, radii = W.c3_obstacles(256)
    S = SphereSet(ctx, centers, radii)
    E = EdgeSet(t)
    E.upload(src, dst, parent)
    from rrtqx_3d_b200 import _abi as A
    fast = E.add_sweep(S, np.arange(256, dtype=np.int32), W.ROBOT_RADIUS, W.DELTA)          # edge-centric (default)
    full = E.add_sweep(S, np.arange(256, dtype=np.int32), W.ROBOT_RADIUS, W.DELTA, flags=A.SWEEP_STATS)
    fe, fn = full.fetch()
    qe, qn = fast.fetch()
    assert np.array_equal(qe, fe) and np.array_equal(qn, fn)                                # both kernels agree
    n_eh, n_nh, n_cand, n_tests = full.sizes()
    assert len(fe) == n_eh > 1e6 and len(fn) == n_nh
    # linearity: OR over obstacles == union of disjoint obstacle groups; statistics add up
    ue, un, cand, tests = set(), set(), 0, 0
    for lo in range(0, 256, 64):
        part = E.add_sweep(S, np.arange(lo, lo + 64, dtype=np.int32), W.ROBOT_RADIUS, W.DELTA, flags=A.SWEEP_STATS)
        pe, pn = part.fetch()
        ue.update(pe.tolist()); un.update(pn.tolist())
        cand += part.sizes()[2]; tests += part.sizes()[3]
    assert ue == set(fe.tolist()) and un == set(fn.tolist()) and (cand, tests) == (n_cand, n_tests)
    # every blocked edge collides with some obstacle (edge check over all obstacles); sample vs the oracle
    flags = edge_check_batch(t, S, src, dst, W.ROBOT_RADIUS)
    assert flags[fe].all()
    sph, ns = oracle.make_spheres(centers, radii)
    sample = np.arange(0, len(src), 9973)
    want = np.zeros(len(sample), dtype=np.uint8)
    oracle.lib().orc_edge_check_batch(sph, ns, oracle._p(np.ascontiguousarray(pts), oracle.c_f64p), 3,
                                      oracle._p(np.ascontiguousarray(src[sample]), oracle.c_i32p),
                                      oracle._p(np.ascontiguousarray(dst[sample]), oracle.c_i32p), 0, len(sample),
                                      W.ROBOT_RADIUS, 0, oracle._p(want, oracle.c_u8p), 8)
    assert np.array_equal(flags[sample], want)
    # the whole sweep against the oracle's addNewObstacle loop (start-node filter, out-edge and parent-edge tests,
    # orphan rule): every obstacle swept on the CPU over the SAME 10^6-node tree and 9.85 M-edge graph, a few
    # obstacles per host thread; blocked-edge set, orphan set and both statistics must equal the GPU's
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor
    orc = oracle.KDTree(3)
    orc.insert_batch(pts)
    order = np.argsort(src, kind="stable")
    row_ptr = np.zeros(len(pts) + 1, dtype=np.int64)
    np.add.at(row_ptr, src + 1, 1)
    row_ptr = np.cumsum(row_ptr)
    col = np.ascontiguousarray(dst[order])
    eid = order.astype(np.int32)
    par = np.ascontiguousarray(parent, dtype=np.int32)
    L = oracle.lib()
    cap = len(src) + 8

    def sweep_some(obs):
        be = np.zeros(cap, dtype=np.int32)
        on = np.zeros(len(pts) + 8, dtype=np.int32)
        eflag = np.zeros(len(src), dtype=np.uint8)
        nflag = np.zeros(len(pts), dtype=np.uint8)
        nc_sum = nt_sum = 0
        for o in obs:
            nb, no, nc, nt = C.c_int64(0), C.c_int64(0), C.c_int64(0), C.c_int64(0)
            rc = L.orc_obstacle_add_sweep(orc.h, C.byref(sph[o]), W.ROBOT_RADIUS, W.DELTA, oracle._p(row_ptr, oracle.c_i64p),
                                          oracle._p(col, oracle.c_i32p), oracle._p(par, oracle.c_i32p), 0,
                                          oracle._p(be, oracle.c_i32p), C.byref(nb), cap, oracle._p(on, oracle.c_i32p),
                                          C.byref(no), len(on), C.byref(nc), C.byref(nt))
            assert rc == 0
            eflag[eid[be[:nb.value]]] = 1
            nflag[on[:no.value]] = 1
            nc_sum += nc.value
            nt_sum += nt.value
        return eflag, nflag, nc_sum, nt_sum

    workers = 16
    with ThreadPoolExecutor(workers) as pool:
        parts = list(pool.map(sweep_some, [range(k, 256, workers) for k in range(workers)]))
    eflag = np.zeros(len(src), dtype=np.uint8)
    nflag = np.zeros(len(pts), dtype=np.uint8)
    for pe, pn, _, _ in parts:
        eflag |= pe
        nflag |= pn
    assert np.array_equal(np.flatnonzero(eflag), np.sort(fe))
    assert np.array_equal(np.flatnonzero(nflag), np.sort(fn))
    assert (sum(p[2] for p in parts), sum(p[3] for p in parts)) == (n_cand, n_tests)


def test_c4_full_size_wrap_queries_and_dubins(ctx):
    """BASELINE config C4 at full size: 200k poses in [-50,50]^2 x {0} x [0,2pi), theta wrapping.
    Range queries (ghost identities through the pair kernel): symmetry of the neighbour relation (q in N(p) <=>
    p in N(q), also across the seam) on the whole batch + oracle sample; Dubins solver on 200k edges: oracle
    sample at 1e-9 and the size-independent invariants (length >= chord, first row = start)."""
    import math
    from rrtqx_3d_b200.device import dubins_trajectory_batch
    two_pi = 2.0 * math.pi
    lo, hi = [-50.0, -50.0, 0.0, 0.0], [50.0, 50.0, 0.0, two_pi]
    n = 200_000
    pts = W.uniform_points(4, n, lo, hi)
    t = DeviceTree(ctx, 4, wraps=[3], wrap_points=[two_pi])
    t.insert_batch(pts)
    r = 1.06
    res, total = t.range_query(pts, r, want_dist=True)       # every node queries its own neighbourhood
    counts, offsets = res.layout()
    idx, dist = res.fetch()
    assert total == counts.sum() and counts.min() >= 1       # every node finds itself (distance 0)
    order = np.argsort(offsets, kind="stable")
    src = np.empty(total, dtype=np.int64)
    src[:] = np.repeat(order, counts[order])
    # symmetric relation: the multiset of (min, max) pairs has every off-diagonal pair exactly twice
    a, b = np.minimum(src, idx), np.maximum(src, idx)
    off = a != b
    key = a[off] * n + b[off]
    uniq, cnt = np.unique(key, return_counts=True)
    assert np.all(cnt == 2)
    assert np.count_nonzero(~off) == n
    # pairs across the seam exist (theta difference > pi but wrapped distance < r)
    dth = np.abs(pts[src[off], 3] - pts[idx[off], 3])
    assert np.count_nonzero(dth > math.pi) > 100
    # oracle sample
    orc = oracle.KDTree(4, wraps=[3], wrap_points=[two_pi])
    orc.insert_batch(pts)
    for q in range(0, n, 997):
        oi, ok = orc.find_within_range(r, pts[q])
        orc.empty(oi)
        g = idx[offsets[q]:offsets[q] + counts[q]]
        gd = dist[offsets[q]:offsets[q] + counts[q]]
        go, oo = np.argsort(g, kind="stable"), np.argsort(oi, kind="stable")
        assert np.array_equal(g[go], oi[oo])
        assert np.array_equal(gd[go].view(np.uint64), ok[oo].view(np.uint64))
    # Dubins edges between neighbours: 200k edges
    e_src = src[off][:200_000]
    e_dst = idx[off][:200_000].astype(np.int64)
    s4, g4 = np.ascontiguousarray(pts[e_src]), np.ascontiguousarray(pts[e_dst])
    dres = dubins_trajectory_batch(ctx, s4, g4, 1.0)
    ddist, dtyp, dptr, dxy = dres.fetch()
    chord = np.hypot(s4[:, 0] - g4[:, 0], s4[:, 1] - g4[:, 1])
    ok_ = np.isfinite(ddist)
    assert ok_.mean() > 0.99 and np.all(ddist[ok_] >= chord[ok_] - 1e-9)
    first = dxy[dptr[:-1][ok_]]
    assert np.allclose(first, s4[ok_, :2], atol=1e-9)
    assert np.all((dtyp >= -1) & (dtyp <= 5))
    for e in range(0, 200_000, 101):
        od, ot, otr = oracle.dubins_trajectory(s4[e], g4[e], 1.0)
        if not math.isfinite(od):
            assert od == ddist[e] or (math.isnan(od) and math.isnan(ddist[e]))
            continue
        assert abs(od - ddist[e]) <= 1e-9 * max(1.0, od)
        if ot == dtyp[e]:
            tr = dxy[dptr[e]:dptr[e + 1]]
            assert len(tr) == len(otr) and np.allclose(tr, otr, rtol=1e-9, atol=1e-9)
