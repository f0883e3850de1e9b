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
