"""GPU parity of the paths specific to the v5 range kernel (csrc/range_v5.cuh): the FP32 filter's error
band, the exact fall-back routine (filter unusable / octet table or hit buffer overflow), odd batches,
non-finite queries.  Bar as everywhere: index sets and keys bit-exact against the CPU oracle."""
import os

import numpy as np
import pytest

import oracle
from rrtqx_3d_b200 import workloads as W
from rrtqx_3d_b200.device import DeviceTree

pytestmark = pytest.mark.gpu


def _check(t, orc, qs, r, ranges=None, stride=1):
    res, total = t.range_query(qs, r if ranges is None else 0.0, ranges=ranges)
    lists = res.lists()
    assert total == sum(len(i) for i, _ in lists)
    for qi in range(0, len(qs), stride):
        gi, gd = lists[qi]
        rr = r if ranges is None else ranges[qi]
        oi, ok = orc.find_within_range(rr, qs[qi])
        orc.empty(oi)
        go, oo = np.argsort(gi, kind="stable"), np.argsort(oi, kind="stable")
        assert np.array_equal(gi[go], oi[oo]), f"query {qi}: sets differ ({len(gi)} vs {len(oi)})"
        assert np.array_equal(gd[go].view(np.uint64), ok[oo].view(np.uint64)), f"query {qi}: keys differ"
    return total


def _tree(ctx, pts, d=3):
    orc = oracle.KDTree(d)
    orc.insert_batch(pts)
    t = DeviceTree(ctx, d)
    t.insert_batch(pts)
    return t, orc


def test_points_inside_the_filter_error_band(ctx):
    """Shells of points at distance r(1 + eps), eps from 1e-16 to 1e-5 on both sides of r: every trip that
    meets them lands in the band |s'| <= M and must be re-decided in exact FP64."""
    rng = np.random.default_rng(7)
    r = 2.0
    base, qs, _ = W.c2_workload(20000, 3000)   # > 2048 queries: sorted-query path
    centers = qs[:40]
    shell = []
    for c in centers:
        u = rng.normal(size=(60, 3))
        u /= np.linalg.norm(u, axis=1)[:, None]
        eps = np.concatenate([[0.0], 10.0 ** rng.uniform(-16, -5, 59)]) * rng.choice([-1.0, 1.0], 60)
        shell.append(c + u * (r * (1.0 + eps))[:, None])
    pts = np.ascontiguousarray(np.vstack([base] + shell))
    t, orc = _tree(ctx, pts)
    _check(t, orc, qs[:400], r)


@pytest.mark.parametrize("offset,extent,r", [(1.0e6, 40.0, 2.0), (-3.0e9, 40.0, 3.0), (0.0, 1.0e8, 2.5e6), (5.0e5, 1.0e7, 3.0)])
def test_large_coordinates_and_extents(ctx, offset, extent, r):
    """FP32 records are relative to the grid origin: a far-away box keeps the filter usable, a huge extent
    with a small radius makes it unusable (exact routine); results are exact either way."""
    rng = np.random.default_rng(11)
    pts = np.ascontiguousarray(offset + rng.random((30000, 3)) * extent)
    qs = np.ascontiguousarray(offset + rng.random((2500, 3)) * extent)
    if extent > 1e6 and r < 100:   # sparse: put queries next to points so the sets are not all empty
        qs[:500] = pts[:500] + rng.normal(size=(500, 3)) * r * 0.4
    t, orc = _tree(ctx, pts)
    total = _check(t, orc, qs, r, stride=7)
    assert total > 0


def test_forced_small_variant_overflows_table_and_buffers(ctx):
    """K far above the capacity of the variant: octet table / hit buffers overflow and the pair is redone
    by the exact routine."""
    pts, qs, _ = W.c2_workload(60000, 2100)
    t, orc = _tree(ctx, pts)
    os.environ["RRTQX_FUSED_VARIANT"] = "0"
    ctx.reload_tuning()
    try:
        for r in (6.5, 9.0):       # ~1100 and ~2900 neighbours per query
            _check(t, orc, qs, r, stride=53)
    finally:
        del os.environ["RRTQX_FUSED_VARIANT"]
        ctx.reload_tuning()
    _check(t, orc, qs, 6.5, stride=53)   # and through the variant the library picks itself


def test_odd_batches_and_single_query(ctx):
    pts, qs, _ = W.c2_workload(5000, 2049)
    t, orc = _tree(ctx, pts)
    for nq in (1, 2, 3, 2049):
        _check(t, orc, np.ascontiguousarray(qs[:nq]), 3.0, stride=17 if nq > 100 else 1)


def test_non_finite_queries_never_hit(ctx):
    pts, qs, _ = W.c2_workload(5000, 2200)
    t, orc = _tree(ctx, pts)
    qs[3, 1] = np.nan
    qs[10, 0] = np.inf
    qs[11, 2] = -np.inf
    qs[2100] = np.nan
    res, _ = t.range_query(qs, 3.0)
    counts, _ = res.layout()
    for q in (3, 10, 11, 2100):
        assert counts[q] == 0
    _check(t, orc, qs[:64], 3.0)


@pytest.mark.parametrize("d", [2, 4])
def test_other_dimensions_large_batch(ctx, d):
    lo = [-20.0, -20.0, -20.0, 0.0][:d] if d < 4 else [-20.0, -20.0, -20.0, 0.0]
    hi = [20.0, 20.0, 20.0, 2 * np.pi][:d] if d < 4 else [20.0, 20.0, 20.0, 2 * np.pi]
    pts = W.uniform_points(21, 40000, lo, hi)
    qs = W.uniform_points(22, 2600, lo, hi)
    t, orc = _tree(ctx, pts, d)
    _check(t, orc, qs, 2.2 if d == 4 else 1.1, stride=11)


def test_wrap_tree_through_ghost_pairs_large_batch(ctx):
    """Dubins-style 4-D tree with theta wrapping (period 2 pi): for r <= pi the batch runs through the pair kernel as
    (real, ghost) virtual queries; for r > pi through the two-pass kernel with explicit dedup.  Same sets and keys
    as the reference's ghost iteration (ghostPoint.jl:60-111, kdTree_general.jl:903-916)."""
    two_pi = 2.0 * np.pi
    lo, hi = [-20.0, -20.0, 0.0, 0.0], [20.0, 20.0, 0.0, two_pi]
    pts = W.uniform_points(31, 40000, lo, hi)
    qs = W.uniform_points(32, 3000, lo, hi)
    qs[:100, 3] = np.linspace(0.0, 0.2, 100)              # near the seam, both sides
    qs[100:200, 3] = two_pi - np.linspace(0.0, 0.2, 100)
    qs[200] = pts[0]                                       # on the root
    qs[201, 3] = np.pi                                     # exactly at the ghost switch (q < P/2 is false)
    orc = oracle.KDTree(4, wraps=[3], wrap_points=[two_pi])
    orc.insert_batch(pts)
    t = DeviceTree(ctx, 4, wraps=[3], wrap_points=[two_pi])
    t.insert_batch(pts)
    for r in (0.8, 2.0, 3.1, 3.3):                         # 3.3 > pi: two-pass path
        _check(t, orc, qs, r, stride=3)
    # odd batch, tiny batch, and a tail of single inserts
    _check(t, orc, np.ascontiguousarray(qs[:2049]), 1.5, stride=11)
    _check(t, orc, np.ascontiguousarray(qs[:3]), 1.5)
    extra = W.uniform_points(33, 300, lo, hi)
    for p in extra:
        orc.insert(p)
        t.insert(p)
    _check(t, orc, qs, 2.0, stride=5)
