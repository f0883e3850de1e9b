"""CPU: numpy model of the cover lists used by the two-stage collision kernels (csrc/collide_queue.cuh,
sphere_grid_kernel's cover set-up in csrc/collision.cuh).  The kernels are checked against the oracle on the GPU;
this file checks the two geometric claims the scheme rests on, independent of any device:

  1. completeness: every obstacle o with |c_o - mid| <= thr_o + half is on the list of the cover cell that holds
     `mid`, for every edge with half <= cap -- also for midpoints outside the grid (border cells are unbounded);
  2. the 16-bit coarsening of the threshold packed next to the obstacle number only ever enlarges it."""
import numpy as np

COV_DIM = 32


def _cover_params(centers, thr):
    lo, hi = centers.min(axis=0), centers.max(axis=0)
    thr_max = thr.max()
    pad = 1.25 * thr_max
    clo = lo - pad
    ext2 = (hi - lo) + 2.0 * pad
    cell = ext2 / COV_DIM
    cinv = COV_DIM / ext2
    cap = float(np.min(0.5 * cell))
    margin = 1e-9 * (np.abs(centers).max() + ext2.max() + thr_max)
    return clo, cell, cinv, cap, margin


def _cell(v, clo, cinv):
    return np.clip(np.floor((v - clo) * cinv), 0, COV_DIM - 1).astype(np.int64)   # sg_cell: floor, clamp


def _register(centers, thr, clo, cell, cinv, cap, margin):
    """lists[cell] = obstacles with dist(c_o, box(cell)) <= thr_o + cap (+ margins); border boxes unbounded."""
    lists = {}
    idx = np.arange(COV_DIM)
    for o, (c, t) in enumerate(zip(centers, thr)):
        rho = (t + cap) * (1.0 + 1e-9) + margin
        a, b = _cell(c - rho, clo, cinv), _cell(c + rho, clo, cinv)
        gaps = []
        for d in range(3):
            L = np.where(idx == 0, -np.inf, clo[d] + idx * cell[d])
            H = np.where(idx == COV_DIM - 1, np.inf, clo[d] + (idx + 1) * cell[d])
            gaps.append(np.maximum(np.maximum(L - c[d], c[d] - H), 0.0))
        for iz in range(a[2], b[2] + 1):
            for iy in range(a[1], b[1] + 1):
                for ix in range(a[0], b[0] + 1):
                    if gaps[0][ix] ** 2 + gaps[1][iy] ** 2 + gaps[2][iz] ** 2 <= rho * rho:
                        lists.setdefault((iz * COV_DIM + iy) * COV_DIM + ix, []).append(o)
    return lists


def test_cover_lists_contain_every_reachable_obstacle():
    rng = np.random.default_rng(3)
    for trial, (n_obs, rmin, rmax) in enumerate([(60, 0.5, 3.0), (200, 0.05, 0.4), (5, 2.0, 9.0), (1, 1.0, 1.0)]):
        centers = rng.uniform(-20, 20, (n_obs, 3))
        thr = rng.uniform(rmin, rmax, n_obs) + 0.5
        clo, cell, cinv, cap, margin = _cover_params(centers, thr)
        lists = _register(centers, thr, clo, cell, cinv, cap, margin)
        # midpoints: inside the grid, around the obstacles, on cell faces, and far outside the grid
        near = centers[rng.integers(n_obs, size=600)] + rng.normal(size=(600, 3)) * thr.mean()
        mids = np.vstack([rng.uniform(-25, 25, (3000, 3)), near, rng.uniform(-200, 200, (300, 3)),
                          clo + cell * rng.integers(0, COV_DIM + 1, (300, 3))])
        half = rng.uniform(0.0, cap, len(mids))
        half[::7] = cap
        cells = _cell(mids, clo, cinv)
        flat = (cells[:, 2] * COV_DIM + cells[:, 1]) * COV_DIM + cells[:, 0]
        d = np.linalg.norm(mids[:, None, :] - centers[None, :, :], axis=2)
        reach = d <= thr[None, :] + half[:, None]
        for m in np.nonzero(reach.any(axis=1))[0]:
            have = set(lists.get(int(flat[m]), []))
            need = set(np.nonzero(reach[m])[0].tolist())
            assert need <= have, (trial, m, need - have)
        assert reach.any()


def test_packed_threshold_is_never_smaller():
    rng = np.random.default_rng(4)
    t = np.concatenate([rng.uniform(0, 50, 5000), -rng.uniform(0, 5, 500), [0.0, 1.0, 2.5, 3.4028235e38, 1e-45]]).astype(np.float32)
    b = t.view(np.uint32).astype(np.uint64)
    up = np.where(b >> 31 != 0, b & 0xFFFF0000, (b + 0xFFFF) & 0xFFFF0000).astype(np.uint32)
    packed = (up | np.uint32(0x1234)).view(np.uint32)          # obstacle number in the low 16 bits
    back = (packed & np.uint32(0xFFFF0000)).view(np.float32)
    assert np.all(back >= t)                                   # +inf for values next to FLT_MAX: never rejects
    assert np.all((packed & np.uint32(0xFFFF)) == 0x1234)
    finite = np.isfinite(back) & (t > 1e-30)
    assert np.all(back[finite] <= t[finite] * (1 + 2.0 ** -7))  # and at most one 16-bit step larger
