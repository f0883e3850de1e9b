"""ctypes loader for the CPU oracle (oracle/rrtqx_oracle.c).

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; never from rrtqx_3d_b200/.
PARITY UNPINNED against a running reference (no Julia here) -- see the header
of rrtqx_oracle.h for what pins it instead.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "librrtqx_oracle.so")

c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)
c_f64p = C.POINTER(C.c_double)
c_u8p = C.POINTER(C.c_uint8)


def build(force: bool = False) -> str:
    """Compile the oracle with gcc if the .so is missing or stale."""
    src = os.path.join(_HERE, "rrtqx_oracle.c")
    hdr = os.path.join(_HERE, "rrtqx_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr))
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "librrtqx_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


class Sphere(C.Structure):
    """orc_sphere: SphereObstacle (DRRT_data_structures.jl:267-307), flattened."""
    _fields_ = [("pos", C.c_double * 3), ("radius", C.c_double),
                ("start_time", C.c_double), ("life_span", C.c_double),
                ("unused", C.c_uint8), ("pad", C.c_uint8 * 7)]


class Obstacle2D(C.Structure):
    _fields_ = [("kind", C.c_int32), ("n_vert", C.c_int32),
                ("pos", C.c_double * 2), ("radius", C.c_double),
                ("life_span", C.c_double), ("unused", C.c_uint8),
                ("pad", C.c_uint8 * 7), ("poly", c_f64p)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    build()
    L = C.CDLL(_SO)
    vp = C.c_void_p
    f64, i64, i32, cint = C.c_double, C.c_int64, C.c_int32, C.c_int
    sig = {
        "orc_euclid": (f64, [c_f64p, c_f64p, cint]),
        "orc_r3sdist": (f64, [c_f64p, c_f64p]),
        "orc_sqrt_threshold": (f64, [f64]),
        "orc_shrinking_ball": (f64, [f64, f64, i64, cint]),
        "orc_kd_new": (vp, [cint, cint, c_i32p, c_f64p]),
        "orc_kd_free": (None, [vp]),
        "orc_kd_insert": (i64, [vp, c_f64p]),
        "orc_kd_insert_batch": (None, [vp, c_f64p, i64]),
        "orc_kd_size": (i64, [vp]),
        "orc_kd_dim": (cint, [vp]),
        "orc_kd_fields": (None, [vp, c_i32p, c_i32p, c_i32p, c_i32p]),
        "orc_kd_dist_evals": (C.c_uint64, [vp]),
        "orc_kd_find_nearest": (i64, [vp, c_f64p, c_f64p]),
        "orc_kd_find_nearest_naive": (i64, [vp, c_f64p, c_f64p]),
        "orc_kd_find_within_range": (i64, [vp, f64, c_f64p, c_u8p, c_i32p, c_f64p, i64, i64]),
        "orc_kd_empty_range_list": (None, [c_u8p, c_i32p, i64]),
        "orc_kd_find_within_range_naive": (i64, [vp, f64, c_f64p, c_i32p, c_f64p, i64]),
        "orc_kd_range_batch": (i64, [vp, f64, c_f64p, i64, i64, c_i32p, c_i64p, c_i32p, c_f64p, i64, cint]),
        "orc_kd_range_batch_once": (i64, [vp, f64, c_f64p, i64, i64, c_i32p, c_i64p, cint, C.POINTER(vp)]),
        "orc_range_lists_copy": (None, [vp, c_i32p, c_f64p]),
        "orc_range_lists_free": (None, [vp]),
        "orc_kd_nearest_batch": (None, [vp, c_f64p, i64, i64, c_i32p, c_f64p, cint]),
        "orc_dist_point_to_segment": (f64, [c_f64p, c_f64p, c_f64p, cint, cint]),
        "orc_edge_check_sphere": (cint, [C.POINTER(Sphere), c_f64p, c_f64p, f64, cint]),
        "orc_edge_check_all": (cint, [C.POINTER(Sphere), i64, cint, c_f64p, c_f64p, f64, cint]),
        "orc_quick_check": (cint, [C.POINTER(Sphere), i64, c_f64p]),
        "orc_point_check": (cint, [C.POINTER(Sphere), i64, cint, c_f64p, f64, c_f64p]),
        "orc_point_check_3d": (cint, [C.POINTER(Sphere), i64, cint, c_f64p, f64, c_f64p]),
        "orc_edge_check_batch": (None, [C.POINTER(Sphere), i64, c_f64p, cint, c_i32p, c_i32p, i64, i64, f64, cint, c_u8p, cint]),
        "orc_obstacle_add_sweep": (cint, [vp, C.POINTER(Sphere), f64, f64, c_i64p, c_i32p, c_i32p, cint,
                                           c_i32p, c_i64p, i64, c_i32p, c_i64p, i64, c_i64p, c_i64p]),
        "orc_obstacle_remove_sweep": (cint, [vp, C.POINTER(Sphere), cint, C.POINTER(Sphere), i64, f64, f64,
                                              c_i64p, c_i32p, c_u8p, cint, c_i32p, c_i64p, i64, c_i32p, c_i64p, i64]),
        "orc_dist2_point_segment_2d": (f64, [c_f64p, c_f64p, c_f64p]),
        "orc_segment_dist2_2d": (f64, [c_f64p, c_f64p, c_f64p, c_f64p]),
        "orc_polygon_bound": (None, [c_f64p, i32, c_f64p, c_f64p, c_f64p]),
        "orc_edge_check_2d": (cint, [C.POINTER(Obstacle2D), c_f64p, c_f64p, f64]),
        "orc_edge_check_dubins": (cint, [C.POINTER(Obstacle2D), c_f64p, c_f64p, c_f64p, i32, f64, f64]),
        "orc_obstacle_add_sweep_2d": (cint, [vp, C.POINTER(Obstacle2D), cint, f64, f64, f64, c_i64p, c_i32p, c_i32p,
                                             c_i64p, c_f64p, c_i32p, c_i64p, i64, c_i32p, c_i64p, i64, c_i64p, c_i64p]),
        "orc_obstacle_remove_sweep_2d": (cint, [vp, C.POINTER(Obstacle2D), cint, C.POINTER(Obstacle2D), i64, f64, f64,
                                                f64, c_i64p, c_i32p, c_i64p, c_f64p, c_u8p, c_i32p, c_i64p, i64,
                                                c_i32p, c_i64p, i64]),
        "orc_dubins_trajectory": (cint, [c_f64p, c_f64p, f64, c_f64p, c_i32p, c_f64p, i32]),
        "orc_saturate_dubins": (None, [c_f64p, c_f64p, f64]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def _f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


def _p(a: np.ndarray, ty):
    return a.ctypes.data_as(ty)


def make_spheres(centers, radii, unused=None, life_span=None, start_time=None):
    """Build an orc_sphere array from SoA inputs (list order = array order)."""
    centers = _f64(centers).reshape(-1, 3)
    n = centers.shape[0]
    radii = np.broadcast_to(_f64(radii), (n,))
    arr = (Sphere * max(n, 1))()
    for i in range(n):
        arr[i].pos[0], arr[i].pos[1], arr[i].pos[2] = centers[i]
        arr[i].radius = radii[i]
        arr[i].start_time = 0.0 if start_time is None else float(start_time[i])
        arr[i].life_span = float("inf") if life_span is None else float(life_span[i])
        arr[i].unused = 0 if unused is None else int(bool(unused[i]))
    return arr, n


class KDTree:
    """Pointer-style insertion-order kd tree (kdTree_general.jl) on the CPU."""

    def __init__(self, d: int, wraps=(), wrap_points=()):
        self.L = lib()
        self.d = int(d)
        w = np.asarray(list(wraps), dtype=np.int32)
        wp = np.asarray(list(wrap_points), dtype=np.float64)
        self.h = self.L.orc_kd_new(self.d, len(w), _p(w, c_i32p), _p(wp, c_f64p))
        if not self.h:
            raise ValueError("bad kd tree parameters")
        self._marks = np.zeros(0, dtype=np.uint8)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.L.orc_kd_free(self.h)
                self.h = None
        except Exception:
            pass

    def __len__(self):
        return int(self.L.orc_kd_size(self.h))

    def insert(self, pos) -> int:
        p = _f64(pos).reshape(-1)
        assert p.size == self.d
        return int(self.L.orc_kd_insert(self.h, _p(p, c_f64p)))

    def insert_batch(self, pos):
        p = _f64(pos).reshape(-1, self.d)
        self.L.orc_kd_insert_batch(self.h, _p(p, c_f64p), p.shape[0])

    def fields(self):
        n = len(self)
        out = [np.empty(n, dtype=np.int32) for _ in range(4)]
        self.L.orc_kd_fields(self.h, *[_p(a, c_i32p) for a in out])
        return tuple(out)  # parent, childL, childR, split

    def dist_evals(self) -> int:
        return int(self.L.orc_kd_dist_evals(self.h))

    def _ensure_marks(self):
        n = len(self)
        if self._marks.size < n:
            m = np.zeros(max(n, 2 * self._marks.size), dtype=np.uint8)
            m[: self._marks.size] = self._marks
            self._marks = m
        return self._marks

    def find_nearest(self, q, naive=False):
        q = _f64(q).reshape(-1)
        d = C.c_double(0.0)
        fn = self.L.orc_kd_find_nearest_naive if naive else self.L.orc_kd_find_nearest
        i = fn(self.h, _p(q, c_f64p), C.byref(d))
        return int(i), d.value

    def find_within_range(self, r, q, prev=None):
        """Returns (idx, key) in push order.  prev=(idx,key) continues a list
        (kdFindMoreWithinRange); the caller must empty() the result."""
        q = _f64(q).reshape(-1)
        marks = self._ensure_marks()
        n = len(self)
        idx = np.empty(max(n, 1), dtype=np.int32)
        key = np.empty(max(n, 1), dtype=np.float64)
        ln = 0
        if prev is not None:
            ln = len(prev[0])
            idx[:ln] = prev[0]
            key[:ln] = prev[1]
        ln2 = self.L.orc_kd_find_within_range(self.h, float(r), _p(q, c_f64p), _p(marks, c_u8p),
                                              _p(idx, c_i32p), _p(key, c_f64p), idx.size, ln)
        assert ln2 >= 0
        return idx[:ln2].copy(), key[:ln2].copy()

    def empty(self, idx):
        idx = np.ascontiguousarray(idx, dtype=np.int32)
        self.L.orc_kd_empty_range_list(_p(self._marks, c_u8p), _p(idx, c_i32p), idx.size)

    def find_within_range_naive(self, r, q):
        q = _f64(q).reshape(-1)
        n = len(self)
        idx = np.empty(max(n, 1), dtype=np.int32)
        key = np.empty(max(n, 1), dtype=np.float64)
        ln = self.L.orc_kd_find_within_range_naive(self.h, float(r), _p(q, c_f64p), _p(idx, c_i32p),
                                                   _p(key, c_f64p), idx.size)
        return idx[:ln].copy(), key[:ln].copy()

    def range_batch(self, r, queries, want_lists=True, nthreads=1):
        """CSR results for all queries: (counts, offsets, idx, key).  ONE tree traversal per query, like the
        reference's kdFindWithinRange (hits go to per-thread growing buffers and are copied out afterwards)."""
        q = _f64(queries).reshape(-1, self.d)
        nq = q.shape[0]
        counts = np.zeros(nq, dtype=np.int32)
        if not want_lists:
            self.L.orc_kd_range_batch(self.h, float(r), _p(q, c_f64p), 0, nq, _p(counts, c_i32p),
                                      None, None, None, 0, nthreads)
            return counts, None, None, None
        offsets = np.zeros(nq + 1, dtype=np.int64)
        lists = C.c_void_p()
        total = self.L.orc_kd_range_batch_once(self.h, float(r), _p(q, c_f64p), 0, nq, _p(counts, c_i32p),
                                               _p(offsets, c_i64p), nthreads, C.byref(lists))
        idx = np.empty(max(total, 1), dtype=np.int32)
        key = np.empty(max(total, 1), dtype=np.float64)
        self.L.orc_range_lists_copy(lists, _p(idx, c_i32p), _p(key, c_f64p))
        self.L.orc_range_lists_free(lists)
        return counts, offsets, idx[:total], key[:total]

    def range_batch_two_pass(self, r, queries, nthreads=1):
        """Count pass + fill pass (two traversals per query); kept to cross-check range_batch."""
        q = _f64(queries).reshape(-1, self.d)
        nq = q.shape[0]
        counts = np.zeros(nq, dtype=np.int32)
        total = self.L.orc_kd_range_batch(self.h, float(r), _p(q, c_f64p), 0, nq, _p(counts, c_i32p),
                                          None, None, None, 0, nthreads)
        offsets = np.zeros(nq + 1, dtype=np.int64)
        idx = np.empty(max(total, 1), dtype=np.int32)
        key = np.empty(max(total, 1), dtype=np.float64)
        t2 = self.L.orc_kd_range_batch(self.h, float(r), _p(q, c_f64p), 0, nq, _p(counts, c_i32p),
                                       _p(offsets, c_i64p), _p(idx, c_i32p), _p(key, c_f64p), idx.size,
                                       nthreads)
        assert t2 == total
        return counts, offsets, idx[:total], key[:total]

    def nearest_batch(self, queries, nthreads=1):
        q = _f64(queries).reshape(-1, self.d)
        nq = q.shape[0]
        idx = np.empty(nq, dtype=np.int32)
        dist = np.empty(nq, dtype=np.float64)
        self.L.orc_kd_nearest_batch(self.h, _p(q, c_f64p), 0, nq, _p(idx, c_i32p), _p(dist, c_f64p), nthreads)
        return idx, dist


def dubins_trajectory(start4, goal4, r_min, cap=512):
    """calculateTrajectory(S, edge::DubinsEdge) restated (rrtqx_oracle.c: orc_dubins_trajectory):
    returns (dist, type, trajectory rows x 2)."""
    s = np.ascontiguousarray(start4, dtype=np.float64)
    g = np.ascontiguousarray(goal4, dtype=np.float64)
    dist = C.c_double(0.0)
    typ = C.c_int32(0)
    buf = np.empty((cap, 2), dtype=np.float64)
    n = lib().orc_dubins_trajectory(_p(s, c_f64p), _p(g, c_f64p), float(r_min), C.byref(dist),
                                    C.byref(typ), _p(buf, c_f64p), cap)
    assert n <= cap
    return dist.value, typ.value, buf[:n].copy()


def saturate_dubins(new_point4, closest4, delta):
    p = np.array(new_point4, dtype=np.float64, copy=True)
    c = np.ascontiguousarray(closest4, dtype=np.float64)
    lib().orc_saturate_dubins(_p(p, c_f64p), _p(c, c_f64p), float(delta))
    return p
