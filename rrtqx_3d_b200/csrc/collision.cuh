// collision.cuh -- sphere-obstacle table and the exact segment/point-vs-sphere
// predicates of the QX-3D generation (DRRT_Q.jl:1205-1210, 1402-1590,
// 1775-1826), shared by the batched checks and the obstacle sweeps.
#pragma once
#include "common.cuh"

struct rrtqx_spheres {
  rrtqx_ctx *ctx = nullptr;
  int64_t n = 0;
  // SoA-in-AoS: (cx, cy, cz, radius) as one 32-byte record + active byte
  rrtqx::DevBuf<double4> rec;
  rrtqx::DevBuf<uint8_t> active;
};

namespace rrtqx {

#ifdef __CUDACC__

// Per-edge invariants of distancePointToSegment (DRRT_Q.jl:1205-1210):
// edgeLen = dist(startPoint,endPoint) and b = endPoint - startPoint.
struct SegPre {
  double sx, sy, sz;
  double bx, by, bz;
  double len;
  // conservative cull helpers (midpoint and half length, any rounding is fine:
  // the cull margin dwarfs it)
  double mx, my, mz, half;
  bool cullable;  // false when the reference result does not follow geometry
                  // (len == 0 -> t = NaN -> "collides with every active
                  // obstacle", or non-finite inputs)
};

__device__ __forceinline__ SegPre seg_prepare(double sx, double sy, double sz, double ex, double ey, double ez) {
  SegPre p;
  p.sx = sx; p.sy = sy; p.sz = sz;
  // dist(startPoint,endPoint): (s-e).^2 summed left to right
  double dx = __dsub_rn(sx, ex), dy = __dsub_rn(sy, ey), dz = __dsub_rn(sz, ez);
  double s = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  p.len = __dsqrt_rn(s);
  p.bx = __dsub_rn(ex, sx); p.by = __dsub_rn(ey, sy); p.bz = __dsub_rn(ez, sz);
  p.mx = 0.5 * (sx + ex); p.my = 0.5 * (sy + ey); p.mz = 0.5 * (sz + ez);
  p.half = 0.5 * p.len;
  p.cullable = (p.len > 0.0) && isfinite(p.len) && isfinite(p.mx) && isfinite(p.my) && isfinite(p.mz);
  return p;
}

// Radicand of dist(point, closePt) for obstacle centre (px,py,pz); the edge
// collides with the sphere iff !(sqrt(result) > robotRadius + radius)
// (explicitEdgeCheck3D DRRT_Q.jl:1789-1794).
template <bool FMA_DOT>
__device__ __forceinline__ double seg_point_radicand(const SegPre &e, double px, double py, double pz) {
  double ax = __dsub_rn(px, e.sx), ay = __dsub_rn(py, e.sy), az = __dsub_rn(pz, e.sz);
  double dot;
  if (FMA_DOT) {
    dot = __fma_rn(e.bx, ax, 0.0);
    dot = __fma_rn(e.by, ay, dot);
    dot = __fma_rn(e.bz, az, dot);
  } else {  // OpenBLAS ddot scalar tail: dot = 0.0; dot += y[i]*x[i]
    dot = __dadd_rn(0.0, __dmul_rn(e.bx, ax));
    dot = __dadd_rn(dot, __dmul_rn(e.by, ay));
    dot = __dadd_rn(dot, __dmul_rn(e.bz, az));
  }
  double t = jl_max(0.0, jl_min(1.0, __ddiv_rn(dot, e.len)));
  double cx = __dadd_rn(e.sx, __dmul_rn(t, e.bx));
  double cy = __dadd_rn(e.sy, __dmul_rn(t, e.by));
  double cz = __dadd_rn(e.sz, __dmul_rn(t, e.bz));
  double ux = __dsub_rn(px, cx), uy = __dsub_rn(py, cy), uz = __dsub_rn(pz, cz);
  return __dadd_rn(__dadd_rn(__dmul_rn(ux, ux), __dmul_rn(uy, uy)), __dmul_rn(uz, uz));
}

// Exact collision decision for one (segment, sphere) pair.
//   thr    = robotRadius + radius          (rounded once, that order)
//   thr_le = sqrt_thresh_le(thr)           (sqrt(s) > thr  <=>  s > thr_le)
template <bool FMA_DOT>
__device__ __forceinline__ bool seg_sphere_collide(const SegPre &e, double px, double py, double pz, double thr,
                                                   double thr_le) {
  // Conservative reject: closePt always lies on the segment, so the
  // reference's D is >= the true point-segment distance >= |p - mid| - len/2.
  if (e.cullable) {
    double ux = px - e.mx, uy = py - e.my, uz = pz - e.mz;
    double lim = (thr + e.half) * (1.0 + 1e-9) + 1e-300;
    if (ux * ux + uy * uy + uz * uz > lim * lim && isfinite(lim)) return false;
  }
  double s = seg_point_radicand<FMA_DOT>(e, px, py, pz);
  (void)thr;
  return !(s > thr_le);  // NaN radicand -> collides, as !(NaN > x) in the reference
}

#endif  // __CUDACC__

}  // namespace rrtqx
