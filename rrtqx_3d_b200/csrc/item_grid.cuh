// item_grid.cuh -- obstacle-centric form of "which resident edges collide with these sphere obstacles"
// (explicitEdgeCheck over an obstacle list, DRRT_Q.jl:1775-1826; the edge part of addNewObstacle, DRRT_Q.jl:3220-3290).
//
// The edge-centric kernels (collide_queue.cuh) let every edge look for its obstacles: 10^7 threads, each with its own
// list look-up, its own handful of candidates and its own queue traffic -- 335 warp-instructions per 32 edges, most of
// them bookkeeping.  A resident edge set allows the opposite decomposition, because its items (out-edges, then one
// parent edge per node) do not change between obstacle events:
//
//   * at (re)build time the items are counting-sorted by the cell of their midpoint in a uniform grid (about 32 items
//     per cell) into sorted SoA records: the FP32 reject record (midpoint, half-length bound), the exact end points
//     (three 16-byte arrays) and the item number; every cell also keeps hmax = the largest half length of its items;
//   * a sweep enumerates WORK UNITS (obstacle o, cell c) with dist(c_o, box(c)) <= thr_o + hmax[c] (+ slack): only those
//     cells can hold an item that collides with o, because the reference's closest point lies ON the segment, so a
//     collision needs |c_o - mid| <= thr_o + len/2 (DRRT_Q.jl:1205-1210);
//   * one warp per unit streams the cell's records -- contiguous, coalesced, the obstacle in registers -- applies the
//     FP32 conservative reject and runs the exact FP64 test (distancePointToSegment, 2 sqrt + 1 div) in place: inside a
//     unit most items do collide, so there is no divergence worth a second stage, no pair list and no queue.
//
// A single new obstacle touches the few hundred cells around it instead of all 10^7 items; 256 obstacles at once (C3)
// touch each item 0.5 times on average.  Items the bound does not cover -- zero-length edges and non-finite end points,
// which the reference collides with EVERY active obstacle (t = 0/0 = NaN, !(NaN > x)) -- are kept on a separate list and
// tested against every obstacle of the call; obstacles with a non-finite centre or threshold take every cell.
// The result is an OR over (item, obstacle) pairs of exactly the predicate the edge-centric kernels evaluate.
#pragma once
#include "collision.cuh"
#include "scan.cuh"

namespace rrtqx {

#ifdef __CUDACC__

constexpr int IG_MAX_DIM = 128;
constexpr int64_t IG_MAX_UNITS = (int64_t)1 << 23;  // work units a sweep may list (8 bytes each); more -> fall back

struct ItemGridView {
  int nx, ny, nz;
  double lo[3], inv[3], cell[3];
  double slack;                 // absolute slack of the cell-box bound (FP32 midpoints, cell function roundings)
  int64_t n_sorted;             // items in the grid
  int64_t n_degenerate;         // items on the "every obstacle" list
  const int32_t *cell_start;    // nx*ny*nz + 1
  const float *hmax;            // per cell: upper bound of the half lengths of its items
  const float4 *frec;           // sorted: (fl32 mid.xyz, half-length bound)
  const double2 *ex0, *ex1, *ex2;  // sorted: start.xy | start.z, end.x | end.yz
  const int32_t *item;          // sorted position -> item number
  const int32_t *degenerate;    // item numbers
};

struct ItemGridBufs {
  DevBuf<int32_t> cell_of, cell_start, cursor, item, degenerate, counters, scan_tmp;
  DevBuf<float> hmax;
  DevBuf<float4> frec;
  DevBuf<double2> ex0, ex1, ex2;
  DevBuf<double> bbox;  // 6 doubles
  int nx = 1, ny = 1, nz = 1;
  double lo[3] = {0, 0, 0}, inv[3] = {0, 0, 0}, cell[3] = {1, 1, 1}, slack = 0.0;
  int64_t n_sorted = 0, n_degenerate = 0, n_items = -1;
  bool valid = false;
  ItemGridView view() const {
    ItemGridView v;
    v.nx = nx; v.ny = ny; v.nz = nz;
    for (int c = 0; c < 3; ++c) { v.lo[c] = lo[c]; v.inv[c] = inv[c]; v.cell[c] = cell[c]; }
    v.slack = slack;
    v.n_sorted = n_sorted; v.n_degenerate = n_degenerate;
    v.cell_start = cell_start.p; v.hmax = hmax.p; v.frec = frec.p;
    v.ex0 = ex0.p; v.ex1 = ex1.p; v.ex2 = ex2.p; v.item = item.p; v.degenerate = degenerate.p;
    return v;
  }
};

// ---- build ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool ig_endpoints(const double4 *pos, int64_t n_nodes, const int32_t *src, const int32_t *dst,
                                             int64_t n_edges, const int32_t *parent, int64_t i, double4 &a, double4 &b) {
  int v, w;
  if (i >= n_edges) {
    v = (int)(i - n_edges);
    w = parent ? parent[v] : -1;
  } else {
    v = src[i];
    w = dst[i];
  }
  if (w < 0) return false;
  a = pos[v];
  b = pos[w];
  return true;
}

// midpoint / squared length of an item exactly as the per-call kernels derive them (collide_queue.cuh)
struct IgItem {
  double mx, my, mz, s2;
  bool has, degenerate;
};
__device__ __forceinline__ IgItem ig_item(const double4 &a, const double4 &b) {
  IgItem t;
  const double dx = __dsub_rn(a.x, b.x), dy = __dsub_rn(a.y, b.y), dz = __dsub_rn(a.z, b.z);
  t.s2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
  t.mx = 0.5 * (a.x + b.x); t.my = 0.5 * (a.y + b.y); t.mz = 0.5 * (a.z + b.z);
  t.has = true;
  t.degenerate = !(t.s2 > 0.0) || !isfinite(t.s2) || !isfinite(t.mx + t.my + t.mz);
  return t;
}

// pass 1: bounding box of the finite midpoints (block partials -> 6 atomics on ordered-integer images of doubles)
__device__ __forceinline__ unsigned long long ig_ord(double x) {  // monotone map double -> uint64
  const unsigned long long b = (unsigned long long)__double_as_longlong(x);
  return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}
__host__ __device__ inline double ig_unord(unsigned long long u) {
  const unsigned long long b = (u >> 63) ? (u & 0x7fffffffffffffffull) : ~u;
#ifdef __CUDA_ARCH__
  return __longlong_as_double((long long)b);
#else
  double d;
  memcpy(&d, &b, sizeof(d));
  return d;
#endif
}
static __global__ void __launch_bounds__(256)
ig_bbox_kernel(const double4 *__restrict__ pos, int64_t n_nodes, const int32_t *__restrict__ src,
               const int32_t *__restrict__ dst, int64_t n_edges, const int32_t *__restrict__ parent,
               unsigned long long *__restrict__ box /* min xyz, max xyz as ordered integers */) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  double mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
  double4 a, b;
  if (i < n_edges + n_nodes && ig_endpoints(pos, n_nodes, src, dst, n_edges, parent, i, a, b)) {
    const IgItem t = ig_item(a, b);
    if (!t.degenerate) { mn[0] = mx[0] = t.mx; mn[1] = mx[1] = t.my; mn[2] = mx[2] = t.mz; }
  }
#pragma unroll
  for (int c = 0; c < 3; ++c)
    for (int o = 16; o > 0; o >>= 1) {
      mn[c] = fmin(mn[c], __shfl_xor_sync(FULL, mn[c], o));
      mx[c] = fmax(mx[c], __shfl_xor_sync(FULL, mx[c], o));
    }
  if ((threadIdx.x & 31) == 0)
#pragma unroll
    for (int c = 0; c < 3; ++c) {
      if (mn[c] <= mx[c]) {
        atomicMin(&box[c], ig_ord(mn[c]));
        atomicMax(&box[3 + c], ig_ord(mx[c]));
      }
    }
}

__device__ __forceinline__ int ig_cell(double v, double lo, double inv, int n) {
  const int c = __double2int_rd((v - lo) * inv);  // floor, saturating; NaN -> 0
  return min(max(c, 0), n - 1);
}

// pass 2: cell of every item (-1: no edge, -2: degenerate), histogram, hmax
static __global__ void __launch_bounds__(256)
ig_assign_kernel(const double4 *__restrict__ pos, int64_t n_nodes, const int32_t *__restrict__ src,
                 const int32_t *__restrict__ dst, int64_t n_edges, const int32_t *__restrict__ parent, ItemGridView G,
                 int32_t *__restrict__ cell_of, int32_t *__restrict__ hist, unsigned *__restrict__ hmax_bits,
                 int32_t *__restrict__ degenerate, int32_t *__restrict__ counters /* [0] degenerate items */) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_edges + n_nodes) return;
  double4 a, b;
  if (!ig_endpoints(pos, n_nodes, src, dst, n_edges, parent, i, a, b)) { cell_of[i] = -1; return; }
  const IgItem t = ig_item(a, b);
  if (t.degenerate) {
    cell_of[i] = -2;
    degenerate[atomicAdd(&counters[0], 1)] = (int32_t)i;
    return;
  }
  // the record's midpoint is the FP32 one: it decides the cell, so the cell box bounds what the reject test sees
  const float fx = __double2float_rn(t.mx), fy = __double2float_rn(t.my), fz = __double2float_rn(t.mz);
  const int c = (ig_cell(fz, G.lo[2], G.inv[2], G.nz) * G.ny + ig_cell(fy, G.lo[1], G.inv[1], G.ny)) * G.nx +
                ig_cell(fx, G.lo[0], G.inv[0], G.nx);
  cell_of[i] = c;
  atomicAdd(&hist[c], 1);
  float rt;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rt) : "f"(__double2float_ru(t.s2)));
  const float h = 0.5f * (rt * 1.000001f) + 1e-19f;  // >= len / 2
  atomicMax(&hmax_bits[c], __float_as_uint(h));      // h > 0: the bit patterns order like the values
  atomicMax(&hmax_bits[G.nx * G.ny * G.nz], __float_as_uint(h));  // and the largest of all (obstacle cell boxes)
}

// pass 3: scatter into sorted order
static __global__ void __launch_bounds__(256)
ig_scatter_kernel(const double4 *__restrict__ pos, int64_t n_nodes, const int32_t *__restrict__ src,
                  const int32_t *__restrict__ dst, int64_t n_edges, const int32_t *__restrict__ parent,
                  const int32_t *__restrict__ cell_of, const int32_t *__restrict__ cell_start, int32_t *__restrict__ cursor,
                  float4 *__restrict__ frec, double2 *__restrict__ ex0, double2 *__restrict__ ex1, double2 *__restrict__ ex2,
                  int32_t *__restrict__ item) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_edges + n_nodes) return;
  const int c = cell_of[i];
  if (c < 0) return;
  double4 a, b;
  ig_endpoints(pos, n_nodes, src, dst, n_edges, parent, i, a, b);
  const IgItem t = ig_item(a, b);
  const int j = cell_start[c] + atomicAdd(&cursor[c], 1);
  float rt;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(rt) : "f"(__double2float_ru(t.s2)));
  frec[j] = make_float4(__double2float_rn(t.mx), __double2float_rn(t.my), __double2float_rn(t.mz),
                        0.5f * (rt * 1.000001f) + 1e-19f);
  ex0[j] = make_double2(a.x, a.y);
  ex1[j] = make_double2(a.z, b.x);
  ex2[j] = make_double2(b.y, b.z);
  item[j] = (int32_t)i;
}

// ---- sweep ------------------------------------------------------------------------------------------------------
struct IgObstacle {      // per obstacle of the call
  double cx, cy, cz, thr, thr_le;
  double ext_t, ext_r;   // start-node filter: T_lt(searchRange), searchRange (unused by the plain check)
  float4 f;              // FP32 reject record: centre, threshold rounded up
  float cmax;            // max |centre component| (error bound of the FP32 reject)
  int finite;            // 0: non-finite centre / threshold -> every cell, never rejected
};

// lower bound of |c_o - m| over the (FP32) midpoints m stored in cell (x, y, z); border cells are unbounded outwards
__device__ __forceinline__ double ig_box_dist2(const ItemGridView &G, const IgObstacle &o, int x, int y, int z) {
  const double c[3] = {o.cx, o.cy, o.cz};
  const int k[3] = {x, y, z}, n[3] = {G.nx, G.ny, G.nz};
  double d2 = 0.0;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    const double L = k[a] == 0 ? -INFINITY : G.lo[a] + k[a] * G.cell[a];
    const double H = k[a] == n[a] - 1 ? INFINITY : G.lo[a] + (k[a] + 1) * G.cell[a];
    const double g = fmax(fmax(L - c[a], c[a] - H) - G.slack, 0.0);
    d2 += g * g;
  }
  return d2;
}

// cells obstacle o can reach: its box in cell coordinates (all cells when it is not finite)
__device__ __forceinline__ void ig_cell_box(const ItemGridView &G, const IgObstacle &o, float hmax_all, int lo_[3], int n_[3]) {
  const int n[3] = {G.nx, G.ny, G.nz};
  const double c[3] = {o.cx, o.cy, o.cz};
  const double R = (o.thr + (double)hmax_all) * (1.0 + 1e-6) + G.slack;
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    if (!o.finite) { lo_[a] = 0; n_[a] = n[a]; continue; }
    const int a0 = ig_cell(c[a] - R, G.lo[a], G.inv[a], n[a]), a1 = ig_cell(c[a] + R, G.lo[a], G.inv[a], n[a]);
    lo_[a] = a0; n_[a] = a1 - a0 + 1;
  }
}

__device__ __forceinline__ bool ig_unit_wanted(const ItemGridView &G, const IgObstacle &o, int x, int y, int z) {
  const int c = (z * G.ny + y) * G.nx + x;
  if (G.cell_start[c + 1] == G.cell_start[c]) return false;
  if (!o.finite) return true;
  const double lim = (o.thr + (double)G.hmax[c]) * (1.0 + 1e-6) + G.slack;
  return !(ig_box_dist2(G, o, x, y, z) > lim * lim);
}

// one block per obstacle.  WRITE = false: count the units; WRITE = true: list them at off[o] ..
template <bool WRITE>
static __global__ void __launch_bounds__(256)
ig_units_kernel(ItemGridView G, const IgObstacle *__restrict__ obs, const float *__restrict__ hmax_all,
                int32_t *__restrict__ cnt, const int64_t *__restrict__ off, uint2 *__restrict__ units,
                const int32_t *__restrict__ overflow, const int32_t *__restrict__ n_obs) {
  __shared__ int s_n;
  if (WRITE && *overflow) return;
  const int o = blockIdx.x;
  if (o >= *n_obs) {  // the grid is sized for the upper bound of the obstacle count
    if (!WRITE && threadIdx.x == 0) cnt[o] = 0;
    return;
  }
  const IgObstacle ob = obs[o];
  if (threadIdx.x == 0) s_n = 0;
  __syncthreads();
  int lo_[3], n_[3];
  ig_cell_box(G, ob, *hmax_all, lo_, n_);
  const int total = n_[0] * n_[1] * n_[2];
  int mine = 0;
  for (int k = threadIdx.x; k < total; k += blockDim.x) {
    const int x = lo_[0] + k % n_[0], y = lo_[1] + (k / n_[0]) % n_[1], z = lo_[2] + k / (n_[0] * n_[1]);
    if (ig_unit_wanted(G, ob, x, y, z)) {
      if (WRITE) units[off[o] + atomicAdd(&s_n, 1)] = make_uint2((unsigned)o, (unsigned)((z * G.ny + y) * G.nx + x));
      else ++mine;
    }
  }
  if (!WRITE) {
    for (int s = 16; s > 0; s >>= 1) mine += __shfl_xor_sync(FULL, mine, s);
    if ((threadIdx.x & 31) == 0 && mine) atomicAdd(&s_n, mine);
    __syncthreads();
    if (threadIdx.x == 0) cnt[o] = s_n;
  }
}

// exclusive scan of the per-obstacle unit counts (one block; a call lists at most 2^24 obstacles)
static __global__ void __launch_bounds__(1024)
ig_offsets_kernel(const int32_t *__restrict__ cnt, int n_obs, int64_t *__restrict__ off, int64_t *__restrict__ total,
                  int32_t *__restrict__ overflow, int64_t cap) {
  __shared__ long long sm[33];
  __shared__ long long carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  for (int base = 0; base < n_obs; base += blockDim.x) {
    const int i = base + threadIdx.x;
    const long long v = i < n_obs ? cnt[i] : 0;
    long long tot;
    const long long ex = block_exclusive_scan<long long>(v, sm, &tot);
    if (i < n_obs) off[i] = ex + carry_s;
    __syncthreads();
    if (threadIdx.x == 0) carry_s += tot;
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    *total = carry_s;
    *overflow = carry_s > cap ? 1 : 0;
  }
}

// Sink: what to do with a colliding (item, obstacle) pair.
//   bool accept(const IgObstacle &o, const double a[3], int item)   extra condition (the sweeps' start-node filter)
//   void mark(int item)
// one warp per unit, grid-stride over the unit list
template <bool FMA_DOT, class Sink>
static __global__ void __launch_bounds__(256)
ig_test_kernel(ItemGridView G, const IgObstacle *__restrict__ obs, const uint2 *__restrict__ units,
               const int64_t *__restrict__ total, const int32_t *__restrict__ overflow, Sink S) {
  if (*overflow) return;
  const int lane = threadIdx.x & 31;
  const int64_t n_units = *total;
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t u = warp0; u < n_units; u += n_warps) {
    const uint2 un = units[u];
    const IgObstacle ob = obs[un.x];
    const int j0 = G.cell_start[un.y], j1 = G.cell_start[un.y + 1];
    for (int j = j0 + lane; j < j1; j += 32) {
      const float4 fr = G.frec[j];
      if (ob.finite) {  // FP32 conservative reject (collision.cuh: seg_reject_f32)
        SegF32 sf;
        sf.mx = fr.x; sf.my = fr.y; sf.mz = fr.z; sf.half = fr.w;
        sf.bound = 3.0e-7f * (ob.cmax + fmaxf(fabsf(fr.x), fmaxf(fabsf(fr.y), fabsf(fr.z))));
        sf.ok = isfinite(sf.bound) && isfinite(sf.half);
        if (seg_reject_f32(sf, ob.f)) continue;
      }
      const double2 p0 = G.ex0[j], p1 = G.ex1[j], p2 = G.ex2[j];
      const SegPre pre = seg_prepare(p0.x, p0.y, p1.x, p1.y, p2.x, p2.y);
      if (seg_sphere_collide_exact<FMA_DOT>(pre, ob.cx, ob.cy, ob.cz, ob.thr_le)) {
        const double a[3] = {p0.x, p0.y, p1.x};
        const int it = G.item[j];
        if (S.accept(ob, a, it)) S.mark(it);
      }
    }
  }
}

// degenerate items: every obstacle of the call, exact test (NaN radicand -> collides, as in the reference)
template <bool FMA_DOT, class Sink>
static __global__ void __launch_bounds__(128)
ig_degenerate_kernel(ItemGridView G, const double4 *__restrict__ pos, int64_t n_nodes, const int32_t *__restrict__ src,
                     const int32_t *__restrict__ dst, int64_t n_edges, const int32_t *__restrict__ parent,
                     const IgObstacle *__restrict__ obs, const int32_t *__restrict__ n_obs_dev, Sink S) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= G.n_degenerate) return;
  const int n_obs = *n_obs_dev;
  const int it = G.degenerate[k];
  double4 a4, b4;
  if (!ig_endpoints(pos, n_nodes, src, dst, n_edges, parent, it, a4, b4)) return;
  const SegPre pre = seg_prepare(a4.x, a4.y, a4.z, b4.x, b4.y, b4.z);
  const double a[3] = {a4.x, a4.y, a4.z};
  for (int o = 0; o < n_obs; ++o) {
    const IgObstacle ob = obs[o];
    if (seg_sphere_collide_exact<FMA_DOT>(pre, ob.cx, ob.cy, ob.cz, ob.thr_le) && S.accept(ob, a, it)) {
      S.mark(it);
      return;
    }
  }
}

// obstacle records of a call from (centre+R, (thr, thr_le)[, (T_lt(searchRange), searchRange)])
static __global__ void ig_obstacles_kernel(const double4 *__restrict__ rec, const double2 *__restrict__ thr,
                                           const double2 *__restrict__ ext, const int32_t *__restrict__ n_live, int n_fixed,
                                           IgObstacle *__restrict__ out, int32_t *__restrict__ n_out) {
  const int n = n_live ? *n_live : n_fixed;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *n_out = n;
  if (i >= n) return;
  const double4 r = rec[i];
  const double2 t = thr[i];
  IgObstacle o;
  o.cx = r.x; o.cy = r.y; o.cz = r.z; o.thr = t.x; o.thr_le = t.y;
  o.ext_t = ext ? ext[i].x : 0.0;
  o.ext_r = ext ? ext[i].y : 0.0;
  o.f = make_float4(__double2float_rn(r.x), __double2float_rn(r.y), __double2float_rn(r.z), __double2float_ru(t.x));
  o.cmax = __double2float_ru(fmax(fabs(r.x), fmax(fabs(r.y), fabs(r.z))));
  o.finite = (isfinite(r.x) && isfinite(r.y) && isfinite(r.z) && isfinite(t.x)) ? 1 : 0;
  out[i] = o;
}

// ---- host side --------------------------------------------------------------------------------------------------
struct IgRunBufs {  // per context (rrtqx_ctx::scratch)
  DevBuf<unsigned char> obs;   // IgObstacle[n]
  DevBuf<int32_t> cnt, flags;  // units per obstacle | [0] overflow, [1] obstacles of the call
  DevBuf<int64_t> off;         // unit offsets per obstacle, [n] = total
  DevBuf<uint2> units;
};
inline char g_ig_run_tag = 0;
static inline IgRunBufs &ig_run_bufs(rrtqx_ctx *ctx) { return ctx->scratch.get<IgRunBufs>(&g_ig_run_tag); }

// Sort the items of an edge set into the grid.  Synchronises the stream twice (bounding box, degenerate count): it
// runs with the edge-set (re)build, not inside a sweep.
static inline void item_grid_build(rrtqx_ctx *ctx, ItemGridBufs &B, const double4 *pos, int64_t n_nodes, const int32_t *src,
                                   const int32_t *dst, int64_t n_edges, const int32_t *parent) {
  cudaStream_t st = ctx->stream;
  const int64_t n_items = n_edges + n_nodes;
  B.valid = false;
  B.n_items = n_items;
  B.n_sorted = B.n_degenerate = 0;
  if (n_items <= 0 || n_items >= ((int64_t)1 << 31)) return;
  const int TB = 256;
  const unsigned blocks = (unsigned)div_up(n_items, TB);
  B.bbox.ensure(8, st);
  unsigned long long init[6] = {~0ull, ~0ull, ~0ull, 0ull, 0ull, 0ull};
  RQ_CUDA(cudaMemcpyAsync(B.bbox.p, init, sizeof(init), cudaMemcpyHostToDevice, st));
  ig_bbox_kernel<<<blocks, TB, 0, st>>>(pos, n_nodes, src, dst, n_edges, parent, (unsigned long long *)B.bbox.p);
  post_launch(ctx);
  unsigned long long box[6];
  RQ_CUDA(cudaMemcpyAsync(box, B.bbox.p, sizeof(box), cudaMemcpyDeviceToHost, st));
  RQ_CUDA(cudaStreamSynchronize(st));
  double mn[3], mx[3];
  bool any = true;
  for (int c = 0; c < 3; ++c) {
    if (box[c] == ~0ull) { any = false; break; }
    mn[c] = ig_unord(box[c]); mx[c] = ig_unord(box[3 + c]);
  }
  // grid shape: about 32 items per cell, cubic cells, at most IG_MAX_DIM per axis
  int dims[3] = {1, 1, 1};
  double span = 0.0, mag = 0.0;
  if (any) {
    double ext[3], vol = 1.0;
    int nz_dims = 0;
    for (int c = 0; c < 3; ++c) {
      ext[c] = mx[c] - mn[c];
      span = std::max(span, ext[c]);
      mag = std::max(mag, std::max(std::fabs(mn[c]), std::fabs(mx[c])));
      if (ext[c] > 0.0 && std::isfinite(ext[c])) { vol *= ext[c]; nz_dims++; }
    }
    const double target_cells = std::max(1.0, (double)n_items / 32.0);
    const double cell_len = nz_dims ? std::pow(vol / target_cells, 1.0 / nz_dims) : 1.0;
    for (int c = 0; c < 3; ++c) {
      if (ext[c] > 0.0 && std::isfinite(ext[c]) && cell_len > 0.0 && std::isfinite(cell_len))
        dims[c] = (int)std::min<double>(IG_MAX_DIM, std::max(1.0, std::floor(ext[c] / cell_len) + 1.0));
      B.lo[c] = mn[c];
      B.cell[c] = dims[c] > 1 ? ext[c] / dims[c] : std::max(ext[c], 1.0);
      B.inv[c] = dims[c] > 1 ? dims[c] / ext[c] : 0.0;
    }
  } else {
    for (int c = 0; c < 3; ++c) { B.lo[c] = 0.0; B.cell[c] = 1.0; B.inv[c] = 0.0; }
  }
  B.nx = dims[0]; B.ny = dims[1]; B.nz = dims[2];
  // FP32 midpoints (relative 2^-24 of their magnitude), the roundings of the cell function and of the box corners
  B.slack = 1e-6 * (mag + span) + 1e-300;
  const int64_t ncell = (int64_t)B.nx * B.ny * B.nz;
  B.cell_of.ensure((size_t)n_items + 1, st);
  B.cursor.ensure((size_t)ncell + 2, st);
  B.cell_start.ensure((size_t)ncell + 2, st);
  B.hmax.ensure((size_t)ncell + 2, st);
  B.degenerate.ensure((size_t)n_items + 1, st);
  B.counters.ensure(4, st);
  RQ_CUDA(cudaMemsetAsync(B.cursor.p, 0, sizeof(int32_t) * ((size_t)ncell + 2), st));
  RQ_CUDA(cudaMemsetAsync(B.hmax.p, 0, sizeof(float) * ((size_t)ncell + 2), st));
  RQ_CUDA(cudaMemsetAsync(B.counters.p, 0, sizeof(int32_t) * 4, st));
  ItemGridView G = B.view();
  ig_assign_kernel<<<blocks, TB, 0, st>>>(pos, n_nodes, src, dst, n_edges, parent, G, B.cell_of.p, B.cursor.p,
                                          (unsigned *)B.hmax.p, B.degenerate.p, B.counters.p);
  post_launch(ctx);
  exclusive_scan<int32_t, int32_t>(ctx, B.cursor.p, ncell, B.cell_start.p, B.scan_tmp);
  int32_t counts[2] = {0, 0};
  RQ_CUDA(cudaMemcpyAsync(&counts[0], B.counters.p, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  RQ_CUDA(cudaMemcpyAsync(&counts[1], B.cell_start.p + ncell, sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  RQ_CUDA(cudaStreamSynchronize(st));
  B.n_degenerate = counts[0];
  B.n_sorted = counts[1];
  B.frec.ensure((size_t)B.n_sorted + 1, st);
  B.ex0.ensure((size_t)B.n_sorted + 1, st);
  B.ex1.ensure((size_t)B.n_sorted + 1, st);
  B.ex2.ensure((size_t)B.n_sorted + 1, st);
  B.item.ensure((size_t)B.n_sorted + 1, st);
  RQ_CUDA(cudaMemsetAsync(B.cursor.p, 0, sizeof(int32_t) * ((size_t)ncell + 2), st));
  ig_scatter_kernel<<<blocks, TB, 0, st>>>(pos, n_nodes, src, dst, n_edges, parent, B.cell_of.p, B.cell_start.p, B.cursor.p,
                                           B.frec.p, B.ex0.p, B.ex1.p, B.ex2.p, B.item.p);
  post_launch(ctx);
  B.valid = true;
}

// The obstacle-centric check of every item of a grid against the obstacles (rec, thr[, ext]) of a call (live count on
// the device in *n_live, or n_upper when n_live is NULL).  Everything is queued on the stream; *overflow_dev (device
// flag) reads 1 afterwards if the call listed more than IG_MAX_UNITS work units -- then NOTHING was marked and the
// caller repeats the call through the edge-centric kernels.
template <bool FMA_DOT, class Sink>
static inline const int32_t *item_grid_run(rrtqx_ctx *ctx, const ItemGridBufs &B, const double4 *rec, const double2 *thr,
                                           const double2 *ext, const int32_t *n_live, int n_upper, const Sink &S,
                                           const double4 *pos, int64_t n_nodes, const int32_t *src, const int32_t *dst,
                                           int64_t n_edges, const int32_t *parent) {
  cudaStream_t st = ctx->stream;
  IgRunBufs &R = ig_run_bufs(ctx);
  R.obs.ensure(sizeof(IgObstacle) * ((size_t)n_upper + 1), st);
  R.cnt.ensure((size_t)n_upper + 1, st);
  R.off.ensure((size_t)n_upper + 2, st);
  R.flags.ensure(4, st);
  R.units.ensure((size_t)IG_MAX_UNITS, st);
  IgObstacle *obs = (IgObstacle *)R.obs.p;
  int32_t *overflow = R.flags.p, *n_obs = R.flags.p + 1;
  const ItemGridView G = B.view();
  const float *hmax_all = B.hmax.p + (int64_t)B.nx * B.ny * B.nz;
  if (n_upper <= 0) {
    RQ_CUDA(cudaMemsetAsync(R.flags.p, 0, sizeof(int32_t) * 4, st));
    return overflow;
  }
  ig_obstacles_kernel<<<div_up(n_upper, 128), 128, 0, st>>>(rec, thr, ext, n_live, n_upper, obs, n_obs);
  ig_units_kernel<false><<<n_upper, 256, 0, st>>>(G, obs, hmax_all, R.cnt.p, nullptr, nullptr, overflow, n_obs);
  ig_offsets_kernel<<<1, 1024, 0, st>>>(R.cnt.p, n_upper, R.off.p, R.off.p + n_upper, overflow, IG_MAX_UNITS);
  ig_units_kernel<true><<<n_upper, 256, 0, st>>>(G, obs, hmax_all, R.cnt.p, R.off.p, R.units.p, overflow, n_obs);
  ig_test_kernel<FMA_DOT, Sink><<<ctx->sm_count * 8, 256, 0, st>>>(G, obs, R.units.p, R.off.p + n_upper, overflow, S);
  post_launch(ctx, 5);
  if (B.n_degenerate > 0) {
    ig_degenerate_kernel<FMA_DOT, Sink><<<div_up(B.n_degenerate, 128), 128, 0, st>>>(G, pos, n_nodes, src, dst, n_edges, parent,
                                                                                     obs, n_obs, S);
    post_launch(ctx);
  }
  return overflow;
}

#endif  // __CUDACC__

}  // namespace rrtqx
