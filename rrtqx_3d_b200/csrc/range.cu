// range.cu -- batched kdFindWithinRange (kdTree_general.jl:889-919) and
// kdFindNearest (:357-385) over the cell-sorted SoA index.
//
// Result semantics: the reference's pruned kd
// traversal returns exactly { n : fl(sqrt(s(q,n))) < r } (root admitted with
// <=), s being the left-to-right radicand of euclidianDist.  We evaluate the
// same s for every candidate the grid cannot exclude, and decide membership
// with the sqrt-free but equivalent test s < T_lt(r).  sqrt is paid on hits
// only (the returned JList key).
#include <cstdlib>

#include "objects.cuh"
#include "scan.cuh"

namespace rrtqx {

// ------------------------------------------------- sort queries by grid cell
template <int D>
__global__ void query_key_kernel(GridView g, const double *__restrict__ q, int64_t nq, int32_t *__restrict__ key,
                                 int32_t *__restrict__ hist) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  const double *r = q + i * D;
  int cx = cell_of(r[0], g.lo[0], g.inv[0], g.nx);
  int cy = cell_of(r[1], g.lo[1], g.inv[1], g.ny);
  int cz = D >= 3 ? cell_of(r[2], g.lo[2], g.inv[2], g.nz) : 0;
  int c = (cz * g.ny + cy) * g.nx + cx;
  key[i] = c;
  atomicAdd(&hist[c], 1);
}

__global__ void query_scatter_kernel(const int32_t *__restrict__ key, int64_t nq, const int32_t *__restrict__ start,
                                     int32_t *__restrict__ cursor, int32_t *__restrict__ order) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= nq) return;
  int c = key[i];
  order[start[c] + atomicAdd(&cursor[c], 1)] = (int32_t)i;
}

__global__ void iota_kernel(int32_t *__restrict__ a, int64_t n) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) a[i] = (int32_t)i;
}

// Order in which warps pick up queries: grouped by the cell that contains the
// query so that neighbouring warps touch the same candidate slices (L1/L2 reuse).
template <int D>
static void sort_queries(rrtqx_tree *t, rrtqx_range_result *r, const double *dq, int64_t nq) {
  rrtqx_ctx *ctx = t->ctx;
  cudaStream_t st = ctx->stream;
  const int TB = 256;
  r->qorder.ensure((size_t)nq, st);
  const int ncell = t->nx * t->ny * t->nz;
  if (t->n_sorted == 0 || nq < 2048) {
    iota_kernel<<<div_up(nq, TB), TB, 0, st>>>(r->qorder.p, nq);
    post_launch(ctx);
    return;
  }
  r->qkey.ensure((size_t)nq, st);
  r->qhist.ensure((size_t)ncell + 1, st);
  r->qstart.ensure((size_t)ncell + 1, st);
  RQ_CUDA(cudaMemsetAsync(r->qhist.p, 0, sizeof(int32_t) * ((size_t)ncell + 1), st));
  GridView g = t->view();
  query_key_kernel<D><<<div_up(nq, TB), TB, 0, st>>>(g, dq, nq, r->qkey.p, r->qhist.p);
  post_launch(ctx);
  exclusive_scan<int32_t, int32_t>(ctx, r->qhist.p, ncell, r->qstart.p, r->scan_tmp32);
  RQ_CUDA(cudaMemsetAsync(r->qhist.p, 0, sizeof(int32_t) * ((size_t)ncell + 1), st));
  query_scatter_kernel<<<div_up(nq, TB), TB, 0, st>>>(r->qkey.p, nq, r->qstart.p, r->qhist.p, r->qorder.p);
  post_launch(ctx);
}

// --------------------------------------------------------- query identities
// The real query point plus its ghost identities in the order of
// getNextGhostPoint (ghostPoint.jl:60-111): pattern p = 1 .. 2^w-1, bit b of p
// wraps dimension wraps[w-1-b]; a ghost is used iff
// !(euclid(closestUnwrappedPoint, ghost) > bound) (:104).
template <int D>
struct Identities {
  double q[1 << MAX_WRAPS][D];
  int count;
};

template <int D>
__device__ inline bool make_ghost(const WrapInfo &w, const double *q, int pattern, double bound, double *ghost) {
  double closest[D];
#pragma unroll
  for (int k = 0; k < D; ++k) { ghost[k] = q[k]; closest[k] = q[k]; }
  for (int b = 0; b < w.num_wraps; ++b) {
    if (!((pattern >> b) & 1)) continue;
    int wi = w.num_wraps - 1 - b;
    int dim = w.wraps[wi];
    double P = w.wrap_points[wi];
    double gv, cv;
    if (q[dim] < P / 2.0) { gv = __dadd_rn(q[dim], P); cv = P; }   // :82-85
    else                  { gv = __dsub_rn(q[dim], P); cv = 0.0; } // :86-89
#pragma unroll
    for (int k = 0; k < D; ++k) if (k == dim) { ghost[k] = gv; closest[k] = cv; }
  }
  double s = sqdist<D>(closest, ghost[0], ghost[1], D >= 3 ? ghost[2] : 0.0, D >= 4 ? ghost[3] : 0.0);
  return !(__dsqrt_rn(s) > bound);
}

// ------------------------------------------------------------ range kernel
struct RowSpan { int start, end; };

// Conservative slice of row (cy,cz) that can contain points within r of q.
template <int D>
__device__ __forceinline__ RowSpan row_span(const GridView &g, const double *q, double r_infl, double r2_infl, int cy,
                                           int cz) {
  // lower bounds of |p.y-q.y| and |p.z-q.z| for points stored in this row;
  // boundary rows extend to infinity (clamped cells hold everything beyond).
  double dy = 0.0, dz = 0.0;
  {
    double ylo = g.lo[1] + cy * g.cell[1], yhi = ylo + g.cell[1];
    double a = (cy == 0) ? -INFINITY : ylo - q[1];
    double b = (cy == g.ny - 1) ? -INFINITY : q[1] - yhi;
    dy = fmax(0.0, fmax(a, b) - 1e-9 * g.cell[1] - 1e-12 * fabs(q[1]));
  }
  if (D >= 3) {
    double zlo = g.lo[2] + cz * g.cell[2], zhi = zlo + g.cell[2];
    double a = (cz == 0) ? -INFINITY : zlo - q[2];
    double b = (cz == g.nz - 1) ? -INFINITY : q[2] - zhi;
    dz = fmax(0.0, fmax(a, b) - 1e-9 * g.cell[2] - 1e-12 * fabs(q[2]));
  }
  RowSpan s;
  double rem = r2_infl - dy * dy - dz * dz;
  if (!(rem >= 0.0)) { s.start = 0; s.end = 0; return s; }
  double xr = isinf(rem) ? rem : sqrt(rem) * (1.0 + 1e-9) + 1e-12 * fabs(q[0]);
  (void)r_infl;
  int cxa = cell_of(q[0] - xr, g.lo[0], g.inv[0], g.nx);
  int cxb = cell_of(q[0] + xr, g.lo[0], g.inv[0], g.nx);
  int base = (cz * g.ny + cy) * g.nx;
  s.start = g.cell_start[base + cxa];
  s.end = g.cell_start[base + cxb + 1];
  return s;
}

// One warp per query.  FILL = false: count hits.  FILL = true: write idx/dist
// at offsets[q].  WRAP: tree has wrap-around dimensions (ghost identities).
template <int D, bool FILL, bool WRAP>
__global__ void __launch_bounds__(256)
range_query_kernel(GridView g, WrapInfo wrap, const double *__restrict__ queries, const int32_t *__restrict__ qorder,
                   int64_t nq, double r_uniform, const double *__restrict__ ranges, int32_t *__restrict__ counts,
                   const int64_t *__restrict__ offsets, int32_t *__restrict__ out_idx, double *__restrict__ out_dist) {
  const int lane = lane_id();
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const unsigned lt = lanemask_lt();

  for (int64_t qi = warp0; qi < nq; qi += nwarps) {
    const int qid = qorder[qi];
    const double r = ranges ? ranges[qid] : r_uniform;
    double q0[D];
#pragma unroll
    for (int k = 0; k < D; ++k) q0[k] = queries[(int64_t)qid * D + k];
    const double T = sqrt_thresh_lt(r);
    int64_t wbase = FILL ? offsets[qid] : 0;  // next free output slot
    int cnt = 0;                              // per-lane hit count (COUNT mode)

    // The root (node 0) is admitted with <= by the real identity only (:896-898).
    bool root_hit = false;
    {
      double4 p0 = g.pos[0];
      double s0 = sqdist<D>(q0, p0.x, p0.y, p0.z, p0.w);
      root_hit = (__dsqrt_rn(s0) <= r);
      if (root_hit) {
        if (FILL) {
          if (lane == 0) {
            out_idx[wbase] = 0;
            if (out_dist) out_dist[wbase] = __dsqrt_rn(s0);
          }
          wbase += 1;
        } else if (lane == 0) {
          cnt += 1;
        }
      }
    }
    if (!(r > 0.0)) {  // r <= 0 or NaN: no strict hit is possible
      if (!FILL) {
        cnt = __reduce_add_sync(FULL, cnt);
        if (lane == 0) counts[qid] = cnt;
      }
      continue;
    }

    const int n_ident = WRAP ? (1 << wrap.num_wraps) : 1;
    double qprev[WRAP ? ((1 << MAX_WRAPS) - 1) : 1][D];  // identities already searched
    int n_prev = 0;
    for (int ident = 0; ident < n_ident; ++ident) {
      double q[D];
      if (ident == 0) {
#pragma unroll
        for (int k = 0; k < D; ++k) q[k] = q0[k];
      } else {
        if (!make_ghost<D>(wrap, q0, ident, r, q)) continue;
      }
      // candidate test for one point; returns true if this identity is the
      // first to reach it
      auto test = [&](double px, double py, double pz, double pw, int node, double &s) -> bool {
        s = sqdist<D>(q, px, py, pz, pw);
        if (!(s < T)) return false;
        if (node == 0 && root_hit) return false;  // already listed
        if (WRAP) {
          for (int m = 0; m < n_prev; ++m) {
            double sm = sqdist<D>(qprev[m], px, py, pz, pw);
            if (sm < T) return false;  // an earlier identity found it (addToRangeList dedup, :765-771)
          }
        }
        return true;
      };
      auto emit = [&](bool hit, int node, double s) {
        if (FILL) {
          unsigned m = __ballot_sync(FULL, hit);
          if (hit) {
            int64_t o = wbase + __popc(m & lt);
            out_idx[o] = node;
            if (out_dist) out_dist[o] = __dsqrt_rn(s);
          }
          wbase += __popc(m);
        } else {
          cnt += hit ? 1 : 0;
        }
      };

      if (g.n_sorted > 0) {
        const double r_infl = r * (1.0 + 1e-9);
        const double r2_infl = r_infl * r_infl;
        const int cy0 = cell_of(q[1] - r_infl, g.lo[1], g.inv[1], g.ny);
        const int cy1 = cell_of(q[1] + r_infl, g.lo[1], g.inv[1], g.ny);
        const int cz0 = D >= 3 ? cell_of(q[2] - r_infl, g.lo[2], g.inv[2], g.nz) : 0;
        const int cz1 = D >= 3 ? cell_of(q[2] + r_infl, g.lo[2], g.inv[2], g.nz) : 0;
        const int wy = cy1 - cy0 + 1;
        const int nrows = wy * (cz1 - cz0 + 1);
        for (int rb = 0; rb < nrows; rb += 32) {
          RowSpan sp{0, 0};
          int row = rb + lane;
          if (row < nrows) sp = row_span<D>(g, q, r_infl, r2_infl, cy0 + row % wy, cz0 + row / wy);
          const int lim = min(32, nrows - rb);
          for (int t = 0; t < lim; ++t) {
            const int a = __shfl_sync(FULL, sp.start, t);
            const int b = __shfl_sync(FULL, sp.end, t);
            for (int j0 = a; j0 < b; j0 += 32) {
              const int j = j0 + lane;
              bool hit = false;
              int node = 0;
              double s = 0.0;
              if (j < b) {
                node = g.sperm[j];
                hit = test(g.sx[j], g.sy[j], D >= 3 ? g.sz[j] : 0.0, D >= 4 ? g.sw[j] : 0.0, node, s);
              }
              emit(hit, node, s);
            }
          }
        }
      }
      // unsorted tail of recent inserts
      for (int j0 = g.n_sorted; j0 < g.n_total; j0 += 32) {
        const int j = j0 + lane;
        bool hit = false;
        double s = 0.0;
        if (j < g.n_total) {
          double4 p = g.pos[j];
          hit = test(p.x, p.y, p.z, p.w, j, s);
        }
        emit(hit, j, s);
      }
      if (WRAP) {
#pragma unroll
        for (int k = 0; k < D; ++k) qprev[n_prev][k] = q[k];
        n_prev++;
      }
    }
    if (!FILL) {
      cnt = __reduce_add_sync(FULL, cnt);
      if (lane == 0) counts[qid] = cnt;
    }
  }
}

template <int D>
static void launch_range(rrtqx_tree *t, bool fill, const double *dq, const int32_t *qorder, int64_t nq, double r,
                         const double *ranges, int32_t *counts, const int64_t *offsets, int32_t *idx, double *dist) {
  rrtqx_ctx *ctx = t->ctx;
  GridView g = t->view();
  const int TB = 256;
  int64_t warps_needed = nq;
  int blocks = (int)std::min<int64_t>((warps_needed * 32 + TB - 1) / TB, (int64_t)ctx->sm_count * 8);
  if (blocks < 1) blocks = 1;
  const bool wrap = t->wrap.num_wraps > 0;
#define RQ_LAUNCH(FILL, WRAP)                                                                          \
  range_query_kernel<D, FILL, WRAP><<<blocks, TB, 0, ctx->stream>>>(g, t->wrap, dq, qorder, nq, r, ranges, counts, \
                                                                    offsets, idx, dist)
  if (fill) { if (wrap) RQ_LAUNCH(true, true); else RQ_LAUNCH(true, false); }
  else      { if (wrap) RQ_LAUNCH(false, true); else RQ_LAUNCH(false, false); }
#undef RQ_LAUNCH
  post_launch(ctx);
}


// ------------------------------------------------- fused single-pass kernel
// One warp per group of QN consecutive (cell-sorted) queries.  Every candidate
// is loaded once and tested against all QN queries (register tiling: halves the
// L1/LSU traffic per test); hits are recorded as slot numbers in a per-warp
// shared-memory buffer; when the scan of a group ends the warp reserves the
// exact output range with ONE atomic on a global cursor and flushes densely:
// coalesced 32-wide stores, and sqrt evaluated on hits only, all lanes busy.
// No count pass: the tree is traversed once.  Blocks take chunks of consecutive
// sorted queries so that the warps of a block share their candidate rows in L1.
constexpr int FUSED_CAP = 768;      // buffered hits per query before the direct-write fallback
constexpr int FUSED_CHUNK = 256;    // queries per block-level work item
constexpr int FUSED_WARPS = 8;

template <int D, int QN>
struct Group {
  double q[QN][D];
  double T[QN];   // strict threshold on the radicand
  double r[QN];
  int qid[QN];
};

// Scan all candidates of the group.  DIRECT = false: append hit slots to buf
// (up to FUSED_CAP per query), cnt[] = number of hits.  DIRECT = true: write
// results at base[k] + ordinal (used when a query overflowed its buffer).
// Slots: j >= 0 is a position in the cell-sorted arrays, j < 0 encodes node
// -(j+1) of the unsorted tail.
template <int D, int QN, bool DIRECT>
__device__ __forceinline__ void scan_group(const GridView &g, const Group<D, QN> &G, int lane, unsigned lt,
                                           int *__restrict__ buf, int (&cnt)[QN], const int64_t (&base)[QN],
                                           int32_t *__restrict__ out_idx, double *__restrict__ out_dist) {
#pragma unroll
  for (int k = 0; k < QN; ++k) cnt[k] = 0;

  auto visit = [&](bool valid, int slot, double px, double py, double pz, double pw) {
#pragma unroll
    for (int k = 0; k < QN; ++k) {
      const double s = sqdist<D>(G.q[k], px, py, pz, pw);
      const bool hit = valid && (s < G.T[k]);
      const unsigned m = __ballot_sync(FULL, hit);
      if (hit) {
        const int o = cnt[k] + __popc(m & lt);
        if (DIRECT) {
          out_idx[base[k] + o] = slot >= 0 ? g.sperm[slot] : -(slot + 1);
          if (out_dist) out_dist[base[k] + o] = __dsqrt_rn(s);
        } else if (o < FUSED_CAP) {
          buf[k * FUSED_CAP + o] = slot;
        }
      }
      cnt[k] += __popc(m);
    }
  };

  if (g.n_sorted > 0) {
    // union of the groups' row ranges
    int cy0 = 0x7fffffff, cy1 = -1, cz0 = 0x7fffffff, cz1 = -1;
#pragma unroll
    for (int k = 0; k < QN; ++k) {
      if (!(G.r[k] > 0.0)) continue;
      const double ri = G.r[k] * (1.0 + 1e-9);
      cy0 = min(cy0, cell_of(G.q[k][1] - ri, g.lo[1], g.inv[1], g.ny));
      cy1 = max(cy1, cell_of(G.q[k][1] + ri, g.lo[1], g.inv[1], g.ny));
      if (D >= 3) {
        cz0 = min(cz0, cell_of(G.q[k][2] - ri, g.lo[2], g.inv[2], g.nz));
        cz1 = max(cz1, cell_of(G.q[k][2] + ri, g.lo[2], g.inv[2], g.nz));
      } else {
        cz0 = 0; cz1 = 0;
      }
    }
    if (cy1 >= cy0 && cz1 >= cz0) {
      const int wy = cy1 - cy0 + 1;
      const int nrows = wy * (cz1 - cz0 + 1);
      for (int rb = 0; rb < nrows; rb += 32) {
        int sa = 0x7fffffff, sb = 0;  // union span of this lane's row over the group
        const int row = rb + lane;
        if (row < nrows) {
          const int cy = cy0 + row % wy, cz = cz0 + row / wy;
#pragma unroll
          for (int k = 0; k < QN; ++k) {
            if (!(G.r[k] > 0.0)) continue;
            const double ri = G.r[k] * (1.0 + 1e-9);
            RowSpan sp = row_span<D>(g, G.q[k], ri, ri * ri, cy, cz);
            if (sp.end > sp.start) { sa = min(sa, sp.start); sb = max(sb, sp.end); }
          }
        }
        if (sb <= sa) { sa = 0; sb = 0; }
        const int lim = min(32, nrows - rb);
        for (int t = 0; t < lim; ++t) {
          const int a = __shfl_sync(FULL, sa, t);
          const int b = __shfl_sync(FULL, sb, t);
          int j0 = a;
          // two candidates per lane per trip: independent dependency chains
          for (; j0 + 32 < b; j0 += 64) {
            const int ja = j0 + lane, jb = j0 + 32 + lane;
            const bool vb = jb < b;
            const double xa = g.sx[ja], ya = g.sy[ja], za = D >= 3 ? g.sz[ja] : 0.0, wa = D >= 4 ? g.sw[ja] : 0.0;
            const int jbc = vb ? jb : ja;
            const double xb = g.sx[jbc], yb = g.sy[jbc], zb = D >= 3 ? g.sz[jbc] : 0.0, wb = D >= 4 ? g.sw[jbc] : 0.0;
            visit(true, ja, xa, ya, za, wa);
            visit(vb, jb, xb, yb, zb, wb);
          }
          if (j0 < b) {
            const int j = j0 + lane;
            const bool v = j < b;
            const int jc = v ? j : a;
            visit(v, j, g.sx[jc], g.sy[jc], D >= 3 ? g.sz[jc] : 0.0, D >= 4 ? g.sw[jc] : 0.0);
          }
        }
      }
    }
  }
  for (int j0 = g.n_sorted; j0 < g.n_total; j0 += 32) {  // unsorted tail of recent inserts
    const int j = j0 + lane;
    const bool v = j < g.n_total;
    double4 p = make_double4(0, 0, 0, 0);
    if (v) p = g.pos[j];
    visit(v, -(j + 1), p.x, p.y, p.z, p.w);
  }
}

template <int D, int QN>
__global__ void __launch_bounds__(FUSED_WARPS * 32, 2)
range_fused_kernel(GridView g, const double *__restrict__ queries, const int32_t *__restrict__ qorder, int64_t nq,
                   double r_uniform, const double *__restrict__ ranges, int32_t *__restrict__ counts,
                   int64_t *__restrict__ offsets, int32_t *__restrict__ out_idx, double *__restrict__ out_dist,
                   unsigned long long cap, unsigned long long *__restrict__ cursor /* [0] output, [1] chunk */,
                   int write_lists) {
  extern __shared__ int s_buf[];  // [FUSED_WARPS][QN][FUSED_CAP]
  __shared__ unsigned long long s_chunk;
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const unsigned lt = lanemask_lt();
  int *buf = s_buf + warp * QN * FUSED_CAP;

  for (;;) {
    __syncthreads();
    if (threadIdx.x == 0) s_chunk = atomicAdd(&cursor[1], 1ull);
    __syncthreads();
    const int64_t chunk_base = (int64_t)s_chunk * FUSED_CHUNK;
    if (chunk_base >= nq) break;

    for (int p = warp; p * QN < FUSED_CHUNK; p += FUSED_WARPS) {
      const int64_t qfirst = chunk_base + (int64_t)p * QN;
      if (qfirst >= nq) break;
      Group<D, QN> G;
      bool root_extra[QN];
      double root_s = 0.0;
#pragma unroll
      for (int k = 0; k < QN; ++k) {
        const bool have = qfirst + k < nq;
        G.qid[k] = have ? qorder[qfirst + k] : -1;
        const int qq = have ? G.qid[k] : G.qid[0];
#pragma unroll
        for (int c = 0; c < D; ++c) G.q[k][c] = queries[(int64_t)qq * D + c];
        G.r[k] = have ? (ranges ? ranges[qq] : r_uniform) : -1.0;
        G.T[k] = sqrt_thresh_lt(G.r[k]);
        // root (node 0) is admitted with <= (kdTree_general.jl:896-898): it needs
        // an explicit entry only when it sits exactly at distance r
        const double4 p0 = g.pos[0];
        const double s0 = sqdist<D>(G.q[k], p0.x, p0.y, p0.z, p0.w);
        root_extra[k] = have && (__dsqrt_rn(s0) <= G.r[k]) && !(s0 < G.T[k]);
        if (root_extra[k]) root_s = s0;
      }
      int cnt[QN];
      int64_t base[QN];
#pragma unroll
      for (int k = 0; k < QN; ++k) base[k] = 0;
      scan_group<D, QN, false>(g, G, lane, lt, buf, cnt, base, nullptr, nullptr);

      // reserve the exact output range of the group: one atomic per group
      unsigned long long need = 0;
#pragma unroll
      for (int k = 0; k < QN; ++k) need += (unsigned long long)(cnt[k] + (root_extra[k] ? 1 : 0));
      unsigned long long b0 = 0;
      if (lane == 0) b0 = atomicAdd(&cursor[0], need);
      b0 = __shfl_sync(FULL, b0, 0);
      bool overflow = false;
#pragma unroll
      for (int k = 0; k < QN; ++k) {
        base[k] = (int64_t)b0;
        b0 += (unsigned long long)(cnt[k] + (root_extra[k] ? 1 : 0));
        if (cnt[k] > FUSED_CAP) overflow = true;
        if (lane == 0 && G.qid[k] >= 0) {
          counts[G.qid[k]] = cnt[k] + (root_extra[k] ? 1 : 0);
          offsets[G.qid[k]] = base[k];
        }
      }
      if (!write_lists || b0 > cap) continue;  // counts only, or the lists do not fit (host retries)
      if (overflow) {
        int cnt2[QN];
        scan_group<D, QN, true>(g, G, lane, lt, buf, cnt2, base, out_idx, out_dist);
      } else {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < QN; ++k) {
          for (int h = lane; h < cnt[k]; h += 32) {  // dense flush: all lanes hold a hit
            const int slot = buf[k * FUSED_CAP + h];
            double px, py, pz = 0.0, pw = 0.0;
            int node;
            if (slot >= 0) {
              px = g.sx[slot]; py = g.sy[slot];
              if (D >= 3) pz = g.sz[slot];
              if (D >= 4) pw = g.sw[slot];
              node = g.sperm[slot];
            } else {
              node = -(slot + 1);
              const double4 pp = g.pos[node];
              px = pp.x; py = pp.y; pz = pp.z; pw = pp.w;
            }
            out_idx[base[k] + h] = node;
            if (out_dist) out_dist[base[k] + h] = __dsqrt_rn(sqdist<D>(G.q[k], px, py, pz, pw));
          }
        }
        __syncwarp();
      }
#pragma unroll
      for (int k = 0; k < QN; ++k)
        if (root_extra[k] && lane == 0) {
          out_idx[base[k] + cnt[k]] = 0;
          if (out_dist) out_dist[base[k] + cnt[k]] = __dsqrt_rn(root_s);
        }
    }
  }
}

template <int D>
static void range_query_fused(rrtqx_tree *t, const double *dq, const double *dr, int64_t nq, double r, uint32_t flags,
                              rrtqx_range_result *res) {
  rrtqx_ctx *ctx = t->ctx;
  cudaStream_t st = ctx->stream;
  constexpr int QN = 2;
  const bool count_only = flags & RRTQX_RANGE_COUNT_ONLY;
  const bool want_dist = flags & RRTQX_RANGE_WANT_DIST;
  res->cursor.ensure(4, st);
  {
    PhaseScope p2(ctx, "range_sort");
    sort_queries<D>(t, res, dq, nq);
  }
  const size_t smem = (size_t)FUSED_WARPS * QN * FUSED_CAP * sizeof(int);
  static bool attr_set[5] = {false, false, false, false, false};
  if (!attr_set[D]) {
    RQ_CUDA(cudaFuncSetAttribute(range_fused_kernel<D, QN>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    attr_set[D] = true;
  }
  const int blocks = (int)std::min<int64_t>((nq + FUSED_CHUNK - 1) / FUSED_CHUNK, (int64_t)ctx->sm_count * 2);
  // Output capacity: grow-only buffers sized by the previous results; if the
  // lists do not fit, the pass still yields exact counts + total and is
  // repeated once with the exact size (first call / growing workloads only).
  size_t cap = count_only ? 0 : res->idx.cap;
  if (!count_only && want_dist) cap = std::min(cap, res->dist.cap);
  if (!count_only && cap == 0) {
    cap = (size_t)nq * 64;
    res->idx.ensure(cap, st, 0, 1.0);
    if (want_dist) res->dist.ensure(cap, st, 0, 1.0);
    cap = want_dist ? std::min(res->idx.cap, res->dist.cap) : res->idx.cap;
  }
  unsigned long long total = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    GridView g = t->view();
    RQ_CUDA(cudaMemsetAsync(res->cursor.p, 0, 4 * sizeof(unsigned long long), st));
    {
      PhaseScope p2(ctx, "range_fill");
      range_fused_kernel<D, QN><<<blocks, FUSED_WARPS * 32, smem, st>>>(
          g, dq, res->qorder.p, nq, r, dr, res->counts.p, res->offsets.p, res->idx.p,
          want_dist ? res->dist.p : nullptr, (unsigned long long)cap, res->cursor.p, count_only ? 0 : 1);
      post_launch(ctx);
    }
    RQ_CUDA(cudaMemcpyAsync(&total, res->cursor.p, sizeof(total), cudaMemcpyDeviceToHost, st));
    RQ_CUDA(cudaStreamSynchronize(st));
    if (count_only || total <= cap) break;
    RQ_REQUIRE(attempt == 0, "internal: range output did not fit after resizing");
    cap = (size_t)total + (size_t)(total / 32) + 1024;
    res->idx.ensure(cap, st, 0, 1.0);
    if (want_dist) res->dist.ensure(cap, st, 0, 1.0);
    cap = want_dist ? std::min(res->idx.cap, res->dist.cap) : res->idx.cap;
  }
  res->n_queries = nq;
  res->total = (int64_t)total;
  res->has_lists = !count_only;
  res->has_dist = !count_only && want_dist;
}

template <int D>
static void range_query_impl(rrtqx_tree *t, const double *queries, int64_t nq, double r, const double *ranges,
                             uint32_t flags, rrtqx_range_result *res) {
  rrtqx_ctx *ctx = t->ctx;
  cudaStream_t st = ctx->stream;
  tree_prepare_query(t);
  const double *dq = to_device(ctx, queries, (size_t)nq * D, ctx->stage_f64);
  const double *dr = ranges ? to_device(ctx, ranges, (size_t)nq, ctx->stage_f64b) : nullptr;
  PhaseScope ph(ctx, "range_query");
  res->counts.ensure((size_t)nq + 1, st);
  res->offsets.ensure((size_t)nq + 1, st);
  if (t->wrap.num_wraps == 0 && !getenv("RRTQX_RANGE_TWO_PASS")) {
    range_query_fused<D>(t, dq, dr, nq, r, flags, res);
    return;
  }
  {
    PhaseScope p2(ctx, "range_sort");
    sort_queries<D>(t, res, dq, nq);
  }
  {
    PhaseScope p2(ctx, "range_count");
    launch_range<D>(t, false, dq, res->qorder.p, nq, r, dr, res->counts.p, nullptr, nullptr, nullptr);
  }
  {
    PhaseScope p2(ctx, "range_scan");
    exclusive_scan<int32_t, int64_t>(ctx, res->counts.p, nq, res->offsets.p, res->scan_tmp64);
  }
  int64_t total = 0;
  RQ_CUDA(cudaMemcpyAsync(&total, res->offsets.p + nq, sizeof(int64_t), cudaMemcpyDeviceToHost, st));
  RQ_CUDA(cudaStreamSynchronize(st));
  res->n_queries = nq;
  res->total = total;
  res->has_dist = false;
  res->has_lists = false;
  if (flags & RRTQX_RANGE_COUNT_ONLY) return;
  const bool want_dist = flags & RRTQX_RANGE_WANT_DIST;
  res->idx.ensure((size_t)total + 1, st, 0, 1.0);
  if (want_dist) res->dist.ensure((size_t)total + 1, st, 0, 1.0);
  {
    PhaseScope p2(ctx, "range_fill");
    launch_range<D>(t, true, dq, res->qorder.p, nq, r, dr, res->counts.p, res->offsets.p, res->idx.p,
                    want_dist ? res->dist.p : nullptr);
  }
  res->has_dist = want_dist;
  res->has_lists = true;
}

void range_query(rrtqx_tree *t, const double *queries, int64_t nq, double r, const double *ranges, uint32_t flags,
                 rrtqx_range_result *res) {
  RQ_REQUIRE(nq >= 0 && nq < (int64_t)0x7fffffff, "n_queries out of range");
  if (t->n == 0) throw Error(RRTQX_ERR_EMPTY_TREE, "range query on an empty tree (reference dereferences an undefined root)");
  if (nq == 0) {
    res->n_queries = 0;
    res->total = 0;
    res->has_lists = true;
    res->has_dist = flags & RRTQX_RANGE_WANT_DIST;
    return;
  }
  switch (t->d) {
    case 2: range_query_impl<2>(t, queries, nq, r, ranges, flags, res); break;
    case 3: range_query_impl<3>(t, queries, nq, r, ranges, flags, res); break;
    case 4: range_query_impl<4>(t, queries, nq, r, ranges, flags, res); break;
    default: throw Error(RRTQX_ERR_UNSUPPORTED, "d must be 2, 3 or 4");
  }
}

// ---------------------------------------------------------- nearest kernel
// One warp per query; cube of cells of growing half-width rho around the
// query's cell until no unseen cell can hold a closer point.
template <int D, bool WRAP>
__global__ void __launch_bounds__(256)
nearest_kernel(GridView g, WrapInfo wrap, const double *__restrict__ queries, const int32_t *__restrict__ qorder,
               int64_t nq, int32_t *__restrict__ out_idx, double *__restrict__ out_dist) {
  const int lane = lane_id();
  const int64_t warp0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;

  for (int64_t qi = warp0; qi < nq; qi += nwarps) {
    const int qid = qorder[qi];
    double q0[D];
#pragma unroll
    for (int k = 0; k < D; ++k) q0[k] = queries[(int64_t)qid * D + k];

    // best over all identities so far (distance = sqrt of radicand; strict <
    // replacement, kdTree_general.jl:375-379)
    double best_dist = INFINITY;
    int best_node = -1;

    const int n_ident = WRAP ? (1 << wrap.num_wraps) : 1;
    for (int ident = 0; ident < n_ident; ++ident) {
      double q[D];
      if (ident == 0) {
#pragma unroll
        for (int k = 0; k < D; ++k) q[k] = q0[k];
      } else {
        if (!make_ghost<D>(wrap, q0, ident, best_dist, q)) continue;
      }
      // per-lane running minimum of the radicand for this identity
      double bs = INFINITY;
      int bn = 0x7fffffff;
      auto consider = [&](double px, double py, double pz, double pw, int node) {
        double s = sqdist<D>(q, px, py, pz, pw);
        if (s < bs || (s == bs && node < bn)) { bs = s; bn = node; }
      };
      auto warp_min = [&]() {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          double os = __shfl_xor_sync(FULL, bs, o);
          int on = __shfl_xor_sync(FULL, bn, o);
          if (os < bs || (os == bs && on < bn)) { bs = os; bn = on; }
        }
      };
      // tail first (also seeds the bound)
      for (int j = g.n_sorted + lane; j < g.n_total; j += 32) {
        double4 p = g.pos[j];
        consider(p.x, p.y, p.z, p.w, j);
      }
      if (g.n_sorted > 0) {
        const int cx = cell_of(q[0], g.lo[0], g.inv[0], g.nx);
        const int cy = cell_of(q[1], g.lo[1], g.inv[1], g.ny);
        const int cz = D >= 3 ? cell_of(q[2], g.lo[2], g.inv[2], g.nz) : 0;
        const int max_rho = max(g.nx, max(g.ny, g.nz));
        for (int rho = 0; rho <= max_rho; ++rho) {
          const int x0 = max(cx - rho, 0), x1 = min(cx + rho, g.nx - 1);
          const int y0 = max(cy - rho, 0), y1 = min(cy + rho, g.ny - 1);
          const int z0 = max(cz - rho, 0), z1 = min(cz + rho, g.nz - 1);
          for (int z = z0; z <= z1; ++z)
            for (int y = y0; y <= y1; ++y) {
              const int base = (z * g.ny + y) * g.nx;
              // shell only: interior rows contribute just their two end cells
              const bool edge_row = (rho == 0) || (y == cy - rho) || (y == cy + rho) ||
                                    (D >= 3 && ((z == cz - rho) || (z == cz + rho)));
              if (edge_row) {
                const int a = g.cell_start[base + x0], b = g.cell_start[base + x1 + 1];
                for (int j = a + lane; j < b; j += 32)
                  consider(g.sx[j], g.sy[j], D >= 3 ? g.sz[j] : 0.0, D >= 4 ? g.sw[j] : 0.0, g.sperm[j]);
              } else {
                if (cx - rho >= 0) {
                  const int a = g.cell_start[base + cx - rho], b = g.cell_start[base + cx - rho + 1];
                  for (int j = a + lane; j < b; j += 32)
                    consider(g.sx[j], g.sy[j], D >= 3 ? g.sz[j] : 0.0, D >= 4 ? g.sw[j] : 0.0, g.sperm[j]);
                }
                if (cx + rho <= g.nx - 1) {
                  const int a = g.cell_start[base + cx + rho], b = g.cell_start[base + cx + rho + 1];
                  for (int j = a + lane; j < b; j += 32)
                    consider(g.sx[j], g.sy[j], D >= 3 ? g.sz[j] : 0.0, D >= 4 ? g.sw[j] : 0.0, g.sperm[j]);
                }
              }
            }
          warp_min();
          // lower bound on the distance of any point outside the scanned cube
          // (faces at the grid boundary do not count: nothing lies beyond).
          double bound = INFINITY;
          {
            double f;
            if (cx - rho > 0)          { f = q[0] - (g.lo[0] + (cx - rho) * g.cell[0]); bound = fmin(bound, f); }
            if (cx + rho < g.nx - 1)   { f = (g.lo[0] + (cx + rho + 1) * g.cell[0]) - q[0]; bound = fmin(bound, f); }
            if (cy - rho > 0)          { f = q[1] - (g.lo[1] + (cy - rho) * g.cell[1]); bound = fmin(bound, f); }
            if (cy + rho < g.ny - 1)   { f = (g.lo[1] + (cy + rho + 1) * g.cell[1]) - q[1]; bound = fmin(bound, f); }
            if (D >= 3) {
              if (cz - rho > 0)        { f = q[2] - (g.lo[2] + (cz - rho) * g.cell[2]); bound = fmin(bound, f); }
              if (cz + rho < g.nz - 1) { f = (g.lo[2] + (cz + rho + 1) * g.cell[2]) - q[2]; bound = fmin(bound, f); }
            }
          }
          if (isinf(bound) && bound > 0) break;  // whole grid scanned
          // conservative: shrink the bound by the cell-geometry slack
          double slack = 1e-9 * fmax(g.cell[0], fmax(g.cell[1], g.cell[2])) +
                         1e-12 * (fabs(q[0]) + fabs(q[1]) + (D >= 3 ? fabs(q[2]) : 0.0));
          double bd = bound - slack;
          if (bd > 0.0 && bd * bd * (1.0 - 1e-9) > bs) break;
        }
      } else {
        warp_min();
      }
      if (g.n_sorted > 0 && g.n_total > g.n_sorted) warp_min();
      double dd = __dsqrt_rn(bs);
      if (bn != 0x7fffffff && (best_node < 0 || dd < best_dist)) {
        best_dist = dd;
        best_node = bn;
      }
    }
    if (lane == 0) {
      out_idx[qid] = best_node;
      if (out_dist) out_dist[qid] = best_dist;
    }
  }
}

struct NearestScratch {
  rrtqx_range_result sortbuf;  // reuses the query sort buffers
  DevBuf<int32_t> idx;
  DevBuf<double> dist;
};

template <int D>
static void nearest_impl(rrtqx_tree *t, rrtqx_range_result *sortbuf, const double *queries, int64_t nq,
                         int32_t *idx_out, double *dist_out, DevBuf<int32_t> &idx_stage, DevBuf<double> &dist_stage) {
  rrtqx_ctx *ctx = t->ctx;
  cudaStream_t st = ctx->stream;
  tree_prepare_query(t);
  const double *dq = to_device(ctx, queries, (size_t)nq * D, ctx->stage_f64);
  const bool idx_dev = is_device_ptr(idx_out), dist_dev = dist_out && is_device_ptr(dist_out);
  int32_t *didx = idx_out;
  double *ddist = dist_out;
  if (!idx_dev) { idx_stage.ensure((size_t)nq, st); didx = idx_stage.p; }
  if (dist_out && !dist_dev) { dist_stage.ensure((size_t)nq, st); ddist = dist_stage.p; }
  {
    PhaseScope ph(ctx, "nearest");
    sort_queries<D>(t, sortbuf, dq, nq);
    GridView g = t->view();
    const int TB = 256;
    int blocks = (int)std::min<int64_t>((nq * 32 + TB - 1) / TB, (int64_t)ctx->sm_count * 8);
    if (blocks < 1) blocks = 1;
    if (t->wrap.num_wraps > 0)
      nearest_kernel<D, true><<<blocks, TB, 0, st>>>(g, t->wrap, dq, sortbuf->qorder.p, nq, didx, ddist);
    else
      nearest_kernel<D, false><<<blocks, TB, 0, st>>>(g, t->wrap, dq, sortbuf->qorder.p, nq, didx, ddist);
    post_launch(ctx);
  }
  if (!idx_dev) from_device(ctx, idx_out, didx, (size_t)nq);
  if (dist_out && !dist_dev) from_device(ctx, dist_out, ddist, (size_t)nq);
  RQ_CUDA(cudaStreamSynchronize(st));
}

void nearest_query(rrtqx_tree *t, rrtqx_range_result *sortbuf, const double *queries, int64_t nq, int32_t *idx_out,
                   double *dist_out, DevBuf<int32_t> &idx_stage, DevBuf<double> &dist_stage) {
  RQ_REQUIRE(nq >= 0 && nq < (int64_t)0x7fffffff, "n_queries out of range");
  RQ_REQUIRE(idx_out != nullptr || nq == 0, "idx_out is NULL");
  if (t->n == 0) throw Error(RRTQX_ERR_EMPTY_TREE, "nearest query on an empty tree (reference dereferences an undefined root)");
  if (nq == 0) return;
  switch (t->d) {
    case 2: nearest_impl<2>(t, sortbuf, queries, nq, idx_out, dist_out, idx_stage, dist_stage); break;
    case 3: nearest_impl<3>(t, sortbuf, queries, nq, idx_out, dist_out, idx_stage, dist_stage); break;
    case 4: nearest_impl<4>(t, sortbuf, queries, nq, idx_out, dist_out, idx_stage, dist_stage); break;
    default: throw Error(RRTQX_ERR_UNSUPPORTED, "d must be 2, 3 or 4");
  }
}

}  // namespace rrtqx
