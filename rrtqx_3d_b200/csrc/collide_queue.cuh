// collide_queue.cuh -- two-stage form of "does edge i collide with any binned sphere obstacle"
// (explicitEdgeCheck over the obstacle list, DRRT_Q.jl:1775-1826; the edge part of addNewObstacle,
// DRRT_Q.jl:3220-3290), used by the large-batch edge check and the edge-centric add sweep.
//
// The thread-per-edge kernels ran the long exact FP64 test (2 sqrt + 1 div, DRRT_Q.jl:1205-1210) under
// divergence: any lane that survived the cheap reject made the whole warp wait.  Here the stages are separate
// kernels.  pq_collect_kernel: one thread per edge finds the candidate obstacles of ITS edge and applies the
// FP32 conservative reject; surviving (edge, obstacle) pairs go through a per-block queue in shared memory to
// a global pair list (one atomic per block).  pq_test_kernel: one thread per pair runs the exact test (edge
// invariants recomputed from the endpoints) and marks the edge -- every lane busy, high occupancy.  The result
// is an OR over pairs, so order and duplicates do not matter.  Edges the scheme does not fit -- degenerate ones
// (no reject possible: the reference collides them with every active obstacle) and edges with more than
// PQ_KEEP surviving candidates -- are listed as items and decided by pq_slow_kernel with early exit.
//
// Cover lists: edges whose half length is at most G.cov_cap (half a cover cell) do not walk the coarse rows;
// they read the list of the one cover cell that holds their midpoint.  Obstacle o is on the list of cell c iff
// dist(c_o, box(c)) <= thr_o + cov_cap (+ margin), which contains every obstacle with |c_o - mid| <= thr_o +
// half for any midpoint in the cell; border cells extend to infinity, so midpoints outside the grid are
// covered too.  C3: about 1 candidate per edge instead of 16.  Longer edges and over-budget obstacle sets
// fall back to the coarse rows.
#pragma once
#include "collision.cuh"
#include "scan.cuh"
#include <cstdlib>
#include <map>

namespace rrtqx {

#ifdef __CUDACC__

// ---------------------------------------------------------------- cover lists
struct SphCoverBufs {
  DevBuf<int32_t> cnt, start, list, scan_tmp;
  DevBuf<float4> list_f;                // FP32 reject record of each list entry (saves one dependent load)
  DevBuf<uint2> pairs;                  // (item, obstacle) pairs that survived the reject
  DevBuf<unsigned> slow;                // items decided by the slow kernel
  DevBuf<unsigned long long> n_pairs;   // [0] pairs, [1] slow items
};

// One warp per binned obstacle; FILL = false counts the cells it belongs to, FILL = true writes the lists.
template <bool FILL>
__global__ void __launch_bounds__(256)
cover_register_kernel(const double4 *__restrict__ rec2, const double2 *__restrict__ thr2, const int32_t *__restrict__ cstart,
                      const SphGrid *__restrict__ Gp, int n_upper, int32_t *__restrict__ cnt,
                      const int32_t *__restrict__ start, int32_t *__restrict__ list, const float4 *__restrict__ frec2,
                      float4 *__restrict__ list_f) {
  const int o = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5), lane = threadIdx.x & 31;
  if (o >= n_upper) return;
  const SphGrid &G = *Gp;
  if (!G.cov_on) return;
  if (o >= cstart[G.nx * G.ny * G.nz]) return;  // the non-finite ones stay in the "always" bucket
  const double4 r = rec2[o];
  const double c[3] = {r.x, r.y, r.z};
  const double rho = (thr2[o].x + G.cov_cap) * (1.0 + 1e-9) + G.cov_margin;
  int a[3], n[3];
  for (int d = 0; d < 3; ++d) {
    a[d] = sg_cell(c[d] - rho, G.clo[d], G.cinv[d], COV_DIM);
    n[d] = sg_cell(c[d] + rho, G.clo[d], G.cinv[d], COV_DIM) - a[d] + 1;
  }
  const int total = n[0] * n[1] * n[2];
  for (int k = lane; k < total; k += 32) {
    const int i[3] = {a[0] + k % n[0], a[1] + (k / n[0]) % n[1], a[2] + k / (n[0] * n[1])};
    double d2 = 0.0;
    for (int d = 0; d < 3; ++d) {
      const double L = i[d] == 0 ? -INFINITY : G.clo[d] + i[d] * G.ccell[d];
      const double H = i[d] == COV_DIM - 1 ? INFINITY : G.clo[d] + (i[d] + 1) * G.ccell[d];
      const double g = fmax(fmax(L - c[d], c[d] - H), 0.0);
      d2 += g * g;
    }
    if (d2 <= rho * rho) {
      const int cell = (i[2] * COV_DIM + i[1]) * COV_DIM + i[0];
      const int p = atomicAdd(&cnt[cell], 1);
      if (FILL) {
        list[start[cell] + p] = o;
        list_f[start[cell] + p] = frec2[o];
      }
    }
  }
}

static __global__ void cover_finalize_kernel(SphGrid *G, const int32_t *__restrict__ start) {
  if (start[COV_CELLS] > COV_BUDGET) G->cov_on = 0;
}

// After sphere_grid_kernel(..., cover = 1) on the same stream.  n_upper: upper bound of the table size.
static inline void build_sphere_cover(rrtqx_ctx *ctx, SphCoverBufs &B, const double4 *rec2, const double2 *thr2,
                                      const float4 *frec2, const int32_t *cstart, SphGrid *dG, int n_upper) {
  cudaStream_t st = ctx->stream;
  B.cnt.ensure(COV_CELLS + 1, st);
  B.start.ensure(COV_CELLS + 2, st);
  B.list.ensure(COV_BUDGET, st);
  B.list_f.ensure(COV_BUDGET, st);
  RQ_CUDA(cudaMemsetAsync(B.cnt.p, 0, (COV_CELLS + 1) * sizeof(int32_t), st));
  const unsigned blocks = (unsigned)div_up((int64_t)n_upper * 32, (int64_t)256);
  cover_register_kernel<false><<<blocks, 256, 0, st>>>(rec2, thr2, cstart, dG, n_upper, B.cnt.p, nullptr, nullptr, nullptr, nullptr);
  exclusive_scan<int32_t, int32_t>(ctx, B.cnt.p, COV_CELLS, B.start.p, B.scan_tmp);
  cover_finalize_kernel<<<1, 1, 0, st>>>(dG, B.start.p);
  RQ_CUDA(cudaMemsetAsync(B.cnt.p, 0, (COV_CELLS + 1) * sizeof(int32_t), st));
  cover_register_kernel<true><<<blocks, 256, 0, st>>>(rec2, thr2, cstart, dG, n_upper, B.cnt.p, B.start.p, B.list.p, frec2, B.list_f.p);
  post_launch(ctx, 3);
}

static inline SphCoverBufs &cover_bufs(rrtqx_ctx *ctx) {
  static std::map<rrtqx_ctx *, SphCoverBufs *> m;  // per context and translation unit, leaked at exit by design
  auto it = m.find(ctx);
  if (it == m.end()) it = m.emplace(ctx, new SphCoverBufs()).first;
  return *it->second;
}

// Batches below this many items keep the thread-per-edge kernels (the cover build is ~10 small launches).
static inline int64_t cover_min_items() {
  static const int64_t v = [] {
    const char *e = getenv("RRTQX_COVER_MIN_ITEMS");
    return e ? (int64_t)atoll(e) : (int64_t)16384;
  }();
  return getenv("RRTQX_EDGE_NO_QUEUE") ? INT64_MAX : v;
}

// ---------------------------------------------------------------- pair queue
constexpr int PQ_THREADS = 128;  // collect kernel block
constexpr int PQ_KEEP = 2;       // pairs an item may put on the list; items with more go to the slow list

// Src supplies the items:
//   bool endpoints(int64_t i, double a[3], double b[3], int &v)   false: item has no edge (node without parent)
//   void clear(int64_t i)                                           before any test of item i
//   bool accept(int o, const double4 &rec, const double a[3], int v)  extra condition on a colliding pair
//   void mark(int64_t i)                                            some accepted obstacle collides with item i
//
// Stage 1: one thread per item.  Candidates that survive the FP32 reject are kept in registers (at most
// PQ_KEEP); at the end the thread publishes them as pairs through the block queue, or -- degenerate edge
// (no reject possible, the reference collides it with every active obstacle) or more than PQ_KEEP survivors
// -- puts the ITEM on the slow list.  No exact arithmetic here, so the kernel stays small; the lists cannot
// overflow (<= PQ_KEEP pairs and <= 1 slow entry per item).
template <class Src>
__global__ void __launch_bounds__(PQ_THREADS)
pq_collect_kernel(Src S, int64_t n_items, const float4 *__restrict__ frec, const int32_t *__restrict__ cstart,
                  const int32_t *__restrict__ cov_start, const int32_t *__restrict__ cov_list,
                  const float4 *__restrict__ cov_frec, const SphGrid *__restrict__ Gp, uint2 *__restrict__ pairs,
                  unsigned *__restrict__ slow, unsigned long long *__restrict__ counters /* [0] pairs, [1] slow items */) {
  __shared__ SphGrid G;
  __shared__ uint2 q[PQ_THREADS * PQ_KEEP];
  __shared__ unsigned sq[PQ_THREADS];
  __shared__ unsigned qn, sn;
  __shared__ unsigned long long qbase, sbase;
  if (threadIdx.x == 0) { G = *Gp; qn = 0; sn = 0; }
  __syncthreads();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  int nk = 0;
  int keep[PQ_KEEP];
#pragma unroll
  for (int k = 0; k < PQ_KEEP; ++k) keep[k] = 0;
  bool to_slow = false;
  {
    double a[3], b[3];
    int v;
    if (i < n_items && S.endpoints(i, a, b, v)) {
      S.clear(i);
      const SegPre pre = seg_prepare(a[0], a[1], a[2], b[0], b[1], b[2]);
      if (!pre.cullable) {
        to_slow = true;
      } else {
        const SegF32 sf = seg_f32(pre, G.cmax);
        auto visit = [&](int o, const float4 f) {
          if (seg_reject_f32(sf, f)) return;
#pragma unroll
          for (int k = 0; k < PQ_KEEP; ++k)
            if (nk == k) keep[k] = o;
          ++nk;
        };
        const int ncell = G.nx * G.ny * G.nz;
        if (G.cov_on && pre.half <= G.cov_cap) {
          const int c = (sg_cell(pre.mz, G.clo[2], G.cinv[2], COV_DIM) * COV_DIM + sg_cell(pre.my, G.clo[1], G.cinv[1], COV_DIM)) * COV_DIM +
                        sg_cell(pre.mx, G.clo[0], G.cinv[0], COV_DIM);
          const int k1 = cov_start[c + 1];
          for (int k = cov_start[c]; k < k1; ++k) visit(cov_list[k], cov_frec[k]);
        } else {
          const double R = (pre.half + G.thr_max) * (1.0 + 1e-9) + 1e-300;
          const int x0 = sg_cell(pre.mx - R, G.lo[0], G.inv[0], G.nx), x1 = sg_cell(pre.mx + R, G.lo[0], G.inv[0], G.nx);
          const int y0 = sg_cell(pre.my - R, G.lo[1], G.inv[1], G.ny), y1 = sg_cell(pre.my + R, G.lo[1], G.inv[1], G.ny);
          const int z0 = sg_cell(pre.mz - R, G.lo[2], G.inv[2], G.nz), z1 = sg_cell(pre.mz + R, G.lo[2], G.inv[2], G.nz);
          for (int z = z0; z <= z1; ++z)
            for (int y = y0; y <= y1; ++y) {
              const int cb = (z * G.ny + y) * G.nx;
              const int o1 = cstart[cb + x1 + 1];
              for (int o = cstart[cb + x0]; o < o1; ++o) visit(o, frec[o]);
            }
        }
        const int o1 = cstart[ncell + 1];
        for (int o = cstart[ncell]; o < o1; ++o) visit(o, frec[o]);  // non-finite obstacles: met by every edge
        to_slow = nk > PQ_KEEP;
      }
    }
  }
  if (to_slow) {
    sq[atomicAdd(&sn, 1u)] = (unsigned)i;
  } else if (nk) {
    const unsigned k0 = atomicAdd(&qn, (unsigned)nk);
#pragma unroll
    for (int k = 0; k < PQ_KEEP; ++k)
      if (k < nk) q[k0 + k] = make_uint2((unsigned)i, (unsigned)keep[k]);
  }
  __syncthreads();
  const unsigned n = qn, ns = sn;
  if (threadIdx.x == 0 && n) qbase = atomicAdd(&counters[0], (unsigned long long)n);
  if (threadIdx.x == 32 && ns) sbase = atomicAdd(&counters[1], (unsigned long long)ns);
  __syncthreads();
  for (unsigned k = threadIdx.x; k < n; k += blockDim.x) pairs[qbase + k] = q[k];
  if (threadIdx.x < ns) slow[sbase + threadIdx.x] = sq[threadIdx.x];
}

// Stage 2: one thread per pair, exact test (edge invariants recomputed from the endpoints).
template <bool FMA_DOT, class Src>
__global__ void __launch_bounds__(256)
pq_test_kernel(Src S, const double4 *__restrict__ rec, const double2 *__restrict__ thr, const uint2 *__restrict__ pairs,
               const unsigned long long *__restrict__ counters) {
  const unsigned long long n = counters[0];
  for (unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; p < n;
       p += (unsigned long long)gridDim.x * blockDim.x) {
    const uint2 e = pairs[p];
    double a[3], b[3];
    int v;
    S.endpoints((int64_t)e.x, a, b, v);
    const SegPre pre = seg_prepare(a[0], a[1], a[2], b[0], b[1], b[2]);
    const double4 r = rec[e.y];
    if (seg_sphere_collide_exact<FMA_DOT>(pre, r.x, r.y, r.z, thr[e.y].y) && S.accept((int)e.y, r, a, v)) S.mark((int64_t)e.x);
  }
}

// Slow items: one thread per item, every candidate decided in place, first hit ends it (the shape of the
// thread-per-edge kernels; few items get here).
template <bool FMA_DOT, class Src>
__global__ void __launch_bounds__(128)
pq_slow_kernel(Src S, const double4 *__restrict__ rec, const double2 *__restrict__ thr, const float4 *__restrict__ frec,
               const int32_t *__restrict__ cstart, const SphGrid *__restrict__ Gp, const unsigned *__restrict__ slow,
               const unsigned long long *__restrict__ counters) {
  __shared__ SphGrid G;
  if (threadIdx.x == 0) G = *Gp;
  __syncthreads();
  const unsigned long long n = counters[1];
  for (unsigned long long p = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; p < n;
       p += (unsigned long long)gridDim.x * blockDim.x) {
    const int64_t i = (int64_t)slow[p];
    double a[3], b[3];
    int v;
    S.endpoints(i, a, b, v);
    const SegPre pre = seg_prepare(a[0], a[1], a[2], b[0], b[1], b[2]);
    const SegF32 sf = seg_f32(pre, G.cmax);
    bool hit = false;
    auto run = [&](int lo, int hi) {
      for (int o = lo; o < hi && !hit; ++o) {
        if (seg_reject_f32(sf, frec[o])) continue;  // never rejects a degenerate edge (sf.ok false)
        const double4 r = rec[o];
        hit = seg_sphere_collide_exact<FMA_DOT>(pre, r.x, r.y, r.z, thr[o].y) && S.accept(o, r, a, v);
      }
    };
    const int ncell = G.nx * G.ny * G.nz;
    if (!pre.cullable) {
      run(0, G.n_total);
    } else {
      const double R = (pre.half + G.thr_max) * (1.0 + 1e-9) + 1e-300;
      const int x0 = sg_cell(pre.mx - R, G.lo[0], G.inv[0], G.nx), x1 = sg_cell(pre.mx + R, G.lo[0], G.inv[0], G.nx);
      const int y0 = sg_cell(pre.my - R, G.lo[1], G.inv[1], G.ny), y1 = sg_cell(pre.my + R, G.lo[1], G.inv[1], G.ny);
      const int z0 = sg_cell(pre.mz - R, G.lo[2], G.inv[2], G.nz), z1 = sg_cell(pre.mz + R, G.lo[2], G.inv[2], G.nz);
      for (int z = z0; z <= z1 && !hit; ++z)
        for (int y = y0; y <= y1 && !hit; ++y) {
          const int cb = (z * G.ny + y) * G.nx;
          run(cstart[cb + x0], cstart[cb + x1 + 1]);
        }
      if (!hit) run(cstart[ncell], cstart[ncell + 1]);
    }
    if (hit) S.mark(i);
  }
}

// The three kernels on the context's stream.  Requires n_items < 2^32 (32-bit item numbers in the lists).
template <bool FMA_DOT, class Src>
static inline void pq_launch(rrtqx_ctx *ctx, SphCoverBufs &B, const Src &S, int64_t n_items, const double4 *rec,
                             const double2 *thr, const float4 *frec, const int32_t *cstart, const SphGrid *dG) {
  cudaStream_t st = ctx->stream;
  B.pairs.ensure((size_t)n_items * PQ_KEEP + 1, st);
  B.slow.ensure((size_t)n_items + 1, st);
  B.n_pairs.ensure(2, st);
  RQ_CUDA(cudaMemsetAsync(B.n_pairs.p, 0, 2 * sizeof(unsigned long long), st));
  pq_collect_kernel<Src><<<(unsigned)div_up(n_items, (int64_t)PQ_THREADS), PQ_THREADS, 0, st>>>(S, n_items, frec, cstart, B.start.p, B.list.p,
                                                                                                  B.list_f.p, dG, B.pairs.p, B.slow.p, B.n_pairs.p);
  const unsigned tblocks = (unsigned)std::min<int64_t>(div_up(n_items * PQ_KEEP, (int64_t)256), (int64_t)ctx->sm_count * 8);
  pq_test_kernel<FMA_DOT, Src><<<tblocks, 256, 0, st>>>(S, rec, thr, B.pairs.p, B.n_pairs.p);
  const unsigned sblocks = (unsigned)std::min<int64_t>(div_up(n_items, (int64_t)128), (int64_t)ctx->sm_count * 4);
  pq_slow_kernel<FMA_DOT, Src><<<sblocks, 128, 0, st>>>(S, rec, thr, frec, cstart, dG, B.slow.p, B.n_pairs.p);
  post_launch(ctx, 3);
}

#endif  // __CUDACC__

}  // namespace rrtqx
