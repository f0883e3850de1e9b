// collide_queue.cuh -- warp-queue form of "does edge i collide with any binned sphere obstacle"
// (explicitEdgeCheck over the obstacle list, DRRT_Q.jl:1775-1826; the edge part of addNewObstacle,
// DRRT_Q.jl:3220-3290), used by the large-batch edge check and the edge-centric add sweep.
//
// The thread-per-edge kernels ran the long exact FP64 test (2 sqrt + 1 div, DRRT_Q.jl:1205-1210) under
// divergence: any lane that survived the cheap reject made the whole warp wait.  Here the two stages are
// separated.  Stage A: each lane walks the grid cells around ITS edge and applies the FP32 conservative
// reject; surviving (edge, obstacle) pairs are ballot-compacted into a per-warp queue in shared memory.
// Stage B: whenever 32 pairs are queued, all 32 lanes run one exact test each (the edge invariants are
// recomputed from the endpoints: cheaper than carrying 7 doubles through shared memory per pair).  A warp
// works through CQ_BATCHES batches of 32 edges and carries the queue across them, so only its last drain
// is partial.  The result is an OR over pairs, so the order of tests does not matter; an edge already
// known to collide may be tested again (no early exit across queued pairs), which only costs time.
//
// Cover lists: edges whose half length is at most G.cov_cap (half a cover cell) do not walk the coarse rows;
// they read the list of the one cover cell that holds their midpoint.  Obstacle o is on the list of cell c iff
// dist(c_o, box(c)) <= thr_o + cov_cap (+ margin), which contains every obstacle with |c_o - mid| <= thr_o +
// half for any midpoint in the cell; border cells extend to infinity, so midpoints outside the grid are
// covered too.  C3: about 1 candidate per edge instead of 16.  Longer edges and over-budget obstacle sets
// fall back to the coarse rows; degenerate edges meet the whole table.
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
};

// One warp per binned obstacle; FILL = false counts the cells it belongs to, FILL = true writes the lists.
template <bool FILL>
__global__ void __launch_bounds__(256)
cover_register_kernel(const double4 *__restrict__ rec2, const double2 *__restrict__ thr2, const int32_t *__restrict__ cstart,
                      const SphGrid *__restrict__ Gp, int n_upper, int32_t *__restrict__ cnt,
                      const int32_t *__restrict__ start, int32_t *__restrict__ list) {
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
      if (FILL) list[start[cell] + p] = o;
    }
  }
}

static __global__ void cover_finalize_kernel(SphGrid *G, const int32_t *__restrict__ start) {
  if (start[COV_CELLS] > COV_BUDGET) G->cov_on = 0;
}

// After sphere_grid_kernel(..., cover = 1) on the same stream.  n_upper: upper bound of the table size.
static inline void build_sphere_cover(rrtqx_ctx *ctx, SphCoverBufs &B, const double4 *rec2, const double2 *thr2,
                                      const int32_t *cstart, SphGrid *dG, int n_upper) {
  cudaStream_t st = ctx->stream;
  B.cnt.ensure(COV_CELLS + 1, st);
  B.start.ensure(COV_CELLS + 2, st);
  B.list.ensure(COV_BUDGET, st);
  RQ_CUDA(cudaMemsetAsync(B.cnt.p, 0, (COV_CELLS + 1) * sizeof(int32_t), st));
  const unsigned blocks = (unsigned)div_up((int64_t)n_upper * 32, (int64_t)256);
  cover_register_kernel<false><<<blocks, 256, 0, st>>>(rec2, thr2, cstart, dG, n_upper, B.cnt.p, nullptr, nullptr);
  exclusive_scan<int32_t, int32_t>(ctx, B.cnt.p, COV_CELLS, B.start.p, B.scan_tmp);
  cover_finalize_kernel<<<1, 1, 0, st>>>(dG, B.start.p);
  RQ_CUDA(cudaMemsetAsync(B.cnt.p, 0, (COV_CELLS + 1) * sizeof(int32_t), st));
  cover_register_kernel<true><<<blocks, 256, 0, st>>>(rec2, thr2, cstart, dG, n_upper, B.cnt.p, B.start.p, B.list.p);
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

// ---------------------------------------------------------------- warp queue
constexpr int CQ_BATCHES = 8;  // batches of 32 edges per warp
constexpr int CQ_CAP = 64;     // queue entries per warp: < 32 carried + <= 32 pushed per step

// Src supplies the items:
//   bool endpoints(int64_t i, double a[3], double b[3], int &v)   false: item has no edge (node without parent)
//   void clear(int64_t i)                                           before any test of item i
//   bool accept(int o, const double4 &rec, const double a[3], int v)  extra condition on a colliding pair
//   void mark(int64_t i)                                            some accepted obstacle collides with item i
template <bool FMA_DOT, class Src>
__device__ __forceinline__ void cq_run(const Src &S, int64_t n_items, const SphGrid &G,
                                       const double4 *__restrict__ rec, const double2 *__restrict__ thr,
                                       const float4 *__restrict__ frec, const int32_t *__restrict__ cstart,
                                       const int32_t *__restrict__ cov_start, const int32_t *__restrict__ cov_list,
                                       int2 *queue /* this warp's CQ_CAP entries in shared memory */) {
  const int lane = threadIdx.x & 31;
  const unsigned lt = lanemask_lt();
  const int64_t warp_id = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t base = warp_id * (32 * CQ_BATCHES);
  if (base >= n_items) return;  // the whole warp
  const int ncell = G.nx * G.ny * G.nz;
  int qn = 0;

  auto drain = [&](int first, int cnt) {
    if (lane < cnt) {
      const int2 en = queue[first + lane];
      const int64_t i = base + en.x;
      double a[3], b[3];
      int v;
      S.endpoints(i, a, b, v);
      const SegPre pre = seg_prepare(a[0], a[1], a[2], b[0], b[1], b[2]);
      const double4 r = rec[en.y];
      if (seg_sphere_collide_exact<FMA_DOT>(pre, r.x, r.y, r.z, thr[en.y].y) && S.accept(en.y, r, a, v)) S.mark(i);
    }
  };

  for (int bt = 0; bt < CQ_BATCHES; ++bt) {
    const int rel = bt * 32 + lane;
    const int64_t i = base + rel;
    if (base + bt * 32 >= n_items) break;  // warp-uniform
    // per-lane iterator over the candidate obstacles of the edge: the cover list of the midpoint's cell for a
    // short edge, else rows (y, z) of coarse cells x0..x1; then the "always tested" bucket.  A degenerate edge
    // meets the whole table (the reference collides it with every active obstacle, and the FP32 reject never
    // fires for it)
    int o = 0, oend = 0, x0 = 0, x1 = 0, y = 0, y0 = 0, y1 = -1, z = 0, z1 = -1;
    int stage = 3;  // 0 coarse rows, 1 always bucket, 2 whole table, 3 done, 4 cover list
    bool ind = false;
    SegF32 sf;
    sf.ok = false;
    sf.mx = sf.my = sf.mz = sf.half = sf.bound = 0.0f;
    {
      double a[3], b[3];
      int v;
      if (i < n_items && S.endpoints(i, a, b, v)) {
        S.clear(i);
        const SegPre pre = seg_prepare(a[0], a[1], a[2], b[0], b[1], b[2]);
        sf = seg_f32(pre, G.cmax);
        if (!pre.cullable) {
          stage = 2;
        } else if (G.cov_on && pre.half <= G.cov_cap) {
          x0 = (sg_cell(pre.mz, G.clo[2], G.cinv[2], COV_DIM) * COV_DIM + sg_cell(pre.my, G.clo[1], G.cinv[1], COV_DIM)) * COV_DIM +
               sg_cell(pre.mx, G.clo[0], G.cinv[0], COV_DIM);
          stage = 4;
        } else {
          const double R = (pre.half + G.thr_max) * (1.0 + 1e-9) + 1e-300;
          x0 = sg_cell(pre.mx - R, G.lo[0], G.inv[0], G.nx); x1 = sg_cell(pre.mx + R, G.lo[0], G.inv[0], G.nx);
          y0 = sg_cell(pre.my - R, G.lo[1], G.inv[1], G.ny); y1 = sg_cell(pre.my + R, G.lo[1], G.inv[1], G.ny);
          z = sg_cell(pre.mz - R, G.lo[2], G.inv[2], G.nz);  z1 = sg_cell(pre.mz + R, G.lo[2], G.inv[2], G.nz);
          y = y0;
          stage = 0;
        }
      }
    }
    while (true) {
      while (o >= oend && stage != 3) {  // next non-empty row of this lane
        if (stage == 4) {
          o = cov_start[x0];
          oend = cov_start[x0 + 1];
          ind = true;
          stage = 1;
        } else if (stage == 0) {
          const int cb = (z * G.ny + y) * G.nx;
          o = cstart[cb + x0];
          oend = cstart[cb + x1 + 1];
          if (++y > y1) { y = y0; if (++z > z1) stage = 1; }
        } else if (stage == 1) {
          o = cstart[ncell];
          oend = cstart[ncell + 1];
          ind = false;
          stage = 3;
        } else {
          o = 0;
          oend = G.n_total;
          stage = 3;
        }
      }
      const bool have = o < oend;
      if (!__any_sync(FULL, have)) break;
      bool keep = false;
      int oo = 0;
      if (have) {
        oo = ind ? cov_list[o] : o;
        keep = !seg_reject_f32(sf, frec[oo]);
        ++o;
      }
      const unsigned m = __ballot_sync(FULL, keep);
      if (m) {
        if (keep) queue[qn + __popc(m & lt)] = make_int2(rel, oo);
        qn += __popc(m);
        if (qn >= 32) {
          __syncwarp();
          qn -= 32;
          drain(qn, 32);
          __syncwarp();
        }
      }
    }
  }
  __syncwarp();
  if (qn) drain(0, qn);
}

#endif  // __CUDACC__

}  // namespace rrtqx
