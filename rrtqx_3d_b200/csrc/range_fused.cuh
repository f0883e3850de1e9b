// range_fused.cuh -- single-pass batched kdFindWithinRange (included by range.cu,
// inside namespace rrtqx).
//
// One warp per group of QN = 2 consecutive (cell-sorted) queries:
//   * the candidate region is the set of grid rows (runs of cells along x)
//     that can hold a point within r; it is computed conservatively in FP32
//     cell units (margins far above the rounding error), never deciding a
//     result -- membership is always the exact FP64 test s < T_lt(r);
//   * rows are short (fine cells give tight culling), so they are flattened
//     into a shared-memory table of "octets" (start slot, 1..8 valid points)
//     and each 8-lane sub-group takes one octet per trip: every sub-group is
//     busy whatever the row lengths are; 4 trips of loads are in flight before
//     the first test;
//   * candidates are read from the SoA arrays (an octet = one contiguous 64-byte run per coordinate);
//   * every candidate is loaded once and tested against both queries;
//   * hits are recorded as slot numbers in a per-warp shared-memory buffer;
//     when the scan ends the warp reserves the exact output range with ONE
//     atomic on a global cursor and flushes densely: coalesced 32-wide
//     stores, sqrt evaluated on hits only with all lanes busy.
// The tree is traversed once (no count pass).  Blocks own chunks of consecutive
// sorted queries (supercell-major order) so their warps share candidate rows in L1.

constexpr int FUSED_TAB = 256;   // octet-table entries per warp
constexpr int FUSED_ROUNDS = 4;  // groups each warp takes from one block-level chunk
#ifndef FUSED_U
#define FUSED_U 4
#endif

__device__ __forceinline__ void sts32(unsigned addr, int v) {
  asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ int lds32(unsigned addr) {
  int v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
  return v;
}
// keeps a value in a register (stops the compiler from re-deriving it from
// special registers inside the hot loop)
__device__ __forceinline__ unsigned pin_reg(unsigned v) {
  unsigned r;
  asm volatile("mov.u32 %0, %1;" : "=r"(r) : "r"(v));
  return r;
}

#ifdef RRTQX_LEGACY
template <int D, int QN>
struct Group {
  double q[QN][D];
  double T[QN];     // strict threshold on the radicand: hit <=> s < T
  float ff[QN][3];  // query position in (fractional) cell coordinates
  float r2f[QN];    // inflated r^2 (FP32, conservative)
  float rf[QN];     // inflated r
  int qid[QN];
  bool live[QN];    // query exists and r > 0
};

#endif

struct FGrid {
  float inv[3], cell[3];
};

__device__ __forceinline__ int clampi(float v, int n) {  // floor + clamp to [0, n-1]; NaN -> 0
  return (int)fminf(fmaxf(floorf(v), 0.0f), (float)(n - 1));
}

// Conservative cell range [cxa, cxb] of row (cy,cz) that can hold a point within r of the query at
// ff (cell units); false if the row is out of reach.  (A single evaluation for the bounding box of
// the group's queries was tried: sphere (+) box inflates the region by 30-60 % -- slower.)
template <int D>
__device__ __forceinline__ bool row_cells(const GridView &g, const FGrid &fg, const float *ff, float r2f, int cy, int cz,
                                          int &cxa, int &cxb) {
  const float *fflo = ff, *ffhi = ff;
  // lower bounds of |p.y - q.y|, |p.z - q.z| in cell units; boundary rows also
  // hold the points clamped in from beyond the grid, so they give no bound on
  // that side.  1e-3 cells of slack covers the FP32 error of ff (coordinates
  // up to 1024 cells: ulp 6e-5) and the FP64 rounding of cell_of().
  const float ninf = -INFINITY;
  float dyc = fmaxf(cy == 0 ? ninf : (float)cy - ffhi[1], cy == g.ny - 1 ? ninf : fflo[1] - (float)(cy + 1));
  dyc = fmaxf(dyc - 1e-3f, 0.0f);
  const float dy = dyc * fg.cell[1];
  float rem = r2f - dy * dy;
  if (D >= 3) {
    float dzc = fmaxf(cz == 0 ? ninf : (float)cz - ffhi[2], cz == g.nz - 1 ? ninf : fflo[2] - (float)(cz + 1));
    dzc = fmaxf(dzc - 1e-3f, 0.0f);
    const float dz = dzc * fg.cell[2];
    rem -= dz * dz;
  }
  if (isinf(r2f)) {  // unbounded query: whole row
    cxa = 0;
    cxb = g.nx - 1;
    return true;
  }
  if (!(rem >= 0.0f)) return false;
  const float xc = sqrtf(rem) * (1.0f + 1e-5f) * fg.inv[0] + 1e-3f;
  cxa = clampi(fflo[0] - xc, g.nx);
  cxb = clampi(ffhi[0] + xc, g.nx);
  return true;
}

// candidate record of sorted slot j
#ifdef RRTQX_LEGACY  // previous generation of the fused kernel (A/B: make EXTRA=-DRRTQX_LEGACY, RRTQX_RANGE_KERNEL=4)
template <int D>
struct Cand {
  double x, y, z, w;
};
template <int D>
__device__ __forceinline__ Cand<D> load_cand(const GridView &g, int j) {
  Cand<D> c;
  // SoA arrays: an octet is one contiguous 64-byte run per coordinate (fewest L1 wavefronts per trip)
  c.x = g.sx[j];
  c.y = g.sy[j];
  c.z = D >= 3 ? g.sz[j] : 0.0;
  c.w = D >= 4 ? g.sw[j] : 0.0;
  return c;
}

// Scan all candidates of the group.  DIRECT = false: append hit slots to the
// warp's shared buffer (up to CAP per query), cnt[] = number of hits.
// DIRECT = true: write results at base[k] + ordinal (a query overflowed its
// buffer).  Slots: j >= 0 is a position in the cell-sorted arrays, j < 0
// encodes node -(j+1) of the unsorted tail.
template <int D, int QN, int CAP, bool DIRECT>
__device__ __forceinline__ void scan_group(const GridView &g, const FGrid &fg, const Group<D, QN> &G, int lane,
                                           unsigned lt, unsigned sbuf, unsigned stab, int (&cnt)[QN],
                                           const int64_t (&base)[QN], int32_t *__restrict__ out_idx,
                                           double *__restrict__ out_dist) {
#pragma unroll
  for (int k = 0; k < QN; ++k) cnt[k] = 0;

  auto visit = [&](bool valid, int slot, double px, double py, double pz, double pw) {
#pragma unroll
    for (int k = 0; k < QN; ++k) {
      const double s = sqdist<D>(G.q[k], px, py, pz, pw);
      const bool hit = valid && (s < G.T[k]);
      const unsigned m = __ballot_sync(FULL, hit);
      const int o = cnt[k] + __popc(m & lt);
      if (DIRECT) {
        if (hit) {
          out_idx[base[k] + o] = slot >= 0 ? g.sperm[slot] : -(slot + 1);
          if (out_dist) out_dist[base[k] + o] = __dsqrt_rn(s);
        }
      } else {
        if (hit && o < CAP) sts32(sbuf + 4u * (unsigned)(k * CAP + o), slot);
      }
      cnt[k] += __popc(m);
    }
  };

  // walk the octet table: sub-group grp takes entry 4*it + grp
  const int grp = lane >> 3, sub = lane & 7;
  auto process_table = [&](int n_ent) {
    __syncwarp();
    // U trips per loop iteration: all candidate loads are issued before the
    // first test, so one memory latency is paid per U trips, not per trip.
    constexpr int U = FUSED_U;
    for (int e0 = grp; e0 < n_ent + grp; e0 += 4 * U) {  // uniform trip count: (e0 - grp) < n_ent
      int jj[U];
      bool vv[U];
      Cand<D> c[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int e = e0 + 4 * u;
        const int ent = e < n_ent ? lds32(stab + 4u * (unsigned)e) : 0;
        vv[u] = sub < (ent & 15);
        jj[u] = vv[u] ? (ent >> 4) + sub : 0;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) c[u] = load_cand<D>(g, jj[u]);
#pragma unroll
      for (int u = 0; u < U; ++u) {
        if (e0 - grp + 4 * u < n_ent)  // warp-uniform: skip trips past the end of the table
          visit(vv[u], jj[u], c[u].x, c[u].y, c[u].z, c[u].w);
      }
    }
    __syncwarp();
  };

  bool any_live = false;
#pragma unroll
  for (int k = 0; k < QN; ++k) any_live |= G.live[k];
  if (g.n_sorted > 0 && any_live) {
    // union of the group's row ranges
    int cy0 = 0x7fffffff, cy1 = -1, cz0 = 0x7fffffff, cz1 = -1;
#pragma unroll
    for (int k = 0; k < QN; ++k) {
      if (!G.live[k]) continue;
      const float ry = G.rf[k] * fg.inv[1] + 1e-3f;
      cy0 = min(cy0, clampi(G.ff[k][1] - ry, g.ny));
      cy1 = max(cy1, clampi(G.ff[k][1] + ry, g.ny));
      if (D >= 3) {
        const float rz = G.rf[k] * fg.inv[2] + 1e-3f;
        cz0 = min(cz0, clampi(G.ff[k][2] - rz, g.nz));
        cz1 = max(cz1, clampi(G.ff[k][2] + rz, g.nz));
      } else {
        cz0 = 0;
        cz1 = 0;
      }
    }
    const int wy = cy1 - cy0 + 1;
    const int nrows = wy * (cz1 - cz0 + 1);
    int tot = 0;  // entries currently in the table
    for (int rb = 0; rb < nrows; rb += 32) {
      // lane <-> row: reachable cells of the row for the group's query box
      int sa = 0, sb = 0;
      const int row = rb + lane;
      if (row < nrows) {
        const int cy = cy0 + row % wy, cz = cz0 + row / wy;
        int ca = 0x7fffffff, cb = -1;
#pragma unroll
        for (int k = 0; k < QN; ++k) {
          if (!G.live[k]) continue;
          int a, b;
          if (row_cells<D>(g, fg, G.ff[k], G.r2f[k], cy, cz, a, b)) { ca = min(ca, a); cb = max(cb, b); }
        }
        if (cb >= ca) {
          const int rbase = (cz * g.ny + cy) * g.nx;
          sa = g.cell_start[rbase + ca];
          sb = g.cell_start[rbase + cb + 1];
        }
      }
      // emit this batch's octets, draining the table whenever it fills up
      int todo = (sb - sa + 7) >> 3;  // octets of my row still to emit
      for (;;) {
        int incl = todo;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int t = __shfl_up_sync(FULL, incl, o);
          if (lane >= o) incl += t;
        }
        const int sum = __shfl_sync(FULL, incl, 31);
        if (sum == 0) break;
        const int pos = incl - todo;
        const int room = FUSED_TAB - tot;
        const int mine = max(0, min(todo, room - pos));
        for (int k = 0; k < mine; ++k) {
          const int n = min(8, sb - sa);
          sts32(stab + 4u * (unsigned)(tot + pos + k), (sa << 4) | n);
          sa += 8;
        }
        todo -= mine;
        tot += min(sum, room);
        if (sum <= room) break;
        process_table(tot);
        tot = 0;
      }
    }
    if (tot > 0) process_table(tot);
  }
  if (any_live)
    for (int j0 = g.n_sorted; j0 < g.n_total; j0 += 32) {  // unsorted tail of recent inserts
      const int j = j0 + lane;
      const bool v = j < g.n_total;
      double4 p = make_double4(0, 0, 0, 0);
      if (v) p = g.pos[j];
      visit(v, -(j + 1), p.x, p.y, p.z, p.w);
    }
}

// Dense flush of one query's buffered hits: all lanes hold a hit.  Full groups of 4 x 32 hits run
// without predicates, the remainder is predicated.
template <int D, bool TAIL>
__device__ __forceinline__ void flush_hits(const GridView &g, const double *q, unsigned sbuf_k, int n, int lane,
                                           int32_t *__restrict__ oi, double *__restrict__ od) {
  constexpr int U = 4;
  auto fetch = [&](int slot, int &node, double &px, double &py, double &pz, double &pw) {
    if (!TAIL || slot >= 0) {
      // SoA gathers: consecutive hits sit in consecutive slots, so these are nearly coalesced.  (A packed
      // 32-byte (x,y,z,node) record per slot was tried for scan and flush: more L1 wavefronts, slower.)
      px = g.sx[slot]; py = g.sy[slot]; pz = D >= 3 ? g.sz[slot] : 0.0; pw = D >= 4 ? g.sw[slot] : 0.0;
      node = g.sperm[slot];
    } else {
      node = -(slot + 1);
      const double4 pp = g.pos[node];
      px = pp.x; py = pp.y; pz = pp.z; pw = pp.w;
    }
  };
  const int n_full = (n / (32 * U)) * (32 * U);
  for (int h0 = lane; h0 < n_full; h0 += 32 * U) {
    int node[U];
    double px[U], py[U], pz[U], pw[U];
#pragma unroll
    for (int u = 0; u < U; ++u) fetch(lds32(sbuf_k + 4u * (unsigned)(h0 + 32 * u)), node[u], px[u], py[u], pz[u], pw[u]);
#pragma unroll
    for (int u = 0; u < U; ++u) {
      oi[h0 + 32 * u] = node[u];
      if (od) od[h0 + 32 * u] = __dsqrt_rn(sqdist<D>(q, px[u], py[u], pz[u], pw[u]));
    }
  }
  for (int h = n_full + lane; h < n; h += 32) {
    int node;
    double px, py, pz, pw;
    fetch(lds32(sbuf_k + 4u * (unsigned)h), node, px, py, pz, pw);
    oi[h] = node;
    if (od) od[h] = __dsqrt_rn(sqdist<D>(q, px, py, pz, pw));
  }
}

template <int D, int QN, int NW, int CAP>
__global__ void __launch_bounds__(NW * 32, 1)
range_fused_kernel(GridView g, const double *__restrict__ queries, const double *__restrict__ qsorted,
                   const int32_t *__restrict__ qorder, int64_t nq,
                   double r_uniform, double T_uniform, const double *__restrict__ ranges,
                   const double *__restrict__ Tq, int32_t *__restrict__ counts, int64_t *__restrict__ offsets,
                   int32_t *__restrict__ out_idx, double *__restrict__ out_dist, unsigned long long cap,
                   unsigned long long *__restrict__ cursor, int write_lists) {
  extern __shared__ int s_buf[];  // [NW][QN][CAP] hit slots, then [NW][FUSED_TAB] octet tables
  const int lane = lane_id(), warp = threadIdx.x >> 5;
  const unsigned lt = lanemask_lt();
  const unsigned s0 = (unsigned)__cvta_generic_to_shared(s_buf);
  const unsigned sbuf = pin_reg(s0 + 4u * (unsigned)(warp * QN * CAP));
  const unsigned stab = pin_reg(s0 + 4u * (unsigned)(NW * QN * CAP + warp * FUSED_TAB));
  FGrid fg;
#pragma unroll
  for (int c = 0; c < 3; ++c) { fg.inv[c] = (float)g.inv[c]; fg.cell[c] = (float)g.cell[c]; }
  const double4 p0 = g.pos[0];
  const bool has_tail = g.n_total > g.n_sorted;

  constexpr int CHUNK = NW * QN * FUSED_ROUNDS;
  const int64_t n_chunks = (nq + CHUNK - 1) / CHUNK;
  for (int64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
    const int64_t chunk_base = chunk * CHUNK;
    for (int rnd = 0; rnd < FUSED_ROUNDS; ++rnd) {
      const int64_t qfirst = chunk_base + (int64_t)(rnd * NW + warp) * QN;
      if (qfirst >= nq) break;
      Group<D, QN> G;
      bool root_extra[QN];
      double root_s = 0.0;
#pragma unroll
      for (int k = 0; k < QN; ++k) {
        const bool have = qfirst + k < nq;
        G.qid[k] = have ? qorder[qfirst + k] : -1;
        const int qq = have ? G.qid[k] : G.qid[0];
        // coordinates: from the sorted copy when the batch was sorted (coalesced), else through qorder
        const double *qsrc = qsorted ? qsorted + (have ? qfirst + k : qfirst) * D : queries + (int64_t)qq * D;
#pragma unroll
        for (int c = 0; c < D; ++c) G.q[k][c] = qsrc[c];
        const double r = have ? (ranges ? ranges[qq] : r_uniform) : -1.0;
        G.T[k] = have ? (Tq ? Tq[qq] : T_uniform) : -1.0;
        G.live[k] = have && (r > 0.0);
        const double ri = r * (1.0 + 1e-9);
        G.rf[k] = __double2float_ru(ri) * (1.0f + 1e-6f);
        G.r2f[k] = __double2float_ru(ri * ri) * (1.0f + 1e-5f);
#pragma unroll
        for (int c = 0; c < 3; ++c) G.ff[k][c] = c < D ? (float)((G.q[k][c] - g.lo[c]) * g.inv[c]) : 0.0f;
        // root (node 0) is admitted with <= (kdTree_general.jl:896-898): it needs
        // an explicit entry only when it sits exactly at distance r
        const double sr = sqdist<D>(G.q[k], p0.x, p0.y, p0.z, p0.w);
        root_extra[k] = have && !(sr < G.T[k]) && (__dsqrt_rn(sr) <= r);
        if (root_extra[k]) root_s = sr;
      }
      int cnt[QN];
      int64_t base[QN];
#pragma unroll
      for (int k = 0; k < QN; ++k) base[k] = 0;
      scan_group<D, QN, CAP, false>(g, fg, G, lane, lt, sbuf, stab, cnt, base, nullptr, nullptr);

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
        if (cnt[k] > CAP) overflow = true;
        if (lane == 0 && G.qid[k] >= 0) {
          counts[G.qid[k]] = cnt[k] + (root_extra[k] ? 1 : 0);
          offsets[G.qid[k]] = base[k];
        }
      }
      if (!write_lists || b0 > cap) continue;  // counts only, or the lists do not fit (host retries)
      if (overflow) {
        int cnt2[QN];
        scan_group<D, QN, CAP, true>(g, fg, G, lane, lt, sbuf, stab, cnt2, base, out_idx, out_dist);
      } else {
        __syncwarp();
#pragma unroll
        for (int k = 0; k < QN; ++k) {
          const unsigned sb_k = sbuf + 4u * (unsigned)(k * CAP);
          if (has_tail)
            flush_hits<D, true>(g, G.q[k], sb_k, cnt[k], lane, out_idx + base[k], out_dist ? out_dist + base[k] : nullptr);
          else
            flush_hits<D, false>(g, G.q[k], sb_k, cnt[k], lane, out_idx + base[k], out_dist ? out_dist + base[k] : nullptr);
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

#endif  // RRTQX_LEGACY

__global__ void thresh_kernel(const double *__restrict__ ranges, int64_t n, double *__restrict__ Tq) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) Tq[i] = sqrt_thresh_lt(ranges[i]);
}

// Host twin of sqrt_thresh_lt (std::sqrt is correctly rounded, like __dsqrt_rn).
static double host_sqrt_thresh_lt(double r) {
  if (r != r) return r;
  if (r <= 0.0) return 0.0;
  if (std::isinf(r)) return r;
  double t = r * r;
  if (std::isinf(t)) {
    t = 1.7976931348623157e308;
    if (std::sqrt(t) < r) return INFINITY;
  }
  while (t > 0.0 && std::sqrt(t) >= r) t = std::nextafter(t, -INFINITY);
  while (std::sqrt(t) < r) t = std::nextafter(t, INFINITY);
  return t;
}

#include "range_v5.cuh"

#ifdef RRTQX_LEGACY
// Kernel variants: more warps per SM when the expected neighbour count is small, bigger hit buffers
// (fewer warps) when it is large.  Shared memory per block = NW * (2*CAP + FUSED_TAB) * 4 bytes.
template <int D, int NW, int CAP>
static void launch_fused(rrtqx_ctx *ctx, const GridView &g, const double *dq, const double *dqs, const int32_t *qorder, int64_t nq,
                         double r, double T, const double *dr, const double *dT, int32_t *counts, int64_t *offsets,
                         int32_t *idx, double *dist, unsigned long long cap, unsigned long long *cursor,
                         int write_lists) {
  constexpr int QN = 2;
  const size_t smem = (size_t)NW * (QN * CAP + FUSED_TAB) * sizeof(int);
  if (ctx->smem_attr_done.insert((const void *)range_fused_kernel<D, QN, NW, CAP>).second)
    RQ_CUDA(cudaFuncSetAttribute(range_fused_kernel<D, QN, NW, CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  constexpr int CHUNK = NW * QN * FUSED_ROUNDS;
  const int64_t n_chunks = (nq + CHUNK - 1) / CHUNK;
  const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(n_chunks, (int64_t)ctx->sm_count));
  range_fused_kernel<D, QN, NW, CAP><<<blocks, NW * 32, smem, ctx->stream>>>(g, dq, dqs, qorder, nq, r, T, dr, dT, counts,
                                                                             offsets, idx, dist, cap, cursor, write_lists);
  post_launch(ctx);
}

#endif  // RRTQX_LEGACY

template <int D>
static void range_query_fused(rrtqx_tree *t, const double *dq, const double *dr, int64_t nq, double r, uint32_t flags,
                              rrtqx_range_result *res) {
  rrtqx_ctx *ctx = t->ctx;
  cudaStream_t st = ctx->stream;
  const bool count_only = flags & RRTQX_RANGE_COUNT_ONLY;
  const bool want_dist = flags & RRTQX_RANGE_WANT_DIST;
  res->cursor.ensure(4, st);
  {
    PhaseScope p2(ctx, "range_sort");
    sort_queries<D>(t, res, dq, nq);
  }
  // wrap-around tree (one wrap dimension, r <= period / 2, checked by the caller): expand into (real, ghost) pairs
  const bool ghost = t->wrap.num_wraps == 1;
  const double *kq = dq, *kqs = res->qbins > 0 ? res->qsorted.p : nullptr;
  const int32_t *korder = res->qorder.p;
  int32_t *kcounts = res->counts.p;
  int64_t *koffsets = res->offsets.p;
  int64_t knq = nq;
  if (ghost) {
    res->vq.ensure((size_t)nq * 2 * D, st);
    res->vorder.ensure((size_t)nq * 2, st);
    res->vcounts.ensure((size_t)nq * 2 + 1, st);
    res->voffsets.ensure((size_t)nq * 2 + 1, st);
    ghost_expand_kernel<D><<<div_up(nq, 256), 256, 0, st>>>(t->wrap, dq, kqs, res->qorder.p, nq, r, res->vq.p, res->vorder.p);
    post_launch(ctx);
    kq = nullptr; kqs = res->vq.p; korder = res->vorder.p; kcounts = res->vcounts.p; koffsets = res->voffsets.p; knq = 2 * nq;
  }
  const double T = host_sqrt_thresh_lt(r);
  const double *dT = nullptr;
  if (dr) {
    res->tq.ensure((size_t)nq, st);
    thresh_kernel<<<div_up(nq, 256), 256, 0, st>>>(dr, nq, res->tq.p);
    post_launch(ctx);
    dT = res->tq.p;
  }
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
  // Expected neighbours per query (only selects the kernel variant): the uniform-density estimate
  // n * V_ball(r) / V_grid with a 15 % margin; per-query radii take the large-buffer variant.
  double k_est = 1e9;
  if (!dr && r > 0.0 && std::isfinite(r)) {
    double vol = 1.0;
    int dims = 0;
    const int nd[3] = {t->nx, t->ny, t->nz};
    for (int c = 0; c < 3; ++c)
      if (nd[c] > 1) { vol *= nd[c] * t->cell[c]; dims++; }
    const double ball = dims == 3 ? 4.18879 * r * r * r : (dims == 2 ? 3.14159 * r * r : 2.0 * r);
    k_est = dims ? 1.15 * (double)t->n * ball / vol : (double)t->n;
  } else if (!dr) {
    k_est = 0.0;
  }
  int variant = k_est <= 576.0 ? 0 : (k_est <= 1280.0 ? 1 : (k_est <= 2048.0 ? 2 : 3));
  if (ctx->tune.fused_variant >= 0) variant = ctx->tune.fused_variant;
#ifdef RRTQX_LEGACY
  const int kernel_gen = ctx->tune.range_kernel;
#endif
  unsigned long long total = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    GridView g = t->view();
    RQ_CUDA(cudaMemsetAsync(res->cursor.p, 0, 4 * sizeof(unsigned long long), st));
    {
      PhaseScope p2(ctx, "range_fill");
#ifdef RRTQX_LEGACY
#define RQ_FUSED(NW_, CAP_)                                                                                          \
  launch_fused<D, NW_, CAP_>(ctx, g, dq, res->qbins > 0 ? res->qsorted.p : nullptr, res->qorder.p, nq, r, T, dr, dT, res->counts.p, res->offsets.p, res->idx.p, \
                             want_dist ? res->dist.p : nullptr, (unsigned long long)cap, res->cursor.p,             \
                             count_only ? 0 : 1)
#endif
#define RQ_V5(NW_, CAP_, TAB_)                                                                                       \
  launch_v5<D, NW_, CAP_, TAB_>(ctx, g, kq, kqs, korder, knq, r, T, dr, dT, kcounts, koffsets, res->idx.p, \
                                want_dist ? res->dist.p : nullptr, (unsigned long long)cap, res->cursor.p, count_only ? 0 : 1, ghost ? 1 : 0)
#ifdef RRTQX_LEGACY
      if (kernel_gen >= 5 || ghost)
#endif
      {
        // v5 (FP32 filter scan, packed exact records, 16-bit hit codes): hit buffers leave 32*V5_U entries of
        // head-room; the octet table must hold a whole pair's region (else the pair takes the exact routine)
        const int v5_nw = ctx->tune.v5_nw;
#ifdef RRTQX_EXPERIMENT
        if (variant == 0 && v5_nw == 32) RQ_V5(32, 704, 256);
        else
#endif
        if (variant == 0 && v5_nw == 28) RQ_V5(28, 704, 256);
        else if (variant == 0 && v5_nw == 20) RQ_V5(20, 704, 256);
        else if (variant == 0) RQ_V5(24, 704, 256);     // 24 warps/SM (80 registers), 94 KB shared
        else if (variant == 1) RQ_V5(24, 1408, 512);    // 186 KB
        else if (variant == 2) RQ_V5(16, 2176, 1024);   // 205 KB
        else RQ_V5(8, 4224, 2048);                      // 201 KB
      }
#ifdef RRTQX_LEGACY
      else
      if (variant == 0) RQ_FUSED(28, 576);        // 28 warps/SM (72 registers), 158 KB shared; measured best of 20/24/28/32
      else if (variant == 1) RQ_FUSED(20, 1280);  // 20 warps/SM, 225 KB
      else if (variant == 2) RQ_FUSED(12, 2048);  // 12 warps/SM, 209 KB
      else RQ_FUSED(6, 4096);                     //  6 warps/SM, 203 KB: up to 4096 neighbours buffered per query
#undef RQ_FUSED
#endif
#undef RQ_V5
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
  if (ghost) {
    ghost_merge_kernel<<<div_up(nq, 256), 256, 0, st>>>(res->vcounts.p, res->voffsets.p, nq, res->counts.p, res->offsets.p);
    post_launch(ctx);
  }
  res->n_queries = nq;
  res->total = (int64_t)total;
  res->has_lists = !count_only;
  res->has_dist = !count_only && want_dist;
}
