// comm.cu -- multi-GPU entry points of the C ABI.
//
// The path shards embarrassingly: the tree and the obstacle set are replicated on every GPU, rank g owns a
// contiguous slice of a query / edge batch, and the ONLY exchange is the gather of fixed-size results (bit-packed
// per-edge flags, per-query counts).  Two ways to drive it:
//   * ONE process holding N contexts (rrtqx_comm_init_local) -- what a Julia host does, the reference being one
//     process (rrtqx.jl:331).  With peer access between all devices the flag gather is fused into the packing
//     kernel: each rank packs 32 flags per word and STORES the word straight into every peer's result buffer
//     over NVLink (no collective call, no staging copy); without peer access NCCL gathers the packed words.
//   * one process per GPU (rrtqx_comm_init_rank; torchrun) -- NCCL all-gather of the packed words, in place.
// NCCL is loaded with dlopen (libnccl.so.2: the copy already in the process when the host is PyTorch, the system
// one otherwise), so the library itself has no link-time dependency on it and single-GPU users never touch it.
#include <dlfcn.h>
#include <nccl.h>

#include <vector>

#include "objects.cuh"

using namespace rrtqx;

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitRankConfig)(ncclComm_t *, int, ncclUniqueId, int, ncclConfig_t *) = nullptr;   // optional
  ncclResult_t (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GetVersion)(int *) = nullptr;
};

static NcclApi &nccl_api() {
  static NcclApi api;
  static std::mutex mu;
  std::lock_guard<std::mutex> lk(mu);
  if (api.lib) return api;
  void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!h) throw Error(RRTQX_ERR_UNSUPPORTED, std::string("NCCL is not available (dlopen libnccl.so.2): ") + dlerror());
  auto sym = [&](const char *n) {
    void *p = dlsym(h, n);
    if (!p) throw Error(RRTQX_ERR_UNSUPPORTED, std::string("NCCL symbol missing: ") + n);
    return p;
  };
  api.GetUniqueId = (decltype(api.GetUniqueId))sym("ncclGetUniqueId");
  api.CommInitRank = (decltype(api.CommInitRank))sym("ncclCommInitRank");
  api.CommInitRankConfig = (decltype(api.CommInitRankConfig))dlsym(h, "ncclCommInitRankConfig");
  api.CommInitAll = (decltype(api.CommInitAll))sym("ncclCommInitAll");
  api.CommDestroy = (decltype(api.CommDestroy))sym("ncclCommDestroy");
  api.AllGather = (decltype(api.AllGather))sym("ncclAllGather");
  api.GroupStart = (decltype(api.GroupStart))sym("ncclGroupStart");
  api.GroupEnd = (decltype(api.GroupEnd))sym("ncclGroupEnd");
  api.GetErrorString = (decltype(api.GetErrorString))sym("ncclGetErrorString");
  api.GetVersion = (decltype(api.GetVersion))sym("ncclGetVersion");
  api.lib = h;
  return api;
}

#define RQ_NCCL(expr)                                                                                    \
  do {                                                                                                   \
    ncclResult_t _r = (expr);                                                                            \
    if (_r != ncclSuccess)                                                                               \
      throw Error(RRTQX_ERR_CUDA, std::string(#expr " failed: ") + nccl_api().GetErrorString(_r));       \
  } while (0)

struct rrtqx_comm {
  int n_ranks = 0;   // ranks of the whole job
  int n_local = 0;   // ranks this process drives (N in the one-process form, 1 under torchrun)
  int rank0 = 0;     // global rank of local rank 0
  bool p2p = false;  // every local device can store into every other local device's memory (one-process form)
  std::vector<rrtqx_ctx *> ctx;
  std::vector<ncclComm_t> nccl;
  std::vector<cudaStream_t> side;    // one side stream per local rank: gathers that must not delay the next step
  std::vector<cudaEvent_t> ev_main, ev_side;
  // per local rank: byte flags of its slice, and the device array of the peers' result pointers for the fused store
  std::vector<DevBuf<uint8_t> *> flag;
  std::vector<DevBuf<unsigned long long> *> peers;
  std::vector<std::vector<unsigned long long>> peers_host;  // what `peers` holds (uploaded only when it changes)
};

namespace rrtqx {

// 32 byte flags -> one word, stored into the result buffer of EVERY destination (the rank's own and, over NVLink,
// its peers'): the gather of the sharded check is this kernel's store, not a collective.
__global__ void __launch_bounds__(256)
pack_scatter_kernel(const uint8_t *__restrict__ flag, int64_t n_flags, int64_t word0, int64_t n_words,
                    const unsigned long long *__restrict__ dests, int n_dest) {
  const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  const int64_t base = w * 32;
  unsigned word = 0;
  if (base + 32 <= n_flags) {  // flag buffers come from cudaMalloc and base is a multiple of 32
    const uint4 a = *reinterpret_cast<const uint4 *>(flag + base), b = *reinterpret_cast<const uint4 *>(flag + base + 16);
    const unsigned v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
    for (int k = 0; k < 8; ++k)
#pragma unroll
      for (int j = 0; j < 4; ++j) word |= ((v[k] >> (8 * j)) & 0xffu) ? (1u << (4 * k + j)) : 0u;
  } else {
    for (int j = 0; j < 32 && base + j < n_flags; ++j) word |= flag[base + j] ? (1u << j) : 0u;
  }
  for (int r = 0; r < n_dest; ++r) reinterpret_cast<unsigned *>(dests[r])[word0 + w] = word;
}

}  // namespace rrtqx

namespace {
template <typename F>
rrtqx_status guarded_c(rrtqx_ctx *ctx, F &&f) {
  try {
    f();
    return RRTQX_OK;
  } catch (const Error &e) {
    if (ctx) ctx->err = e.what();
    return e.code;
  } catch (const std::exception &e) {
    if (ctx) ctx->err = e.what();
    return RRTQX_ERR_INVALID;
  }
}

rrtqx_comm *new_comm(int n_ranks, int n_local, int rank0, rrtqx_ctx *const *ctxs) {
  rrtqx_comm *c = new rrtqx_comm();
  c->n_ranks = n_ranks; c->n_local = n_local; c->rank0 = rank0;
  for (int i = 0; i < n_local; ++i) {
    c->ctx.push_back(ctxs[i]);
    RQ_CUDA(cudaSetDevice(ctxs[i]->device));
    cudaStream_t s;
    cudaEvent_t a, b;
    RQ_CUDA(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    RQ_CUDA(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
    RQ_CUDA(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
    c->side.push_back(s); c->ev_main.push_back(a); c->ev_side.push_back(b);
    c->flag.push_back(new DevBuf<uint8_t>());
    c->peers.push_back(new DevBuf<unsigned long long>());
    c->peers_host.emplace_back();
  }
  return c;
}
}  // namespace

extern "C" {

rrtqx_status rrtqx_comm_init_local(rrtqx_ctx *const *ctxs, int32_t n, rrtqx_comm **out) {
  if (!ctxs || n <= 0 || !out || !ctxs[0]) return RRTQX_ERR_INVALID;
  return guarded_c(ctxs[0], [&] {
    *out = nullptr;
    std::vector<int> devs;
    for (int i = 0; i < n; ++i) {
      RQ_REQUIRE(ctxs[i] != nullptr, "NULL context");
      for (int j = 0; j < i; ++j) RQ_REQUIRE(ctxs[j]->device != ctxs[i]->device, "one context per device, please");
      devs.push_back(ctxs[i]->device);
    }
    rrtqx_comm *c = new_comm(n, n, 0, ctxs);
    c->nccl.resize(n);
    if (n > 1) RQ_NCCL(nccl_api().CommInitAll(c->nccl.data(), n, devs.data()));
    // peer access, both directions of every pair; "already enabled" is fine
    bool p2p = n > 1;
    for (int i = 0; i < n && p2p; ++i)
      for (int j = 0; j < n && p2p; ++j) {
        if (i == j) continue;
        int can = 0;
        RQ_CUDA(cudaDeviceCanAccessPeer(&can, devs[i], devs[j]));
        if (!can) { p2p = false; break; }
        RQ_CUDA(cudaSetDevice(devs[i]));
        cudaError_t e = cudaDeviceEnablePeerAccess(devs[j], 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) p2p = false;
        cudaGetLastError();
      }
    c->p2p = p2p;
    *out = c;
  });
}

rrtqx_status rrtqx_comm_unique_id(void *id128) {
  if (!id128) return RRTQX_ERR_INVALID;
  return guarded_c(nullptr, [&] {
    static_assert(sizeof(ncclUniqueId) == 128, "the ABI passes NCCL unique ids as 128 bytes");
    ncclUniqueId id;
    RQ_NCCL(nccl_api().GetUniqueId(&id));
    memcpy(id128, &id, sizeof(id));
  });
}

rrtqx_status rrtqx_comm_init_rank(rrtqx_ctx *ctx, const void *id128, int32_t rank, int32_t n_ranks, rrtqx_comm **out) {
  if (!ctx || !id128 || !out) return RRTQX_ERR_INVALID;
  return guarded_c(ctx, [&] {
    *out = nullptr;
    RQ_REQUIRE(n_ranks >= 1 && rank >= 0 && rank < n_ranks, "rank out of range");
    rrtqx_comm *c = new_comm(n_ranks, 1, rank, &ctx);
    c->nccl.resize(1);
    if (n_ranks > 1) {
      ncclUniqueId id;
      memcpy(&id, id128, sizeof(id));
      RQ_CUDA(cudaSetDevice(ctx->device));
      // RRTQX_NCCL_MAX_CTAS > 0 caps the CTAs NCCL may use (diagnosis: the gathers share the GPU with persistent
      // kernels that own every SM).  Measured on 8 B200s (scripts/exp_scale8.sh): caps of 1, 2, 4 make the C2 step
      // 1.4-3 % and the sharded edge batch 27-135 % slower than NCCL's default, so the default is no cap.  The config
      // struct of the compile-time header is accepted by newer NCCL builds (its size / version fields exist for that).
      NcclApi &N = nccl_api();
      if (N.CommInitRankConfig && ctx->tune.nccl_max_ctas > 0) {
        ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
        cfg.minCTAs = 1;
        cfg.maxCTAs = ctx->tune.nccl_max_ctas;
        RQ_NCCL(N.CommInitRankConfig(&c->nccl[0], n_ranks, id, rank, &cfg));
      } else {
        RQ_NCCL(N.CommInitRank(&c->nccl[0], n_ranks, id, rank));
      }
    }
    *out = c;
  });
}

rrtqx_status rrtqx_comm_destroy(rrtqx_comm *c) {
  if (!c) return RRTQX_OK;
  for (int i = 0; i < c->n_local; ++i) {
    if (handle_live(c->ctx[i])) cudaSetDevice(c->ctx[i]->device);
    cudaStreamSynchronize(c->side[i]);
    if (c->n_ranks > 1 && c->nccl[i]) nccl_api().CommDestroy(c->nccl[i]);
    cudaStreamDestroy(c->side[i]);
    cudaEventDestroy(c->ev_main[i]);
    cudaEventDestroy(c->ev_side[i]);
    delete c->flag[i];
    delete c->peers[i];
  }
  cudaGetLastError();
  delete c;
  return RRTQX_OK;
}

rrtqx_status rrtqx_comm_info(const rrtqx_comm *c, int32_t *n_ranks, int32_t *n_local, int32_t *first_rank,
                             int32_t *peer_stores, int32_t *nccl_version) {
  if (!c) return RRTQX_ERR_INVALID;
  if (n_ranks) *n_ranks = c->n_ranks;
  if (n_local) *n_local = c->n_local;
  if (first_rank) *first_rank = c->rank0;
  if (peer_stores) *peer_stores = c->p2p ? 1 : 0;
  if (nccl_version) {
    *nccl_version = 0;
    if (c->n_ranks > 1) { int v = 0; if (nccl_api().GetVersion(&v) == ncclSuccess) *nccl_version = v; }
  }
  return RRTQX_OK;
}

// words of the packed flag array of a sharded check: G equal shards of ceil(ceil(n / 32) / G) words
rrtqx_status rrtqx_comm_packed_words(const rrtqx_comm *c, int64_t n_edges, int64_t *words_total, int64_t *words_per_rank) {
  if (!c || n_edges < 0) return RRTQX_ERR_INVALID;
  const int64_t nw = (n_edges + 31) / 32, per = (nw + c->n_ranks - 1) / c->n_ranks;
  if (words_total) *words_total = per * c->n_ranks;
  if (words_per_rank) *words_per_rank = per;
  return RRTQX_OK;
}

rrtqx_status rrtqx_comm_allgather(rrtqx_comm *c, const void *const *send, void *const *recv, int64_t bytes_per_rank,
                                  int32_t on_side_stream) {
  if (!c || !send || !recv) return RRTQX_ERR_INVALID;
  return guarded_c(c->ctx[0], [&] {
    RQ_REQUIRE(bytes_per_rank >= 0, "bytes_per_rank is negative");
    if (bytes_per_rank == 0) return;
    if (c->n_ranks == 1) {
      RQ_CUDA(cudaSetDevice(c->ctx[0]->device));
      if (send[0] != recv[0]) RQ_CUDA(cudaMemcpyAsync(recv[0], send[0], (size_t)bytes_per_rank, cudaMemcpyDeviceToDevice, c->ctx[0]->stream));
      return;
    }
    NcclApi &N = nccl_api();
    if (on_side_stream)  // the gather waits for what the main stream has queued so far, and nothing waits for it
      for (int i = 0; i < c->n_local; ++i) {
        RQ_CUDA(cudaSetDevice(c->ctx[i]->device));
        RQ_CUDA(cudaEventRecord(c->ev_main[i], c->ctx[i]->stream));
        RQ_CUDA(cudaStreamWaitEvent(c->side[i], c->ev_main[i], 0));
      }
    RQ_NCCL(N.GroupStart());
    for (int i = 0; i < c->n_local; ++i)
      RQ_NCCL(N.AllGather(send[i], recv[i], (size_t)bytes_per_rank, ncclUint8, c->nccl[i],
                          on_side_stream ? c->side[i] : c->ctx[i]->stream));
    RQ_NCCL(N.GroupEnd());
    for (int i = 0; i < c->n_local; ++i) c->ctx[i]->launches += 1;
  });
}

// the main streams wait for every gather queued on the side streams so far
rrtqx_status rrtqx_comm_join(rrtqx_comm *c) {
  if (!c) return RRTQX_ERR_INVALID;
  return guarded_c(c->ctx[0], [&] {
    for (int i = 0; i < c->n_local; ++i) {
      RQ_CUDA(cudaSetDevice(c->ctx[i]->device));
      RQ_CUDA(cudaEventRecord(c->ev_side[i], c->side[i]));
      RQ_CUDA(cudaStreamWaitEvent(c->ctx[i]->stream, c->ev_side[i], 0));
    }
  });
}

rrtqx_status rrtqx_edge_check_batch_sharded(rrtqx_comm *c, rrtqx_tree *const *trees, const rrtqx_spheres *const *spheres,
                                            const int32_t *const *src, const int32_t *const *dst, int64_t n_edges,
                                            double robot_radius, uint32_t flags, uint32_t *const *packed_out) {
  if (!c || !trees || !spheres || !src || !dst || !packed_out) return RRTQX_ERR_INVALID;
  return guarded_c(c->ctx[0], [&] {
    RQ_REQUIRE(n_edges >= 0, "n_edges is negative");
    const int G = c->n_ranks, L = c->n_local;
    const int64_t nw = (n_edges + 31) / 32, per = (nw + G - 1) / G;
    std::vector<const int32_t *> bad((size_t)L, nullptr);
    for (int i = 0; i < L; ++i) {
      RQ_REQUIRE(trees[i] && spheres[i] && packed_out[i], "NULL handle / buffer");
      RQ_REQUIRE(trees[i]->ctx == c->ctx[i] && spheres[i]->ctx == c->ctx[i], "tree / obstacle set of another context");
      RQ_REQUIRE(trees[i]->d == 3, "SimpleEdge checks need a 3-D tree (explicitEdgeCheck3D)");
      RQ_REQUIRE(is_device_ptr(src[i]) && is_device_ptr(dst[i]) && is_device_ptr(packed_out[i]),
                 "the sharded check works on device-resident arrays (replicated edge list, result words)");
    }
    // destinations of each rank's words: every local rank's buffer when peers can be stored to directly, otherwise
    // the rank's own buffer (NCCL moves the words afterwards).  Uploaded only when the caller's buffers change.
    const bool direct = c->p2p && L == G && G > 1;
    for (int i = 0; i < L; ++i) {
      std::vector<unsigned long long> dests;
      if (direct) for (int j = 0; j < L; ++j) dests.push_back((unsigned long long)(uintptr_t)packed_out[j]);
      else dests.push_back((unsigned long long)(uintptr_t)packed_out[i]);
      if (dests != c->peers_host[i]) {
        rrtqx_ctx *ctx = c->ctx[i];
        RQ_CUDA(cudaSetDevice(ctx->device));
        c->peers[i]->ensure(dests.size() + 1, ctx->stream);
        RQ_CUDA(cudaMemcpyAsync(c->peers[i]->p, dests.data(), sizeof(unsigned long long) * dests.size(), cudaMemcpyHostToDevice, ctx->stream));
        RQ_CUDA(cudaStreamSynchronize(ctx->stream));  // dests is a local vector
        c->peers_host[i] = dests;
      }
    }
    // queue the slice of every local rank before waiting for any of them
    for (int i = 0; i < L; ++i) {
      rrtqx_ctx *ctx = c->ctx[i];
      RQ_CUDA(cudaSetDevice(ctx->device));
      const int g = c->rank0 + i;
      const int64_t w0 = std::min<int64_t>(per * g, nw), w1 = std::min<int64_t>(per * (g + 1), nw);
      const int64_t e0 = w0 * 32, e1 = std::min<int64_t>(w1 * 32, n_edges), cnt = e1 - e0;
      c->flag[i]->ensure((size_t)std::max<int64_t>(per * 32, 32), ctx->stream);
      if (cnt > 0)
        edge_check_launch(ctx, trees[i], spheres[i], src[i] + e0, dst[i] + e0, nullptr, nullptr, cnt, robot_radius, flags,
                          c->flag[i]->p, &bad[i]);
      if (w1 > w0) {
        pack_scatter_kernel<<<div_up(w1 - w0, 256), 256, 0, ctx->stream>>>(c->flag[i]->p, cnt, w0, w1 - w0, c->peers[i]->p,
                                                                           (int)c->peers_host[i].size());
        post_launch(ctx);
      }
    }
    if (G > 1 && !direct) {  // in-place all-gather of the equal word shards
      NcclApi &N = nccl_api();
      RQ_NCCL(N.GroupStart());
      for (int i = 0; i < L; ++i)
        RQ_NCCL(N.AllGather(packed_out[i] + per * (c->rank0 + i), packed_out[i], (size_t)per * 4, ncclUint8, c->nccl[i], c->ctx[i]->stream));
      RQ_NCCL(N.GroupEnd());
    }
    int32_t n_bad = 0;
    for (int i = 0; i < L; ++i) {
      RQ_CUDA(cudaSetDevice(c->ctx[i]->device));
      int32_t b = 0;
      if (bad[i]) RQ_CUDA(cudaMemcpyAsync(&b, bad[i], sizeof(int32_t), cudaMemcpyDeviceToHost, c->ctx[i]->stream));
      RQ_CUDA(cudaStreamSynchronize(c->ctx[i]->stream));
      n_bad += b;
    }
    RQ_REQUIRE(n_bad == 0, "edge endpoint out of range");
  });
}

}  // extern "C"
