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
#include <unistd.h>

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
  // One-process-per-GPU form on ONE node: a result window per rank, mapped into every other rank's address space
  // through CUDA IPC.  A gather is then peer copies (copy engines, no SM) or the packing kernel's own stores straight
  // into every window, one epoch word per (destination, source) pair, and a local copy out of the own window --
  // no collective kernel that would share the SMs with the persistent range kernel.
  struct Ipc {
    bool on = false;
    size_t cap = 0;                           // bytes per slot; two slots alternate by epoch (double buffering)
    unsigned char *win = nullptr;             // own window [2][cap]
    unsigned long long *flags = nullptr;      // own epoch words [2][n_ranks]
    std::vector<unsigned char *> peer_win;    // [n_ranks], own entry = win
    std::vector<unsigned long long *> peer_flags;
    DevBuf<unsigned long long> d_peer_win, d_peer_flags;   // the same pointers on the device (for the kernels)
    DevBuf<unsigned int> counter;                         // blocks of the gather kernel that have finished (self-resetting)
    unsigned long long epoch = 0;
    // the epochs of a rank form ONE sequence whatever stream carries them (the double buffering relies on it): an
    // operation queued on another stream than its predecessor first waits for that predecessor
    cudaStream_t last_stream = nullptr;
    cudaEvent_t ev_last = nullptr;
  } ipc;
};

namespace rrtqx {

// 32 byte flags -> one word, stored into the result buffer of EVERY destination (the rank's own and, over NVLink,
// its peers'): the gather of the sharded check is this kernel's store, not a collective.
__global__ void __launch_bounds__(256)
pack_scatter_kernel(const uint8_t *__restrict__ flag, int64_t n_flags, int64_t word0, int64_t n_words,
                    const unsigned long long *__restrict__ dests, int n_dest, unsigned long long dest_byte_offset) {
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
  for (int r = 0; r < n_dest; ++r) reinterpret_cast<unsigned *>(dests[r] + dest_byte_offset)[word0 + w] = word;
}

// Epoch words of the IPC windows, one launch behind the copies / the packing kernel on the same stream: thread j
// tells rank j that this rank's part of epoch e has landed in its window (system-scope release store), then waits
// until rank j has said so for this rank's window.  Every rank signals before it waits, so there is no cycle.  The
// wait is bounded (about 10 s): a dead peer must not hang the GPU; the next call reports it through *timed_out.
__global__ void ipc_exchange_kernel(const unsigned long long *__restrict__ peer_flags, const unsigned long long *__restrict__ flags,
                                    int n_ranks, int slot, int rank, unsigned long long epoch, int32_t *__restrict__ timed_out) {
  const int j = threadIdx.x;
  if (j >= n_ranks) return;
  __threadfence_system();
  unsigned long long *f = reinterpret_cast<unsigned long long *>(peer_flags[j]) + (size_t)slot * n_ranks + rank;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + (size_t)slot * n_ranks + j) : "memory");
    if (v >= epoch) break;
    __nanosleep(100);
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 10000000000ull) { *timed_out = 1; break; }
  }
  __threadfence_system();
}


// The gather of one rank's part in ONE launch: a few small blocks (128 threads, <= 32 registers: they fit next to a
// resident block of the persistent range kernel) store the part into slot `slot` of EVERY rank's window over NVLink;
// the last block to finish runs the epoch exchange above.  (Eight cudaMemcpyAsync calls instead cost 40 us of host
// time per step at N = 8 -- more than the NCCL call they replaced.)
__global__ void __launch_bounds__(128, 16)
ipc_gather_kernel(const unsigned char *__restrict__ send, size_t bytes, const unsigned long long *__restrict__ peer_win,
                  size_t dst_offset, int n_ranks, unsigned int *__restrict__ counter,
                  const unsigned long long *__restrict__ peer_flags, const unsigned long long *__restrict__ flags, int slot,
                  int rank, unsigned long long epoch, int32_t *__restrict__ timed_out) {
  __shared__ int s_last;
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x, nthr = (size_t)gridDim.x * blockDim.x;
  if ((((uintptr_t)send | bytes | dst_offset) & 15u) == 0) {
    const uint4 *src = reinterpret_cast<const uint4 *>(send);
    for (size_t i = tid; i < bytes / 16; i += nthr) {
      const uint4 v = src[i];
      for (int j = 0; j < n_ranks; ++j) reinterpret_cast<uint4 *>(peer_win[j] + dst_offset)[i] = v;
    }
  } else {
    for (size_t i = tid; i < bytes; i += nthr) {
      const unsigned char v = send[i];
      for (int j = 0; j < n_ranks; ++j) reinterpret_cast<unsigned char *>(peer_win[j] + dst_offset)[i] = v;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(counter, 1u);
    s_last = prev == gridDim.x - 1;
    if (s_last) *counter = 0u;
  }
  __syncthreads();
  if (!s_last) return;
  const int j = threadIdx.x;
  if (j >= n_ranks) return;
  __threadfence_system();
  unsigned long long *f = reinterpret_cast<unsigned long long *>(peer_flags[j]) + (size_t)slot * n_ranks + rank;
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(f), "l"(epoch) : "memory");
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + (size_t)slot * n_ranks + j) : "memory");
    if (v >= epoch) break;
    __nanosleep(100);
    unsigned long long t1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t1));
    if (t1 - t0 > 10000000000ull) { *timed_out = 1; break; }
  }
  __threadfence_system();
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

void ipc_teardown(rrtqx_comm *c) {
  rrtqx_comm::Ipc &I = c->ipc;
  const int me = c->rank0;
  for (size_t j = 0; j < I.peer_win.size(); ++j) {
    if ((int)j == me) continue;
    if (I.peer_win[j]) cudaIpcCloseMemHandle(I.peer_win[j]);
    if (I.peer_flags[j]) cudaIpcCloseMemHandle(I.peer_flags[j]);
  }
  if (I.win) cudaFree(I.win);
  if (I.flags) cudaFree(I.flags);
  I.win = nullptr; I.flags = nullptr;
  if (I.ev_last) { cudaEventDestroy(I.ev_last); I.ev_last = nullptr; }
  I.last_stream = nullptr;
  I.peer_win.clear(); I.peer_flags.clear();
  I.on = false;
  cudaGetLastError();
}

// Collective over the ranks of a one-process-per-GPU communicator: allocate the window, exchange the IPC handles
// through NCCL, map the peers' windows.  Every rank ends with the same answer (a second exchange agrees on it); on
// "no" the windows are released and the gathers stay on NCCL.
void ipc_setup(rrtqx_comm *c, size_t cap) {
  rrtqx_ctx *ctx = c->ctx[0];
  cudaStream_t st = ctx->stream;
  rrtqx_comm::Ipc &I = c->ipc;
  const int G = c->n_ranks, me = c->rank0;
  NcclApi &N = nccl_api();
  struct Card { cudaIpcMemHandle_t win, flags; char host[64]; int32_t ok, device; unsigned char pad[56]; };
  static_assert(sizeof(Card) == 256, "IPC exchange card");
  Card mine;
  memset(&mine, 0, sizeof(mine));
  mine.device = ctx->device;
  gethostname(mine.host, sizeof(mine.host) - 1);
  bool ok = true;
  if (cudaMalloc((void **)&I.win, 2 * cap) != cudaSuccess) ok = false;
  if (ok && cudaMalloc((void **)&I.flags, 2 * sizeof(unsigned long long) * (size_t)G) != cudaSuccess) ok = false;
  if (ok) ok = cudaMemset(I.flags, 0, 2 * sizeof(unsigned long long) * (size_t)G) == cudaSuccess;
  if (ok) ok = cudaIpcGetMemHandle(&mine.win, I.win) == cudaSuccess && cudaIpcGetMemHandle(&mine.flags, I.flags) == cudaSuccess;
  cudaGetLastError();
  mine.ok = ok ? 1 : 0;
  DevBuf<unsigned char> stage;
  stage.ensure(sizeof(Card) * (size_t)G, st);
  std::vector<Card> all((size_t)G);
  auto exchange = [&]() {
    RQ_CUDA(cudaMemcpyAsync(stage.p + sizeof(Card) * (size_t)me, &mine, sizeof(Card), cudaMemcpyHostToDevice, st));
    RQ_NCCL(N.AllGather(stage.p + sizeof(Card) * (size_t)me, stage.p, sizeof(Card), ncclUint8, c->nccl[0], st));
    RQ_CUDA(cudaMemcpyAsync(all.data(), stage.p, sizeof(Card) * (size_t)G, cudaMemcpyDeviceToHost, st));
    RQ_CUDA(cudaStreamSynchronize(st));
  };
  exchange();
  for (int j = 0; j < G; ++j) ok = ok && all[j].ok && strncmp(all[j].host, mine.host, sizeof(mine.host)) == 0;
  I.peer_win.assign((size_t)G, nullptr);
  I.peer_flags.assign((size_t)G, nullptr);
  if (ok) {
    for (int j = 0; j < G && ok; ++j) {
      if (j == me) { I.peer_win[j] = I.win; I.peer_flags[j] = I.flags; continue; }
      int can = 0;
      if (cudaDeviceCanAccessPeer(&can, ctx->device, all[j].device) != cudaSuccess || !can) { ok = false; break; }
      void *pw = nullptr, *pf = nullptr;
      if (cudaIpcOpenMemHandle(&pw, all[j].win, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
      I.peer_win[j] = (unsigned char *)pw;
      if (cudaIpcOpenMemHandle(&pf, all[j].flags, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) { ok = false; break; }
      I.peer_flags[j] = (unsigned long long *)pf;
    }
    cudaGetLastError();
  }
  mine.ok = ok ? 1 : 0;
  exchange();   // agree: a window is used only if EVERY rank mapped every window
  for (int j = 0; j < G; ++j) ok = ok && all[j].ok;
  if (!ok) { ipc_teardown(c); return; }
  std::vector<unsigned long long> pw((size_t)G), pf((size_t)G);
  for (int j = 0; j < G; ++j) { pw[j] = (unsigned long long)(uintptr_t)I.peer_win[j]; pf[j] = (unsigned long long)(uintptr_t)I.peer_flags[j]; }
  I.d_peer_win.ensure((size_t)G, st);
  I.d_peer_flags.ensure((size_t)G, st);
  I.counter.ensure(4, st);
  RQ_CUDA(cudaMemsetAsync(I.counter.p, 0, 4 * sizeof(unsigned int), st));
  RQ_CUDA(cudaMemcpyAsync(I.d_peer_win.p, pw.data(), sizeof(unsigned long long) * (size_t)G, cudaMemcpyHostToDevice, st));
  RQ_CUDA(cudaMemcpyAsync(I.d_peer_flags.p, pf.data(), sizeof(unsigned long long) * (size_t)G, cudaMemcpyHostToDevice, st));
  RQ_CUDA(cudaStreamSynchronize(st));
  I.cap = cap;
  I.epoch = 0;
  I.on = true;
}

// start of an IPC gather on stream s: order it behind the previous one, draw the epoch and its slot
unsigned long long ipc_begin(rrtqx_comm *c, cudaStream_t s, int *slot) {
  rrtqx_comm::Ipc &I = c->ipc;
  if (!I.ev_last) RQ_CUDA(cudaEventCreateWithFlags(&I.ev_last, cudaEventDisableTiming));
  if (I.last_stream && I.last_stream != s) RQ_CUDA(cudaStreamWaitEvent(s, I.ev_last, 0));
  const unsigned long long e = ++I.epoch;
  *slot = (int)(e & 1ull);
  return e;
}

// a wait that ran into its 10 s bound (dead or wedged peer) fails the NEXT call on the communicator
void ipc_check(rrtqx_comm *c) {
  if (!c->ipc.on) return;
  HostMail &M = host_mail(c->ctx[0]);
  if (*(volatile int32_t *)&M.h->aux[1]) {
    M.h->aux[1] = 0;
    throw Error(RRTQX_ERR_CUDA, "a rank did not publish its part of a gather within 10 s (IPC window)");
  }
}

// Behind the writers of epoch e on stream s (peer copies or the packing kernel): publish, wait for every rank's part,
// copy the own window out.  Returns nothing; a timed-out wait is reported by ipc_check().
void ipc_finish(rrtqx_comm *c, cudaStream_t s, int slot, unsigned long long e, void *recv, size_t total_bytes) {
  rrtqx_comm::Ipc &I = c->ipc;
  const int G = c->n_ranks;
  HostMail &M = host_mail(c->ctx[0]);
  ipc_exchange_kernel<<<1, 32, 0, s>>>(I.d_peer_flags.p, I.flags, G, slot, c->rank0, e, &M.d->aux[1]);
  RQ_CUDA(cudaMemcpyAsync(recv, I.win + (size_t)slot * I.cap, total_bytes, cudaMemcpyDeviceToDevice, s));
  RQ_CUDA(cudaEventRecord(I.ev_last, s));
  I.last_stream = s;
  post_launch(c->ctx[0], 1);
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
      if (!ctx->tune.comm_no_ipc && n_ranks <= 32) ipc_setup(c, (size_t)ctx->tune.comm_ipc_mb << 20);
    }
    *out = c;
  });
}

rrtqx_status rrtqx_comm_destroy(rrtqx_comm *c) {
  if (!c) return RRTQX_OK;
  if (c->ipc.win) {
    if (handle_live(c->ctx[0])) { cudaSetDevice(c->ctx[0]->device); cudaStreamSynchronize(c->ctx[0]->stream); }
    cudaStreamSynchronize(c->side[0]);
    ipc_teardown(c);
  }
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
  if (peer_stores) *peer_stores = (c->p2p || c->ipc.on) ? 1 : 0;
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
    ipc_check(c);
    if (c->ipc.on && c->n_local == 1 && (size_t)bytes_per_rank * (size_t)c->n_ranks <= c->ipc.cap) {
      // copy engines: this rank's part goes into slot (epoch & 1) of every rank's window, then epoch words, then the
      // own window is copied out -- no kernel of a collective library on the SMs
      rrtqx_comm::Ipc &I = c->ipc;
      rrtqx_ctx *ctx = c->ctx[0];
      RQ_CUDA(cudaSetDevice(ctx->device));
      cudaStream_t s = ctx->stream;
      if (on_side_stream) {
        RQ_CUDA(cudaEventRecord(c->ev_main[0], ctx->stream));
        RQ_CUDA(cudaStreamWaitEvent(c->side[0], c->ev_main[0], 0));
        s = c->side[0];
      }
      int slot = 0;
      const unsigned long long e = ipc_begin(c, s, &slot);
      HostMail &M = host_mail(ctx);
      const int blocks = (int)std::max<int64_t>(1, std::min<int64_t>(32, bytes_per_rank / (16 * 128)));
      ipc_gather_kernel<<<blocks, 128, 0, s>>>((const unsigned char *)send[0], (size_t)bytes_per_rank, I.d_peer_win.p,
                                               (size_t)slot * I.cap + (size_t)c->rank0 * (size_t)bytes_per_rank, c->n_ranks,
                                               I.counter.p, I.d_peer_flags.p, I.flags, slot, c->rank0, e, &M.d->aux[1]);
      RQ_CUDA(cudaMemcpyAsync(recv[0], I.win + (size_t)slot * I.cap, (size_t)bytes_per_rank * (size_t)c->n_ranks,
                              cudaMemcpyDeviceToDevice, s));
      RQ_CUDA(cudaEventRecord(I.ev_last, s));
      I.last_stream = s;
      post_launch(ctx, 1);
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
    // one process per GPU with IPC windows: the packing kernel stores into slot (epoch & 1) of every rank's window
    const bool ipc_direct = !direct && c->ipc.on && L == 1 && G > 1 && (size_t)per * 4u * (size_t)G <= c->ipc.cap;
    ipc_check(c);
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
      unsigned long long e = 0;
      int slot = 0;
      if (ipc_direct) e = ipc_begin(c, ctx->stream, &slot);
      // every word of the rank's shard is written, also the padding words behind the last edge (zero): the gathered
      // array is then fully defined whatever the caller's buffer or the window held before
      if (per > 0) {
        if (ipc_direct)
          pack_scatter_kernel<<<div_up(per, 256), 256, 0, ctx->stream>>>(c->flag[i]->p, std::max<int64_t>(cnt, 0), per * g, per,
                                                                         c->ipc.d_peer_win.p, G, (unsigned long long)slot * c->ipc.cap);
        else
          pack_scatter_kernel<<<div_up(per, 256), 256, 0, ctx->stream>>>(c->flag[i]->p, std::max<int64_t>(cnt, 0), per * g, per,
                                                                         c->peers[i]->p, (int)c->peers_host[i].size(), 0ull);
        post_launch(ctx);
      }
      if (ipc_direct) ipc_finish(c, ctx->stream, slot, e, packed_out[i], (size_t)per * 4u * (size_t)G);
    }
    if (G > 1 && !direct && !ipc_direct) {  // in-place all-gather of the equal word shards
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
