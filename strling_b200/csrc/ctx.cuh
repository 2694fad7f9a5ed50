// Internal definition of strgpu_ctx shared by api.cu (context, scan, cluster entry points) and comm.cu (NCCL entry points).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <mutex>

#include "cluster_kernels.cuh"
#include "scan_kernels.cuh"
#include "strgpu.h"

namespace strgpu_internal {

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
};

struct Slot {
  cudaStream_t stream = nullptr;
  cudaEvent_t done = nullptr;
  DevBuf seq, nmask, xmask, segs, out, list;  // list: survivor / stage lists of the scan kernels (scan_scratch_words)
  bool busy = false;
  strgpu_repeat *host_out = nullptr;
  uint32_t n_seg = 0;
  int *d_status = nullptr;
  int *h_status = nullptr;  // pinned
};

struct Comm;  // comm.cu: the NCCL communicator of a sharded job and its exchange buffers
struct Decode;  // decode_kernels.cu: stream and buffers of strgpu_inflate_bgzf

// The cluster path is ~100 short kernels; a caller that repeats a call with the same arguments (a pipeline clustering batch
// after batch into the same buffers) gets it replayed as ONE CUDA graph launch: first call direct (it sizes the workspace),
// second call captured, from then on replayed until an argument or a workspace buffer changes.  STRGPU_NO_GRAPH=1 disables it.
struct GraphSlot {
  cudaGraphExec_t exec = nullptr;
  uint64_t key = 0, ws_gen = 0;          // what the captured graph was built for
  uint64_t sized_key = 0, sized_gen = 0; // arguments of the last direct run
  uint64_t launches = 0;
  bool broken = false;                   // capture failed once: stay on direct launches
};

}  // namespace strgpu_internal

struct strgpu_ctx {
  int device = -1;
  int sm_count = 0;
  uint16_t *d_thr = nullptr;
  uint16_t *d_luts = nullptr;
  int variant = 0;
  bool thr_set = false;
  strgpu_internal::Slot slots[STRGPU_SLOTS];
  int *d_status_dev = nullptr;  // sticky status for strgpu_scan_device launches
  strgpu::ClusterWorkspace cluster_ws;
  cudaStream_t cluster_stream = nullptr;
  strgpu_internal::DevBuf cl_in, cl_out, cl_loci;
  strgpu_internal::DevBuf dev_list;                   // survivor list for strgpu_scan_device launches
  cudaEvent_t dev_list_done = nullptr;
  uint32_t *d_cl_n = nullptr;
  uint64_t launches = 0;
  char err[512] = {0};
  // submit and wait may be called from two different host threads (a producer that stages batches and a consumer that
  // replays results): slot bookkeeping, the launch counter and the error text are guarded by this mutex
  std::mutex mu;
  std::mutex err_mu;
  strgpu_internal::Comm *comm = nullptr;
  strgpu_internal::Decode *decode = nullptr;
  strgpu_internal::GraphSlot cluster_graph;
};

namespace strgpu_internal {

inline int fail(strgpu_ctx *ctx, int status, const char *fmt, ...) {
  if (ctx) {
    std::lock_guard<std::mutex> lk(ctx->err_mu);
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
    va_end(ap);
  }
  return status;
}

#define CU(ctx, call)                                                                                                    \
  do {                                                                                                                   \
    cudaError_t e_ = (call);                                                                                             \
    if (e_ != cudaSuccess) return strgpu_internal::fail(ctx, STRGPU_ERR_CUDA, "%s: %s", #call, cudaGetErrorString(e_)); \
  } while (0)

inline int ensure(strgpu_ctx *ctx, DevBuf &b, size_t bytes) {
  if (bytes <= b.cap) return STRGPU_OK;
  if (b.p) CU(ctx, cudaFree(b.p));
  b.p = nullptr;
  b.cap = 0;
  size_t cap = bytes + bytes / 4 + 256;
  CU(ctx, cudaMalloc(&b.p, cap));
  b.cap = cap;
  return STRGPU_OK;
}

void comm_release(strgpu_ctx *ctx);   // comm.cu
void decode_release(strgpu_ctx *ctx); // decode_kernels.cu

inline uint64_t hash_bytes(const void *p, size_t n, uint64_t h = 0xcbf29ce484222325ull) {
  const unsigned char *b = static_cast<const unsigned char *>(p);
  for (size_t i = 0; i < n; i++) h = (h ^ b[i]) * 0x100000001b3ull;
  return h ? h : 1;
}

// enqueue(stream, &launches, capturing) -> strgpu_status enqueues the work on a stream; see GraphSlot
template <typename F>
int run_graphed(strgpu_ctx *ctx, GraphSlot &g, uint64_t key, cudaStream_t user, F enqueue) {
  static const bool no_graph = getenv("STRGPU_NO_GRAPH") != nullptr;
  uint64_t l = 0;
  if (no_graph || g.broken) {
    const int rc = enqueue(user, &l, false);
    ctx->launches += l;
    return rc;
  }
  if (g.exec && g.key == key && g.ws_gen == ctx->cluster_ws.gen) {
    CU(ctx, cudaGraphLaunch(g.exec, user));
    ctx->launches += g.launches;
    return STRGPU_OK;
  }
  if (g.exec) {   // stale: other arguments, or a workspace buffer moved
    CU(ctx, cudaStreamSynchronize(user));
    cudaGraphExecDestroy(g.exec);
    g.exec = nullptr;
  }
  if (g.sized_key != key || g.sized_gen != ctx->cluster_ws.gen) {   // first call with these arguments: direct
    const int rc = enqueue(user, &l, false);
    ctx->launches += l;
    g.sized_key = key;
    g.sized_gen = ctx->cluster_ws.gen;
    return rc;
  }
  // second call: capture on the context's own stream (the caller's may be the legacy default stream, which cannot be captured)
  cudaStream_t cs = ctx->cluster_stream;
  if (cudaStreamBeginCapture(cs, cudaStreamCaptureModeThreadLocal) != cudaSuccess) {
    cudaGetLastError();
    g.broken = true;
    const int rc = enqueue(user, &l, false);
    ctx->launches += l;
    return rc;
  }
  const int rc = enqueue(cs, &l, true);
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(cs, &graph);
  if (rc != STRGPU_OK || e != cudaSuccess || !graph || cudaGraphInstantiate(&g.exec, graph, 0) != cudaSuccess) {
    cudaGetLastError();
    if (graph) cudaGraphDestroy(graph);
    g.exec = nullptr;
    g.broken = true;
    uint64_t l2 = 0;
    const int rc2 = enqueue(user, &l2, false);
    ctx->launches += l2;
    return rc2;
  }
  cudaGraphDestroy(graph);
  g.key = key;
  g.ws_gen = ctx->cluster_ws.gen;
  g.launches = l;
  CU(ctx, cudaGraphLaunch(g.exec, user));
  ctx->launches += l;
  return STRGPU_OK;
}

}  // namespace strgpu_internal
