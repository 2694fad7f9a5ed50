// Internal definition of strgpu_ctx shared by api.cu (context, scan, cluster entry points) and comm.cu (NCCL entry points).
#pragma once
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
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

}  // namespace strgpu_internal
