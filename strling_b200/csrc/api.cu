// C ABI of libstrgpu.so (include/strgpu.h): context, device buffers, streams, the submit/wait pipeline.
// There is deliberately no CPU code path for the scan or the cluster kernels: without an sm_100 device
// every compute entry point fails with STRGPU_ERR_NO_DEVICE.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <algorithm>
#include <mutex>
#include <new>
#include <numeric>
#include <vector>

#include "cluster_kernels.cuh"
#include "scan_kernels.cuh"
#include "strgpu.h"

static_assert(sizeof(strgpu_segment) == 8, "strgpu_segment layout");
static_assert(sizeof(strgpu_repeat) == 8, "strgpu_repeat layout");
static_assert(sizeof(strgpu_tread) == 24, "strgpu_tread layout");
static_assert(sizeof(strgpu_bounds) == 48, "strgpu_bounds layout");
static_assert(sizeof(strgpu_cluster_params) == 16, "strgpu_cluster_params layout");
static_assert(sizeof(strgpu_locus) == 24, "strgpu_locus layout");

#include "ctx.cuh"

using namespace strgpu_internal;

extern "C" {

const char *strgpu_version(void) { return "strling-b200 0.1.0 (STRling 0.6.0 hot path, sm_100a)"; }

const char *strgpu_error_string(int status) {
  switch (status) {
    case STRGPU_OK: return "ok";
    case STRGPU_ERR_INVALID: return "invalid argument";
    case STRGPU_ERR_CUDA: return "CUDA error";
    case STRGPU_ERR_NO_DEVICE: return "no sm_100 CUDA device (no CPU fallback exists)";
    case STRGPU_ERR_TOO_LONG: return "segment longer than supported";
    case STRGPU_ERR_BUSY: return "no free submit slot";
    case STRGPU_ERR_TICKET: return "bad ticket";
    case STRGPU_ERR_OVERFLOW: return "output capacity too small";
    case STRGPU_ERR_DATA: return "malformed input data";
    default: return "unknown status";
  }
}

static int create_impl(strgpu_ctx **out, int device);

int strgpu_create(strgpu_ctx **out, int device) {
  // a context that failed half way is destroyed here: on error *out is NULL and there is nothing for the caller to free
  const int rc = create_impl(out, device);
  if (rc != STRGPU_OK && out && *out) {
    strgpu_destroy(*out);
    *out = nullptr;
  }
  return rc;
}

static int create_impl(strgpu_ctx **out, int device) {
  if (!out) return STRGPU_ERR_INVALID;
  *out = nullptr;
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0 || device < 0 || device >= n) return STRGPU_ERR_NO_DEVICE;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return STRGPU_ERR_NO_DEVICE;
  if (prop.major != 10) return STRGPU_ERR_NO_DEVICE;  // kernels are built for sm_100a only
  strgpu_ctx *ctx = new (std::nothrow) strgpu_ctx();
  if (!ctx) return STRGPU_ERR_INVALID;
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  *out = ctx;
  CU(ctx, cudaSetDevice(device));
  CU(ctx, cudaMalloc(&ctx->d_thr, strgpu::kThrEntries * sizeof(uint16_t)));
  CU(ctx, cudaMalloc(&ctx->d_status_dev, sizeof(int)));
  {
    uint16_t luts[strgpu::kLaneLutEntries];
    strgpu::build_lane_luts(luts);
    CU(ctx, cudaMalloc(&ctx->d_luts, sizeof(luts)));
    CU(ctx, cudaMemcpy(ctx->d_luts, luts, sizeof(luts), cudaMemcpyHostToDevice));
    const char *v = getenv("STRGPU_SCAN_VARIANT");
    ctx->variant = v ? atoi(v) : 0;
  }
  CU(ctx, cudaMemset(ctx->d_status_dev, 0, sizeof(int)));
  for (auto &s : ctx->slots) {
    CU(ctx, cudaStreamCreateWithFlags(&s.stream, cudaStreamNonBlocking));
    CU(ctx, cudaEventCreateWithFlags(&s.done, cudaEventDisableTiming));
    CU(ctx, cudaMalloc(&s.d_status, sizeof(int)));
    CU(ctx, cudaMallocHost(&s.h_status, sizeof(int)));
  }
  CU(ctx, cudaStreamCreateWithFlags(&ctx->cluster_stream, cudaStreamNonBlocking));
  CU(ctx, cudaEventCreateWithFlags(&ctx->dev_list_done, cudaEventDisableTiming));
  CU(ctx, cudaMalloc(&ctx->d_cl_n, sizeof(uint32_t)));
  const double dflt[3] = {0.8, 0.8 - 0.07, 0.6};  // extract.nim:255,208,242 defaults
  return strgpu_set_proportions(ctx, dflt, 3);
}

void strgpu_destroy(strgpu_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  for (auto &s : ctx->slots) {
    if (s.stream) cudaStreamSynchronize(s.stream);
    for (DevBuf *b : {&s.seq, &s.nmask, &s.xmask, &s.segs, &s.out, &s.list})
      if (b->p) cudaFree(b->p);
    if (s.d_status) cudaFree(s.d_status);
    if (s.h_status) cudaFreeHost(s.h_status);
    if (s.done) cudaEventDestroy(s.done);
    if (s.stream) cudaStreamDestroy(s.stream);
  }
  if (ctx->cluster_graph.exec) cudaGraphExecDestroy(ctx->cluster_graph.exec);
  decode_release(ctx);
  strgpu::free_workspace(ctx->cluster_ws);
  if (ctx->cl_in.p) cudaFree(ctx->cl_in.p);
  if (ctx->cl_out.p) cudaFree(ctx->cl_out.p);
  if (ctx->cl_loci.p) cudaFree(ctx->cl_loci.p);
  if (ctx->dev_list.p) cudaFree(ctx->dev_list.p);
  if (ctx->dev_list_done) cudaEventDestroy(ctx->dev_list_done);
  if (ctx->d_cl_n) cudaFree(ctx->d_cl_n);
  if (ctx->cluster_stream) cudaStreamDestroy(ctx->cluster_stream);
  if (ctx->d_thr) cudaFree(ctx->d_thr);
  if (ctx->d_luts) cudaFree(ctx->d_luts);
  if (ctx->d_status_dev) cudaFree(ctx->d_status_dev);
  strgpu_internal::comm_release(ctx);
  delete ctx;
}

const char *strgpu_last_error(const strgpu_ctx *ctx) { return ctx ? ctx->err : "null ctx"; }

uint64_t strgpu_launch_count(const strgpu_ctx *ctx) { return ctx ? ctx->launches : 0; }

int strgpu_host_alloc(void **ptr, size_t bytes) {
  if (!ptr) return STRGPU_ERR_INVALID;
  return cudaMallocHost(ptr, bytes ? bytes : 1) == cudaSuccess ? STRGPU_OK : STRGPU_ERR_CUDA;
}
void strgpu_host_free(void *ptr) {
  if (ptr) cudaFreeHost(ptr);
}

int strgpu_set_proportions(strgpu_ctx *ctx, const double *p, int n) {
  if (!ctx || !p || n < 1 || n > STRGPU_MAX_PCLASS) return fail(ctx, STRGPU_ERR_INVALID, "set_proportions: n=%d", n);
  CU(ctx, cudaSetDevice(ctx->device));
  uint16_t *h = (uint16_t *)malloc(strgpu::kThrEntries * sizeof(uint16_t));
  if (!h) return fail(ctx, STRGPU_ERR_INVALID, "out of host memory");
  for (int cls = 0; cls <= STRGPU_MAX_PCLASS; cls++) {
    // classes past n repeat the last given one; the extra class is the 0.12 give-up bound (utils.nim:251)
    const double pv = (cls == STRGPU_MAX_PCLASS) ? 0.12 : p[cls < n ? cls : n - 1];
    for (int k = 2; k <= 6; k++)
      for (int len = 0; len < strgpu::kThrLen; len++) {
        // (read.len.float * p / k.float).int : fp64 multiply, fp64 divide, truncate (utils.nim:251,259)
        volatile double prod = (double)len * pv;
        volatile double q = prod / (double)k;
        long v = (long)q;
        if (v < 0) v = 0;
        if (v > 65535) v = 65535;
        h[(cls * 5 + (k - 2)) * strgpu::kThrLen + len] = (uint16_t)v;
      }
  }
  for (int cls = 0; cls < STRGPU_MAX_PCLASS; cls++)
    for (int len = 0; len < strgpu::kThrLen; len++) {
      uint16_t m = 65535;
      for (int k = 2; k <= 6; k++) m = std::min(m, h[(cls * 5 + (k - 2)) * strgpu::kThrLen + len]);
      h[strgpu::kThrMinOff + cls * strgpu::kThrLen + len] = m;
      for (int k = 2; k <= 6; k++) {
        const long v = ((long)h[(cls * 5 + (k - 2)) * strgpu::kThrLen + len] + 1) * (k - 1);
        h[strgpu::kThrFiltOff + (cls * strgpu::kThrLen + len) * 8 + (k - 2)] = (uint16_t)(v > 65535 ? 65535 : v);
      }
      for (int j = 5; j < 8; j++) h[strgpu::kThrFiltOff + (cls * strgpu::kThrLen + len) * 8 + j] = 65535;
    }
  // all slots idle? thresholds are read by in-flight kernels, so drain first
  for (auto &s : ctx->slots) CU(ctx, cudaStreamSynchronize(s.stream));
  cudaError_t e = cudaMemcpy(ctx->d_thr, h, strgpu::kThrEntries * sizeof(uint16_t), cudaMemcpyHostToDevice);
  free(h);
  if (e != cudaSuccess) return fail(ctx, STRGPU_ERR_CUDA, "thresholds upload: %s", cudaGetErrorString(e));
  ctx->thr_set = true;
  return STRGPU_OK;
}

size_t strgpu_seq2_bytes(uint64_t n_bases) { return (size_t)((n_bases + 3) / 4) + 8; }
size_t strgpu_nmask_bytes(uint64_t n_bases) { return (size_t)((n_bases + 31) / 32) * 4 + 8; }

}  // extern "C"

namespace {

// Common part of the two asynchronous submits: claim a slot, copy the batch in, run the scan, copy the results out.
// `uniform` != nullptr: the first uniform->n_reads segments are implicit whole reads, `segs` holds the n_desc others.
int submit_batch(strgpu_ctx *ctx, const uint8_t *seq2, uint64_t n_bases, const uint32_t *nmask, const uint32_t *xmask,
                 const strgpu_segment *segs, uint32_t n_desc, const strgpu::UniformReads *uniform, uint32_t max_len,
                 strgpu_repeat *out, int *ticket) {
  int si = -1;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    for (int i = 0; i < STRGPU_SLOTS; i++)
      if (!ctx->slots[i].busy) { si = i; break; }
    if (si >= 0) ctx->slots[si].busy = true;   // claimed: no other submit takes it; scan_wait releases it
  }
  if (si < 0) return fail(ctx, STRGPU_ERR_BUSY, "all %d submit slots in flight", STRGPU_SLOTS);
  Slot &s = ctx->slots[si];
  const uint32_t n_seg = (uniform ? uniform->n_reads : 0u) + n_desc;
  auto body = [&]() -> int {
    CU(ctx, cudaSetDevice(ctx->device));
    const size_t seq_bytes = (size_t)((n_bases + 3) / 4);
    const size_t nm_bytes = nmask ? (size_t)((n_bases + 31) / 32) * 4 : 0;
    const bool use_x = nmask && xmask;
    int rc;
    if ((rc = ensure(ctx, s.seq, seq_bytes + 16))) return rc;
    if ((rc = ensure(ctx, s.nmask, nm_bytes + 16))) return rc;
    if (use_x && (rc = ensure(ctx, s.xmask, nm_bytes + 16))) return rc;
    if ((rc = ensure(ctx, s.segs, (size_t)n_desc * sizeof(strgpu_segment) + 16))) return rc;
    if ((rc = ensure(ctx, s.out, (size_t)n_seg * sizeof(strgpu_repeat) + 16))) return rc;
    if ((rc = ensure(ctx, s.list, strgpu::scan_scratch_words(n_seg) * sizeof(uint32_t)))) return rc;
    s.n_seg = n_seg;
    s.host_out = out;
    if (n_seg) {
      CU(ctx, cudaMemcpyAsync(s.seq.p, seq2, seq_bytes, cudaMemcpyHostToDevice, s.stream));
      CU(ctx, cudaMemsetAsync((char *)s.seq.p + seq_bytes, 0, 16, s.stream));
      if (nmask) {
        CU(ctx, cudaMemcpyAsync(s.nmask.p, nmask, nm_bytes, cudaMemcpyHostToDevice, s.stream));
        CU(ctx, cudaMemsetAsync((char *)s.nmask.p + nm_bytes, 0, 16, s.stream));
      }
      if (use_x) {
        CU(ctx, cudaMemcpyAsync(s.xmask.p, xmask, nm_bytes, cudaMemcpyHostToDevice, s.stream));
        CU(ctx, cudaMemsetAsync((char *)s.xmask.p + nm_bytes, 0, 16, s.stream));
      }
      if (n_desc) CU(ctx, cudaMemcpyAsync(s.segs.p, segs, (size_t)n_desc * sizeof(strgpu_segment), cudaMemcpyHostToDevice, s.stream));
      CU(ctx, cudaMemsetAsync(s.d_status, 0, sizeof(int), s.stream));
      CU(ctx, strgpu::launch_repeat_scan((const uint32_t *)s.seq.p, nmask ? (const uint32_t *)s.nmask.p : nullptr,
                                         use_x ? (const uint32_t *)s.xmask.p : nullptr, (const strgpu_segment *)s.segs.p, n_seg,
                                         max_len, ctx->d_thr, ctx->d_luts, (strgpu_repeat *)s.out.p, s.d_status, ctx->sm_count,
                                         ctx->variant, s.stream, uniform, (uint32_t *)s.list.p));
      CU(ctx, cudaMemcpyAsync(out, s.out.p, (size_t)n_seg * sizeof(strgpu_repeat), cudaMemcpyDeviceToHost, s.stream));
      CU(ctx, cudaMemcpyAsync(s.h_status, s.d_status, sizeof(int), cudaMemcpyDeviceToHost, s.stream));
    } else {
      *s.h_status = 0;
    }
    CU(ctx, cudaEventRecord(s.done, s.stream));
    return STRGPU_OK;
  };
  const int rc = body();
  std::lock_guard<std::mutex> lk(ctx->mu);
  if (rc != STRGPU_OK) {
    s.busy = false;
    return rc;
  }
  if (n_seg) ctx->launches += strgpu::scan_launches(max_len, ctx->variant);
  *ticket = si;
  return STRGPU_OK;
}

}  // namespace

extern "C" {

int strgpu_scan_submit(strgpu_ctx *ctx, const uint8_t *seq2, uint64_t n_bases, const uint32_t *nmask, const uint32_t *xmask,
                       const strgpu_segment *segs, uint32_t n_seg, uint32_t max_len, strgpu_repeat *out, int *ticket) {
  if (!ctx || !ticket || (n_seg && (!seq2 || !segs || !out))) return fail(ctx, STRGPU_ERR_INVALID, "scan_submit: null argument");
  if (max_len > STRGPU_MAX_SEGMENT_LEN) return fail(ctx, STRGPU_ERR_TOO_LONG, "max_len %u > %d", max_len, STRGPU_MAX_SEGMENT_LEN);
  return submit_batch(ctx, seq2, n_bases, nmask, xmask, segs, n_seg, nullptr, max_len, out, ticket);
}

int strgpu_scan_reads_submit(strgpu_ctx *ctx, const uint8_t *seq2, uint32_t n_reads, uint32_t read_len, uint32_t stride_bases,
                             uint32_t pclass, const uint32_t *nmask, const uint32_t *xmask, const strgpu_segment *extra,
                             uint32_t n_extra, uint32_t extra_max_len, strgpu_repeat *out, int *ticket) {
  if (!ctx || !ticket || ((n_reads || n_extra) && (!seq2 || !out)) || (n_extra && !extra))
    return fail(ctx, STRGPU_ERR_INVALID, "scan_reads_submit: null argument");
  if ((stride_bases & 3u) || stride_bases < read_len || pclass >= STRGPU_MAX_PCLASS)
    return fail(ctx, STRGPU_ERR_INVALID, "scan_reads_submit: stride %u / read_len %u / pclass %u", stride_bases, read_len, pclass);
  if ((uint64_t)n_reads * stride_bases > 0xffffffffull) return fail(ctx, STRGPU_ERR_INVALID, "scan_reads_submit: batch addresses more than 2^32 bases");
  if (read_len > STRGPU_MAX_SEGMENT_LEN || extra_max_len > STRGPU_MAX_SEGMENT_LEN)
    return fail(ctx, STRGPU_ERR_TOO_LONG, "read_len %u / extra_max_len %u > %d", read_len, extra_max_len, STRGPU_MAX_SEGMENT_LEN);
  if (read_len > (uint32_t)strgpu::kShortMaxLen || extra_max_len > (uint32_t)strgpu::kShortMaxLen) {
    // long reads: expand into ordinary descriptors on the host and take the general path
    std::vector<strgpu_segment> all((size_t)n_reads + n_extra);
    for (uint32_t i = 0; i < n_reads; i++) {
      bool has_n = false;
      if (nmask)
        for (uint64_t b = (uint64_t)i * stride_bases; b < (uint64_t)i * stride_bases + read_len && !has_n; b++)
          has_n = (nmask[b >> 5] >> (b & 31)) & 1u;
      all[i] = strgpu_segment{i * stride_bases, (uint16_t)read_len, (uint8_t)pclass, (uint8_t)(has_n ? STRGPU_SEG_HAS_N : 0)};
    }
    for (uint32_t i = 0; i < n_extra; i++) all[n_reads + i] = extra[i];
    // the descriptor array is copied to the device before this function returns (pageable memory: the copy is staged)
    return submit_batch(ctx, seq2, (uint64_t)n_reads * stride_bases, nmask, xmask, all.data(), n_reads + n_extra, nullptr,
                        read_len > extra_max_len ? read_len : extra_max_len, out, ticket);
  }
  const strgpu::UniformReads u{n_reads, read_len, stride_bases, pclass};
  return submit_batch(ctx, seq2, (uint64_t)n_reads * stride_bases, nmask, xmask, extra, n_extra, &u, read_len, out, ticket);
}

int strgpu_scan_wait(strgpu_ctx *ctx, int ticket) {
  if (!ctx || ticket < 0 || ticket >= STRGPU_SLOTS) return fail(ctx, STRGPU_ERR_TICKET, "scan_wait: bad ticket %d", ticket);
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->slots[ticket].busy) return fail(ctx, STRGPU_ERR_TICKET, "scan_wait: bad ticket %d", ticket);
  }
  Slot &s = ctx->slots[ticket];
  cudaError_t e = cudaEventSynchronize(s.done);
  const int st = *s.h_status;
  {
    std::lock_guard<std::mutex> lk(ctx->mu);
    s.busy = false;
  }
  if (e != cudaSuccess) return fail(ctx, STRGPU_ERR_CUDA, "scan_wait: %s", cudaGetErrorString(e));
  if (st != 0) return fail(ctx, st, "scan: %s", strgpu_error_string(st));
  return STRGPU_OK;
}

int strgpu_scan(strgpu_ctx *ctx, const uint8_t *seq2, uint64_t n_bases, const uint32_t *nmask, const uint32_t *xmask,
                const strgpu_segment *segs, uint32_t n_seg, uint32_t max_len, strgpu_repeat *out) {
  int t = -1;
  int rc = strgpu_scan_submit(ctx, seq2, n_bases, nmask, xmask, segs, n_seg, max_len, out, &t);
  if (rc) return rc;
  return strgpu_scan_wait(ctx, t);
}

int strgpu_scan_device(strgpu_ctx *ctx, const void *d_seq2, const void *d_nmask, const void *d_xmask, const void *d_segs,
                       uint32_t n_seg, uint32_t max_len, void *d_out, void *cuda_stream) {
  if (!ctx || (n_seg && (!d_seq2 || !d_segs || !d_out))) return fail(ctx, STRGPU_ERR_INVALID, "scan_device: null argument");
  if (max_len > STRGPU_MAX_SEGMENT_LEN) return fail(ctx, STRGPU_ERR_TOO_LONG, "max_len %u > %d", max_len, STRGPU_MAX_SEGMENT_LEN);
  if (((uintptr_t)d_seq2 & 3) || ((uintptr_t)d_nmask & 3) || ((uintptr_t)d_xmask & 3) || ((uintptr_t)d_segs & 7) || ((uintptr_t)d_out & 7))
    return fail(ctx, STRGPU_ERR_INVALID, "scan_device: misaligned device pointer");
  if (n_seg == 0) return STRGPU_OK;
  CU(ctx, cudaSetDevice(ctx->device));
  // the scratch lists are one buffer per context: launches that use it are chained through an event, so calls on
  // different streams stay correct (they serialise)
  int rc;
  if ((rc = ensure(ctx, ctx->dev_list, strgpu::scan_scratch_words(n_seg) * sizeof(uint32_t)))) return rc;
  CU(ctx, cudaStreamWaitEvent((cudaStream_t)cuda_stream, ctx->dev_list_done, 0));
  CU(ctx, strgpu::launch_repeat_scan((const uint32_t *)d_seq2, (const uint32_t *)d_nmask, d_nmask ? (const uint32_t *)d_xmask : nullptr,
                                     (const strgpu_segment *)d_segs, n_seg, max_len, ctx->d_thr, ctx->d_luts, (strgpu_repeat *)d_out,
                                     ctx->d_status_dev, ctx->sm_count, ctx->variant, (cudaStream_t)cuda_stream, nullptr,
                                     (uint32_t *)ctx->dev_list.p));
  CU(ctx, cudaEventRecord(ctx->dev_list_done, (cudaStream_t)cuda_stream));
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->launches += strgpu::scan_launches(max_len, ctx->variant);
  return STRGPU_OK;
}

int strgpu_scan_reads_device(strgpu_ctx *ctx, const void *d_seq2, uint32_t n_reads, uint32_t read_len, uint32_t stride_bases,
                             uint32_t pclass, const void *d_nmask, const void *d_xmask, const void *d_extra, uint32_t n_extra,
                             uint32_t extra_max_len, void *d_out, void *cuda_stream) {
  if (!ctx || ((n_reads || n_extra) && (!d_seq2 || !d_out)) || (n_extra && !d_extra))
    return fail(ctx, STRGPU_ERR_INVALID, "scan_reads_device: null argument");
  if ((stride_bases & 3u) || stride_bases < read_len || pclass >= STRGPU_MAX_PCLASS)
    return fail(ctx, STRGPU_ERR_INVALID, "scan_reads_device: stride %u / read_len %u / pclass %u", stride_bases, read_len, pclass);
  if ((uint64_t)n_reads * stride_bases > 0xffffffffull) return fail(ctx, STRGPU_ERR_INVALID, "scan_reads_device: batch addresses more than 2^32 bases");
  if (read_len > (uint32_t)strgpu::kShortMaxLen || extra_max_len > (uint32_t)strgpu::kShortMaxLen)
    return fail(ctx, STRGPU_ERR_TOO_LONG, "scan_reads_device: read_len %u / extra_max_len %u > %d (use strgpu_scan_device)", read_len,
                extra_max_len, strgpu::kShortMaxLen);
  if (((uintptr_t)d_seq2 & 3) || ((uintptr_t)d_nmask & 3) || ((uintptr_t)d_xmask & 3) || ((uintptr_t)d_extra & 7) || ((uintptr_t)d_out & 7))
    return fail(ctx, STRGPU_ERR_INVALID, "scan_reads_device: misaligned device pointer");
  const uint32_t n_seg = n_reads + n_extra;
  if (n_seg == 0) return STRGPU_OK;
  CU(ctx, cudaSetDevice(ctx->device));
  int rc;
  if ((rc = ensure(ctx, ctx->dev_list, strgpu::scan_scratch_words(n_seg) * sizeof(uint32_t)))) return rc;
  CU(ctx, cudaStreamWaitEvent((cudaStream_t)cuda_stream, ctx->dev_list_done, 0));
  const strgpu::UniformReads u{n_reads, read_len, stride_bases, pclass};
  CU(ctx, strgpu::launch_repeat_scan((const uint32_t *)d_seq2, (const uint32_t *)d_nmask, d_nmask ? (const uint32_t *)d_xmask : nullptr,
                                     (const strgpu_segment *)d_extra, n_seg, read_len > extra_max_len ? read_len : extra_max_len,
                                     ctx->d_thr, ctx->d_luts, (strgpu_repeat *)d_out, ctx->d_status_dev, ctx->sm_count, ctx->variant,
                                     (cudaStream_t)cuda_stream, &u, (uint32_t *)ctx->dev_list.p));
  CU(ctx, cudaEventRecord(ctx->dev_list_done, (cudaStream_t)cuda_stream));
  std::lock_guard<std::mutex> lk(ctx->mu);
  ctx->launches += strgpu::scan_launches(read_len, ctx->variant);
  return STRGPU_OK;
}

int strgpu_device_status(strgpu_ctx *ctx, void *cuda_stream) {
  if (!ctx) return STRGPU_ERR_INVALID;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaStreamSynchronize((cudaStream_t)cuda_stream));
  int st = 0;
  CU(ctx, cudaMemcpy(&st, ctx->d_status_dev, sizeof(int), cudaMemcpyDeviceToHost));
  if (st != 0) {
    CU(ctx, cudaMemset(ctx->d_status_dev, 0, sizeof(int)));
    return fail(ctx, st, "device: %s", strgpu_error_string(st));
  }
  return STRGPU_OK;
}

int strgpu_cluster_device(strgpu_ctx *ctx, const void *d_treads, uint32_t n, const strgpu_cluster_params *params, void *d_out,
                          uint32_t cap, void *d_n_out, void *cuda_stream) {
  if (!ctx || !params || !d_n_out || (n && !d_treads) || (cap && !d_out))
    return fail(ctx, STRGPU_ERR_INVALID, "cluster_device: null argument");
  if (((uintptr_t)d_treads & 7) || ((uintptr_t)d_out & 7) || ((uintptr_t)d_n_out & 3))
    return fail(ctx, STRGPU_ERR_INVALID, "cluster_device: misaligned device pointer");
  CU(ctx, cudaSetDevice(ctx->device));
  const strgpu_cluster_params p = *params;
  uint64_t key = hash_bytes(&d_treads, sizeof(d_treads));
  key = hash_bytes(&n, sizeof(n), key);
  key = hash_bytes(&p, sizeof(p), key);          // 16 bytes, no padding
  key = hash_bytes(&d_out, sizeof(d_out), key);
  key = hash_bytes(&cap, sizeof(cap), key);
  key = hash_bytes(&d_n_out, sizeof(d_n_out), key);
  return run_graphed(ctx, ctx->cluster_graph, key, (cudaStream_t)cuda_stream, [&](cudaStream_t st, uint64_t *launches, bool) -> int {
    CU(ctx, strgpu::run_cluster(ctx->cluster_ws, (const strgpu_tread *)d_treads, n, p, (strgpu_bounds *)d_out, cap, (uint32_t *)d_n_out, st,
                                launches));
    return STRGPU_OK;
  });
}

int strgpu_cluster(strgpu_ctx *ctx, const strgpu_tread *treads, uint32_t n, const strgpu_cluster_params *params,
                   strgpu_bounds *out, uint32_t cap, uint32_t *n_out) {
  return strgpu_cluster_loci(ctx, treads, n, params, nullptr, 0, out, cap, n_out);
}

int strgpu_cluster_loci(strgpu_ctx *ctx, const strgpu_tread *treads, uint32_t n, const strgpu_cluster_params *params,
                        strgpu_locus *loci, uint32_t n_loci, strgpu_bounds *out, uint32_t cap, uint32_t *n_out) {
  if (!ctx || !params || !n_out || (n && !treads) || (cap && !out) || (n_loci && !loci))
    return fail(ctx, STRGPU_ERR_INVALID, "cluster: null argument");
  *n_out = 0;
  CU(ctx, cudaSetDevice(ctx->device));
  int rc;
  if ((rc = ensure(ctx, ctx->cl_in, (size_t)n * sizeof(strgpu_tread) + 16))) return rc;
  if ((rc = ensure(ctx, ctx->cl_out, (size_t)cap * sizeof(strgpu_bounds) + 16))) return rc;
  cudaStream_t st = ctx->cluster_stream;
  if (n) CU(ctx, cudaMemcpyAsync(ctx->cl_in.p, treads, (size_t)n * sizeof(strgpu_tread), cudaMemcpyHostToDevice, st));
  strgpu::LociArgs la;
  std::vector<strgpu::DevLocus> hl;
  std::vector<uint32_t> chain_start;
  if (n_loci) {
    // group the loci by bucket key, keeping file order inside a bucket (the reference handles them sequentially)
    std::vector<uint32_t> order(n_loci);
    std::iota(order.begin(), order.end(), 0u);
    hl.resize(n_loci);
    std::vector<strgpu::DevLocus> key(n_loci);
    for (uint32_t i = 0; i < n_loci; i++) {
      key[i].hi = strgpu::tid_key_host(loci[i].tid);
      key[i].mid = strgpu::unit_rank_host(loci[i].repeat);
      key[i].left_most = loci[i].left_most;
      key[i].right_most = loci[i].right_most;
      key[i].orig = i;
      loci[i].n_left = loci[i].n_right = loci[i].n_total = 0;
    }
    std::stable_sort(order.begin(), order.end(), [&](uint32_t a, uint32_t b) {
      return key[a].hi != key[b].hi ? key[a].hi < key[b].hi : key[a].mid < key[b].mid;
    });
    for (uint32_t i = 0; i < n_loci; i++) {
      hl[i] = key[order[i]];
      if (i == 0 || hl[i].hi != hl[i - 1].hi || hl[i].mid != hl[i - 1].mid) chain_start.push_back(i);
    }
    chain_start.push_back(n_loci);
    if ((rc = ensure(ctx, ctx->cl_loci, hl.size() * sizeof(strgpu::DevLocus) + chain_start.size() * 4 + (size_t)n_loci * 6 + 64))) return rc;
    char *base = (char *)ctx->cl_loci.p;
    const size_t off_chain = (hl.size() * sizeof(strgpu::DevLocus) + 15) & ~(size_t)15;
    const size_t off_counts = (off_chain + chain_start.size() * 4 + 15) & ~(size_t)15;
    CU(ctx, cudaMemcpyAsync(base, hl.data(), hl.size() * sizeof(strgpu::DevLocus), cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemcpyAsync(base + off_chain, chain_start.data(), chain_start.size() * 4, cudaMemcpyHostToDevice, st));
    CU(ctx, cudaMemsetAsync(base + off_counts, 0, (size_t)n_loci * 6, st));
    la.d_loci = (const strgpu::DevLocus *)base;
    la.d_chain_start = (const uint32_t *)(base + off_chain);
    la.n_chains = (uint32_t)chain_start.size() - 1;
    la.d_counts = (uint16_t *)(base + off_counts);
  }
  if (n == 0 && n_loci == 0) return STRGPU_OK;
  CU(ctx, strgpu::run_cluster(ctx->cluster_ws, (const strgpu_tread *)ctx->cl_in.p, n, *params, (strgpu_bounds *)ctx->cl_out.p, cap,
                              ctx->d_cl_n, st, &ctx->launches, n_loci && n ? &la : nullptr));
  if (n_loci && n) {
    std::vector<uint16_t> counts((size_t)n_loci * 3);
    CU(ctx, cudaMemcpyAsync(counts.data(), la.d_counts, counts.size() * 2, cudaMemcpyDeviceToHost, st));
    CU(ctx, cudaStreamSynchronize(st));
    for (uint32_t i = 0; i < n_loci; i++) {
      loci[i].n_left = counts[3 * i];
      loci[i].n_right = counts[3 * i + 1];
      loci[i].n_total = counts[3 * i + 2];
    }
  }
  uint32_t produced = 0;
  CU(ctx, cudaMemcpyAsync(&produced, ctx->d_cl_n, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
  CU(ctx, cudaStreamSynchronize(st));
  *n_out = produced;
  const uint32_t take = produced < cap ? produced : cap;
  if (take) CU(ctx, cudaMemcpy(out, ctx->cl_out.p, (size_t)take * sizeof(strgpu_bounds), cudaMemcpyDeviceToHost));
  if (produced > cap) return fail(ctx, STRGPU_ERR_OVERFLOW, "cluster: %u records produced, capacity %u", produced, cap);
  return STRGPU_OK;
}

}  // extern "C"
