// BGZF inflate on the GPU ("optionally GPU inflate"): `strling extract --gpu-inflate` ships the COMPRESSED
// blocks of a batch over PCIe (about a third of the bytes), inflates them here and copies the inflated records back for
// the host's record walk / staging, taking the dominant host cost (inflate: ~80 % of the CPU time of the extract pipeline)
// off the CPU cores.  Stands in for the htslib inflate the reference reaches through hts-nim (extract.nim:275-329).
//
// One BGZF block (<= 64 KiB, one complete DEFLATE stream) per CTA of one warp; lane 0 decodes with the SAME code the host
// uses (host/inflate_fast.hpp compiled as a device function: bit buffer, 11-bit / 8-bit decode tables in shared memory,
// byte-wise loads and copies because device memory accesses must be aligned), so the CPU tests of that decoder cover the
// arithmetic; what is CUDA-specific is only this file.  DEFLATE is sequential inside a block, the parallelism is across the
// thousands of blocks of a batch: ~15 KB of tables per CTA lets 15 CTAs share an SM, 2220 blocks in flight on 148 SMs.
#include <cuda_runtime.h>

#include <cstdint>

#include "../host/inflate_fast.hpp"
#include "ctx.cuh"

static_assert(sizeof(strgpu_bgzf_block) == 24, "strgpu_bgzf_block layout");

namespace strgpu_internal {

struct Decode {
  cudaStream_t stream = nullptr;
  DevBuf comp, out, blocks;
  int *d_status = nullptr;   // [0]: 1 + index of the first block that failed (0: none), [1]: its decoder status
  int *h_status = nullptr;   // pinned
  std::mutex mu;
};

__global__ void __launch_bounds__(32) inflate_bgzf_blocks(const uint8_t *__restrict__ comp, const strgpu_bgzf_block *__restrict__ blocks, uint32_t n_blocks,
                                                          uint8_t *__restrict__ out, uint64_t out_base, int *status) {
  __shared__ strling::infl::Tables tables;
  if (threadIdx.x != 0) return;
  const uint32_t b = blockIdx.x;
  if (b >= n_blocks) return;
  const strgpu_bgzf_block blk = blocks[b];
  if (blk.isize == 0) return;
  tables.fixed_built = false;
  const int rc = strling::infl::inflate_block(tables, comp + blk.in_off, blk.csize, out + (blk.out_off - out_base), blk.isize);
  if (rc != strling::infl::kOk && atomicCAS(&status[0], 0, (int)b + 1) == 0) status[1] = rc;
}

void decode_release(strgpu_ctx *ctx) {
  Decode *d = ctx->decode;
  if (!d) return;
  if (d->stream) cudaStreamSynchronize(d->stream);
  for (DevBuf *b : {&d->comp, &d->out, &d->blocks})
    if (b->p) cudaFree(b->p);
  if (d->d_status) cudaFree(d->d_status);
  if (d->h_status) cudaFreeHost(d->h_status);
  if (d->stream) cudaStreamDestroy(d->stream);
  delete d;
  ctx->decode = nullptr;
}

}  // namespace strgpu_internal

using namespace strgpu_internal;

extern "C" int strgpu_inflate_bgzf(strgpu_ctx *ctx, const uint8_t *comp, size_t comp_bytes, const strgpu_bgzf_block *blocks, uint32_t n_blocks, uint8_t *out,
                                   size_t out_bytes) {
  if (!ctx) return STRGPU_ERR_INVALID;
  if (n_blocks == 0) return STRGPU_OK;
  if (!comp || !blocks || !out) return fail(ctx, STRGPU_ERR_INVALID, "inflate_bgzf: null argument");
  uint64_t lo = UINT64_MAX, hi = 0;
  for (uint32_t i = 0; i < n_blocks; i++) {
    const strgpu_bgzf_block &b = blocks[i];
    if (b.isize > 65536u || b.in_off > comp_bytes || (uint64_t)b.csize > comp_bytes - b.in_off || b.out_off > out_bytes ||
        (uint64_t)b.isize > out_bytes - b.out_off)
      return fail(ctx, STRGPU_ERR_INVALID, "inflate_bgzf: block %u lies outside the buffers", i);
    if (!b.isize) continue;
    lo = b.out_off < lo ? b.out_off : lo;
    hi = b.out_off + b.isize > hi ? b.out_off + b.isize : hi;
  }
  if (hi <= lo) return STRGPU_OK;  // empty blocks only
  CU(ctx, cudaSetDevice(ctx->device));
  if (!ctx->decode) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->decode) {
      Decode *d = new (std::nothrow) Decode();
      if (!d) return fail(ctx, STRGPU_ERR_INVALID, "inflate_bgzf: out of memory");
      ctx->decode = d;
      CU(ctx, cudaStreamCreateWithFlags(&d->stream, cudaStreamNonBlocking));
      CU(ctx, cudaMalloc(&d->d_status, 2 * sizeof(int)));
      CU(ctx, cudaMallocHost(&d->h_status, 2 * sizeof(int)));
    }
  }
  Decode *d = ctx->decode;
  std::lock_guard<std::mutex> lk(d->mu);  // one inflate call at a time per context; scans on the submit slots run beside it
  int rc;
  if ((rc = ensure(ctx, d->comp, comp_bytes + 64))) return rc;
  if ((rc = ensure(ctx, d->out, (size_t)(hi - lo) + 64))) return rc;
  if ((rc = ensure(ctx, d->blocks, (size_t)n_blocks * sizeof(strgpu_bgzf_block)))) return rc;
  cudaStream_t st = d->stream;
  CU(ctx, cudaMemsetAsync(d->d_status, 0, 2 * sizeof(int), st));
  CU(ctx, cudaMemsetAsync(static_cast<uint8_t *>(d->comp.p) + comp_bytes, 0, 64, st));  // the decoder may read 8 bytes past a stream
  CU(ctx, cudaMemcpyAsync(d->comp.p, comp, comp_bytes, cudaMemcpyHostToDevice, st));
  CU(ctx, cudaMemcpyAsync(d->blocks.p, blocks, (size_t)n_blocks * sizeof(strgpu_bgzf_block), cudaMemcpyHostToDevice, st));
  inflate_bgzf_blocks<<<n_blocks, 32, 0, st>>>(static_cast<const uint8_t *>(d->comp.p), static_cast<const strgpu_bgzf_block *>(d->blocks.p), n_blocks,
                                               static_cast<uint8_t *>(d->out.p), lo, d->d_status);
  CU(ctx, cudaGetLastError());
  CU(ctx, cudaMemcpyAsync(out + lo, d->out.p, (size_t)(hi - lo), cudaMemcpyDeviceToHost, st));
  CU(ctx, cudaMemcpyAsync(d->h_status, d->d_status, 2 * sizeof(int), cudaMemcpyDeviceToHost, st));
  CU(ctx, cudaStreamSynchronize(st));
  {
    std::lock_guard<std::mutex> lk2(ctx->mu);
    ctx->launches += 1;
  }
  if (d->h_status[0] != 0)
    return fail(ctx, STRGPU_ERR_DATA, "inflate_bgzf: block %d is not a valid DEFLATE stream of the stated sizes (decoder status %d)", d->h_status[0] - 1,
                d->h_status[1]);
  return STRGPU_OK;
}
