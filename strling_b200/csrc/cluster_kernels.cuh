// Launch interface of the cluster kernels (K2 tread_sort, K3 cluster_chain, K4 cluster_bounds).  Internal to libstrgpu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "strgpu.h"

namespace strgpu {

struct SortRec {  // 16-byte radix-sort record: key fields + the tread's index in input order
  uint32_t hi;    // tid + 1 (unplaced reads first)
  uint32_t mid;   // repeat unit as a base-6 number of its chars, memcmp order (0 < A < C < G < T)
  uint32_t pos;
  uint32_t idx;
};

// Growable device workspace owned by the ctx; every pointer is device memory.
struct ClusterWorkspace {
  void *buf[20] = {nullptr};
  size_t cap[20] = {0};
  uint64_t gen = 0;   // bumped whenever a buffer is reallocated: captured CUDA graphs that hold the old pointers are stale
};

// Loci for assign_reads_locus, already grouped on the host: loci of one bucket are consecutive ("chain") and keep
// their file order; chain_start has n_chains + 1 entries.
struct DevLocus {
  uint32_t hi, mid;              // bucket key in sort-record encoding
  uint32_t left_most, right_most;
  uint32_t orig;                 // index in the caller's array
};
struct LociArgs {
  const DevLocus *d_loci = nullptr;
  const uint32_t *d_chain_start = nullptr;
  uint32_t n_chains = 0;
  uint16_t *d_counts = nullptr;  // [orig][3] = n_left, n_right, n_total
};
uint32_t unit_rank_host(const char repeat[6]);
uint32_t tid_key_host(int32_t tid);

// Enqueues the whole cluster path for n treads already in device memory on `stream`; nothing is read back and nothing
// synchronises (sizes only the device knows stay in device memory).  Returns cudaSuccess or the failing call's error;
// *launches is incremented per kernel launch.  d_n_out receives the number of records produced (may exceed cap).
// n must be below 2^29.
cudaError_t run_cluster(ClusterWorkspace &ws, const strgpu_tread *d_treads, uint32_t n, const strgpu_cluster_params &p,
                        strgpu_bounds *d_out, uint32_t cap, uint32_t *d_n_out, cudaStream_t stream, uint64_t *launches,
                        const LociArgs *loci = nullptr, const uint32_t *d_n_in = nullptr);
// d_n_in (optional, device): the actual record count when only the device knows it; n is then the upper bound that sizes
// grids and workspace (sharded clustering: the records a rank receives from its peers)

// helpers of the sharded path (comm.cu)
cudaError_t scan_u32(ClusterWorkspace &ws, const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *d_total, cudaStream_t st,
                     uint64_t *launches);
cudaError_t sort_bounds_device(ClusterWorkspace &ws, const strgpu_bounds *d_in, uint32_t n_max, const uint32_t *d_n, strgpu_bounds *d_out,
                               uint32_t cap, uint32_t *d_n_out, cudaStream_t st, uint64_t *launches);

void free_workspace(ClusterWorkspace &ws);

}  // namespace strgpu
