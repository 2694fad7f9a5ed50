// K2 tread_sort, K3 cluster_chain, K4 cluster_bounds: STRling's cluster loop (call.nim:118-130,223-235;
// merge.nim:125-187; cluster.nim:175-374; callclusters.nim:52-66) on sm_100a.  Integer work only.
//
//   K2  stable LSD radix sort of 16-byte (tid, unit, position, index) records: one warp owns one contiguous
//       chunk, ranks 32 records at a time with match.any, so equal keys keep `.bin` order exactly like the
//       reference's bucket append + stable sort by position (call.nim:124-130).  The bits that vary across the
//       batch (position span, unit span, tid span) are concatenated into one virtual key on the DEVICE (sort_plan), so
//       only ceil(varying bits / 8) passes do work and the host never reads anything back: every kernel of the
//       cluster path takes its sizes from device memory and the whole path is enqueued without a synchronisation.
//   K3  next(i) = first read NOT absorbed by a cluster started at read i (trcluster, cluster.nim:323-362) is a
//       pure function of i: <= 8 explicit steps while the median-of-first-9 still moves, then one binary search.
//       Bucket heads then chase next() to mark cluster starts; the chase is cut into 256-record pieces (exit tables per
//       piece, a piece-to-piece hop per bucket, a marking sweep per piece) so it stays parallel for huge buckets.  A piece
//       is handled by one warp with its next[] slice in shared memory: exits and marks by pointer doubling (8 rounds).
//   K4  one thread per chained cluster: trim, left/right_most, min_support + anchor test, split_cluster, bounds,
//       filters, has_per_sample_reads.  CountTable.largest ties follow Nim's slot order (hashWangYi1 + linear
//       probing + growth), emulated in a per-cluster scratch region.
#include "cluster_kernels.cuh"

#include <algorithm>
#include <cstdio>
#include <mutex>

namespace strgpu {

namespace {

constexpr uint32_t kFull = 0xffffffffu;
constexpr int SOFT_LEFT = 0, SOFT_RIGHT = 1, SOFT_NONE = 3;

// ------------------------------------------------------------------------------------------- scan
constexpr int kScanThreads = 256;
constexpr int kScanItems = 8;
constexpr int kScanTile = kScanThreads * kScanItems;

__device__ __forceinline__ uint32_t block_exclusive_scan(uint32_t v, uint32_t *total) {
  // exclusive scan of one value per thread across a 256-thread block
  __shared__ uint32_t warp_sums[kScanThreads / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(kFull, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    uint32_t w = lane < kScanThreads / 32 ? warp_sums[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(kFull, w, o);
      if (lane >= o) w += t;
    }
    if (lane < kScanThreads / 32) warp_sums[lane] = w;
  }
  __syncthreads();
  const uint32_t base = warp ? warp_sums[warp - 1] : 0;
  if (total) *total = warp_sums[kScanThreads / 32 - 1];
  __syncthreads();
  return base + inc - v;
}

// Exclusive scan of n uint32 in two launches.  n comes from device memory (*d_n, capped at n_max); `gate` (may be null)
// names a device flag: when it is zero both kernels return at once (an inactive radix pass).
//   scan_block_sums : per-tile totals; the LAST block to finish (ticket) scans the totals in place and writes the grand total
//   scan_apply      : rescans each tile with its offset
__global__ void __launch_bounds__(kScanThreads) scan_block_sums(const uint32_t *in, const uint32_t *d_n, uint32_t n_max, uint32_t *block_sums,
                                                                uint32_t *ticket, uint32_t *total_out, const uint32_t *gate) {
  if (gate && *gate == 0u) return;
  const uint32_t n = min(*d_n, n_max);
  const uint32_t n_blocks = (n + kScanTile - 1) / kScanTile;
  if (blockIdx.x >= n_blocks) {
    if (n == 0 && blockIdx.x == 0 && threadIdx.x == 0 && total_out) *total_out = 0;
    return;
  }
  const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; j++)
    if (base + j < n) s += in[base + j];
  uint32_t total;
  block_exclusive_scan(s, &total);
  __shared__ bool last;
  if (threadIdx.x == 0) {
    block_sums[blockIdx.x] = total;
    __threadfence();
    last = atomicAdd(ticket, 1u) == n_blocks - 1u;
  }
  __syncthreads();
  if (!last) return;
  __threadfence();
  uint32_t carry = 0;
  for (uint32_t b0 = 0; b0 < n_blocks; b0 += kScanThreads) {
    const uint32_t i = b0 + threadIdx.x;
    const uint32_t v = i < n_blocks ? reinterpret_cast<volatile uint32_t *>(block_sums)[i] : 0;
    uint32_t tot;
    const uint32_t ex = block_exclusive_scan(v, &tot);
    if (i < n_blocks) block_sums[i] = carry + ex;
    carry += tot;
  }
  if (threadIdx.x == 0) {
    if (total_out) *total_out = carry;
    *ticket = 0;   // ready for the next scan on this stream
  }
}

__global__ void __launch_bounds__(kScanThreads) scan_apply(const uint32_t *in, const uint32_t *d_n, uint32_t n_max, const uint32_t *block_offsets,
                                                           uint32_t *out, const uint32_t *gate) {
  if (gate && *gate == 0u) return;
  const uint32_t n = min(*d_n, n_max);
  const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  if (blockIdx.x * kScanTile >= n) return;
  uint32_t v[kScanItems];
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; j++) {
    v[j] = (base + j < n) ? in[base + j] : 0;
    s += v[j];
  }
  uint32_t ex = block_exclusive_scan(s, nullptr) + block_offsets[blockIdx.x];
#pragma unroll
  for (int j = 0; j < kScanItems; j++) {
    if (base + j < n) out[base + j] = ex;
    ex += v[j];
  }
}

// ------------------------------------------------------------------------------------------- K2 sort
// Sort keys are kept NARROW because the radix sort only walks the bits that vary: tid + 1 (unplaced reads, tid -1, sort first
// as 0; a BAM has no other negative tid -- one would sort last) and the unit as a base-6 number of its six characters
// (0 < A < C < G < T < anything else: memcmp order of the zero-padded array[6, char] for units over ACGT, 16 bits).
__host__ __device__ __forceinline__ uint32_t tid_key(int32_t tid) { return (uint32_t)tid + 1u; }
__device__ __forceinline__ bool key_unplaced(uint32_t hi) { return (int32_t)(hi - 1u) < 0; }
__device__ __forceinline__ uint32_t unit_rank(char c) {
  switch (c) {
    case 0: return 0;
    case 'A': return 1;
    case 'C': return 2;
    case 'G': return 3;
    case 'T': return 4;
    default: return 5;
  }
}

// The top three bits of SortRec::idx carry the tread's split (Soft, cluster.nim:14-20), so that the chain and bounds kernels work
// on the sorted 16-byte records alone (position, split) and never gather the 24-byte treads; n stays below 2^29.
constexpr int kIdxBits = 29;
constexpr uint32_t kIdxMask = (1u << kIdxBits) - 1u;
struct Reads {
  const SortRec *recs;
  const strgpu_tread *treads;   // input order, for the few fields read once per cluster (tid, repeat) or on demand (sample)
  __device__ __forceinline__ uint32_t pos(uint32_t i) const { return recs[i].pos; }
  __device__ __forceinline__ int split(uint32_t i) const { return (int)(recs[i].idx >> kIdxBits); }
  __device__ __forceinline__ const strgpu_tread &tread(uint32_t i) const { return treads[recs[i].idx & kIdxMask]; }
};

// ---- device-side sort plan.  Small device words (d_small), all uint32:
//   [0..2] bits that vary across the batch in pos / unit / tid (make_sort_records)      [3] clusters chained (K3)
//   [5] records entering K3 (n, or what assign_reads_locus left)    [6] scan ticket     [7] scratch bump pointer (K4)
//   [8] 2 * clusters (length of the K4 compaction scan)   [9] length of the radix count matrix   [10] clusters listed for cluster_bounds   [15] which ping-pong buffer holds the sorted records
//   [16 + p] radix pass p does work     [32 + p] its source buffer
//   [48 + f], [51 + f], [54 + f] lowest varying bit / width / offset in the virtual key of field f (0 pos, 1 unit, 2 tid)
constexpr int kSmallWords = 64;
constexpr int kMaxPasses = 10;   // 32 + 16 + 32 varying bits at most, 8 per pass
enum { SM_VAR = 0, SM_NCLUSTERS = 3, SM_NCUR = 5, SM_TICKET = 6, SM_BUMP = 7, SM_N2 = 8, SM_COUNTLEN = 9, SM_HEAVY = 10, SM_FINAL = 15, SM_ACTIVE = 16, SM_SRC = 32,
       SM_LO = 48, SM_WIDTH = 51, SM_OFF = 54 };

// the record count of this run: the host's n, or -- sharded clustering, where only the device knows how many records a rank
// received -- the device value (never more than the host's n, which sizes the grids and the workspace)
__global__ void init_small(uint32_t *small, uint32_t n_host, const uint32_t *d_n_in) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  small[SM_NCUR] = d_n_in ? min(*d_n_in, n_host) : n_host;
}

__global__ void make_sort_records(const strgpu_tread *__restrict__ treads, const uint32_t *__restrict__ small, SortRec *__restrict__ recs,
                                  uint32_t *__restrict__ varbits) {
  const uint32_t n = small[SM_NCUR];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  SortRec r{0, 0, 0, 0};
  uint32_t d0 = 0, d1 = 0, d2 = 0;
  if (i < n) {
    const strgpu_tread t = treads[i];
    r.hi = tid_key(t.tid);
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 6; j++) m = m * 6u + unit_rank(t.repeat[j]);
    r.mid = m;
    r.pos = t.position;
    r.idx = i | ((uint32_t)(t.split < 8 ? t.split : 7) << kIdxBits);   // K3 / K4 read the split from the sort record
    recs[i] = r;
    const strgpu_tread f = treads[0];
    uint32_t fm = 0;
#pragma unroll
    for (int j = 0; j < 6; j++) fm = fm * 6u + unit_rank(f.repeat[j]);
    d0 = r.pos ^ f.position;
    d1 = r.mid ^ fm;
    d2 = r.hi ^ tid_key(f.tid);
  }
  // OR-reduce over the block, then one atomic per block and field
  __shared__ uint32_t blk[3];
  if (threadIdx.x < 3) blk[threadIdx.x] = 0;
  __syncthreads();
  d0 = __reduce_or_sync(kFull, d0);
  d1 = __reduce_or_sync(kFull, d1);
  d2 = __reduce_or_sync(kFull, d2);
  if ((threadIdx.x & 31) == 0) {
    if (d0) atomicOr(&blk[0], d0);
    if (d1) atomicOr(&blk[1], d1);
    if (d2) atomicOr(&blk[2], d2);
  }
  __syncthreads();
  if (threadIdx.x < 3 && blk[threadIdx.x]) atomicOr(&varbits[threadIdx.x], blk[threadIdx.x]);
}

__global__ void sort_plan(uint32_t *small, uint32_t count_len) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  small[SM_COUNTLEN] = count_len;
  uint32_t off = 0;
  for (int f = 0; f < 3; f++) {
    const uint32_t var = small[SM_VAR + f];
    const uint32_t lo = var ? (uint32_t)(__ffs(var) - 1) : 0u;
    const uint32_t width = var ? (uint32_t)(32 - __clz(var)) - lo : 0u;
    small[SM_LO + f] = lo;
    small[SM_WIDTH + f] = width;
    small[SM_OFF + f] = off;
    off += width;
  }
  uint32_t src = 0;
  for (int p = 0; p < kMaxPasses; p++) {
    const uint32_t active = (uint32_t)(8 * p) < off ? 1u : 0u;
    small[SM_ACTIVE + p] = active;
    small[SM_SRC + p] = src;
    src ^= active;
  }
  small[SM_FINAL] = src;
}

// the pass's 8-bit digit of the virtual key: the varying spans of pos, unit and tid, concatenated (pos least significant)
struct DigitPlan {
  uint32_t lo[3], mask[3];
  int rel[3];   // bit offset of the field's span relative to the digit's bit 0
};
__device__ __forceinline__ DigitPlan load_digit_plan(const uint32_t *small, int pass) {
  DigitPlan d;
#pragma unroll
  for (int f = 0; f < 3; f++) {
    const uint32_t w = small[SM_WIDTH + f];
    d.lo[f] = small[SM_LO + f];
    d.mask[f] = w >= 32u ? 0xffffffffu : ((1u << w) - 1u);
    d.rel[f] = (int)small[SM_OFF + f] - 8 * pass;
  }
  return d;
}
__device__ __forceinline__ uint32_t digit_of(const SortRec &r, const DigitPlan &d) {
  uint32_t out = 0;
#pragma unroll
  for (int f = 0; f < 3; f++) {
    const uint32_t v = ((f == 0 ? r.pos : (f == 1 ? r.mid : r.hi)) >> d.lo[f]) & d.mask[f];
    const int rel = d.rel[f];
    if (rel >= 0) out |= rel < 8 ? (v << rel) : 0u;
    else out |= rel > -32 ? (v >> (-rel)) : 0u;
  }
  return out & 0xffu;
}

// One CTA owns one tile of kTile consecutive records.  The scatter first sorts the tile by digit in shared memory (warp w
// ranks its 512 records 32 at a time with match.any, so equal digits keep their order; a per-warp / per-digit prefix makes
// the ranks tile-wide), then copies the tile out position by position: records of one digit leave for consecutive
// addresses, so a pass writes whole sectors instead of one 16-byte record per sector.
constexpr int kSortWarps = 8;
constexpr uint32_t kTile = 4096;
constexpr uint32_t kWarpSpan = kTile / kSortWarps;   // 512 records per warp

__global__ void __launch_bounds__(kSortWarps * 32) radix_histogram(const SortRec *__restrict__ buf0, const SortRec *__restrict__ buf1,
                                                                   uint32_t n_tiles, const uint32_t *__restrict__ small, int pass,
                                                                   uint32_t *__restrict__ counts) {
  if (small[SM_ACTIVE + pass] == 0u) return;
  const SortRec *in = small[SM_SRC + pass] ? buf1 : buf0;
  const DigitPlan dp = load_digit_plan(small, pass);
  __shared__ uint32_t hist[256];
  hist[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t n = small[SM_NCUR];
  const uint32_t beg = min(n, blockIdx.x * kTile), end = min(n, beg + kTile);
  for (uint32_t i = beg + threadIdx.x; i < end; i += kSortWarps * 32) atomicAdd(&hist[digit_of(in[i], dp)], 1u);
  __syncthreads();
  counts[(size_t)threadIdx.x * n_tiles + blockIdx.x] = hist[threadIdx.x];
}

__global__ void __launch_bounds__(kSortWarps * 32) radix_scatter(SortRec *__restrict__ buf0, SortRec *__restrict__ buf1,
                                                                 uint32_t n_tiles, const uint32_t *__restrict__ small, int pass,
                                                                 const uint32_t *__restrict__ offsets) {
  if (small[SM_ACTIVE + pass] == 0u) return;
  const bool flip = small[SM_SRC + pass] != 0u;
  const SortRec *in = flip ? buf1 : buf0;
  SortRec *out = flip ? buf0 : buf1;
  const DigitPlan dp = load_digit_plan(small, pass);
  extern __shared__ __align__(16) unsigned char sort_smem[];
  SortRec *stage = reinterpret_cast<SortRec *>(sort_smem);                                 // kTile records
  uint32_t(*whist)[256] = reinterpret_cast<uint32_t(*)[256]>(sort_smem + kTile * sizeof(SortRec));   // [warp][digit]
  uint32_t *dbase = reinterpret_cast<uint32_t *>(whist + kSortWarps);                      // tile-local first position of a digit
  uint32_t *gbase = dbase + 256;                                                           // its first position in the output
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n = small[SM_NCUR];
  const uint32_t beg = min(n, blockIdx.x * kTile), end = min(n, beg + kTile);
  if (beg >= end) return;   // a tile past the device-side record count
  const uint32_t wbeg = min(end, beg + warp * kWarpSpan), wend = min(end, wbeg + kWarpSpan);
  for (int w = 0; w < kSortWarps; w++) whist[w][tid] = 0;
  __syncthreads();
  // pass A: digit counts per warp
  for (uint32_t i = wbeg + lane; i < wend; i += 32) atomicAdd(&whist[warp][digit_of(in[i], dp)], 1u);
  __syncthreads();
  {
    // digit tid: exclusive prefix over the warps, tile-wide total, then the exclusive scan of the totals
    uint32_t run = 0;
    for (int w = 0; w < kSortWarps; w++) {
      const uint32_t c = whist[w][tid];
      whist[w][tid] = run;
      run += c;
    }
    const uint32_t ex = block_exclusive_scan(run, nullptr);
    dbase[tid] = ex;
    gbase[tid] = offsets[(size_t)tid * n_tiles + blockIdx.x];
    for (int w = 0; w < kSortWarps; w++) whist[w][tid] += ex;
  }
  __syncthreads();
  // pass B: stable ranks, records to their tile-local position
  const uint32_t lane_lt = (1u << lane) - 1u;
  for (uint32_t base = wbeg; base < wend; base += 32) {
    const uint32_t i = base + lane;
    const bool valid = i < wend;
    SortRec r{0, 0, 0, 0};
    if (valid) r = in[i];
    const uint32_t d = valid ? digit_of(r, dp) : (0x100u + (uint32_t)lane);
    const uint32_t grp = __match_any_sync(kFull, d);
    uint32_t local = 0;
    if (valid) local = whist[warp][d] + __popc(grp & lane_lt);  // lane order == input order: stable
    __syncwarp();
    if (valid && lane == 31 - __clz(grp)) whist[warp][d] += __popc(grp);
    __syncwarp();
    if (valid) stage[local] = r;
  }
  __syncthreads();
  for (uint32_t j = tid; j < end - beg; j += kSortWarps * 32) {
    const SortRec r = stage[j];
    const uint32_t d = digit_of(r, dp);
    out[gbase[d] + (j - dbase[d])] = r;
  }
}
constexpr int kScatterSmem = kTile * sizeof(SortRec) + kSortWarps * 256 * 4 + 2 * 256 * 4;

// after assign_reads_locus: the records kept (left in SM_NCLUSTERS by the compaction scan) are what K3 works on
__global__ void set_ncur(uint32_t *small) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  small[SM_NCUR] = small[SM_NCLUSTERS];
  small[SM_NCLUSTERS] = 0;
}

__device__ __forceinline__ const SortRec *sorted_recs(const SortRec *buf0, const SortRec *buf1, const uint32_t *small) {
  return small[SM_FINAL] ? buf1 : buf0;
}

// ------------------------------------------------------------------------------------------- C10 assign_reads_locus
// One thread per chain (= the loci of one bucket, in file order): callclusters.nim:14-50 on the sorted records with
// removal expressed as marks.  Earlier loci change what later loci of the same bucket see, hence the serial chain.
__global__ void assign_loci(const SortRec *__restrict__ buf0, const SortRec *__restrict__ buf1, const uint32_t *__restrict__ small,
                            uint32_t n,
                            const DevLocus *__restrict__ loci, const uint32_t *__restrict__ chain_start, uint32_t n_chains,
                            uint32_t *__restrict__ removed, uint16_t *__restrict__ counts) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chains) return;
  const SortRec *recs = sorted_recs(buf0, buf1, small);
  const DevLocus first = loci[chain_start[c]];
  // bucket [bs, be): records whose (hi, mid) equals the key
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t m = lo + ((hi - lo) >> 1);
    const SortRec r = recs[m];
    if (r.hi < first.hi || (r.hi == first.hi && r.mid < first.mid)) lo = m + 1; else hi = m;
  }
  const uint32_t bs = lo;
  hi = n;
  while (lo < hi) {
    const uint32_t m = lo + ((hi - lo) >> 1);
    const SortRec r = recs[m];
    if (r.hi == first.hi && r.mid == first.mid) lo = m + 1; else hi = m;
  }
  const uint32_t be = lo;
  for (uint32_t j = chain_start[c]; j < chain_start[c + 1]; j++) {
    const DevLocus L = loci[j];
    uint16_t n_left = 0, n_right = 0, n_total = 0;
    if (be > bs) {
      const uint32_t lm1 = L.left_most == 0 ? 0u : L.left_most - 1u;
      uint32_t a = bs, b = be;
      while (a < b) { const uint32_t m = a + ((b - a) >> 1); if (recs[m].pos < lm1) a = m + 1; else b = m; }
      const uint32_t li = a;
      b = be;
      while (a < b) { const uint32_t m = a + ((b - a) >> 1); if (recs[m].pos <= L.right_most) a = m + 1; else b = m; }
      const uint32_t ri = a;
      for (uint32_t i = li; i < ri; i++) {
        if (removed[i]) continue;
        removed[i] = 1;
        n_total++;
        const int sp = (int)(recs[i].idx >> kIdxBits);
        if (sp == SOFT_RIGHT) n_right++; else if (sp == SOFT_LEFT) n_left++;
      }
      for (uint32_t i = ri; i < be; i++)  // the element at `ri` of the shrunken bucket is dropped too (callclusters.nim:35-36)
        if (!removed[i]) { removed[i] = 1; break; }
    }
    counts[3 * L.orig + 0] = n_left;
    counts[3 * L.orig + 1] = n_right;
    counts[3 * L.orig + 2] = n_total;
  }
}

__global__ void invert_flags(const uint32_t *__restrict__ removed, uint32_t n, uint32_t *__restrict__ keep) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) keep[i] = removed[i] ? 0u : 1u;
}

__global__ void compact_sorted(SortRec *__restrict__ buf0, SortRec *__restrict__ buf1, const uint32_t *__restrict__ small,
                               const uint32_t *__restrict__ keep, const uint32_t *__restrict__ dst, uint32_t n) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !keep[i]) return;
  const bool fin = small[SM_FINAL] != 0u;
  const SortRec *recs = fin ? buf1 : buf0;
  SortRec *recs_out = fin ? buf0 : buf1;   // the kept records land in the other ping-pong buffer
  recs_out[dst[i]] = recs[i];
}

__device__ __forceinline__ bool same_bucket(const SortRec &a, const SortRec &b) { return a.hi == b.hi && a.mid == b.mid; }

// recs of K3: the sorted buffer, or the other one when assign_reads_locus compacted into it (flip == 1)
__device__ __forceinline__ const SortRec *k3_recs(const SortRec *buf0, const SortRec *buf1, const uint32_t *small, int flip) {
  return ((small[SM_FINAL] != 0u) != (flip != 0)) ? buf1 : buf0;
}

__global__ void cluster_next(const SortRec *__restrict__ buf0, const SortRec *__restrict__ buf1, const uint32_t *__restrict__ small, int flip,
                             uint32_t max_dist, uint32_t *__restrict__ next, uint32_t *__restrict__ bucket_end, uint32_t *__restrict__ head,
                             uint32_t *__restrict__ entry_flag) {
  const uint32_t n = small[SM_NCUR];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const SortRec *recs = k3_recs(buf0, buf1, small, flip);
  head[i] = 0;
  entry_flag[i] = 0;
  const SortRec me = recs[i];
  // end of my (tid, unit) bucket: first index whose bucket key differs (records are sorted).  Galloping first: most buckets
  // are short, so the search costs log(bucket) probes instead of log(n)
  uint32_t lo = i + 1, hi = n;
  for (uint32_t step = 1; lo < n; step *= 2) {
    const uint32_t probe = min(n - 1u, i + step);
    if (same_bucket(recs[probe], me)) { lo = probe + 1; if (probe == n - 1u) break; }
    else { hi = probe; break; }
  }
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (same_bucket(recs[mid], me)) lo = mid + 1; else hi = mid;
  }
  const uint32_t be = lo;
  bucket_end[i] = be;
  if (key_unplaced(me.hi)) {  // unplaced: the whole bucket is one Cluster (cluster.nim:369-371)
    next[i] = be;
    return;
  }
  uint32_t j = i + 1;
  uint32_t m = 1;
  while (j < be) {
    const uint32_t mm = m < 9u ? m : 9u;
    const uint32_t thr = recs[i + ((mm - 1u) >> 1)].pos + max_dist + 100u;  // posmed + max_dist + 100, uint32 wrap
    if (m >= 9u) {  // the median of the first 9 no longer moves: first position > thr ends the cluster
      uint32_t a = j, b = be;
      for (uint32_t step = 1; a < be; step *= 2) {   // galloping: clusters are short
        const uint32_t probe = min(be - 1u, j + step - 1u);
        if (recs[probe].pos <= thr) { a = probe + 1; if (probe == be - 1u) break; }
        else { b = probe; break; }
      }
      while (a < b) {
        const uint32_t mid = a + ((b - a) >> 1);
        if (recs[mid].pos <= thr) a = mid + 1; else b = mid;
      }
      j = a;
      break;
    }
    if (recs[j].pos <= thr) { j++; m++; } else break;
  }
  next[i] = j;
}

// Cluster starts = the chain bucket_start -> next -> next ... of every bucket.  The walk is cut into pieces of kSeg records
// so that no thread takes more than a few steps (a bucket can hold 10^5 reads):
//   piece_exits : for every record i of a piece, exit[i] = the first index >= the piece's end that the chain through i reaches
//                 (or the index at which the chain leaves its bucket)
//   bucket_entry: one thread per bucket hops piece to piece (entry -> exit[entry]) and flags each piece's entry point
//   piece_mark  : marks the cluster starts inside every piece, from its flagged entry points
// A piece is one warp's job with its slice of next[] in shared memory; both sweeps are pointer doubling (8 rounds of 8
// records per lane) instead of a serial walk at L2 latency.
constexpr uint32_t kSeg = 256;
constexpr int kPieceWarps = 8;
constexpr uint32_t kTerm = 0x80000000u;   // "this value is final" (record counts stay below 2^31)

__global__ void __launch_bounds__(kPieceWarps * 32) piece_exits(const uint32_t *__restrict__ next, const uint32_t *__restrict__ bucket_end,
                                                                const uint32_t *__restrict__ small, uint32_t *__restrict__ exit_of) {
  __shared__ uint32_t ex[kPieceWarps][kSeg];
  const uint32_t n = small[SM_NCUR];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lo = (blockIdx.x * kPieceWarps + warp) * kSeg;
  if (lo >= n) return;
  const uint32_t hi = min(n, lo + kSeg);
  uint32_t *e = ex[warp];
  // a chain never leaves its bucket: next[i] <= bucket_end[i]; it is final when it reaches the piece end or the bucket end
  for (uint32_t t = lane; t < kSeg; t += 32) {
    const uint32_t i = lo + t;
    uint32_t v = kTerm;
    if (i < hi) {
      const uint32_t nx = next[i];
      v = (nx >= hi || nx >= bucket_end[i]) ? (nx | kTerm) : nx;
    }
    e[t] = v;
  }
  __syncwarp();
  for (int round = 0; round < 8; round++) {   // path lengths double every round: 2^8 = kSeg
    uint32_t nv[kSeg / 32];
#pragma unroll
    for (int q = 0; q < (int)(kSeg / 32); q++) {
      const uint32_t v = e[q * 32 + lane];
      nv[q] = (v & kTerm) ? v : e[v - lo];
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < (int)(kSeg / 32); q++) e[q * 32 + lane] = nv[q];
    __syncwarp();
  }
  for (uint32_t t = lane; t < kSeg; t += 32)
    if (lo + t < hi) exit_of[lo + t] = e[t] & ~kTerm;
}

__global__ void bucket_entry(const SortRec *__restrict__ buf0, const SortRec *__restrict__ buf1, const uint32_t *__restrict__ small, int flip,
                             const uint32_t *__restrict__ bucket_end, const uint32_t *__restrict__ exit_of, uint32_t *__restrict__ entry_flag) {
  const uint32_t n = small[SM_NCUR];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const SortRec *recs = k3_recs(buf0, buf1, small, flip);
  if (i > 0 && same_bucket(recs[i - 1], recs[i])) return;  // bucket starts only
  const uint32_t be = bucket_end[i];
  uint32_t h = i;
  while (h < be) {
    entry_flag[h] = 1;   // h is a cluster start and the point where the chain enters (or re-enters) a piece
    h = exit_of[h];
  }
}

__global__ void __launch_bounds__(kPieceWarps * 32) piece_mark(const uint32_t *__restrict__ next, const uint32_t *__restrict__ bucket_end,
                                                               const uint32_t *__restrict__ small, const uint32_t *__restrict__ entry_flag,
                                                               uint32_t *__restrict__ head) {
  __shared__ uint32_t jump[kPieceWarps][kSeg];
  __shared__ uint32_t mark[kPieceWarps][kSeg];
  const uint32_t n = small[SM_NCUR];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t lo = (blockIdx.x * kPieceWarps + warp) * kSeg;
  if (lo >= n) return;
  const uint32_t hi = min(n, lo + kSeg);
  uint32_t *jp = jump[warp], *mk = mark[warp];
  uint32_t any = 0;
  for (uint32_t t = lane; t < kSeg; t += 32) {
    const uint32_t i = lo + t;
    uint32_t j = kTerm, m = 0;
    if (i < hi) {
      const uint32_t nx = next[i];
      j = (nx >= hi || nx >= bucket_end[i]) ? kTerm : nx - lo;   // successor inside the piece and the bucket, else none
      m = entry_flag[i];
    }
    jp[t] = j;
    mk[t] = m;
    any |= m;
  }
  if (__ballot_sync(kFull, any != 0u) == 0u) return;   // no chain enters this piece (warp-uniform)
  __syncwarp();
  for (int round = 0; round < 8; round++) {
    // every marked record marks its 2^round-th successor; then the jump distance doubles
    uint32_t nj[kSeg / 32];
#pragma unroll
    for (int q = 0; q < (int)(kSeg / 32); q++) {
      const uint32_t t = q * 32 + lane;
      const uint32_t j = jp[t];
      if (mk[t] && !(j & kTerm)) mk[j] = 1;     // concurrent writers all store 1
      nj[q] = (j & kTerm) ? j : jp[j];
    }
    __syncwarp();
#pragma unroll
    for (int q = 0; q < (int)(kSeg / 32); q++) jp[q * 32 + lane] = nj[q];
    __syncwarp();
  }
  for (uint32_t t = lane; t < kSeg; t += 32)
    if (lo + t < hi && mk[t]) head[lo + t] = 1;
}

__global__ void cluster_fill(const uint32_t *__restrict__ head, const uint32_t *__restrict__ cid, const uint32_t *__restrict__ next,
                             uint32_t *__restrict__ small, uint32_t *__restrict__ cl_start, uint32_t *__restrict__ cl_end) {
  const uint32_t n = small[SM_NCUR];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) small[SM_N2] = 2u * small[SM_NCLUSTERS];
  if (i >= n || !head[i]) return;
  cl_start[cid[i]] = i;
  cl_end[cid[i]] = next[i];
}

// ------------------------------------------------------------------------------------------- K4 bounds
// Nim 1.6 hashes.nim hashWangYi1 (CountTable[uint32] hashes keys as uint64)
__device__ __forceinline__ uint64_t hi_xor_lo(uint64_t a, uint64_t b) { return __umul64hi(a, b) ^ (a * b); }
__device__ __forceinline__ uint64_t hash_wangyi1(uint64_t x) {
  const uint64_t P0 = 0xa0761d6478bd642fULL, P1 = 0xe7037ed1a0b428dbULL, P58 = 0xeb44accab455d165ULL ^ 8ULL;
  return hi_xor_lo(hi_xor_lo(P0, x ^ P1), P58);
}

struct Slot {
  uint32_t key;
  uint32_t val;
};

enum ClipMode { kSplitLeft, kSplitRight, kBoundsLeft, kBoundsRight };

__device__ __forceinline__ bool clip_selected(int split, uint32_t position, ClipMode mode, uint32_t cm, uint32_t mcd) {
  // bounds(): cluster.nim:193,197 (int32 casts and adds wrap); the elif makes "right" exclude nothing extra
  // because a read has a single split value.
  switch (mode) {
    case kSplitLeft: return split == SOFT_LEFT;
    case kSplitRight: return split == SOFT_RIGHT;
    case kBoundsLeft: return split == SOFT_LEFT && (int32_t)position < (int32_t)(cm + mcd);
    default: return split == SOFT_RIGHT && (int32_t)position > (int32_t)(cm - mcd);
  }
}

struct Largest {
  uint32_t key;
  uint32_t val;
  uint32_t n_distinct;
  uint32_t n_selected;
};

// CountTable over the selected reads' positions in [a, b) + `largest`.  Positions ascend, so equal keys are
// consecutive among the selected reads and first-insertion order is ascending key order.
__device__ Largest count_largest(const Reads &reads, uint32_t a, uint32_t b, ClipMode mode, uint32_t cm,
                                 uint32_t mcd, Slot *scratch) {
  Largest res{0, 0, 0, 0};
  uint32_t cur_key = 0, cur_val = 0, n_at_max = 0;
  for (uint32_t i = a; i < b; i++) {
    const SortRec rr = reads.recs[i];
    const uint32_t r_position = rr.pos;
    if (!clip_selected((int)(rr.idx >> kIdxBits), r_position, mode, cm, mcd)) continue;
    res.n_selected++;
    if (cur_val && r_position == cur_key) { cur_val++; }
    else {
      if (cur_val) {
        if (cur_val > res.val) { res.val = cur_val; res.key = cur_key; n_at_max = 1; }
        else if (cur_val == res.val) n_at_max++;
      }
      cur_key = r_position; cur_val = 1; res.n_distinct++;
    }
  }
  if (cur_val) {
    if (cur_val > res.val) { res.val = cur_val; res.key = cur_key; n_at_max = 1; }
    else if (cur_val == res.val) n_at_max++;
  }
  if (n_at_max <= 1) return res;
  // tie: replay the inserts into an emulated Nim CountTable (initCountTable(8) -> 16 slots; grow x2 when
  // cap*2 < counter*3 or cap - counter < 4, re-inserting in old slot order) and take the first max slot.
  const uint32_t half = 16u + 4u * (b - a);
  Slot *tab = scratch, *alt = scratch + half;
  uint32_t cap = 16, counter = 0;
  for (uint32_t s = 0; s < cap; s++) tab[s] = Slot{0, 0};
  auto raw_insert = [](Slot *t, uint32_t c, uint32_t key, uint32_t val) {
    uint32_t h = (uint32_t)(hash_wangyi1((uint64_t)key) & (uint64_t)(c - 1));
    while (t[h].val != 0) h = (h + 1) & (c - 1);
    t[h] = Slot{key, val};
  };
  auto insert = [&](uint32_t key, uint32_t val) {
    if (cap * 2 < counter * 3 || cap - counter < 4) {
      const uint32_t ncap = cap * 2;
      for (uint32_t s = 0; s < ncap; s++) alt[s] = Slot{0, 0};
      for (uint32_t s = 0; s < cap; s++)
        if (tab[s].val != 0) raw_insert(alt, ncap, tab[s].key, tab[s].val);
      Slot *t = tab; tab = alt; alt = t;
      cap = ncap;
    }
    raw_insert(tab, cap, key, val);
    counter++;
  };
  cur_val = 0;
  for (uint32_t i = a; i < b; i++) {
    const SortRec rr = reads.recs[i];
    const uint32_t r_position = rr.pos;
    if (!clip_selected((int)(rr.idx >> kIdxBits), r_position, mode, cm, mcd)) continue;
    if (cur_val && r_position == cur_key) { cur_val++; }
    else {
      if (cur_val) insert(cur_key, cur_val);
      cur_key = r_position; cur_val = 1;
    }
  }
  if (cur_val) insert(cur_key, cur_val);
  uint32_t mi = 0;
  for (uint32_t h = 1; h < cap; h++)
    if (tab[mi].val < tab[h].val) mi = h;
  res.key = tab[mi].key;
  res.val = tab[mi].val;
  return res;
}

__device__ __forceinline__ uint32_t posmed(const Reads &reads, uint32_t a, uint32_t b) {
  const uint32_t n = b - a;  // cluster.nim:59-62 : reads[int(min(9, n)/2 - 0.5)]
  const uint32_t m = n < 9u ? n : 9u;
  return reads.pos(a + ((m - 1u) >> 1));
}

// merge.nim:18-25 : does any sample own >= supporting reads of [a, b)?  (scratch as an open-addressing counter)
__device__ bool per_sample_support(const Reads &reads, uint32_t a, uint32_t b, int supporting, Slot *scratch) {
  if (supporting <= 0) return true;
  const uint32_t n = b - a;
  if (n < (uint32_t)supporting) return false;
  uint32_t cap = 16;
  while (cap < 2 * n) cap <<= 1;
  for (uint32_t s = 0; s < cap; s++) scratch[s] = Slot{0, 0};
  for (uint32_t i = a; i < b; i++) {
    const uint32_t key = (uint32_t)reads.tread(i).sample;
    uint32_t h = (key * 2654435761u) & (cap - 1);
    while (scratch[h].val != 0 && scratch[h].key != key) h = (h + 1) & (cap - 1);
    scratch[h].key = key;
    if ((int)++scratch[h].val >= supporting) return true;
  }
  return false;
}

// cluster.nim:175-250 + callclusters.nim:52-66.  Returns false when the cluster is dropped.
__device__ bool bounds_of(const Reads &reads, uint32_t a, uint32_t b, uint32_t cl_left_most,
                          uint32_t cl_right_most, const strgpu_cluster_params &p, Slot *scratch, strgpu_bounds &out) {
  const uint32_t n = b - a;
  if (n >= 65535u) return false;  // callclusters.nim:53-55
  const strgpu_tread first = reads.tread(a);
  out.tid = first.tid;
#pragma unroll
  for (int j = 0; j < 6; j++) out.repeat[j] = first.repeat[j];
  const uint32_t cm = reads.pos(a + (n >> 1));
  out.center_mass = cm;
  const uint32_t mcd = p.max_clip_dist;
  const Largest ll = count_largest(reads, a, b, kBoundsLeft, cm, mcd, scratch);
  const Largest rr = count_largest(reads, a, b, kBoundsRight, cm, mcd, scratch);
  out.n_left = (uint16_t)ll.n_selected;
  out.n_right = (uint16_t)rr.n_selected;
  out.n_total = (uint16_t)n;
  uint32_t left = 0, right = 0;
  if (ll.n_distinct > 0 && ll.val > 1) left = ll.key;
  if (rr.n_distinct > 0 && rr.val > 1) right = rr.key;
  if (left == 0) left = cm;
  if (right == 0) right = left + 1;
  if (left >= right) {
    if (out.n_left > 0 && out.n_right > 0) { const uint32_t t = left; left = right; right = t; }
    else left = right - 1;
  }
  // positions ascend, so posns.min()/max() are the first / last read
  uint32_t lm = cl_left_most > 0 ? cl_left_most : reads.pos(a);
  uint32_t rm = cl_right_most > 0 ? cl_right_most : reads.pos(b - 1);
  if (lm > left) lm = left;
  if (rm < right) rm = right;
  out.left = left; out.right = right; out.left_most = lm; out.right_most = rm;
  out.first_read = a; out.n_reads = n; out.reserved = 0;
  if (right - left > 1000u) return false;  // callclusters.nim:57-59
  if (out.n_left < p.min_clip) return false;
  if (out.n_right < p.min_clip) return false;
  if ((uint16_t)(out.n_right + out.n_left) < p.min_clip_total) return false;
  return true;
}

// K4 in two kernels.  cluster_screen (one thread per chained cluster): the cheap part every cluster goes through -- unplaced
// buckets, trim, min_support, has_anchor -- after which most clusters are done (a lone noise read is a cluster too); the
// ones that go on are appended to a dense list.  cluster_bounds (one thread per listed cluster): split_cluster,
// has_per_sample_reads, bounds and the filters.  With the list, the 32 threads of a warp all hold a real cluster instead of
// one of them walking 30 reads while 31 wait.
struct Heavy {
  uint32_t c, a, b;   // cluster id, reads [a, b) after trim
};

__global__ void cluster_screen(const SortRec *__restrict__ buf0, const SortRec *__restrict__ buf1, int flip,
                               const strgpu_tread *__restrict__ treads, const uint32_t *__restrict__ cl_start,
                               const uint32_t *__restrict__ cl_end, uint32_t *__restrict__ small, strgpu_cluster_params p,
                               strgpu_bounds *__restrict__ out2, uint32_t *__restrict__ valid2, Heavy *__restrict__ heavy) {
  const uint32_t n_clusters = small[SM_NCLUSTERS];
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_clusters) return;
  uint32_t a = cl_start[c];
  const uint32_t b = cl_end[c];
  valid2[2 * c] = 0;
  valid2[2 * c + 1] = 0;
  const Reads reads{k3_recs(buf0, buf1, small, flip), treads};
  if (key_unplaced(reads.recs[a].hi)) {  // unplaced bucket: call.nim:226-228 records len per unit; merge.nim:175-176 skips
    if (!p.merge_mode) {
      const strgpu_tread first = reads.tread(a);
      strgpu_bounds u;
      u.tid = -1; u.left = u.left_most = u.right = u.right_most = u.center_mass = 0;
      u.n_left = u.n_right = u.n_total = 0;
#pragma unroll
      for (int j = 0; j < 6; j++) u.repeat[j] = first.repeat[j];
      u.first_read = a; u.n_reads = b - a; u.reserved = 0;
      out2[2 * c] = u;
      valid2[2 * c] = 1;
    }
    return;
  }
  // trim (cluster.nim:252-257): lo is computed once from the untrimmed cluster
  if (b - a > 1) {
    const long long lo_l = (long long)posmed(reads, a, b) - (long long)(p.window + 100u);
    const uint32_t lo = lo_l > 0 ? (uint32_t)lo_l : 0u;
    while (b - a > 1 && reads.pos(a) < lo) a++;
  }
  if ((long long)(b - a) < (long long)p.min_support) return;
  bool anchor = false;
  for (uint32_t i = a; i < b && !anchor; i++) anchor = reads.split(i) == SOFT_NONE;
  if (!anchor) return;
  heavy[atomicAdd(small + SM_HEAVY, 1u)] = Heavy{c, a, b};
}

__global__ void cluster_bounds(const SortRec *__restrict__ buf0, const SortRec *__restrict__ buf1, int flip,
                               const strgpu_tread *__restrict__ treads, uint32_t *__restrict__ small, strgpu_cluster_params p,
                               const Heavy *__restrict__ heavy, Slot *__restrict__ scratch_all, uint32_t scratch_slots,
                               strgpu_bounds *__restrict__ out2, uint32_t *__restrict__ valid2) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= small[SM_HEAVY]) return;
  const Heavy h = heavy[t];
  const uint32_t c = h.c, a = h.a, b = h.b;
  const Reads reads{k3_recs(buf0, buf1, small, flip), treads};
  const uint32_t max_dist = p.window;
  const uint32_t pm = posmed(reads, a, b);
  const uint32_t last = reads.pos(b - 1), firstp = reads.pos(a);
  const uint32_t hi_edge = pm + max_dist, lo_edge = pm - max_dist;  // uint32 wrap (cluster.nim:343-344)
  const uint32_t cl_right_most = last > hi_edge ? last : hi_edge;
  const uint32_t cl_left_most = firstp < lo_edge ? firstp : lo_edge;
  Slot *scratch;
  {
    // scratch for the CountTable replay (16 + 4 * reads slots, twice) and the per-sample counter (<= 4 * reads + 16): bump-allocated,
    // so its size follows the clusters that need it (at most reads / max(1, min_support) of them) instead of the worst case
    const uint32_t need = 32u + 8u * (b - a);
    const uint32_t off = atomicAdd(small + SM_BUMP, need);
    if (off + need > scratch_slots) return;   // cannot happen: the pool is sized for every cluster that can get here (run_cluster)
    scratch = scratch_all + off;
  }
  // split_cluster (cluster.nim:283-320)
  uint32_t sub_a[2] = {a, 0}, sub_b[2] = {b, 0}, sub_lm[2] = {cl_left_most, 0}, sub_rm[2] = {cl_right_most, 0};
  int n_sub = 1;
  {
    const Largest ll = count_largest(reads, a, b, kSplitLeft, 0, 0, scratch);
    const Largest rl = count_largest(reads, a, b, kSplitRight, 0, 0, scratch);
    if (ll.n_distinct > 0 && rl.n_distinct > 0 && rl.key < ll.key && (long long)rl.val >= p.min_support &&
        (long long)ll.val >= p.min_support && (double)ll.val / (double)ll.n_distinct > 0.5 &&
        (double)rl.val / (double)rl.n_distinct > 0.5) {
      const uint32_t mid = (uint32_t)(0.5 + ((double)rl.key + (double)ll.key) / 2.0);
      uint32_t m = a;
      while (m < b && reads.pos(m) < mid) m++;
      sub_a[0] = a; sub_b[0] = m; sub_lm[0] = 0; sub_rm[0] = mid - 1;
      sub_a[1] = m; sub_b[1] = b; sub_lm[1] = mid; sub_rm[1] = 0;
      n_sub = 2;
    }
  }
  for (int s = 0; s < n_sub; s++) {
    if (p.merge_mode && !per_sample_support(reads, sub_a[s], sub_b[s], p.min_support, scratch)) continue;
    strgpu_bounds bd;
    if (bounds_of(reads, sub_a[s], sub_b[s], sub_lm[s], sub_rm[s], p, scratch, bd)) {
      out2[2 * c + s] = bd;
      valid2[2 * c + s] = 1;
    }
  }
}

__global__ void compact_bounds(const strgpu_bounds *__restrict__ out2, const uint32_t *__restrict__ valid2,
                               const uint32_t *__restrict__ dst_idx, const uint32_t *__restrict__ small, strgpu_bounds *__restrict__ out,
                               uint32_t cap) {
  const uint32_t n2 = small[SM_N2];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n2 || !valid2[i]) return;
  const uint32_t d = dst_idx[i];
  if (d < cap) out[d] = out2[i];
}

// ------------------------------------------------------------------------------------------- host driver
enum { WS_RECS_A, WS_RECS_B, WS_COUNTS, WS_BLOCKSUMS, WS_SORTED, WS_NEXT, WS_BEND, WS_HEAD, WS_CID, WS_CLSTART, WS_CLEND,
       WS_SCRATCH, WS_OUT2, WS_VALID2, WS_DST2, WS_SMALL, WS_SORTED_B, WS_ENTRY, WS_HEAVY, WS_SCANAUX, WS_COUNT_ };
static_assert(WS_COUNT_ <= (int)(sizeof(ClusterWorkspace::buf) / sizeof(void *)), "workspace slots");

cudaError_t ws_ensure(ClusterWorkspace &ws, int which, size_t bytes) {
  if (bytes <= ws.cap[which]) return cudaSuccess;
  ws.gen++;
  if (ws.buf[which]) cudaFree(ws.buf[which]);
  ws.buf[which] = nullptr;
  ws.cap[which] = 0;
  const size_t cap = bytes + bytes / 8 + 256;
  cudaError_t e = cudaMalloc(&ws.buf[which], cap);
  if (e == cudaSuccess) ws.cap[which] = cap;
  return e;
}

#define CK(call)                           \
  do {                                     \
    cudaError_t e_ = (call);               \
    if (e_ != cudaSuccess) return e_;      \
  } while (0)

// exclusive scan of *d_n (<= n_max) uint32 (in -> out, may alias), total to *d_total (device); see scan_block_sums
cudaError_t exclusive_scan(ClusterWorkspace &ws, const uint32_t *in, uint32_t *out, const uint32_t *d_n, uint32_t n_max, uint32_t *d_total,
                           const uint32_t *gate, cudaStream_t st, uint64_t *launches) {
  if (n_max == 0) return cudaSuccess;
  const uint32_t blocks = (n_max + kScanTile - 1) / kScanTile;
  uint32_t *bs = (uint32_t *)ws.buf[WS_BLOCKSUMS];
  uint32_t *small = (uint32_t *)ws.buf[WS_SMALL];
  scan_block_sums<<<blocks, kScanThreads, 0, st>>>(in, d_n, n_max, bs, small + SM_TICKET, d_total, gate);
  scan_apply<<<blocks, kScanThreads, 0, st>>>(in, d_n, n_max, bs, out, gate);
  *launches += 2;
  return cudaGetLastError();
}

}  // namespace

uint32_t tid_key_host(int32_t tid) { return tid_key(tid); }

uint32_t unit_rank_host(const char repeat[6]) {
  uint32_t m = 0;
  for (int j = 0; j < 6; j++) {
    uint32_t r;
    switch (repeat[j]) {
      case 0: r = 0; break;
      case 'A': r = 1; break;
      case 'C': r = 2; break;
      case 'G': r = 3; break;
      case 'T': r = 4; break;
      default: r = 5;
    }
    m = m * 6u + r;
  }
  return m;
}

void free_workspace(ClusterWorkspace &ws) {
  for (int i = 0; i < 20; i++) {
    if (ws.buf[i]) cudaFree(ws.buf[i]);
    ws.buf[i] = nullptr;
    ws.cap[i] = 0;
  }
}

namespace {
__global__ void set_words(uint32_t *dst, uint32_t a, uint32_t b) {
  if (threadIdx.x == 0 && blockIdx.x == 0) { dst[0] = a; dst[1] = b; }
}

__global__ void bounds_sort_records(const strgpu_bounds *__restrict__ in, const uint32_t *small, SortRec *__restrict__ recs, uint32_t *varbits) {
  const uint32_t n = small[SM_NCUR];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t d1 = 0, d2 = 0;
  if (i < n) {
    auto key_of = [&](const strgpu_bounds &b, uint32_t &hi, uint32_t &mid) {
      hi = tid_key(b.tid);
      mid = 0;
#pragma unroll
      for (int j = 0; j < 6; j++) mid = mid * 6u + unit_rank(b.repeat[j]);
    };
    SortRec r;
    key_of(in[i], r.hi, r.mid);
    r.pos = 0;
    r.idx = i;
    recs[i] = r;
    uint32_t fh, fm;
    key_of(in[0], fh, fm);
    d1 = r.mid ^ fm;
    d2 = r.hi ^ fh;
  }
  d1 = __reduce_or_sync(kFull, d1);
  d2 = __reduce_or_sync(kFull, d2);
  if ((threadIdx.x & 31) == 0) {
    if (d1) atomicOr(&varbits[1], d1);
    if (d2) atomicOr(&varbits[2], d2);
  }
}

__global__ void bounds_gather(const strgpu_bounds *__restrict__ in, const SortRec *__restrict__ buf0, const SortRec *__restrict__ buf1,
                              const uint32_t *__restrict__ small, strgpu_bounds *__restrict__ out, uint32_t cap, uint32_t *__restrict__ n_out) {
  const uint32_t n = small[SM_NCUR];
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) *n_out = n;
  if (i >= n || i >= cap) return;
  out[i] = in[sorted_recs(buf0, buf1, small)[i].idx];
}
}  // namespace

// exclusive scan of n (host-known) uint32 on `st`, for the callers outside this file (the owner partition of comm.cu)
cudaError_t scan_u32(ClusterWorkspace &ws, const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *d_total, cudaStream_t st,
                     uint64_t *launches) {
  if (n == 0) return cudaSuccess;
  CK(ws_ensure(ws, WS_BLOCKSUMS, ((size_t)(n + kScanTile - 1) / kScanTile + 1) * 4));
  CK(ws_ensure(ws, WS_SCANAUX, 64));
  uint32_t *aux = (uint32_t *)ws.buf[WS_SCANAUX];   // [0] n, [1] ticket
  set_words<<<1, 32, 0, st>>>(aux, n, 0u);
  const uint32_t blocks = (n + kScanTile - 1) / kScanTile;
  uint32_t *bs = (uint32_t *)ws.buf[WS_BLOCKSUMS];
  scan_block_sums<<<blocks, kScanThreads, 0, st>>>(in, aux, n, bs, aux + 1, d_total, nullptr);
  scan_apply<<<blocks, kScanThreads, 0, st>>>(in, aux, n, bs, out, nullptr);
  *launches += 3;
  return cudaGetLastError();
}

// Stable sort of *d_n (<= n_max) cluster records by (tid, unit): the last step of sharded clustering (a bucket lives on one
// rank, so inside a bucket the rank-major concatenation is already in position order).  Writes min(*d_n, cap) records, *d_n_out = *d_n.
cudaError_t sort_bounds_device(ClusterWorkspace &ws, const strgpu_bounds *d_in, uint32_t n_max, const uint32_t *d_n, strgpu_bounds *d_out,
                               uint32_t cap, uint32_t *d_n_out, cudaStream_t st, uint64_t *launches) {
  if (n_max == 0) return cudaMemsetAsync(d_n_out, 0, 4, st);
  const int T = 256;
  const uint32_t nb = (n_max + T - 1) / T;
  const uint32_t n_chunks = (n_max + kTile - 1) / kTile;
  CK(ws_ensure(ws, WS_RECS_A, (size_t)n_max * sizeof(SortRec)));
  CK(ws_ensure(ws, WS_RECS_B, (size_t)n_max * sizeof(SortRec)));
  CK(ws_ensure(ws, WS_SMALL, kSmallWords * 4));
  CK(ws_ensure(ws, WS_COUNTS, (size_t)256 * n_chunks * 4));
  CK(ws_ensure(ws, WS_BLOCKSUMS, ((size_t)(256 * n_chunks + kScanTile - 1) / kScanTile + 1) * 4));
  {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    static std::once_flag attr_once[64];
    std::call_once(attr_once[dev], []() { cudaFuncSetAttribute(radix_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, kScatterSmem); });
  }
  uint32_t *d_small = (uint32_t *)ws.buf[WS_SMALL];
  CK(cudaMemsetAsync(d_small, 0, kSmallWords * 4, st));
  SortRec *ra = (SortRec *)ws.buf[WS_RECS_A], *rb = (SortRec *)ws.buf[WS_RECS_B];
  init_small<<<1, 32, 0, st>>>(d_small, n_max, d_n);
  bounds_sort_records<<<nb, T, 0, st>>>(d_in, d_small, ra, d_small + SM_VAR);
  sort_plan<<<1, 32, 0, st>>>(d_small, 256u * n_chunks);
  *launches += 3;
  uint32_t *counts = (uint32_t *)ws.buf[WS_COUNTS];
  for (int pass = 0; pass < 6; pass++) {   // 16 + 32 key bits at most
    radix_histogram<<<n_chunks, kSortWarps * 32, 0, st>>>(ra, rb, n_chunks, d_small, pass, counts);
    CK(exclusive_scan(ws, counts, counts, d_small + SM_COUNTLEN, 256 * n_chunks, nullptr, d_small + SM_ACTIVE + pass, st, launches));
    radix_scatter<<<n_chunks, kSortWarps * 32, kScatterSmem, st>>>(ra, rb, n_chunks, d_small, pass, counts);
    *launches += 2;
  }
  bounds_gather<<<nb, T, 0, st>>>(d_in, ra, rb, d_small, d_out, cap, d_n_out);
  ++*launches;
  return cudaGetLastError();
}

cudaError_t run_cluster(ClusterWorkspace &ws, const strgpu_tread *d_treads, uint32_t n, const strgpu_cluster_params &p,
                        strgpu_bounds *d_out, uint32_t cap, uint32_t *d_n_out, cudaStream_t st, uint64_t *launches,
                        const LociArgs *loci, const uint32_t *d_n_in) {
  // Everything below is enqueued on `st` and nothing is read back: sizes the host does not know (how many digits vary, how
  // many records assign_reads_locus leaves, how many clusters were chained) stay in device memory (d_small) and the
  // kernels that depend on them are launched for the worst case n.
  if (n == 0) return cudaMemsetAsync(d_n_out, 0, 4, st);
  if (n > kIdxMask) return cudaErrorInvalidValue;   // 2^29 records: the split shares SortRec::idx with the index
  const int T = 256;
  const uint32_t nb = (n + T - 1) / T;
  const uint32_t n_chunks = (n + kTile - 1) / kTile;   // one CTA per tile
  static std::once_flag attr_once[64];
  {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    std::call_once(attr_once[dev], []() { cudaFuncSetAttribute(radix_scatter, cudaFuncAttributeMaxDynamicSharedMemorySize, kScatterSmem); });
  }
  const bool with_loci = loci && loci->n_chains;
  // K4 scratch pool: 32 + 8 * reads slots per cluster that passes the min_support test (at most n / max(1, min_support) of them)
  const size_t scratch_slots = (size_t)8 * n + (size_t)32 * (n / (uint32_t)(p.min_support > 1 ? p.min_support : 1)) + 64;
  const uint32_t scan_max = std::max<uint32_t>(2u * n, 256u * n_chunks);
  CK(ws_ensure(ws, WS_RECS_A, (size_t)n * sizeof(SortRec)));
  CK(ws_ensure(ws, WS_RECS_B, (size_t)n * sizeof(SortRec)));
  CK(ws_ensure(ws, WS_SMALL, kSmallWords * 4));
  CK(ws_ensure(ws, WS_COUNTS, (size_t)256 * n_chunks * 4));
  CK(ws_ensure(ws, WS_BLOCKSUMS, ((size_t)(scan_max + kScanTile - 1) / kScanTile + 1) * 4));
  CK(ws_ensure(ws, WS_NEXT, (size_t)n * 4));
  CK(ws_ensure(ws, WS_BEND, (size_t)n * 4));
  CK(ws_ensure(ws, WS_HEAD, (size_t)n * 4));
  CK(ws_ensure(ws, WS_CID, (size_t)n * 4));
  CK(ws_ensure(ws, WS_ENTRY, (size_t)n * 4));
  CK(ws_ensure(ws, WS_CLSTART, (size_t)n * 4));
  CK(ws_ensure(ws, WS_CLEND, (size_t)n * 4));
  CK(ws_ensure(ws, WS_SCRATCH, scratch_slots * sizeof(Slot)));
  CK(ws_ensure(ws, WS_OUT2, (size_t)2 * n * sizeof(strgpu_bounds)));
  CK(ws_ensure(ws, WS_VALID2, (size_t)2 * n * 4));
  CK(ws_ensure(ws, WS_DST2, (size_t)2 * n * 4));
  uint32_t *d_small = (uint32_t *)ws.buf[WS_SMALL];
  CK(cudaMemsetAsync(d_small, 0, kSmallWords * 4, st));
  SortRec *ra = (SortRec *)ws.buf[WS_RECS_A], *rb = (SortRec *)ws.buf[WS_RECS_B];
  init_small<<<1, 32, 0, st>>>(d_small, n, d_n_in);
  make_sort_records<<<nb, T, 0, st>>>(d_treads, d_small, ra, d_small + SM_VAR);
  sort_plan<<<1, 32, 0, st>>>(d_small, 256u * n_chunks);
  *launches += 3;

  // ---- K2: LSD radix sort over the virtual key (varying bits of position, unit, tid); passes past its width return at once
  uint32_t *counts = (uint32_t *)ws.buf[WS_COUNTS];
  for (int pass = 0; pass < kMaxPasses; pass++) {
    radix_histogram<<<n_chunks, kSortWarps * 32, 0, st>>>(ra, rb, n_chunks, d_small, pass, counts);
    ++*launches;
    CK(exclusive_scan(ws, counts, counts, d_small + SM_COUNTLEN, 256 * n_chunks, nullptr, d_small + SM_ACTIVE + pass, st, launches));
    radix_scatter<<<n_chunks, kSortWarps * 32, kScatterSmem, st>>>(ra, rb, n_chunks, d_small, pass, counts);
    ++*launches;
  }

  // ---- C10: loci take their reads out of the sorted buckets before clustering
  int flip = 0;
  if (with_loci) {
    uint32_t *removed = (uint32_t *)ws.buf[WS_HEAD], *keep = (uint32_t *)ws.buf[WS_NEXT], *dst = (uint32_t *)ws.buf[WS_CID];
    CK(cudaMemsetAsync(removed, 0, (size_t)n * 4, st));
    assign_loci<<<(loci->n_chains + 63) / 64, 64, 0, st>>>(ra, rb, d_small, n, loci->d_loci, loci->d_chain_start, loci->n_chains,
                                                            removed, loci->d_counts);
    invert_flags<<<nb, T, 0, st>>>(removed, n, keep);
    *launches += 2;
    CK(exclusive_scan(ws, keep, dst, d_small + SM_NCUR, n, d_small + SM_NCLUSTERS /* temporary: records kept */, nullptr, st, launches));
    compact_sorted<<<nb, T, 0, st>>>(ra, rb, d_small, keep, dst, n);
    set_ncur<<<1, 32, 0, st>>>(d_small);
    *launches += 2;
    flip = 1;
  }

  // ---- K3: next(i), bucket heads, cluster ids (n = d_small[SM_NCUR] from here on)
  uint32_t *next = (uint32_t *)ws.buf[WS_NEXT], *bend = (uint32_t *)ws.buf[WS_BEND], *head = (uint32_t *)ws.buf[WS_HEAD],
           *cid = (uint32_t *)ws.buf[WS_CID], *entry_flag = (uint32_t *)ws.buf[WS_ENTRY];
  const uint32_t piece_blocks = ((n + kSeg - 1) / kSeg + kPieceWarps - 1) / kPieceWarps;
  cluster_next<<<nb, T, 0, st>>>(ra, rb, d_small, flip, p.window, next, bend, head, entry_flag);
  uint32_t *exit_of = cid;   // cid doubles as exit_of until the scan below
  piece_exits<<<piece_blocks, kPieceWarps * 32, 0, st>>>(next, bend, d_small, exit_of);
  bucket_entry<<<nb, T, 0, st>>>(ra, rb, d_small, flip, bend, exit_of, entry_flag);
  piece_mark<<<piece_blocks, kPieceWarps * 32, 0, st>>>(next, bend, d_small, entry_flag, head);
  *launches += 4;
  CK(exclusive_scan(ws, head, cid, d_small + SM_NCUR, n, d_small + SM_NCLUSTERS, nullptr, st, launches));
  uint32_t *cl_start = (uint32_t *)ws.buf[WS_CLSTART], *cl_end = (uint32_t *)ws.buf[WS_CLEND];
  cluster_fill<<<nb, T, 0, st>>>(head, cid, next, d_small, cl_start, cl_end);
  ++*launches;

  // ---- K4: bounds per cluster (launched for n clusters, the worst case), then ordered compaction
  strgpu_bounds *out2 = (strgpu_bounds *)ws.buf[WS_OUT2];
  uint32_t *valid2 = (uint32_t *)ws.buf[WS_VALID2], *dst2 = (uint32_t *)ws.buf[WS_DST2];
  // clusters that reach cluster_bounds hold >= min_support reads each
  const uint32_t heavy_max = p.min_support > 1 ? n / (uint32_t)p.min_support + 1u : n;
  CK(ws_ensure(ws, WS_HEAVY, (size_t)heavy_max * sizeof(Heavy)));
  Heavy *heavy = (Heavy *)ws.buf[WS_HEAVY];
  cluster_screen<<<(n + 127) / 128, 128, 0, st>>>(ra, rb, flip, d_treads, cl_start, cl_end, d_small, p, out2, valid2, heavy);
  cluster_bounds<<<(heavy_max + 63) / 64, 64, 0, st>>>(ra, rb, flip, d_treads, d_small, p, heavy, (Slot *)ws.buf[WS_SCRATCH],
                                                       (uint32_t)std::min<size_t>(scratch_slots, 0xffffffffu), out2, valid2);
  *launches += 2;
  CK(exclusive_scan(ws, valid2, dst2, d_small + SM_N2, 2 * n, d_n_out, nullptr, st, launches));
  compact_bounds<<<(2 * n + T - 1) / T, T, 0, st>>>(out2, valid2, dst2, d_small, d_out, cap);
  ++*launches;
  return cudaGetLastError();
}

}  // namespace strgpu
