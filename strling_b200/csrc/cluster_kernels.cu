// K2 tread_sort, K3 cluster_chain, K4 cluster_bounds: STRling's cluster loop (call.nim:118-130,223-235;
// merge.nim:125-187; cluster.nim:175-374; callclusters.nim:52-66) on sm_100a.  Integer work only.
//
//   K2  stable LSD radix sort of 16-byte (tid, unit, position, index) records: one warp owns one contiguous
//       chunk, ranks 32 records at a time with match.any, so equal keys keep `.bin` order exactly like the
//       reference's bucket append + stable sort by position (call.nim:124-130).  Digits whose bits do not vary
//       are skipped.
//   K3  next(i) = first read NOT absorbed by a cluster started at read i (trcluster, cluster.nim:323-362) is a
//       pure function of i: <= 8 explicit steps while the median-of-first-9 still moves, then one binary search.
//       Bucket heads then chase next() to mark cluster starts; the chase is cut into 256-record pieces (exit tables per
//       piece, a piece-to-piece hop per bucket, a marking sweep per piece) so it stays parallel for huge buckets.
//   K4  one thread per chained cluster: trim, left/right_most, min_support + anchor test, split_cluster, bounds,
//       filters, has_per_sample_reads.  CountTable.largest ties follow Nim's slot order (hashWangYi1 + linear
//       probing + growth), emulated in a per-cluster scratch region.
#include "cluster_kernels.cuh"

#include <cstdio>

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

__global__ void __launch_bounds__(kScanThreads) scan_block_sums(const uint32_t *in, uint32_t n, uint32_t *block_sums) {
  const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
  uint32_t s = 0;
#pragma unroll
  for (int j = 0; j < kScanItems; j++)
    if (base + j < n) s += in[base + j];
  uint32_t total;
  block_exclusive_scan(s, &total);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = total;
}

__global__ void __launch_bounds__(kScanThreads) scan_single_block(uint32_t *data, uint32_t n, uint32_t *total_out) {
  uint32_t carry = 0;
  for (uint32_t base = 0; base < n; base += kScanThreads) {
    const uint32_t i = base + threadIdx.x;
    const uint32_t v = i < n ? data[i] : 0;
    uint32_t total;
    const uint32_t ex = block_exclusive_scan(v, &total);
    if (i < n) data[i] = carry + ex;
    carry += total;
  }
  if (threadIdx.x == 0 && total_out) *total_out = carry;
}

__global__ void __launch_bounds__(kScanThreads) scan_apply(const uint32_t *in, uint32_t n, const uint32_t *block_offsets,
                                                           uint32_t *out) {
  const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * kScanItems;
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

__global__ void make_sort_records(const strgpu_tread *__restrict__ treads, uint32_t n, SortRec *__restrict__ recs,
                                  uint32_t *__restrict__ varbits) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  SortRec r{0, 0, 0, 0};
  uint32_t d0 = 0, d1 = 0, d2 = 0;
  if (i < n) {
    const strgpu_tread t = treads[i];
    r.hi = (uint32_t)t.tid ^ 0x80000000u;
    uint32_t m = 0;
#pragma unroll
    for (int j = 0; j < 6; j++) m = (m << 3) | unit_rank(t.repeat[j]);
    r.mid = m;
    r.pos = t.position;
    r.idx = i;
    recs[i] = r;
    const strgpu_tread f = treads[0];
    uint32_t fm = 0;
#pragma unroll
    for (int j = 0; j < 6; j++) fm = (fm << 3) | unit_rank(f.repeat[j]);
    d0 = r.pos ^ f.position;
    d1 = r.mid ^ fm;
    d2 = r.hi ^ ((uint32_t)f.tid ^ 0x80000000u);
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

__device__ __forceinline__ uint32_t digit_of(const SortRec &r, int field, int shift) {
  const uint32_t v = field == 0 ? r.pos : (field == 1 ? r.mid : r.hi);
  return (v >> shift) & 0xffu;
}

constexpr int kSortWarps = 8;

__global__ void __launch_bounds__(kSortWarps * 32) radix_histogram(const SortRec *__restrict__ in, uint32_t n, uint32_t chunk,
                                                                   uint32_t n_chunks, int field, int shift,
                                                                   uint32_t *__restrict__ counts) {
  __shared__ uint32_t hist[kSortWarps][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t w = blockIdx.x * kSortWarps + warp;
  for (int d = lane; d < 256; d += 32) hist[warp][d] = 0;
  __syncwarp();
  if (w < n_chunks) {
    const uint32_t beg = w * chunk;
    const uint32_t end = min(n, beg + chunk);
    for (uint32_t i = beg + lane; i < end; i += 32) atomicAdd(&hist[warp][digit_of(in[i], field, shift)], 1u);
    __syncwarp();
    for (int d = lane; d < 256; d += 32) counts[(size_t)d * n_chunks + w] = hist[warp][d];
  }
}

__global__ void __launch_bounds__(kSortWarps * 32) radix_scatter(const SortRec *__restrict__ in, SortRec *__restrict__ out,
                                                                 uint32_t n, uint32_t chunk, uint32_t n_chunks, int field,
                                                                 int shift, const uint32_t *__restrict__ offsets) {
  __shared__ uint32_t off[kSortWarps][256];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t w = blockIdx.x * kSortWarps + warp;
  if (w >= n_chunks) return;
  for (int d = lane; d < 256; d += 32) off[warp][d] = offsets[(size_t)d * n_chunks + w];
  __syncwarp();
  const uint32_t beg = w * chunk;
  const uint32_t end = min(n, beg + chunk);
  const uint32_t lane_lt = (1u << lane) - 1u;
  for (uint32_t base = beg; base < end; base += 32) {
    const uint32_t i = base + lane;
    const bool valid = i < end;
    SortRec r{0, 0, 0, 0};
    if (valid) r = in[i];
    const uint32_t d = valid ? digit_of(r, field, shift) : (0x100u + (uint32_t)lane);
    const uint32_t grp = __match_any_sync(kFull, d);
    uint32_t dst = 0;
    if (valid) dst = off[warp][d] + __popc(grp & lane_lt);  // lane order == input order: stable
    __syncwarp();
    if (valid && lane == 31 - __clz(grp)) off[warp][d] += __popc(grp);
    __syncwarp();
    if (valid) out[dst] = r;
  }
}

__global__ void gather_treads(const strgpu_tread *__restrict__ treads, const SortRec *__restrict__ recs, uint32_t n,
                              strgpu_tread *__restrict__ sorted) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const unsigned long long *src = reinterpret_cast<const unsigned long long *>(treads + recs[i].idx);
  unsigned long long *dst = reinterpret_cast<unsigned long long *>(sorted + i);
  dst[0] = src[0];
  dst[1] = src[1];
  dst[2] = src[2];
}

// ------------------------------------------------------------------------------------------- C10 assign_reads_locus
// One thread per chain (= the loci of one bucket, in file order): callclusters.nim:14-50 on the sorted records with
// removal expressed as marks.  Earlier loci change what later loci of the same bucket see, hence the serial chain.
__global__ void assign_loci(const SortRec *__restrict__ recs, const strgpu_tread *__restrict__ sorted, uint32_t n,
                            const DevLocus *__restrict__ loci, const uint32_t *__restrict__ chain_start, uint32_t n_chains,
                            uint32_t *__restrict__ removed, uint16_t *__restrict__ counts) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_chains) return;
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
        const uint8_t sp = sorted[i].split;
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

__global__ void compact_sorted(const SortRec *__restrict__ recs, const strgpu_tread *__restrict__ sorted, const uint32_t *__restrict__ keep,
                               const uint32_t *__restrict__ dst, uint32_t n, SortRec *__restrict__ recs_out,
                               strgpu_tread *__restrict__ sorted_out) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || !keep[i]) return;
  const uint32_t d = dst[i];
  recs_out[d] = recs[i];
  const unsigned long long *src = reinterpret_cast<const unsigned long long *>(sorted + i);
  unsigned long long *o = reinterpret_cast<unsigned long long *>(sorted_out + d);
  o[0] = src[0]; o[1] = src[1]; o[2] = src[2];
}

// ------------------------------------------------------------------------------------------- K3 chain
__device__ __forceinline__ bool same_bucket(const SortRec &a, const SortRec &b) { return a.hi == b.hi && a.mid == b.mid; }

__global__ void cluster_next(const SortRec *__restrict__ recs, uint32_t n, uint32_t max_dist, uint32_t *__restrict__ next,
                             uint32_t *__restrict__ bucket_end, uint32_t *__restrict__ head) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  head[i] = 0;
  const SortRec me = recs[i];
  // end of my (tid, unit) bucket: first index whose bucket key differs (records are sorted)
  uint32_t lo = i + 1, hi = n;
  while (lo < hi) {
    const uint32_t mid = lo + ((hi - lo) >> 1);
    if (same_bucket(recs[mid], me)) lo = mid + 1; else hi = mid;
  }
  const uint32_t be = lo;
  bucket_end[i] = be;
  if ((int32_t)(me.hi ^ 0x80000000u) < 0) {  // unplaced: the whole bucket is one Cluster (cluster.nim:369-371)
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
// so that no thread takes more than kSeg steps (a bucket can hold 10^5 reads):
//   seg_exits : for every record i of a piece, exit[i] = the first index >= the piece's end that the chain through i reaches
//               (a reverse sweep of the piece; next[i] > i always)
//   seg_entry : one thread per bucket hops piece to piece (entry -> exit[entry]) and records each piece's entry point
//   seg_mark  : one thread per entered piece marks the cluster starts inside it
constexpr uint32_t kSeg = 256;

__global__ void seg_exits(const uint32_t *__restrict__ next, const uint32_t *__restrict__ bucket_end, uint32_t n,
                          uint32_t *__restrict__ exit_of) {
  const uint32_t sgi = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lo = sgi * kSeg;
  if (lo >= n) return;
  const uint32_t hi = min(n, lo + kSeg);
  for (uint32_t i = hi; i-- > lo;) {
    const uint32_t nx = next[i];
    // a chain never leaves its bucket: next[i] <= bucket_end[i]; stop at the piece end or the bucket end
    exit_of[i] = (nx >= hi || nx >= bucket_end[i]) ? nx : exit_of[nx];
  }
}

__global__ void seg_entry(const SortRec *__restrict__ recs, uint32_t n, const uint32_t *__restrict__ bucket_end,
                          const uint32_t *__restrict__ exit_of, uint32_t *__restrict__ entry_of_seg) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (i > 0 && same_bucket(recs[i - 1], recs[i])) return;  // bucket starts only
  const uint32_t be = bucket_end[i];
  uint32_t h = i;
  while (h < be) {
    // several buckets can start inside one piece: keep the piece's entries as a linked list through the records
    entry_of_seg[h] = 1;   // h is a cluster start and the point where the chain enters (or re-enters) its piece
    h = exit_of[h];
  }
}

__global__ void seg_mark(const uint32_t *__restrict__ next, const uint32_t *__restrict__ bucket_end, uint32_t n,
                         const uint32_t *__restrict__ entry_flag, uint32_t *__restrict__ head) {
  const uint32_t sgi = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t lo = sgi * kSeg;
  if (lo >= n) return;
  const uint32_t hi = min(n, lo + kSeg);
  for (uint32_t e = lo; e < hi; e++) {
    if (!entry_flag[e]) continue;          // chains enter the piece here (one per bucket that touches the piece)
    const uint32_t be = bucket_end[e];
    uint32_t h = e;
    while (h < hi && h < be) {
      head[h] = 1;
      h = next[h];
    }
  }
}

__global__ void cluster_fill(const uint32_t *__restrict__ head, const uint32_t *__restrict__ cid, const uint32_t *__restrict__ next,
                             uint32_t n, uint32_t *__restrict__ cl_start, uint32_t *__restrict__ cl_end) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
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

__device__ __forceinline__ bool clip_selected(const strgpu_tread &r, ClipMode mode, uint32_t cm, uint32_t mcd) {
  // bounds(): cluster.nim:193,197 (int32 casts and adds wrap); the elif makes "right" exclude nothing extra
  // because a read has a single split value.
  switch (mode) {
    case kSplitLeft: return r.split == SOFT_LEFT;
    case kSplitRight: return r.split == SOFT_RIGHT;
    case kBoundsLeft: return r.split == SOFT_LEFT && (int32_t)r.position < (int32_t)(cm + mcd);
    default: return r.split == SOFT_RIGHT && (int32_t)r.position > (int32_t)(cm - mcd);
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
__device__ Largest count_largest(const strgpu_tread *__restrict__ reads, uint32_t a, uint32_t b, ClipMode mode, uint32_t cm,
                                 uint32_t mcd, Slot *scratch) {
  Largest res{0, 0, 0, 0};
  uint32_t cur_key = 0, cur_val = 0, n_at_max = 0;
  for (uint32_t i = a; i < b; i++) {
    const strgpu_tread r = reads[i];
    if (!clip_selected(r, mode, cm, mcd)) continue;
    res.n_selected++;
    if (cur_val && r.position == cur_key) { cur_val++; }
    else {
      if (cur_val) {
        if (cur_val > res.val) { res.val = cur_val; res.key = cur_key; n_at_max = 1; }
        else if (cur_val == res.val) n_at_max++;
      }
      cur_key = r.position; cur_val = 1; res.n_distinct++;
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
    const strgpu_tread r = reads[i];
    if (!clip_selected(r, mode, cm, mcd)) continue;
    if (cur_val && r.position == cur_key) { cur_val++; }
    else {
      if (cur_val) insert(cur_key, cur_val);
      cur_key = r.position; cur_val = 1;
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

__device__ __forceinline__ uint32_t posmed(const strgpu_tread *__restrict__ reads, uint32_t a, uint32_t b) {
  const uint32_t n = b - a;  // cluster.nim:59-62 : reads[int(min(9, n)/2 - 0.5)]
  const uint32_t m = n < 9u ? n : 9u;
  return reads[a + ((m - 1u) >> 1)].position;
}

// merge.nim:18-25 : does any sample own >= supporting reads of [a, b)?  (scratch as an open-addressing counter)
__device__ bool per_sample_support(const strgpu_tread *__restrict__ reads, uint32_t a, uint32_t b, int supporting, Slot *scratch) {
  if (supporting <= 0) return true;
  const uint32_t n = b - a;
  if (n < (uint32_t)supporting) return false;
  uint32_t cap = 16;
  while (cap < 2 * n) cap <<= 1;
  for (uint32_t s = 0; s < cap; s++) scratch[s] = Slot{0, 0};
  for (uint32_t i = a; i < b; i++) {
    const uint32_t key = (uint32_t)reads[i].sample;
    uint32_t h = (key * 2654435761u) & (cap - 1);
    while (scratch[h].val != 0 && scratch[h].key != key) h = (h + 1) & (cap - 1);
    scratch[h].key = key;
    if ((int)++scratch[h].val >= supporting) return true;
  }
  return false;
}

// cluster.nim:175-250 + callclusters.nim:52-66.  Returns false when the cluster is dropped.
__device__ bool bounds_of(const strgpu_tread *__restrict__ reads, uint32_t a, uint32_t b, uint32_t cl_left_most,
                          uint32_t cl_right_most, const strgpu_cluster_params &p, Slot *scratch, strgpu_bounds &out) {
  const uint32_t n = b - a;
  if (n >= 65535u) return false;  // callclusters.nim:53-55
  const strgpu_tread first = reads[a];
  out.tid = first.tid;
#pragma unroll
  for (int j = 0; j < 6; j++) out.repeat[j] = first.repeat[j];
  const uint32_t cm = reads[a + (n >> 1)].position;
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
  uint32_t lm = cl_left_most > 0 ? cl_left_most : first.position;
  uint32_t rm = cl_right_most > 0 ? cl_right_most : reads[b - 1].position;
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

__global__ void cluster_bounds(const strgpu_tread *__restrict__ reads, const uint32_t *__restrict__ cl_start,
                               const uint32_t *__restrict__ cl_end, uint32_t n_clusters, strgpu_cluster_params p,
                               Slot *__restrict__ scratch_all, strgpu_bounds *__restrict__ out2, uint32_t *__restrict__ valid2) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_clusters) return;
  uint32_t a = cl_start[c];
  const uint32_t b = cl_end[c];
  Slot *scratch = scratch_all + ((size_t)32 * c + (size_t)8 * a);
  valid2[2 * c] = 0;
  valid2[2 * c + 1] = 0;
  const strgpu_tread first = reads[a];
  if (first.tid < 0) {  // unplaced bucket: call.nim:226-228 records len per unit; merge.nim:175-176 skips
    if (!p.merge_mode) {
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
  const uint32_t max_dist = p.window;
  // trim (cluster.nim:252-257): lo is computed once from the untrimmed cluster
  {
    const long long lo_l = (long long)posmed(reads, a, b) - (long long)(max_dist + 100u);
    const uint32_t lo = lo_l > 0 ? (uint32_t)lo_l : 0u;
    while (b - a > 1 && reads[a].position < lo) a++;
  }
  const uint32_t pm = posmed(reads, a, b);
  const uint32_t last = reads[b - 1].position, firstp = reads[a].position;
  const uint32_t hi_edge = pm + max_dist, lo_edge = pm - max_dist;  // uint32 wrap (cluster.nim:343-344)
  const uint32_t cl_right_most = last > hi_edge ? last : hi_edge;
  const uint32_t cl_left_most = firstp < lo_edge ? firstp : lo_edge;
  if ((long long)(b - a) < (long long)p.min_support) return;
  bool anchor = false;
  for (uint32_t i = a; i < b && !anchor; i++) anchor = reads[i].split == SOFT_NONE;
  if (!anchor) return;
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
      while (m < b && reads[m].position < mid) m++;
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
                               const uint32_t *__restrict__ dst_idx, uint32_t n2, strgpu_bounds *__restrict__ out, uint32_t cap) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n2 || !valid2[i]) return;
  const uint32_t d = dst_idx[i];
  if (d < cap) out[d] = out2[i];
}

// ------------------------------------------------------------------------------------------- host driver
enum { WS_RECS_A, WS_RECS_B, WS_COUNTS, WS_BLOCKSUMS, WS_SORTED, WS_NEXT, WS_BEND, WS_HEAD, WS_CID, WS_CLSTART, WS_CLEND,
       WS_SCRATCH, WS_OUT2, WS_VALID2, WS_DST2, WS_SMALL, WS_SORTED_B };

cudaError_t ws_ensure(ClusterWorkspace &ws, int which, size_t bytes) {
  if (bytes <= ws.cap[which]) return cudaSuccess;
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

// exclusive scan of n uint32 (in -> out, may alias), total to *d_total (device)
cudaError_t exclusive_scan(ClusterWorkspace &ws, const uint32_t *in, uint32_t *out, uint32_t n, uint32_t *d_total,
                           cudaStream_t st, uint64_t *launches) {
  const uint32_t blocks = (n + kScanTile - 1) / kScanTile;
  CK(ws_ensure(ws, WS_BLOCKSUMS, (size_t)(blocks + 1) * 4));
  uint32_t *bs = (uint32_t *)ws.buf[WS_BLOCKSUMS];
  scan_block_sums<<<blocks, kScanThreads, 0, st>>>(in, n, bs);
  scan_single_block<<<1, kScanThreads, 0, st>>>(bs, blocks, d_total);
  scan_apply<<<blocks, kScanThreads, 0, st>>>(in, n, bs, out);
  *launches += 3;
  return cudaGetLastError();
}

}  // namespace

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
    m = (m << 3) | r;
  }
  return m;
}

void free_workspace(ClusterWorkspace &ws) {
  for (int i = 0; i < 17; i++) {
    if (ws.buf[i]) cudaFree(ws.buf[i]);
    ws.buf[i] = nullptr;
    ws.cap[i] = 0;
  }
}

cudaError_t run_cluster(ClusterWorkspace &ws, const strgpu_tread *d_treads, uint32_t n, const strgpu_cluster_params &p,
                        strgpu_bounds *d_out, uint32_t cap, uint32_t *d_n_out, cudaStream_t st, uint64_t *launches,
                        const LociArgs *loci) {
  if (n == 0) return cudaMemsetAsync(d_n_out, 0, 4, st);
  const int T = 256;
  const uint32_t nb = (n + T - 1) / T;
  CK(ws_ensure(ws, WS_RECS_A, (size_t)n * sizeof(SortRec)));
  CK(ws_ensure(ws, WS_RECS_B, (size_t)n * sizeof(SortRec)));
  CK(ws_ensure(ws, WS_SMALL, 64));
  uint32_t *d_small = (uint32_t *)ws.buf[WS_SMALL];  // [0..2] varying bits, [3] n_clusters, [4] n_out
  CK(cudaMemsetAsync(d_small, 0, 64, st));
  SortRec *ra = (SortRec *)ws.buf[WS_RECS_A], *rb = (SortRec *)ws.buf[WS_RECS_B];
  make_sort_records<<<nb, T, 0, st>>>(d_treads, n, ra, d_small);
  ++*launches;
  uint32_t var[3];
  CK(cudaMemcpyAsync(var, d_small, 12, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));

  // ---- K2: LSD radix sort, least significant field first: position, unit, tid
  uint32_t chunk = 1024;
  while ((n + chunk - 1) / chunk > 8192) chunk *= 2;
  const uint32_t n_chunks = (n + chunk - 1) / chunk;
  const uint32_t sort_blocks = (n_chunks + kSortWarps - 1) / kSortWarps;
  CK(ws_ensure(ws, WS_COUNTS, (size_t)256 * n_chunks * 4));
  uint32_t *counts = (uint32_t *)ws.buf[WS_COUNTS];
  for (int field = 0; field < 3; field++)
    for (int shift = 0; shift < 32; shift += 8) {
      if (((var[field] >> shift) & 0xffu) == 0) continue;  // this digit is the same everywhere
      radix_histogram<<<sort_blocks, kSortWarps * 32, 0, st>>>(ra, n, chunk, n_chunks, field, shift, counts);
      ++*launches;
      CK(exclusive_scan(ws, counts, counts, 256 * n_chunks, nullptr, st, launches));
      radix_scatter<<<sort_blocks, kSortWarps * 32, 0, st>>>(ra, rb, n, chunk, n_chunks, field, shift, counts);
      ++*launches;
      SortRec *t = ra; ra = rb; rb = t;
    }
  CK(ws_ensure(ws, WS_SORTED, (size_t)n * sizeof(strgpu_tread)));
  strgpu_tread *sorted = (strgpu_tread *)ws.buf[WS_SORTED];
  gather_treads<<<nb, T, 0, st>>>(d_treads, ra, n, sorted);
  ++*launches;

  // ---- C10: loci take their reads out of the sorted buckets before clustering
  if (loci && loci->n_chains) {
    CK(ws_ensure(ws, WS_HEAD, (size_t)n * 4));
    CK(ws_ensure(ws, WS_NEXT, (size_t)n * 4));
    CK(ws_ensure(ws, WS_CID, (size_t)n * 4));
    CK(ws_ensure(ws, WS_SORTED_B, (size_t)n * sizeof(strgpu_tread)));
    uint32_t *removed = (uint32_t *)ws.buf[WS_HEAD], *keep = (uint32_t *)ws.buf[WS_NEXT], *dst = (uint32_t *)ws.buf[WS_CID];
    CK(cudaMemsetAsync(removed, 0, (size_t)n * 4, st));
    assign_loci<<<(loci->n_chains + 63) / 64, 64, 0, st>>>(ra, sorted, n, loci->d_loci, loci->d_chain_start, loci->n_chains, removed,
                                                            loci->d_counts);
    invert_flags<<<nb, T, 0, st>>>(removed, n, keep);
    *launches += 2;
    CK(exclusive_scan(ws, keep, dst, n, d_small + 5, st, launches));
    strgpu_tread *sorted_b = (strgpu_tread *)ws.buf[WS_SORTED_B];
    compact_sorted<<<nb, T, 0, st>>>(ra, sorted, keep, dst, n, rb, sorted_b);
    ++*launches;
    uint32_t n_kept = 0;
    CK(cudaMemcpyAsync(&n_kept, d_small + 5, 4, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    SortRec *t = ra; ra = rb; rb = t;
    sorted = sorted_b;
    n = n_kept;
    if (n == 0) return cudaMemsetAsync(d_n_out, 0, 4, st);
  }
  const uint32_t nb2 = (n + T - 1) / T;

  // ---- K3: next(i), bucket heads, cluster ids
  CK(ws_ensure(ws, WS_NEXT, (size_t)n * 4));
  CK(ws_ensure(ws, WS_BEND, (size_t)n * 4));
  CK(ws_ensure(ws, WS_HEAD, (size_t)n * 4));
  CK(ws_ensure(ws, WS_CID, (size_t)n * 4));
  uint32_t *next = (uint32_t *)ws.buf[WS_NEXT], *bend = (uint32_t *)ws.buf[WS_BEND], *head = (uint32_t *)ws.buf[WS_HEAD],
           *cid = (uint32_t *)ws.buf[WS_CID];
  cluster_next<<<nb2, T, 0, st>>>(ra, n, p.window, next, bend, head);
  {
    // chain walk in pieces of kSeg records (see seg_exits): cid doubles as exit_of, the scatter buffer rb as entry flags
    uint32_t *exit_of = cid;
    uint32_t *entry_flag = reinterpret_cast<uint32_t *>(rb);
    const uint32_t n_seg = (n + kSeg - 1) / kSeg;
    CK(cudaMemsetAsync(entry_flag, 0, (size_t)n * 4, st));
    seg_exits<<<(n_seg + 63) / 64, 64, 0, st>>>(next, bend, n, exit_of);
    seg_entry<<<nb2, T, 0, st>>>(ra, n, bend, exit_of, entry_flag);
    seg_mark<<<(n_seg + 63) / 64, 64, 0, st>>>(next, bend, n, entry_flag, head);
  }
  *launches += 4;
  CK(exclusive_scan(ws, head, cid, n, d_small + 3, st, launches));
  uint32_t n_clusters = 0;
  CK(cudaMemcpyAsync(&n_clusters, d_small + 3, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  if (n_clusters == 0) return cudaMemsetAsync(d_n_out, 0, 4, st);
  CK(ws_ensure(ws, WS_CLSTART, (size_t)n_clusters * 4));
  CK(ws_ensure(ws, WS_CLEND, (size_t)n_clusters * 4));
  uint32_t *cl_start = (uint32_t *)ws.buf[WS_CLSTART], *cl_end = (uint32_t *)ws.buf[WS_CLEND];
  cluster_fill<<<nb2, T, 0, st>>>(head, cid, next, n, cl_start, cl_end);
  ++*launches;

  // ---- K4: bounds per cluster, then ordered compaction
  CK(ws_ensure(ws, WS_SCRATCH, ((size_t)32 * n_clusters + (size_t)8 * n + 64) * sizeof(Slot)));
  CK(ws_ensure(ws, WS_OUT2, (size_t)2 * n_clusters * sizeof(strgpu_bounds)));
  CK(ws_ensure(ws, WS_VALID2, (size_t)2 * n_clusters * 4));
  CK(ws_ensure(ws, WS_DST2, (size_t)2 * n_clusters * 4));
  strgpu_bounds *out2 = (strgpu_bounds *)ws.buf[WS_OUT2];
  uint32_t *valid2 = (uint32_t *)ws.buf[WS_VALID2], *dst2 = (uint32_t *)ws.buf[WS_DST2];
  const uint32_t cb = (n_clusters + 127) / 128;
  cluster_bounds<<<cb, 128, 0, st>>>(sorted, cl_start, cl_end, n_clusters, p, (Slot *)ws.buf[WS_SCRATCH], out2, valid2);
  ++*launches;
  CK(exclusive_scan(ws, valid2, dst2, 2 * n_clusters, d_n_out, st, launches));
  compact_bounds<<<(2 * n_clusters + T - 1) / T, T, 0, st>>>(out2, valid2, dst2, 2 * n_clusters, d_out, cap);
  ++*launches;
  return cudaGetLastError();
}

}  // namespace strgpu
