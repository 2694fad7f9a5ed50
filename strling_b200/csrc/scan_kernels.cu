// K1 repeat_scan: the per-segment repeat-unit decision of STRling's get_repeat (utils.nim:236-271) on sm_100a.
//
// Mapping (v1, general kernel): one warp per segment, one lane per k-mer window.
//   * the segment's 2-bit bases are re-aligned into a per-warp shared-memory line (base 0 at the top of word 0),
//   * for k = 2..6 every lane extracts its window's 2k-bit code, takes the minimum over the k rotations
//     (slide_by, utils.nim:10-34), and the warp finds the max multiplicity and its order-exact leader
//     (Seq.inc keeps the FIRST code to reach the final maximum: strict `>`, utils.nim:192-195) with
//     match.any + popc ordinals + redux.max; counts carried across 32-window rounds live in a per-warp
//     uint8 table that is un-done (not memset) afterwards,
//   * the phase-aware recount (strutils.count, greedy non-overlapping; utils.nim:254) is a ballot of
//     per-position pattern matches followed by a warp-uniform find-first-set walk,
//   * the score / break / continue ladder of utils.nim:250-265 is evaluated warp-uniformly.
// Integer / bitwise work only: no tensor cores.  Thresholds come from a host-built fp64-exact table.
#include "scan_kernels.cuh"

#include <cstdlib>

namespace strgpu {

namespace {

constexpr uint32_t kFull = 0xffffffffu;

template <int K>
__device__ __forceinline__ uint32_t min_rotation(uint32_t x) {
  constexpr uint32_t kMask = (1u << (2 * K)) - 1u;
  uint32_t m = x;
#pragma unroll
  for (int j = 1; j < K; j++) {
    x = ((x << 2) | (x >> (2 * K - 2))) & kMask;
    m = min(m, x);
  }
  return m;
}

// 2K bits starting at base `i` of the aligned line (big-endian: first base most significant)
template <int K>
__device__ __forceinline__ uint32_t bases_at(const uint32_t *sw, uint32_t i) {
  const uint32_t bit = 2u * i;
  const uint32_t w = bit >> 5;
  return __funnelshift_l(sw[w + 1], sw[w], bit & 31u) >> (32 - 2 * K);
}

// count(read, k, counts[k]) + argmax (utils.nim:197,205-211): M = max multiplicity, leader = its code.
template <int K, int MAXR>
__device__ __forceinline__ void count_k(const uint32_t *sw, uint8_t *tab, int L, int lane, int &M, uint32_t &leader) {
  const int W = L / K;
  const int rounds = (W + 31) >> 5;
  M = 0;
  leader = (1u << (2 * K)) - 1u;  // imax == -1 -> argmax is all ones -> decodes to "GG.." (utils.nim:197,245)
  const uint32_t lane_le = kFull >> (31 - lane);
  uint32_t cc[MAXR];
#pragma unroll
  for (int r = 0; r < MAXR; r++) {
    cc[r] = kFull;
    if (r < rounds) {  // warp-uniform
      const int w = 32 * r + lane;
      const bool valid = w < W;
      const uint32_t c = min_rotation<K>(bases_at<K>(sw, valid ? (uint32_t)(K * w) : 0u));
      if (valid) cc[r] = c;
      const uint32_t grp = __match_any_sync(kFull, valid ? c : (0x80000000u | (uint32_t)lane));
      int base = 0;
      if (rounds > 1) {
        if (valid) base = tab[c];
        __syncwarp();
        if (valid && lane == 31 - __clz(grp)) tab[c] = (uint8_t)(base + __popc(grp));
        __syncwarp();
      }
      const int occ = valid ? base + __popc(grp & lane_le) : 0;
      const int rmax = __reduce_max_sync(kFull, occ);
      if (rmax > M) {  // warp-uniform; the first window (in read order) that reaches the new maximum leads
        M = rmax;
        const uint32_t b = __ballot_sync(kFull, valid && occ == rmax);
        leader = __shfl_sync(kFull, c, __ffs(b) - 1);
      }
    }
  }
  if (rounds > 1) {
#pragma unroll
    for (int r = 0; r < MAXR; r++)
      if (cc[r] != kFull) tab[cc[r]] = 0;
    __syncwarp();
  }
}

// read.count(s): greedy leftmost non-overlapping occurrences of the K-base pattern (utils.nim:254).
// A non-ACGT base never matches (the reference compares raw ASCII while `s` is drawn from CATG).
template <int K, int MAXP>
__device__ __forceinline__ int recount_k(const uint32_t *sw, const uint32_t *nm, bool has_n, int L, int lane,
                                         uint32_t pat) {
  const int npos = L - K + 1;
  int c = 0, next = 0;
#pragma unroll
  for (int r = 0; r < MAXP; r++) {
    if (32 * r < npos) {  // warp-uniform
      const int i = 32 * r + lane;
      const bool valid = i < npos;
      bool eq = valid && (bases_at<K>(sw, valid ? (uint32_t)i : 0u) == pat);
      if (has_n) {
        const uint32_t nb = __funnelshift_r(nm[r], nm[r + 1], lane) & ((1u << K) - 1u);
        eq = eq && (nb == 0u);
      }
      uint32_t m = __ballot_sync(kFull, eq);
      const int rel = next - 32 * r;
      if (rel > 0) m = (rel >= 32) ? 0u : (m & (kFull << rel));
      while (m) {
        const int nx = __ffs(m) - 1 + K;
        c++;
        m = (nx >= 32) ? 0u : (m & (kFull << nx));
        next = 32 * r + nx;
      }
    }
  }
  return c;
}

// one 8-byte store per result record
__device__ __forceinline__ void store_result(strgpu_repeat *out, uint32_t idx, const strgpu_repeat &r) {
  unsigned long long v = 0;
#pragma unroll
  for (int i = 0; i < 6; i++) v |= (unsigned long long)(uint8_t)r.unit[i] << (8 * i);
  v |= (unsigned long long)r.repeat_count << 48;
  reinterpret_cast<unsigned long long *>(out)[idx] = v;
}

struct ScanState {
  int best;
  uint32_t unit_code;
  int unit_k;
  int rc;
};

// one rung of the k = 2..6 ladder (utils.nim:242-265).  Returns false on `break`.
template <int K, int MAXR, int MAXP>
__device__ __forceinline__ bool ladder_step(const uint32_t *sw, const uint32_t *nm, bool has_n, uint8_t *tab, int L,
                                            int lane, int thr_p, int thr_giveup, ScanState &st) {
  int M;
  uint32_t leader;
  count_k<K, MAXR>(sw, tab, L, lane, M, leader);
  int score = M * K;
  if (score <= st.best) return !(M < thr_giveup);
  const int c = recount_k<K, MAXP>(sw, nm, has_n, L, lane, leader);
  score = c * K;
  if (score < st.best) return true;
  st.best = score;
  if (c > thr_p) {
    st.unit_code = leader;
    st.unit_k = K;
    st.rc = c;
  }
  return true;
}

template <int MAXLEN>
struct WarpScratch {
  static constexpr int kSeqWords = (2 * MAXLEN + 31) / 32 + 2;
  static constexpr int kNWords = (MAXLEN + 31) / 32 + 2;
  static constexpr int kTab = (MAXLEN <= kShortMaxLen) ? 256 : 4096;  // k<=4 only needs carries when len <= 160
  uint32_t sw[kSeqWords];
  uint32_t nm[kNWords];
  uint8_t tab[kTab];
  uint16_t wc[(MAXLEN / 2 + 31) / 32 * 32];  // window codes of the current k (compact path only)
};

// decode (kmer.decode with alphabet "CATG") + reduce_repeat (utils.nim:220-233,271), one 8-byte store
__device__ __forceinline__ void emit_result(strgpu_repeat *out, uint32_t s, const ScanState &st) {
  strgpu_repeat res;
#pragma unroll
  for (int i = 0; i < 6; i++) res.unit[i] = 0;
  res.repeat_count = 0;
  if (st.unit_k > 0) {
    const uint32_t alpha = 0x47544143u;  // 'C','A','T','G' little-endian
    bool homo = true;
    const uint32_t first = (st.unit_code >> (2 * (st.unit_k - 1))) & 3u;
#pragma unroll
    for (int j = 0; j < 6; j++) {
      if (j < st.unit_k) {
        const uint32_t b = (st.unit_code >> (2 * (st.unit_k - 1 - j))) & 3u;
        res.unit[j] = (char)((alpha >> (8 * b)) & 0xffu);
        homo = homo && (b == first);
      }
    }
    int rc = st.rc;
    if (homo) {
#pragma unroll
      for (int j = 1; j < 6; j++) res.unit[j] = 0;
      rc *= st.unit_k;
    }
    res.repeat_count = (uint16_t)rc;
  }
  store_result(out, s, res);
}

// One segment on one warp: stage it, run the ladder from rung `start_k` with the given state, store the result.
// ws.tab must be all zero on entry (it is left all zero).
template <int MAXLEN>
__device__ __forceinline__ void warp_scan_segment(WarpScratch<MAXLEN> &ws, const uint32_t *__restrict__ seq,
                                                  const uint32_t *__restrict__ nmask, const strgpu_segment sg, uint32_t s,
                                                  const uint16_t *__restrict__ thr, int lane, int start_k, ScanState st,
                                                  strgpu_repeat *__restrict__ out, int *status) {
  constexpr int MAXR2 = (MAXLEN / 2 + 31) / 32, MAXR3 = (MAXLEN / 3 + 31) / 32, MAXR4 = (MAXLEN / 4 + 31) / 32,
                MAXR5 = (MAXLEN / 5 + 31) / 32, MAXR6 = (MAXLEN / 6 + 31) / 32;
  constexpr int MAXP = (MAXLEN + 31) / 32;
  const int L = sg.len;
  if (L > MAXLEN || L > STRGPU_MAX_SEGMENT_LEN) {  // warp-uniform
    if (lane == 0) {
      atomicExch(status, (int)STRGPU_ERR_TOO_LONG);
      emit_result(out, s, ScanState{-1, 0u, 0, 0});
    }
    return;
  }
  // ---- stage the segment: aligned big-endian words, base 0 at bit 31 of sw[0]
  const int n_words = (2 * L + 31) >> 5;
  __syncwarp();
  for (int l = lane; l < n_words + 1; l += 32) {
    uint32_t v = 0;
    if (l < n_words) {
      const uint32_t g = (sg.base_off >> 4) + (uint32_t)l;
      const uint32_t hi = __byte_perm(seq[g], 0, 0x0123);
      const uint32_t lo = __byte_perm(seq[g + 1], 0, 0x0123);
      v = __funnelshift_l(lo, hi, 2u * (sg.base_off & 15u));
    }
    ws.sw[l] = v;
  }
  const bool has_n = (sg.flags & STRGPU_SEG_HAS_N) != 0;
  int n_count = 0;
  if (has_n) {  // warp-uniform
    const int n_nw = (L + 31) >> 5;
    for (int l = lane; l < n_nw + 1; l += 32) {
      uint32_t v = 0;
      if (l < n_nw) {
        const uint32_t g = (sg.base_off >> 5) + (uint32_t)l;
        v = __funnelshift_r(nmask[g], nmask[g + 1], sg.base_off & 31u);
        const int rem = L - 32 * l;
        if (rem < 32) v &= (1u << rem) - 1u;
      }
      ws.nm[l] = v;
      n_count += __popc(v);
    }
    n_count = __reduce_add_sync(kFull, n_count);
  }
  __syncwarp();

  if (n_count <= 20) {  // utils.nim:238
    const int pclass = sg.pclass < STRGPU_MAX_PCLASS ? sg.pclass : STRGPU_MAX_PCLASS - 1;
    const uint16_t *tp = thr + (size_t)(pclass * 5) * kThrLen + L;
    const uint16_t *tg = thr + (size_t)(STRGPU_MAX_PCLASS * 5) * kThrLen + L;
    bool go = true;
    if (start_k <= 2) go = ladder_step<2, MAXR2, MAXP>(ws.sw, ws.nm, has_n, ws.tab, L, lane, tp[0], tg[0], st);
    if (go && start_k <= 3) go = ladder_step<3, MAXR3, MAXP>(ws.sw, ws.nm, has_n, ws.tab, L, lane, tp[kThrLen], tg[kThrLen], st);
    if (go && start_k <= 4) go = ladder_step<4, MAXR4, MAXP>(ws.sw, ws.nm, has_n, ws.tab, L, lane, tp[2 * kThrLen], tg[2 * kThrLen], st);
    if (go && start_k <= 5) go = ladder_step<5, MAXR5, MAXP>(ws.sw, ws.nm, has_n, ws.tab, L, lane, tp[3 * kThrLen], tg[3 * kThrLen], st);
    if (go) ladder_step<6, MAXR6, MAXP>(ws.sw, ws.nm, has_n, ws.tab, L, lane, tp[4 * kThrLen], tg[4 * kThrLen], st);
  }
  if (lane == 0) emit_result(out, s, st);
}

template <int MAXLEN, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) repeat_scan_warp(const uint32_t *__restrict__ seq,
                                                               const uint32_t *__restrict__ nmask,
                                                               const strgpu_segment *__restrict__ segs, uint32_t n_seg,
                                                               const uint16_t *__restrict__ thr,
                                                               strgpu_repeat *__restrict__ out, int *status) {
  __shared__ WarpScratch<MAXLEN> scratch[WARPS];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  WarpScratch<MAXLEN> &ws = scratch[warp];
  for (int i = lane; i < WarpScratch<MAXLEN>::kTab; i += 32) ws.tab[i] = 0;
  __syncwarp();
  const uint32_t warps_total = gridDim.x * WARPS;
  for (uint32_t s = blockIdx.x * WARPS + warp; s < n_seg; s += warps_total)
    warp_scan_segment<MAXLEN>(ws, seq, nmask, segs[s], s, thr, lane, 2, ScanState{-1, 0u, 0, 0}, out, status);
}

// ================================================================================================
// K1 v3: one LANE per segment (segments of <= 160 bases without non-ACGT bases).
//   * each lane keeps its read as a private column of shared-memory words ([word][lane]: bank == lane, conflict free),
//   * k = 2, 3 and 4 are counted in ONE pass over the read: every window's min-rotation class comes from a small shared
//     LUT and is counted in the lane's private column of uint32 counters ([class][lane], conflict free); the three
//     histograms are independent read-modify-write chains, interleaved by hand so three shared-memory loads are in
//     flight per lane; the running leader is updated exactly like Seq.inc (strict >, utils.nim:192-195),
//   * the ladder (utils.nim:250-265) then runs on the three (M, leader) pairs; the recount is bit-parallel over the
//     ten words (popcount when no two matches can overlap, a run-wise greedy walk otherwise),
//   * lanes that survive to k = 5 (~12 % of random reads) are compacted into a warp-local queue and counted 32 at a
//     time with packed uint8 counters; only k = 6 survivors (~1 %), segments with N and segments longer than 160
//     bases are finished one per warp by the compact warp-per-segment code.
// Every warp owns its shared-memory region and its queues: there is no block-level barrier after start-up.
// ================================================================================================
constexpr int kLaneThreads = 640;               // one CTA of 20 independent warps per SM (shared-memory bound)
constexpr int kLaneWarps = kLaneThreads / 32;
constexpr int kLaneWords = 11;                 // ten words hold 160 bases; one more absorbs the re-alignment shift
constexpr int kCls2 = 10, kCls3 = 24, kCls4 = 70, kCls5 = 208;
constexpr int kCls4Words = (kCls4 + 3) / 4;       // 4-mer classes are counted in packed uint8 (18 words) to keep smem small
constexpr int kTabWords = kCls2 + kCls3 + kCls4Words;  // 52 words per lane; k = 5 reuses them as 208 packed uint8
constexpr int kLut2 = 0, kLut3 = 16, kLut4 = 80, kRev234 = 336, kLut5 = 440, kRev5 = 1464;  // offsets into the uint16 table
constexpr int kLutTotal = 1672;
constexpr int kQueueCap = 64;                  // a batch is taken at 32 entries and a stage adds at most 32
// Q3 (3 words/entry), Q4, Q5, Q6 (2 words/entry), Q2 (filter survivors, 1 word/entry), QW (1 word/entry, cap 32)
constexpr int kQ3Off = 0, kQ4Off = kQueueCap * 3, kQ5Off = kQ4Off + kQueueCap * 2, kQ6Off = kQ5Off + kQueueCap * 2,
              kQ2Off = kQ6Off + kQueueCap * 2, kQWOff = kQ2Off + kQueueCap;
constexpr int kQueueWords = kQWOff + 32;
constexpr int kWarpSmemWords = kTabWords * 32 + kLaneWords * 32 + kQueueWords;
constexpr int kLaneSmemBytes = kLaneWarps * kWarpSmemWords * 4 + kLutTotal * 2 + 16;
constexpr int lane_smem_bytes(int warps) { return warps * kWarpSmemWords * 4 + kLutTotal * 2 + 16; }
static_assert(sizeof(WarpScratch<512>) <= (size_t)kTabWords * 32 * 4, "warp scratch must fit in the warp's counter region");
static_assert(kCls5 / 4 <= kTabWords, "k = 5 counters must fit");

// read.count(s) for one lane: greedy leftmost non-overlapping matches of the K-base pattern (utils.nim:254).
// One copy for all k, kept out of line (instruction-cache footprint).  `scratch` is the lane's counter column (dead by
// now); `magic` = 65536 / K + 1 turns the one division of the slow path into a multiply.
// NW: words that hold the segment (4: segments of <= 64 bases, the soft clips; 10: up to 160 bases).
template <int NW>
__device__ __noinline__ int lane_recount(const uint32_t *rd, uint32_t *scratch, int L, uint32_t pat, int K, uint32_t magic) {
  const int npos = L - K + 1;
  if (npos <= 0) return 0;
  constexpr uint32_t kLow = 0x55555555u;
  uint32_t w[NW + 1], m[NW];
#pragma unroll
  for (int i = 0; i < NW; i++) w[i] = rd[i * 32];
  w[NW] = 0;
#pragma unroll
  for (int i = 0; i < NW; i++) {  // keep positions < npos (position p of word i sits at bit 30 - 2p)
    const int n = npos - 16 * i;
    m[i] = n >= 16 ? kLow : (n <= 0 ? 0u : (kLow & ~((1u << (32 - 2 * n)) - 1u)));
  }
#pragma unroll 1
  for (int j = 0; j < K; j++) {
    const uint32_t rep = ((pat >> (2 * (K - 1 - j))) & 3u) * kLow;
    uint32_t e[NW + 1];
#pragma unroll
    for (int i = 0; i < NW + 1; i++) {
      const uint32_t t = w[i] ^ rep;
      e[i] = ~(t | (t >> 1)) & kLow;          // slot LSB set <=> that base equals pattern base j
    }
#pragma unroll
    for (int i = 0; i < NW; i++) m[i] &= __funnelshift_l(e[i + 1], e[i], 2 * j);
  }
  uint32_t conflict = 0;
#pragma unroll 1
  for (int d = 1; d < K; d++) {
#pragma unroll
    for (int i = 0; i < NW; i++) conflict |= m[i] & __funnelshift_l(i < NW - 1 ? m[i + 1] : 0u, m[i], 2 * d);
  }
  int c = 0;
  if (conflict == 0) {  // no two matches closer than K: every match counts
#pragma unroll
    for (int i = 0; i < NW; i++) c += __popc(m[i]);
    return c;
  }
  if (K == 2) {
    // Self-overlap at K = 2 means a homopolymer pattern XX.  Matches then form runs of consecutive positions and the
    // greedy walk takes every other one from each run's start: ceil(r / 2) = the run's positions that share the
    // parity of its start.  Runs that start on even positions are isolated with one multi-word add (the carry
    // runs through exactly those runs), so the count is two popcounts per word instead of a serial walk.
    uint32_t f[NW], sel = 0;
    int cy = 0;
    c = 0;
#pragma unroll
    for (int i = 0; i < NW; i++) {            // LSB-first: position p of word i -> bits 2p, 2p+1 (both set for a match)
      const uint32_t r = __brev(m[i]);        // match bit of position p now at bit 2p + 1
      f[i] = r | (r >> 1);
    }
    uint32_t prev_top = 0;                     // was the last position of the previous word a match?
#pragma unroll
    for (int i = 0; i < NW; i++) {
      const uint32_t prevm = (f[i] << 2) | (prev_top ? 3u : 0u);   // match state of position p - 1
      const uint32_t starts = f[i] & ~prevm & 0x55555555u;          // low bit of each run's first position
      const uint32_t es = starts & 0x11111111u;                     // runs starting on an even position
      const unsigned long long sum = (unsigned long long)f[i] + es + (unsigned)cy;
      cy = (int)(sum >> 32);
      const uint32_t even_runs = f[i] & ~(uint32_t)sum;             // every bit of a run that started even (carry cleared it)
      const uint32_t odd_runs = f[i] & (uint32_t)sum;               // (a carry entering from the previous word continues a run)
      sel = (even_runs & 0x11111111u) | (odd_runs & 0x44444444u);   // even-start runs count even positions, odd-start odd ones
      c += __popc(sel);
      prev_top = f[i] >> 31;
    }
    return c;
  }
  // other self-overlapping patterns (ACAC.., AAA.. at k >= 3): walk the runs of consecutive match positions
#pragma unroll
  for (int i = 0; i < NW; i++) scratch[i * 32] = m[i];
  int next = 0;  // first position the greedy walk may use
#pragma unroll 1
  for (int i = 0; i < NW; i++) {
    uint32_t mm = scratch[i * 32];
    while (mm) {
      const int hb = 31 - __clz(mm);                       // earliest remaining match of this word
      const uint32_t gap = ~mm & kLow & ((1u << hb) - 1u);  // first non-match slot after it
      const int hb2 = gap ? 31 - __clz(gap) : -2;
      const int p = 16 * i + ((30 - hb) >> 1);
      const int e = 16 * i + ((30 - hb2) >> 1);            // run of consecutive matches [p, e)
      const int s0 = p > next ? p : next;
      if (s0 < e) {
        const int n = (int)(((uint32_t)(e - s0 + K - 1) * magic) >> 16);
        c += n;
        next = s0 + n * K;
      }
      mm = gap ? (mm & ((1u << hb2) - 1u)) : 0u;
    }
  }
  return c;
}

// Compact warp-per-segment path (runtime k, rolled loops) for the few segments the lane path hands off: same
// arithmetic as count_k / recount_k / ladder_step above, written for code size instead of speed.
__device__ __noinline__ void warp_scan_compact(WarpScratch<512> &ws, const uint32_t *__restrict__ seq,
                                               const uint32_t *__restrict__ nmask, const strgpu_segment sg, uint32_t s,
                                               const uint16_t *__restrict__ thr, int lane, int start_k, ScanState st,
                                               strgpu_repeat *__restrict__ out, int *status) {
  const int L = sg.len;
  if (L > STRGPU_MAX_SEGMENT_LEN) {
    if (lane == 0) {
      atomicExch(status, (int)STRGPU_ERR_TOO_LONG);
      emit_result(out, s, ScanState{-1, 0u, 0, 0});
    }
    return;
  }
  const int n_words = (2 * L + 31) >> 5;
  __syncwarp();
  for (int l = lane; l < n_words + 1; l += 32) {
    uint32_t v = 0;
    if (l < n_words) {
      const uint32_t g = (sg.base_off >> 4) + (uint32_t)l;
      v = __funnelshift_l(__byte_perm(seq[g + 1], 0, 0x0123), __byte_perm(seq[g], 0, 0x0123), 2u * (sg.base_off & 15u));
    }
    ws.sw[l] = v;
  }
  const bool has_n = (sg.flags & STRGPU_SEG_HAS_N) != 0;
  int n_count = 0;
  if (has_n) {
    const int n_nw = (L + 31) >> 5;
    for (int l = lane; l < n_nw + 1; l += 32) {
      uint32_t v = 0;
      if (l < n_nw) {
        const uint32_t g = (sg.base_off >> 5) + (uint32_t)l;
        v = __funnelshift_r(nmask[g], nmask[g + 1], sg.base_off & 31u);
        const int rem = L - 32 * l;
        if (rem < 32) v &= (1u << rem) - 1u;
      }
      ws.nm[l] = v;
      n_count += __popc(v);
    }
    n_count = __reduce_add_sync(kFull, n_count);
  }
  __syncwarp();
  if (n_count <= 20) {
    const int pclass = sg.pclass < STRGPU_MAX_PCLASS ? sg.pclass : STRGPU_MAX_PCLASS - 1;
    const uint32_t lane_le = kFull >> (31 - lane);
#pragma unroll 1
    for (int K = start_k; K <= 6; K++) {
      const uint32_t kmask = (1u << (2 * K)) - 1u;
      const int W = L / K;
      const int rounds = (W + 31) >> 5;
      int M = 0;
      uint32_t leader = kmask;
#pragma unroll 1
      for (int r = 0; r < rounds; r++) {
        const int w = 32 * r + lane;
        const bool valid = w < W;
        const uint32_t bit = valid ? 2u * (uint32_t)(K * w) : 0u;
        uint32_t x = __funnelshift_l(ws.sw[(bit >> 5) + 1], ws.sw[bit >> 5], bit & 31u) >> (32 - 2 * K);
        uint32_t c = x;
#pragma unroll 1
        for (int j = 1; j < K; j++) {
          x = ((x << 2) | (x >> (2 * K - 2))) & kmask;
          c = min(c, x);
        }
        ws.wc[32 * r + lane] = valid ? (uint16_t)c : (uint16_t)0xffffu;
        const uint32_t grp = __match_any_sync(kFull, valid ? c : (0x80000000u | (uint32_t)lane));
        int base = 0;
        if (rounds > 1) {
          if (valid) base = ws.tab[c];
          __syncwarp();
          if (valid && lane == 31 - __clz(grp)) ws.tab[c] = (uint8_t)(base + __popc(grp));
          __syncwarp();
        }
        const int occ = valid ? base + __popc(grp & lane_le) : 0;
        const int rmax = __reduce_max_sync(kFull, occ);
        if (rmax > M) {
          M = rmax;
          const uint32_t b = __ballot_sync(kFull, valid && occ == rmax);
          leader = __shfl_sync(kFull, c, __ffs(b) - 1);
        }
      }
      if (rounds > 1) {
#pragma unroll 1
        for (int r = 0; r < rounds; r++) {
          const uint32_t c = ws.wc[32 * r + lane];
          if (c != 0xffffu) ws.tab[c] = 0;
        }
        __syncwarp();
      }
      const int thr_p = thr[(size_t)(pclass * 5 + K - 2) * kThrLen + L];
      const int thr_giveup = thr[(size_t)(STRGPU_MAX_PCLASS * 5 + K - 2) * kThrLen + L];
      int score = M * K;
      if (score <= st.best) {
        if (M < thr_giveup) break;
        continue;
      }
      // recount
      const int npos = L - K + 1;
      int cnt = 0, next = 0;
#pragma unroll 1
      for (int r = 0; 32 * r < npos; r++) {
        const int i = 32 * r + lane;
        const bool valid = i < npos;
        const uint32_t bit = valid ? 2u * (uint32_t)i : 0u;
        const uint32_t x = __funnelshift_l(ws.sw[(bit >> 5) + 1], ws.sw[bit >> 5], bit & 31u) >> (32 - 2 * K);
        bool eq = valid && x == leader;
        if (has_n) eq = eq && ((__funnelshift_r(ws.nm[r], ws.nm[r + 1], lane) & ((1u << K) - 1u)) == 0u);
        uint32_t m = __ballot_sync(kFull, eq);
        const int rel = next - 32 * r;
        if (rel > 0) m = (rel >= 32) ? 0u : (m & (kFull << rel));
        while (m) {
          const int nx = __ffs(m) - 1 + K;
          cnt++;
          m = (nx >= 32) ? 0u : (m & (kFull << nx));
          next = 32 * r + nx;
        }
      }
      score = cnt * K;
      if (score < st.best) continue;
      st.best = score;
      if (cnt > thr_p) {
        st.unit_code = leader;
        st.unit_k = K;
        st.rc = cnt;
      }
    }
  }
  if (lane == 0) emit_result(out, s, st);
}


struct Lead {
  int M;
  uint32_t off;
};
__device__ __forceinline__ uint32_t *slot_at(uint32_t *tab, uint32_t off) {
  return reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(tab) + off);
}

// Time-stamped counters for k = 2 and k = 3: a class's word holds count << 8 | (255 - t), t = index of the window that
// last incremented it.  One update is LDS, LOP3, IADD3, STS -- no per-window leader bookkeeping: after the pass the
// leader is the class with the largest word, because among classes with the maximal count the one whose LAST increment
// came first is the one that reached that count first (Seq.inc keeps the first to reach the maximum, utils.nim:192-195).
__device__ __forceinline__ void stamp2(uint32_t *tab, uint32_t oa, uint32_t ob, uint32_t adda, uint32_t addb) {
  uint32_t *pa = slot_at(tab, oa), *pb = slot_at(tab, ob);   // two independent histograms: both loads before both stores
  const uint32_t va = *pa, vb = *pb;
  *pa = (va | 0xffu) + adda;
  *pb = (vb | 0xffu) + addb;
}
__device__ __forceinline__ void stamp1(uint32_t *tab, uint32_t o, uint32_t add) {
  uint32_t *p = slot_at(tab, o);
  *p = (*p | 0xffu) + add;
}

// count(read, k, counts[k]) for k = 2 and k = 3 in one pass (utils.nim:205-211 twice): the read is walked in groups of
// twelve bases (six 2-mer and four 3-mer windows).  Returns each k's best word << 5 | class (0 when there is no window).
__device__ __forceinline__ void lane_count23(const uint32_t *rd, uint32_t *tab, const uint16_t *lut, int L, uint32_t &best2,
                                             uint32_t &best3) {
#pragma unroll
  for (int c = 0; c < kCls2 + kCls3; c++) tab[c * 32] = 0;
  const int n_groups = L / 12;
  uint32_t add2 = 256u, add3 = 256u;   // 256 - (index of the group's first window)
#pragma unroll 1
  for (int g = 0; g < n_groups; g++) {
    const uint32_t bit = 24u * g;
    const uint32_t x = __funnelshift_l(rd[((bit >> 5) + 1) * 32], rd[(bit >> 5) * 32], bit & 31u) >> 8;  // the group's 24 bits
    const uint32_t a0 = lut[kLut2 + (x >> 20)], a1 = lut[kLut2 + ((x >> 16) & 15u)], a2 = lut[kLut2 + ((x >> 12) & 15u)],
                   a3 = lut[kLut2 + ((x >> 8) & 15u)], a4 = lut[kLut2 + ((x >> 4) & 15u)], a5 = lut[kLut2 + (x & 15u)];
    const uint32_t b0 = lut[kLut3 + (x >> 18)], b1 = lut[kLut3 + ((x >> 12) & 63u)], b2 = lut[kLut3 + ((x >> 6) & 63u)],
                   b3 = lut[kLut3 + (x & 63u)];
    stamp2(tab, a0, b0, add2, add3);
    stamp2(tab, a1, b1, add2 - 1u, add3 - 1u);
    stamp2(tab, a2, b2, add2 - 2u, add3 - 2u);
    stamp2(tab, a3, b3, add2 - 3u, add3 - 3u);
    stamp1(tab, a4, add2 - 4u);
    stamp1(tab, a5, add2 - 5u);
    add2 -= 6u;
    add3 -= 4u;
  }
  // tail: fewer than twelve bases left
  const int r2 = L / 2 - 6 * n_groups, r3 = L / 3 - 4 * n_groups;
  const uint32_t base = 24u * n_groups;
#pragma unroll 1
  for (int t = 0; t < r2; t++) {
    {
      const uint32_t bit = base + 4u * t;
      stamp1(tab, lut[kLut2 + (__funnelshift_l(rd[((bit >> 5) + 1) * 32], rd[(bit >> 5) * 32], bit & 31u) >> 28)], add2 - (uint32_t)t);
    }
    if (t < r3) {
      const uint32_t bit = base + 6u * t;
      stamp1(tab, lut[kLut3 + (__funnelshift_l(rd[((bit >> 5) + 1) * 32], rd[(bit >> 5) * 32], bit & 31u) >> 26)], add3 - (uint32_t)t);
    }
  }
  best2 = 0;
  best3 = 0;
#pragma unroll
  for (int c = 0; c < kCls2; c++) best2 = max(best2, (tab[c * 32] << 5) | (uint32_t)c);
#pragma unroll
  for (int c = 0; c < kCls3; c++) best3 = max(best3, (tab[(kCls2 + c) * 32] << 5) | (uint32_t)c);
}

// count(read, 4, counts[4]) with 70 packed uint8 counters ([class / 4][lane] words 34..51 of the lane's column)
__device__ __forceinline__ void lane_count4(const uint32_t *rd, uint32_t *tab, const uint16_t *lut, int L, int &M, uint32_t &leader) {
#pragma unroll
  for (int c = 0; c < kCls4Words; c++) tab[(kCls2 + kCls3 + c) * 32] = 0;
  M = 0;
  uint32_t lead = 0xffffffffu;
  const int W = L / 4;
  const int n_words = W / 4;
#pragma unroll 1
  for (int wi = 0; wi < n_words; wi++) {
    const uint32_t x = rd[wi * 32];
    const uint32_t e0 = lut[kLut4 + (x >> 24)], e1 = lut[kLut4 + ((x >> 16) & 255u)], e2 = lut[kLut4 + ((x >> 8) & 255u)],
                   e3 = lut[kLut4 + (x & 255u)];
#pragma unroll
    for (int t = 0; t < 4; t += 2) {
      // two windows at a time: both counter words are loaded before either is stored (two loads in flight); if the
      // windows share a word the second update is applied on top of the first
      const uint32_t ea = t == 0 ? e0 : e2, eb = t == 0 ? e1 : e3;
      uint32_t *pa = slot_at(tab, ea & 0xff80u), *pb = slot_at(tab, eb & 0xff80u);
      const uint32_t sa = ea & 31u, sb = eb & 31u;
      const uint32_t va = *pa + (1u << sa);
      const uint32_t vb0 = *pb;
      const uint32_t vb = (pa == pb ? va : vb0) + (1u << sb);
      *pa = va;
      *pb = vb;
      const int ca = (int)((va >> sa) & 0xffu), cb = (int)((vb >> sb) & 0xffu);
      if (ca > M) { M = ca; lead = ea; }   // strict >: the earlier leader keeps ties
      if (cb > M) { M = cb; lead = eb; }
    }
  }
  for (int w = 4 * n_words; w < W; w++) {
    const uint32_t e = lut[kLut4 + ((rd[(w >> 2) * 32] >> (24 - 8 * (w & 3))) & 255u)];
    uint32_t *p = slot_at(tab, e & 0xff80u);
    const uint32_t sh = e & 31u;
    const uint32_t v = *p + (1u << sh);
    *p = v;
    const int c = (int)((v >> sh) & 0xffu);
    if (c > M) { M = c; lead = e; }
  }
  leader = (lead == 0xffffffffu) ? 0xffu : (uint32_t)lut[kRev234 + kCls2 + kCls3 + ((lead >> 7) - (kCls2 + kCls3)) * 4 + ((lead & 31u) >> 3)];
}

// one rung of the ladder (utils.nim:246-265) given this k's count result.  Returns false on `break`.
// wide: recount over ten words (any segment) instead of four (segments of <= 64 bases); callers make it uniform over the lanes
// that recount together so that only one of the two instances runs
__device__ __forceinline__ bool lane_decide(const uint32_t *rd, uint32_t *tab, int L, int K, int M, uint32_t leader, int thr_p,
                                            int thr_giveup, ScanState &st, bool wide) {
  int score = M * K;
  if (score <= st.best) return !(M < thr_giveup);
  const uint32_t magic = 65536u / (uint32_t)K + 1u;
  const int c = wide ? lane_recount<10>(rd, tab, L, leader, K, magic) : lane_recount<4>(rd, tab, L, leader, K, magic);
  score = c * K;
  if (score < st.best) return true;
  st.best = score;
  if (c > thr_p) {
    st.unit_code = leader;
    st.unit_k = K;
    st.rc = c;
  }
  return true;
}

// stage the lane's read: eleven words, re-aligned so that base 0 sits at bit 31 of word 0
__device__ __forceinline__ void lane_stage(const uint32_t *__restrict__ seq, const strgpu_segment &sg, uint32_t *rd) {
  const uint32_t g = sg.base_off >> 4;
  const uint32_t sh = 2u * (sg.base_off & 15u);
  const int n_words = (2 * (int)sg.len + 31) >> 5;
  uint32_t raw[kLaneWords + 1];
#pragma unroll
  for (int j = 0; j < kLaneWords + 1; j++) raw[j] = (j <= n_words) ? __byte_perm(seq[g + j], 0, 0x0123) : 0u;
#pragma unroll
  for (int j = 0; j < kLaneWords; j++) rd[j * 32] = __funnelshift_l(raw[j + 1], raw[j], sh);
}

// count(read, 5, counts[5]) with 208 packed uint8 counters ([class / 4][lane] words)
__device__ __forceinline__ void lane_count5(const uint32_t *rd, uint32_t *tab, const uint16_t *lut, int L, int &M, uint32_t &leader) {
#pragma unroll
  for (int c = 0; c < kCls5 / 4; c++) tab[c * 32] = 0;
  M = 0;
  uint32_t lead = 0xffffffffu;
  const int W = L / 5;
#pragma unroll 2
  for (int j = 0; j < W; j++) {
    const uint32_t bit = 10u * j;
    const uint32_t code = __funnelshift_l(rd[((bit >> 5) + 1) * 32], rd[(bit >> 5) * 32], bit & 31u) >> 22;
    const uint32_t e = lut[kLut5 + code];        // (class / 4) * 128 | (class % 4) * 8
    uint32_t *p = slot_at(tab, e & 0xff80u);
    const uint32_t sh = e & 31u;
    const uint32_t v = *p + (1u << sh);
    *p = v;
    const int c = (int)((v >> sh) & 0xffu);
    if (c > M) { M = c; lead = e; }
  }
  leader = (lead == 0xffffffffu) ? 0x3ffu : (uint32_t)lut[kRev5 + (lead >> 7) * 4 + ((lead & 31u) >> 3)];
}

// count(read, 6, counts[6]) for one lane: at most 26 windows, counted in an open-addressing table of 52 slots
// ({canonical code + 1, count} per uint32 word of the lane's column); the min-rotation is taken in the ALU.
__device__ __forceinline__ void lane_count6(const uint32_t *rd, uint32_t *tab, int L, int &M, uint32_t &leader) {
#pragma unroll
  for (int c = 0; c < kTabWords; c++) tab[c * 32] = 0;
  M = 0;
  leader = 0xfffu;
  const int W = L / 6;
#pragma unroll 1
  for (int j = 0; j < W; j++) {
    const uint32_t bit = 12u * j;
    uint32_t x = __funnelshift_l(rd[((bit >> 5) + 1) * 32], rd[(bit >> 5) * 32], bit & 31u) >> 20;
    const uint32_t c = min_rotation<6>(x);
    const uint32_t key = (c + 1u) << 16;
    uint32_t h = (((c * 40503u) & 0xffffu) * (uint32_t)kTabWords) >> 16;
    int cnt;
    while (true) {
      const uint32_t v = tab[h * 32];
      if (v == 0u) { tab[h * 32] = key | 1u; cnt = 1; break; }
      if ((v & 0xffff0000u) == key) { tab[h * 32] = v + 1u; cnt = (int)(v & 0xffffu) + 1; break; }
      h = (h + 1u == (uint32_t)kTabWords) ? 0u : h + 1u;
    }
    if (cnt > M) { M = cnt; leader = c; }   // strict >: the earlier leader keeps ties
  }
}

// Warp-local FIFOs of segments waiting for their next stage.  Q4..Q6 entries are two words, Q3 entries three:
//   0: segment index      1: best << 24 | repeat_count << 16 | unit_k << 12 | unit_code      2 (Q3 only): M3 | leader3 << 8
// (entries are written after the k = 2 rung, so 0 <= best <= 160 and repeat_count <= 80)
template <int WORDS>
__device__ __forceinline__ void queue_push(uint32_t *buf, int &n, bool want, int lane, uint32_t s, const ScanState &st, uint32_t extra) {
  const uint32_t m = __ballot_sync(kFull, want);
  if (want) {
    uint32_t *e = buf + WORDS * (n + __popc(m & ((1u << lane) - 1u)));
    e[0] = s;
    if (WORDS >= 2) e[1] = ((uint32_t)(st.best & 0xff) << 24) | ((uint32_t)(st.rc & 0xff) << 16) | ((uint32_t)st.unit_k << 12) | (st.unit_code & 0xfffu);
    if (WORDS >= 3) e[2] = extra;
  }
  n += __popc(m);
  __syncwarp();
}
template <int WORDS>
__device__ __forceinline__ void queue_read(const uint32_t *buf, int i, uint32_t &s, ScanState &st, uint32_t &extra) {
  const uint32_t *e = buf + WORDS * i;
  s = e[0];
  const uint32_t a = e[1];
  st.best = (int)(a >> 24);
  st.rc = (int)((a >> 16) & 0xffu);
  st.unit_k = (int)((a >> 12) & 0xfu);
  st.unit_code = a & 0xfffu;
  extra = WORDS >= 3 ? e[2] : 0u;
}

// Segment s of the batch: the first u.n_reads are implicit whole reads of one length on a fixed stride (no descriptor
// is read; non-ACGT bases are detected from the mask itself), the rest come from the descriptor array.
__device__ __forceinline__ strgpu_segment load_segment(const strgpu_segment *__restrict__ segs, const uint32_t *__restrict__ nmask,
                                                       const UniformReads &u, uint32_t s) {
  if (s >= u.n_reads) return segs[s - u.n_reads];
  strgpu_segment sg;
  sg.base_off = s * u.stride;
  sg.len = (uint16_t)u.read_len;
  sg.pclass = (uint8_t)u.pclass;
  sg.flags = 0;
  if (nmask != nullptr) {
    uint32_t any = 0;
    const uint32_t first = sg.base_off >> 5, last = (sg.base_off + u.read_len + 31u) >> 5;
    for (uint32_t w = first; w < last; w++) {
      uint32_t v = nmask[w];
      if (w == first) v &= kFull << (sg.base_off & 31u);
      const uint32_t end = sg.base_off + u.read_len;
      if (w == (end >> 5) && (end & 31u)) v &= (1u << (end & 31u)) - 1u;
      if (w > (end >> 5) || (w == (end >> 5) && !(end & 31u))) v = 0;
      any |= v;
    }
    if (any) sg.flags = STRGPU_SEG_HAS_N;
  }
  return sg;
}

// ------------------------------------------------------------------------------------------------------------------
// Stage 1: the exact pre-filter.  get_repeat (utils.nim:236-271) can only return a unit when, at some rung k, the
// greedy non-overlapping count of a k-mer s (read.count(s), utils.nim:254) exceeds int(L * p / k) (utils.nim:259).  That
// count is at most the number of positions where s starts, which is at most the number of positions where the 2-mer
// s[0..2) starts.  So a segment whose most frequent 2-mer (all L - 1 overlapping positions) occurs at most
// T = min_k int(L * p / k) times returns the empty unit with repeat_count 0 whatever path the ladder takes: it is
// finished here, bit-exactly, without ever being counted.  ~95 % of the reads of a sequencing run end here.
//
// The 16 occurrence counts are taken bit-parallel.  Two 16-base words are merged into dense 32-position planes
// (hi / lo bit of every base; the odd bits of a plane belong to the first word, the even bits to the second), once for
// the bases themselves (P) and once for their successors (Q); E_a = positions holding base a, E_a & (Q == b) =
// positions where the 2-mer ab starts.  Only b = 0..2 are counted, b = 3 follows from popc(E_a).
constexpr uint32_t kOdd = 0xaaaaaaaau, kEven = 0x55555555u;

// valid-position masks of the five word pairs for a segment of L bases (positions 0 .. L - 2 start a 2-mer)
__device__ __forceinline__ void filter_masks(int L, uint32_t (&V)[5]) {
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const int v = L - 1 - 32 * j;
    const int v0 = min(max(v, 0), 16), v1 = min(max(v - 16, 0), 16);
    const uint32_t m0 = v0 >= 16 ? kOdd : (kOdd & ~(kFull >> (2 * v0)));
    const uint32_t m1 = v1 >= 16 ? kEven : (kEven & ~(kFull >> (2 * v1)));
    V[j] = m0 | m1;
  }
}

template <int VARIANT>
__device__ __forceinline__ int lane_filter_max2(const uint32_t (&w)[11], const uint32_t (&V)[5]) {
  uint32_t E[5][4], Qh[5], Ql[5];
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const uint32_t w0 = w[2 * j], w1 = w[2 * j + 1], w2 = w[2 * j + 2];
    const uint32_t n0 = __funnelshift_l(w1, w0, 2), n1 = __funnelshift_l(w2, w1, 2);   // successor of every base
    const uint32_t Ph = (w0 & kOdd) | ((w1 >> 1) & kEven), Pl = ((w0 << 1) & kOdd) | (w1 & kEven);
    Qh[j] = (n0 & kOdd) | ((n1 >> 1) & kEven);
    Ql[j] = ((n0 << 1) & kOdd) | (n1 & kEven);
    E[j][0] = ~Ph & ~Pl & V[j];
    E[j][1] = ~Ph & Pl & V[j];
    E[j][2] = Ph & ~Pl & V[j];
    E[j][3] = Ph & Pl & V[j];
  }
  int best = 0;
#pragma unroll
  for (int a = 0; a < 4; a++) {
    int ca[4];
#pragma unroll
    for (int b = 0; b < 4; b++) {
      uint32_t m[5];
#pragma unroll
      for (int j = 0; j < 5; j++) {
        const uint32_t e = E[j][a];
        m[j] = b == 0 ? (e & ~Qh[j] & ~Ql[j]) : (b == 1 ? (e & ~Qh[j] & Ql[j]) : (b == 2 ? (e & Qh[j] & ~Ql[j]) : e));
      }
      if (VARIANT == 0) {
        ca[b] = __popc(m[0]) + __popc(m[1]) + __popc(m[2]) + __popc(m[3]) + __popc(m[4]);
      } else {  // carry-save: five words -> one "ones" and two "twos" words, three popcounts instead of five
        const uint32_t s1 = m[0] ^ m[1] ^ m[2], c1 = (m[0] & m[1]) | (m[2] & (m[0] | m[1]));
        const uint32_t s2 = s1 ^ m[3] ^ m[4], c2 = (s1 & m[3]) | (m[4] & (s1 | m[3]));
        ca[b] = __popc(s2) + 2 * (__popc(c1) + __popc(c2));
      }
    }
    ca[3] -= ca[0] + ca[1] + ca[2];
    best = max(best, max(max(ca[0], ca[1]), max(ca[2], ca[3])));
  }
  return best;
}

// the lane's segment as eleven aligned words in registers (base 0 at bit 31 of w[0]); w[10] only ever feeds a masked slot
__device__ __forceinline__ void lane_load(const uint32_t *__restrict__ seq, const strgpu_segment &sg, uint32_t (&w)[11]) {
  const uint32_t g = sg.base_off >> 4;
  const uint32_t sh = 2u * (sg.base_off & 15u);
  const int n_words = (2 * (int)sg.len + 31) >> 5;
  uint32_t raw[kLaneWords];
#pragma unroll
  for (int j = 0; j < kLaneWords; j++) raw[j] = (j <= n_words) ? __byte_perm(seq[g + j], 0, 0x0123) : 0u;
#pragma unroll
  for (int j = 0; j < kLaneWords - 1; j++) w[j] = __funnelshift_l(raw[j + 1], raw[j], sh);
  w[kLaneWords - 1] = 0u;
}

// ------------------------------------------------------------------------------------------------------------------
// repeat_prefilter: the pre-filter as a streaming kernel of its own (no shared-memory tables, 64 registers, 32 resident
// warps per SM, so the HBM latency of the read words is covered by occupancy).  Every segment is read once; segments the
// bound finishes get their empty result here, the others -- and segments with non-ACGT bases or more than 160 bases --
// are appended to the survivor list that repeat_scan_lane then works through 32 at a time.
//
// Same counting as lane_filter_max2 with a cheaper successor plane: word j is paired with word j + 5 (odd bits: positions
// 16 j .., even bits: positions 16 (j + 5) ..), so the successor of every position of plane j is plane j shifted up by one
// slot with the top slot pair of plane j + 1 shifted in -- one funnel shift.  CSA of the 16 streams are popcounted through
// carry-save adders (3 POPC + 4 LOP3 instead of 5 POPC): POPC runs on the 16-lane XU pipe, LOP3 on the 64-lane ALU pipe.
constexpr int kPreThreads = 256;
constexpr int kPreCsaDefault = 8;

__device__ __forceinline__ void prefilter_masks(int L, uint32_t (&V)[5]) {
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const int v0 = min(max(L - 1 - 16 * j, 0), 16), v1 = min(max(L - 1 - 16 * (j + 5), 0), 16);
    const uint32_t m0 = v0 >= 16 ? kOdd : (kOdd & ~(kFull >> (2 * v0)));
    const uint32_t m1 = v1 >= 16 ? kEven : (kEven & ~(kFull >> (2 * v1)));
    V[j] = m0 | m1;
  }
}

// returns the largest and the second largest of the 16 occurrence counts
// `one` is the runtime constant 1 (a kernel argument): a * one + b compiles to IMAD on the FMA pipe, which is idle here, instead of
// IADD3 on the ALU pipe, which is the bottleneck (every ALU instruction costs two issue cycles)
template <int CSA>
__device__ __forceinline__ void prefilter_top2(const uint32_t (&w)[10], const uint32_t (&V)[5], int &s1, int &s2, int one) {
  uint32_t Dh[6], Dl[6];
#pragma unroll
  for (int j = 0; j < 5; j++) {
    Dh[j] = (w[j] & kOdd) | (__umulhi(w[j + 5], 0x80000000u) & kEven);   // x >> 1 as a high multiply: FMA pipe, not ALU
    Dl[j] = ((w[j] << 1) & kOdd) | (w[j + 5] & kEven);
  }
  Dh[5] = w[5];        // only its top slot pair is used: base 0 of word 5 (the even bit feeds a slot that is never valid)
  Dl[5] = w[5] << 1;
  uint32_t E[5][4], Qh[5], Ql[5];
#pragma unroll
  for (int j = 0; j < 5; j++) {
    Qh[j] = __funnelshift_l(Dh[j + 1], Dh[j], 2);
    Ql[j] = __funnelshift_l(Dl[j + 1], Dl[j], 2);
    E[j][0] = ~Dh[j] & ~Dl[j] & V[j];
    E[j][1] = ~Dh[j] & Dl[j] & V[j];
    E[j][2] = Dh[j] & ~Dl[j] & V[j];
    E[j][3] = Dh[j] & Dl[j] & V[j];
  }
  s1 = 0;
  s2 = 0;
#pragma unroll
  for (int a = 0; a < 4; a++) {
    int ca[4];
#pragma unroll
    for (int b = 0; b < 4; b++) {
      uint32_t m[5];
#pragma unroll
      for (int j = 0; j < 5; j++) {
        const uint32_t e = E[j][a];
        m[j] = b == 0 ? (e & ~Qh[j] & ~Ql[j]) : (b == 1 ? (e & ~Qh[j] & Ql[j]) : (b == 2 ? (e & Qh[j] & ~Ql[j]) : e));
      }
      if (a * 4 + b >= CSA) {
        ca[b] = (((__popc(m[0]) * one + __popc(m[1])) * one + __popc(m[2])) * one + __popc(m[3])) * one + __popc(m[4]);
      } else {
        const uint32_t x1 = m[0] ^ m[1] ^ m[2], c1 = (m[0] & m[1]) | (m[2] & (m[0] | m[1]));
        const uint32_t x2 = x1 ^ m[3] ^ m[4], c2 = (x1 & m[3]) | (m[4] & (x1 | m[3]));
        ca[b] = (__popc(c1) * one + __popc(c2)) * (one + one) + __popc(x2);
      }
    }
    ca[3] = ((ca[0] * one + ca[1]) * one + ca[2]) * (-one) + ca[3];
    // the two largest of the four, merged into the running pair
    const int m1 = max(ca[0], ca[1]), n1 = min(ca[0], ca[1]), m2 = max(ca[2], ca[3]), n2 = min(ca[2], ca[3]);
    const int t1 = max(m1, m2), t2 = max(min(m1, m2), max(n1, n2));
    s2 = max(min(s1, t1), max(s2, t2));
    s1 = max(s1, t1);
  }
}

// ---- TMA plumbing (sm_90+ PTX): one mbarrier per staging buffer, bulk global -> shared copies completing on it
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}

// The filter decision for one lane-path segment whose words start at `src` (global or shared memory).
// Bound: c non-overlapping occurrences of a k-mer u put c * (k - 1) distinct positions into the occurrence sets of the
// (at most k - 1 distinct) 2-mers inside u, so the k - 1 largest 2-mer counts sum to at least c * (k - 1).  With s1 >= s2 the two
// largest counts that sum is at most s1 + (k - 2) * s2; a unit needs c >= int(L * p / k) + 1 (utils.nim:259).  A segment with
// s1 + (k - 2) * s2 < (int(L * p / k) + 1) * (k - 1) for every k = 2..6 therefore returns the empty unit whatever the ladder does.
struct FilterCache {
  uint32_t V[5];
  int len, pclass;
  int r[5];   // (int(L * p / k) + 1) * (k - 1), k = 2..6
};
// ALL_WORDS: all eleven words may be read whatever L is (staging buffer; words past the segment only feed masked slots)
// aligned (warp-uniform): every lane's segment starts on a word boundary (sh == 0), the re-alignment shifts are skipped
template <int CSA, bool ALL_WORDS>
__device__ __forceinline__ bool prefilter_keep(const uint32_t *src, uint32_t sh, int L, int pclass, const uint16_t *__restrict__ tfilt,
                                               FilterCache &fc, int one, bool aligned) {
  const int n_words = (2 * L + 31) >> 5;
  uint32_t raw[11], w[10];
#pragma unroll
  for (int j = 0; j < 11; j++) raw[j] = (ALL_WORDS || j <= n_words) ? __byte_perm(src[j], 0, 0x0123) : 0u;
  if (aligned) {
#pragma unroll
    for (int j = 0; j < 10; j++) w[j] = raw[j];
  } else {
#pragma unroll
    for (int j = 0; j < 10; j++) w[j] = __funnelshift_l(raw[j + 1], raw[j], sh);
  }
  if (L != fc.len || pclass != fc.pclass) {
    prefilter_masks(L, fc.V);
    const uint4 t = *reinterpret_cast<const uint4 *>(tfilt + (size_t)(pclass * kThrLen + L) * 8);
    fc.r[0] = (int)(t.x & 0xffffu);
    fc.r[1] = (int)(t.x >> 16);
    fc.r[2] = (int)(t.y & 0xffffu);
    fc.r[3] = (int)(t.y >> 16);
    fc.r[4] = (int)(t.z & 0xffffu);
    fc.len = L;
    fc.pclass = pclass;
  }
  int s1, s2;
  prefilter_top2<CSA>(w, fc.V, s1, s2, one);
  return s1 >= fc.r[0] || s1 + s2 >= fc.r[1] || s1 + 2 * s2 >= fc.r[2] || s1 + 3 * s2 >= fc.r[3] || s1 + 4 * s2 >= fc.r[4];
}

// Appends the kept segments of this warp's group to the survivor list (one atomic per warp and class).  The list has two
// ends: list[0] long segments (>= kLongLen bases, and everything bound for the warp path) filling list[kListHdr ..] upwards,
// list[1] short segments filling list[2 + cap - 1 ..] downwards, so that the scan kernel's batches of 32 hold segments of
// similar length (its per-lane loops run for the longest lane of a warp).
constexpr int kLongLen = 96;
constexpr uint32_t kListHdr = 4;   // list[0] long count, list[1] short count, list[2] next group to hand out (scan kernel), list[3] unused
__device__ __forceinline__ void survivors_push(uint32_t *__restrict__ list, uint32_t cap, bool keep, bool is_short, uint32_t s, int lane) {
  const uint32_t km = __ballot_sync(kFull, keep);
  if (km == 0u) return;
  const uint32_t sm = __ballot_sync(kFull, keep && is_short);
  const uint32_t lm = km & ~sm;
  const uint32_t below = (1u << lane) - 1u;
  uint32_t base_l = 0, base_s = 0;
  if (lane == 0) {
    if (lm) base_l = atomicAdd(list, (uint32_t)__popc(lm));
    if (sm) base_s = atomicAdd(list + 1, (uint32_t)__popc(sm));
  }
  base_l = __shfl_sync(kFull, base_l, 0);
  base_s = __shfl_sync(kFull, base_s, 0);
  if (keep) {
    if (is_short) list[kListHdr + cap - 1u - (base_s + __popc(sm & below))] = s;
    else list[kListHdr + base_l + __popc(lm & below)] = s;
  }
}

// Staging: a group of 32 uniform reads is one contiguous span of 8 * stride bytes.  Lane 0 of the warp that owns the group
// arms an mbarrier and issues ONE bulk copy (TMA, cp.async.bulk) of the span into the warp's shared-memory buffer; the copy
// of the next STAGES - 1 groups are in flight while the current one is counted (a ring of buffers per warp), so no warp waits on an HBM
// load with its registers tied up.  16 bytes past the span are copied too (the re-alignment of the last lane reads one word
// beyond its read), which is why the batch's last group -- and everything that is not a uniform read -- takes the LDG path.
constexpr int kStageBytes = 8 * kShortMaxLen + 32;   // 1312: spans of up to 8 * 160 bytes + 16, kept 16-byte aligned
constexpr int kPreWarps = kPreThreads / 32;

template <int CSA, int STAGES>
__global__ void __launch_bounds__(kPreThreads, 4) repeat_prefilter(const uint32_t *__restrict__ seq, const uint32_t *__restrict__ nmask,
                                                                   const strgpu_segment *__restrict__ segs, uint32_t n_seg,
                                                                   const UniformReads u, const uint16_t *__restrict__ thr,
                                                                   strgpu_repeat *__restrict__ out, uint32_t *__restrict__ list,
                                                                   uint32_t n_tma_groups, int one) {
  __shared__ __align__(128) unsigned char stage_buf[kPreWarps][STAGES][kStageBytes];
  __shared__ __align__(8) uint64_t stage_bar[kPreWarps][STAGES];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint16_t *tfilt = thr + kThrFiltOff;
  const uint32_t warps_total = gridDim.x * kPreWarps;
  const uint32_t warp_global = blockIdx.x * kPreWarps + warp;
  FilterCache fc;
  fc.len = -1;
  fc.pclass = -1;

  // ---- part 1: uniform reads, staged through shared memory by bulk copies
  if (n_tma_groups != 0u) {
    if (lane == 0) {
#pragma unroll
      for (int i = 0; i < STAGES; i++) mbar_init(&stage_bar[warp][i], 1);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // make the initialised barriers visible to the copy engine
    }
    __syncwarp();
    const uint32_t span = 8u * u.stride;                 // bytes of 32 reads
    const uint32_t lane_byte = (uint32_t)lane * (u.stride >> 2);
    const bool stride_aligned = (u.stride & 15u) == 0u;   // 16-base stride: every read starts on a word boundary
    const unsigned char *gsrc = reinterpret_cast<const unsigned char *>(seq);
    const int L = (int)u.read_len;
    const int pclass = (int)u.pclass;
    uint32_t parity = 0;   // bit b = phase of buffer b
    int b = 0;
    uint32_t g = warp_global, g_issue = warp_global;
    // STAGES - 1 groups are in flight while one is being counted
#pragma unroll
    for (int i = 0; i < STAGES - 1; i++) {
      if (g_issue < n_tma_groups && lane == 0) {
        mbar_expect_tx(&stage_bar[warp][i], span + 16u);
        bulk_load(stage_buf[warp][i], gsrc + (size_t)g_issue * span, span + 16u, &stage_bar[warp][i]);
      }
      g_issue += warps_total;
    }
    while (g < n_tma_groups) {
      {
        const int bi = b == 0 ? STAGES - 1 : b - 1;     // the buffer the previous iteration finished reading
        if (g_issue < n_tma_groups && lane == 0) {
          mbar_expect_tx(&stage_bar[warp][bi], span + 16u);
          bulk_load(stage_buf[warp][bi], gsrc + (size_t)g_issue * span, span + 16u, &stage_bar[warp][bi]);
        }
        g_issue += warps_total;
      }
      mbar_wait(&stage_bar[warp][b], (parity >> b) & 1u);
      parity ^= 1u << b;
      const uint32_t s = g * 32u + (uint32_t)lane;
      bool has_n = false;
      if (nmask != nullptr) {
        const uint32_t b0 = s * u.stride, b1 = b0 + u.read_len;
        uint32_t any = 0;
        for (uint32_t wd = b0 >> 5; wd <= ((b1 - 1u) >> 5); wd++) {
          uint32_t v = nmask[wd];
          if (wd == (b0 >> 5)) v &= kFull << (b0 & 31u);
          if (wd == ((b1 - 1u) >> 5) && (b1 & 31u)) v &= (1u << (b1 & 31u)) - 1u;
          any |= v;
        }
        has_n = any != 0u;
      }
      bool keep = has_n;
      if (!has_n) {
        const uint32_t *src = reinterpret_cast<const uint32_t *>(stage_buf[warp][b] + (lane_byte & ~3u));
        keep = prefilter_keep<CSA, true>(src, 8u * (lane_byte & 3u), L, pclass, tfilt, fc, one, stride_aligned);
        if (!keep) reinterpret_cast<unsigned long long *>(out)[s] = 0ull;   // empty unit, repeat_count 0
      }
      survivors_push(list, n_seg, keep, !has_n && L < kLongLen, s, lane);
      __syncwarp();   // every lane has read this buffer before the next iteration refills it
      g += warps_total;
      b = b + 1 == STAGES ? 0 : b + 1;
    }
  }

  // ---- part 2: everything else (descriptor segments, the batch's last uniform group): per-lane loads
  const uint32_t first = n_tma_groups * 32u;
  const uint32_t n_groups = (n_seg - first + 31u) / 32u;
  for (uint32_t grp = warp_global; grp < n_groups; grp += warps_total) {
    const uint32_t s = first + grp * 32u + (uint32_t)lane;
    const bool active = s < n_seg;
    strgpu_segment sg{0, 0, 0, 0};
    if (active) sg = load_segment(segs, nmask, u, s);
    const int L = sg.len;
    const bool lane_path = active && L <= kShortMaxLen && !(sg.flags & STRGPU_SEG_HAS_N);
    bool keep = active && !lane_path;   // non-ACGT bases or > 160 bases: the scan kernel's warp path
    if (lane_path) {
      const int pclass = sg.pclass < STRGPU_MAX_PCLASS ? sg.pclass : STRGPU_MAX_PCLASS - 1;
      keep = prefilter_keep<CSA, false>(seq + (sg.base_off >> 4), 2u * (sg.base_off & 15u), L, pclass, tfilt, fc, one, false);
      if (!keep) reinterpret_cast<unsigned long long *>(out)[s] = 0ull;
    }
    survivors_push(list, n_seg, keep, lane_path && L < kLongLen, s, lane);
  }
}

template <int FILTER, int WARPS>
__global__ void __launch_bounds__(WARPS * 32, WARPS == kLaneWarps ? 1 : 2) repeat_scan_lane(const uint32_t *__restrict__ seq, const uint32_t *__restrict__ nmask,
                                                                    const strgpu_segment *__restrict__ segs, uint32_t n_seg,
                                                                    const UniformReads u,
                                                                    const uint16_t *__restrict__ thr, const uint16_t *__restrict__ luts,
                                                                    strgpu_repeat *__restrict__ out, int *status,
                                                                    uint32_t *__restrict__ list) {
  // list == nullptr: every segment of the batch; else the two-ended survivor list of repeat_prefilter (survivors_push):
  // groups of 32 long segments first, then groups of 32 short ones
  extern __shared__ __align__(16) uint32_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n_long = list ? list[0] : n_seg, n_short = list ? list[1] : 0u;
  const uint32_t groups_long = (n_long + 31) / 32;
  uint16_t *lut = reinterpret_cast<uint16_t *>(smem + WARPS * kWarpSmemWords);
  for (int i = tid; i < kLutTotal; i += WARPS * 32) lut[i] = luts[i];
  __syncthreads();
  uint32_t *warp_base = smem + warp * kWarpSmemWords;
  uint32_t *tab = warp_base + lane;                      // [class][lane] counters; warp scratch for the warp path
  uint32_t *rd = warp_base + kTabWords * 32 + lane;      // [word][lane] read columns
  uint32_t *qmem = warp_base + kTabWords * 32 + kLaneWords * 32;
  uint32_t *q2 = qmem + kQ2Off, *q3 = qmem + kQ3Off, *q4 = qmem + kQ4Off, *q5 = qmem + kQ5Off, *q6 = qmem + kQ6Off, *qw = qmem + kQWOff;
  int n2 = 0, n3 = 0, n4 = 0, n5 = 0, n6 = 0;
  const uint16_t *tg = thr + (size_t)(STRGPU_MAX_PCLASS * 5) * kThrLen;
  const uint16_t *tmin = thr + kThrMinOff;
  const uint32_t n_groups = groups_long + (n_short + 31) / 32;
  const uint32_t warps_total = gridDim.x * WARPS;
  // groups are handed out dynamically when they come from the survivor list (their cost varies a lot: a warp takes the
  // next group whenever its queues run low), statically otherwise
  uint32_t *next_group = list ? list + 2 : nullptr;
  uint32_t grp = blockIdx.x * WARPS + warp;
  if (next_group) {
    if (lane == 0) grp = atomicAdd(next_group, 1u);
    grp = __shfl_sync(kFull, grp, 0);
  }
  uint32_t V[5] = {0u, 0u, 0u, 0u, 0u};
  int v_len = -1;
  // Stages: 1 new segments (pre-filter when fused) -> Q2; 2 counts k = 2, 3 and decides the k = 2 rung; 4, 5, 6 count k = 4, 5, 6;
  // R = every recount a rung k >= 3 needs (read.count(s), utils.nim:254) followed by that rung's decision: lanes that need one are
  // pushed to QR (q3) with their k, M and leader instead of recounting divergently inside the counting stage, so recounts run 32
  // lanes wide.  All queues hold <= 64 entries; a stage pops <= 32 and pushes <= 32 into any queue downstream of it, and runs only
  // when those queues have room for a full batch -- except that a counting stage whose recount queue is full recounts in place
  // (the pre-v7 behaviour), which is what makes the schedule deadlock free.  Downstream stages are served first; new segments are
  // taken only when no queue holds a full batch; at the end the queues are drained upstream-first.
  while (true) {
    const bool more = grp < n_groups;
    const bool r_ok = n4 <= 32 && n5 <= 32 && n6 <= 32;   // stage R pushes into Q4 / Q5 / Q6
    int stage;
    if (n3 >= 32 && r_ok) stage = 3;       // keep the recount queue below a full batch so that counting stages can defer into it
    else if (n6 >= 32) stage = 6;
    else if (n5 >= 32) stage = 5;
    else if (n4 >= 32) stage = 4;
    else if (n2 >= 32) stage = 2;
    else if (more) stage = 1;
    else if (n2 > 0) stage = 2;
    else if (n3 > 0 && r_ok) stage = 3;
    else if (n4 > 0) stage = 4;
    else if (n5 > 0) stage = 5;
    else if (n6 > 0) stage = 6;
    else break;

    if (stage == 1) {
      // ---- stage 1: new segments, 32 per pass: the 2-mer pre-filter finishes most of them; survivors go to Q2
      do {
        const bool in_long = grp < groups_long;
        const uint32_t item = (in_long ? grp : grp - groups_long) * 32 + lane;
        if (next_group) {
          uint32_t g = 0;
          if (lane == 0) g = atomicAdd(next_group, 1u);
          grp = __shfl_sync(kFull, g, 0);
        } else {
          grp += warps_total;
        }
        const bool active = item < (in_long ? n_long : n_short);
        const uint32_t s = (list && active) ? (in_long ? list[kListHdr + item] : list[kListHdr + n_seg - 1u - item]) : item;
        strgpu_segment sg{0, 0, 0, 0};
        if (active) sg = load_segment(segs, nmask, u, s);
        const int L = sg.len;
        const bool lane_path = active && L <= kShortMaxLen && !(sg.flags & STRGPU_SEG_HAS_N);
        bool survive = lane_path;
        if (lane_path && FILTER >= 0) {
          uint32_t w[kLaneWords];
          lane_load(seq, sg, w);
          if (L != v_len) {
            filter_masks(L, V);
            v_len = L;
          }
          const int pclass = sg.pclass < STRGPU_MAX_PCLASS ? sg.pclass : STRGPU_MAX_PCLASS - 1;
          survive = lane_filter_max2<(FILTER > 0 ? 1 : 0)>(w, V) > (int)tmin[pclass * kThrLen + L];
          if (!survive) reinterpret_cast<unsigned long long *>(out)[s] = 0ull;   // empty unit, repeat_count 0
        }
        queue_push<1>(q2, n2, survive, lane, s, ScanState{0, 0u, 0, 0}, 0);
        // segments with non-ACGT bases or longer than 160 bases: one at a time on the whole warp
        const uint32_t wm = __ballot_sync(kFull, active && !lane_path);
        if (wm != 0u) {
          if (active && !lane_path) qw[__popc(wm & ((1u << lane) - 1u))] = s;
          const int nw = __popc(wm);
          WarpScratch<512> &ws = *reinterpret_cast<WarpScratch<512> *>(warp_base);
          for (int i = lane; i < WarpScratch<512>::kTab / 4; i += 32) reinterpret_cast<uint32_t *>(ws.tab)[i] = 0;
          __syncwarp();
          for (int e = 0; e < nw; e++) {
            const uint32_t ws_s = qw[e];
            warp_scan_compact(ws, seq, nmask, load_segment(segs, nmask, u, ws_s), ws_s, thr, lane, 2, ScanState{-1, 0u, 0, 0}, out, status);
          }
          __syncwarp();
        }
      } while (grp < n_groups && n2 < 32);
      continue;
    }

    if (stage == 2) {
      // ---- stage 2: count k = 2 and 3; the k = 2 rung (every lane recounts: best is still -1); the k = 3 rung if it needs no recount
      const int nb = n2 < 32 ? n2 : 32;
      const int first = n2 - nb;
      uint32_t s = 0, extra = 0;
      ScanState st{-1, 0u, 0, 0};
      int next_k = 0;  // 0: finished, 3: needs the k = 3 recount (QR), 4: goes on to k = 4
      if (lane < nb) {
        s = q2[first + lane];
        const strgpu_segment sg = load_segment(segs, nmask, u, s);
        const int L = sg.len;
        lane_stage(seq, sg, rd);
        uint32_t best2, best3;
        lane_count23(rd, tab, lut, L, best2, best3);
        const int pclass = sg.pclass < STRGPU_MAX_PCLASS ? sg.pclass : STRGPU_MAX_PCLASS - 1;
        const uint16_t *tp = thr + (size_t)(pclass * 5) * kThrLen + L;
        const int M2 = (int)(best2 >> 13), M3 = (int)(best3 >> 13);
        const uint32_t lead2 = M2 ? (uint32_t)lut[kRev234 + (best2 & 31u)] : 0xfu;
        const uint32_t lead3 = M3 ? (uint32_t)lut[kRev234 + kCls2 + (best3 & 31u)] : 0x3fu;
        extra = (uint32_t)M3 | (lead3 << 8) | (3u << 20);
        const bool wide = __any_sync(nb == 32 ? kFull : ((1u << nb) - 1u), L > 64);
        bool go = lane_decide(rd, tab, L, 2, M2, lead2, tp[0], tg[L], st, wide);
        if (go) {
          if (3 * M3 > st.best) next_k = 3;                       // needs the k = 3 recount
          else go = !(M3 < (int)tg[kThrLen + L]);
        }
        if (go && next_k == 0) next_k = 4;
        if (!go) emit_result(out, s, st);
      }
      __syncwarp();
      n2 = first;
      queue_push<3>(q3, n3, next_k == 3, lane, s, st, extra);
      queue_push<2>(q4, n4, next_k == 4, lane, s, st, 0);
      continue;
    }

    if (stage == 3) {
      // ---- stage R: recounts of any rung k = 3..6, 32 lanes wide; then the rung's decision (utils.nim:254-265)
      const int nb = n3 < 32 ? n3 : 32;
      const int first = n3 - nb;
      uint32_t s = 0, extra = 0;
      ScanState st{-1, 0u, 0, 0};
      int k = 0;
      if (lane < nb) {
        queue_read<3>(q3, first + lane, s, st, extra);
        k = (int)(extra >> 20);
        const uint32_t leader = (extra >> 8) & 0xfffu;
        const strgpu_segment sg = load_segment(segs, nmask, u, s);
        const int L = sg.len;
        lane_stage(seq, sg, rd);
        const int pclass = sg.pclass < STRGPU_MAX_PCLASS ? sg.pclass : STRGPU_MAX_PCLASS - 1;
        const uint32_t magic = 65536u / (uint32_t)k + 1u;
        const bool wide = __any_sync(nb == 32 ? kFull : ((1u << nb) - 1u), L > 64);
        const int c = wide ? lane_recount<10>(rd, tab, L, leader, k, magic) : lane_recount<4>(rd, tab, L, leader, k, magic);
        const int score = c * k;
        if (score >= st.best) {
          st.best = score;
          if (c > (int)thr[(size_t)(pclass * 5 + k - 2) * kThrLen + L]) {
            st.unit_code = leader;
            st.unit_k = k;
            st.rc = c;
          }
        }
        if (k == 6) emit_result(out, s, st);
      }
      __syncwarp();
      n3 = first;
      queue_push<2>(q4, n4, k == 3, lane, s, st, 0);
      queue_push<2>(q5, n5, k == 4, lane, s, st, 0);
      queue_push<2>(q6, n6, k == 5, lane, s, st, 0);
      continue;
    }

    // ---- counting stages k = 4, 5, 6: 32 segments at a time, one per lane
    const int qn = stage == 4 ? n4 : (stage == 5 ? n5 : n6);
    const int nb = qn < 32 ? qn : 32;
    const int first = qn - nb;
    const bool defer = n3 <= 32;            // room for a full batch in the recount queue; else recount in place
    uint32_t s = 0, extra = 0;
    ScanState st{-1, 0u, 0, 0};
    int what = 0;                           // 0: finished (emitted), 1: next counting stage, 2: recount queue
    if (lane < nb) {
      queue_read<2>(stage == 4 ? q4 : (stage == 5 ? q5 : q6), first + lane, s, st, extra);
      const strgpu_segment sg = load_segment(segs, nmask, u, s);
      const int L = sg.len;
      lane_stage(seq, sg, rd);
      const int pclass = sg.pclass < STRGPU_MAX_PCLASS ? sg.pclass : STRGPU_MAX_PCLASS - 1;
      int M;
      uint32_t leader;
      if (stage == 4) lane_count4(rd, tab, lut, L, M, leader);
      else if (stage == 5) lane_count5(rd, tab, lut, L, M, leader);
      else lane_count6(rd, tab, L, M, leader);
      if (M * stage <= st.best) {           // no recount: break or continue (utils.nim:250-253)
        what = (M < (int)tg[(stage - 2) * kThrLen + L]) ? 0 : 1;
      } else if (defer) {
        what = 2;
        extra = (uint32_t)M | (leader << 8) | ((uint32_t)stage << 20);
      } else {
        lane_decide(rd, tab, L, stage, M, leader, thr[(size_t)(pclass * 5 + stage - 2) * kThrLen + L], tg[(stage - 2) * kThrLen + L], st, L > 64);
        what = 1;
      }
      if (what == 0 || (what == 1 && stage == 6)) {
        emit_result(out, s, st);
        what = 0;
      }
    }
    __syncwarp();
    if (stage == 4) { n4 = first; queue_push<2>(q5, n5, what == 1, lane, s, st, 0); }
    else if (stage == 5) { n5 = first; queue_push<2>(q6, n6, what == 1, lane, s, st, 0); }
    else n6 = first;
    queue_push<3>(q3, n3, what == 2, lane, s, st, extra);
  }
}

}  // namespace

// Class LUTs of the lane kernel (uint16 each, kLutTotal entries):
//   [kLut2, kLut3)         byte offset of the uint32 counter of every 2/3-mer code's min-rotation class in the lane's column,
//   [kLut4)                for 4-mer codes: word byte offset | bit shift of the class's packed uint8 counter,
//   [kRev234)              canonical (minimal) code of each of the 10 + 24 + 70 classes,
//   [kLut5)                for 5-mer codes: (class / 4) * 128 | (class % 4) * 8 (word offset | bit shift of its uint8 counter),
//   [kRev5)                canonical code of each of the 208 5-mer classes.
void build_lane_luts(uint16_t *dst) {
  const int lut_off[6] = {0, 0, kLut2, kLut3, kLut4, kLut5};
  const int rev_off[6] = {0, 0, kRev234, kRev234 + kCls2, kRev234 + kCls2 + kCls3, kRev5};
  const int cls_base[6] = {0, 0, 0, kCls2, kCls2 + kCls3, 0};
  for (int k = 2; k <= 5; k++) {
    const int n = 1 << (2 * k);
    const uint32_t mask = (uint32_t)n - 1u;
    auto canon = [&](uint32_t code) {
      uint32_t m = code, x = code;
      for (int j = 1; j < k; j++) {
        x = ((x << 2) | (x >> (2 * k - 2))) & mask;
        if (x < m) m = x;
      }
      return m;
    };
    int n_classes = 0;
    for (int code = 0; code < n; code++)
      if (canon((uint32_t)code) == (uint32_t)code) dst[rev_off[k] + n_classes++] = (uint16_t)code;  // canonical codes ascend
    for (int code = 0; code < n; code++) {
      const uint16_t m = (uint16_t)canon((uint32_t)code);
      int cls = 0;
      while (dst[rev_off[k] + cls] != m) cls++;
      if (k < 4) dst[lut_off[k] + code] = (uint16_t)((cls_base[k] + cls) * 128);
      else dst[lut_off[k] + code] = (uint16_t)(((cls_base[k] + (cls >> 2)) * 128) | ((cls & 3) * 8));  // packed uint8 counters
    }
  }
}

cudaError_t launch_repeat_scan(const uint32_t *d_seq_words, const uint32_t *d_nmask, const strgpu_segment *d_segs,
                               uint32_t n_seg, uint32_t max_len, const uint16_t *d_thr, const uint16_t *d_luts,
                               strgpu_repeat *d_out, int *d_status, int sm_count, int variant, cudaStream_t stream,
                               const UniformReads *uniform, uint32_t *d_list) {
  if (n_seg == 0) return cudaSuccess;
  UniformReads u{0, 0, 0, 0};
  if (uniform) {
    u = *uniform;
    if (u.read_len > (uint32_t)kShortMaxLen) return cudaErrorInvalidValue;  // callers expand long uniform reads into descriptors
    if (variant == 1) variant = 0;
  }
  if (variant < 0 || variant > 7) variant = 0;
  if (d_list == nullptr && (variant == 0 || variant >= 5)) variant = 2;   // no survivor list: fused kernel
  constexpr int kWarps = 8;
  const uint32_t blocks_needed = (n_seg + kWarps - 1) / kWarps;
  if (max_len <= (uint32_t)kShortMaxLen && variant != 1) {
    // 0: repeat_prefilter + repeat_scan_lane over its survivor list (5..7: the same with 0 / 12 / 16 carry-save streams);
    // 2 / 4: one fused kernel (plain / carry-save popcounts); 3: no pre-filter (A/B runs)
    const bool split = variant == 0 || variant >= 5;
    // (A geometry that lets batch i's scan kernel share the SMs with batch i + 1's pre-filter -- 12-warp scan CTAs next to 2 pre-filter
    // CTAs -- was measured and is slower than letting whole kernels of two streams interleave: 2.22e10 vs 2.54e10 reads/s.)
    auto kernel = variant == 2 ? repeat_scan_lane<0, kLaneWarps> : (variant == 4 ? repeat_scan_lane<1, kLaneWarps> : repeat_scan_lane<-1, kLaneWarps>);
    const int lane_warps = kLaneWarps;
    const int lane_smem = lane_smem_bytes(lane_warps);
    // function attributes are per device: remember which (device, variant) pairs have been configured
    static bool configured[64][8] = {{false}};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) dev = 0;
    if (!configured[dev][variant]) {
      cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, lane_smem);
      if (e != cudaSuccess) return e;
      e = cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      if (e != cudaSuccess) return e;
      configured[dev][variant] = true;
    }
    static const bool no_tma = getenv("STRGPU_NO_TMA") != nullptr;   // A/B: per-lane loads for uniform reads too
    if (split) {
      cudaError_t e = cudaMemsetAsync(d_list, 0, 4 * sizeof(uint32_t), stream);
      if (e != cudaSuccess) return e;
      const uint32_t groups = (n_seg + 31) / 32;
      uint32_t pre_grid = (uint32_t)sm_count * 4u;   // 4 resident CTAs of 8 warps per SM, grid-stride over groups of 32
      const uint32_t pre_need = (groups + kPreThreads / 32 - 1) / (kPreThreads / 32);
      if (pre_grid > pre_need) pre_grid = pre_need;
      static const int stages = getenv("STRGPU_STAGES") ? atoi(getenv("STRGPU_STAGES")) : 4;   // A/B: staging depth
      auto pre = variant == 7 ? (stages == 2 ? repeat_prefilter<16, 2> : repeat_prefilter<16, 4>)
                 : variant == 6 ? repeat_prefilter<12, 4>
                 : variant == 5 ? repeat_prefilter<4, 4>
                                : (stages == 2 ? repeat_prefilter<kPreCsaDefault, 2> : repeat_prefilter<kPreCsaDefault, 4>);
      // uniform reads go through the TMA-staged part when their 32-read spans fit the staging buffers (all groups but the
      // batch's last one: the copy reads 16 bytes past its span)
      uint32_t n_tma = 0;
      if (u.n_reads >= 64u && u.read_len >= 1u && 8u * u.stride + 16u <= (uint32_t)kStageBytes && ((uintptr_t)d_seq_words & 15u) == 0 && !no_tma)
        n_tma = u.n_reads / 32u - 1u;
      pre<<<pre_grid, kPreThreads, 0, stream>>>(d_seq_words, d_nmask, d_segs, n_seg, u, d_thr, d_out, d_list, n_tma, 1);
      e = cudaGetLastError();
      if (e != cudaSuccess) return e;
    }
    const uint32_t tiles = (n_seg + lane_warps * 32 - 1) / (lane_warps * 32);
    uint32_t grid = (uint32_t)sm_count;  // one persistent CTA per SM
    if (grid > tiles) grid = tiles;
    kernel<<<grid, lane_warps * 32, lane_smem, stream>>>(d_seq_words, d_nmask, d_segs, n_seg, u, d_thr, d_luts, d_out, d_status,
                                                          split ? d_list : nullptr);
  } else if (max_len <= (uint32_t)kShortMaxLen) {
    uint32_t grid = (uint32_t)sm_count * 8u;  // 8 resident CTAs of 256 threads per SM
    if (grid > blocks_needed) grid = blocks_needed;
    repeat_scan_warp<kShortMaxLen, kWarps><<<grid, kWarps * 32, 0, stream>>>(d_seq_words, d_nmask, d_segs, n_seg, d_thr,
                                                                              d_out, d_status);
  } else {
    uint32_t grid = (uint32_t)sm_count * 4u;
    if (grid > blocks_needed) grid = blocks_needed;
    repeat_scan_warp<512, kWarps><<<grid, kWarps * 32, 0, stream>>>(d_seq_words, d_nmask, d_segs, n_seg, d_thr, d_out,
                                                                     d_status);
  }
  return cudaGetLastError();
}

}  // namespace strgpu
