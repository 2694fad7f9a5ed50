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

#include <algorithm>
#include <cstdlib>
#include <mutex>

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
                                                  const uint32_t *__restrict__ nmask, const uint32_t *__restrict__ xmask,
                                                  const strgpu_segment sg, uint32_t s,
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
      uint32_t v = 0, x = 0;
      if (l < n_nw) {
        const uint32_t g = (sg.base_off >> 5) + (uint32_t)l;
        v = __funnelshift_r(nmask[g], nmask[g + 1], sg.base_off & 31u);
        const int rem = L - 32 * l;
        if (rem < 32) v &= (1u << rem) - 1u;
        // utils.nim:238 counts the literal 'N' only: bases flagged in xmask (IUPAC codes other than N) never match in the
        // recount but do not count towards the gate
        if (xmask != nullptr) x = __funnelshift_r(xmask[g], xmask[g + 1], sg.base_off & 31u);
      }
      ws.nm[l] = v;
      n_count += __popc(v & ~x);
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
                                                               const uint32_t *__restrict__ xmask,
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
    warp_scan_segment<MAXLEN>(ws, seq, nmask, xmask, segs[s], s, thr, lane, 2, ScanState{-1, 0u, 0, 0}, out, status);
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
constexpr int kLaneWords = 11;                 // ten words hold 160 bases; one more absorbs the re-alignment shift
constexpr int kCls2 = 10, kCls3 = 24, kCls4 = 70, kCls5 = 208;
constexpr int kCls4Words = (kCls4 + 3) / 4;       // 4-mer classes are counted in packed uint8 (18 words) to keep smem small
constexpr int kLut2 = 0, kLut3 = 16, kLut4 = 80, kRev234 = 336, kLut5 = 440, kRev5 = 1464;  // offsets into the uint16 table
constexpr int kLutTotal = 1672;                // the k = 2..5 tables (copied to shared memory by the ladder kernels that use them)
constexpr int kCls6 = 700;                     // necklace classes of 6-mers over 4 letters
constexpr int kLut6 = kLutTotal, kRev6 = kLut6 + 4096;   // class rank of every 6-mer code; canonical code of every rank
constexpr int kLut5R = kRev6 + kCls6;                    // class rank of every 5-mer code (its canonical codes: kRev5)
constexpr int kRankWords = (4096 + 1024) / 2;            // the two rank tables as shared-memory words (6-mers first)
static_assert(kLut5R + 1024 == kLaneLutEntries, "LUT size");
static_assert(kLut5R % 2 == 0 && kLut6 % 2 == 0, "tables are copied as 32-bit words");

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
                                               const uint32_t *__restrict__ nmask, const uint32_t *__restrict__ xmask,
                                               const strgpu_segment sg, uint32_t s,
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
      uint32_t v = 0, x = 0;
      if (l < n_nw) {
        const uint32_t g = (sg.base_off >> 5) + (uint32_t)l;
        v = __funnelshift_r(nmask[g], nmask[g + 1], sg.base_off & 31u);
        const int rem = L - 32 * l;
        if (rem < 32) v &= (1u << rem) - 1u;
        // utils.nim:238 counts the literal 'N' only: bases flagged in xmask (IUPAC codes other than N) never match in the
        // recount but do not count towards the gate
        if (xmask != nullptr) x = __funnelshift_r(xmask[g], xmask[g + 1], sg.base_off & 31u);
      }
      ws.nm[l] = v;
      n_count += __popc(v & ~x);
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

// count(read, K, counts[K]) for K = 5 and 6, one lane, without any table: the at most 32 (K = 5) / 26 (K = 6) windows of a
// <= 160-base segment become keys (class rank << 5 | window index; the class rank of a K-mer code comes from a table in
// shared memory), the keys are SORTED in registers -- two independent 16-key merge-exchange networks run side by side on
// packed 16x2 values (VIMNMX.U16x2), then one bitonic merge across the halves -- and a single sweep over the sorted keys
// finds the longest run of equal classes.  Seq.inc keeps the FIRST class to reach the final maximum (strict >,
// utils.nim:192-195): among the classes with the maximal count that is the one whose LAST window comes first, i.e. the
// maximum of (run length so far, 31 - window index) over all keys.  No shared-memory counters, no probing loop, no divergence.
__device__ __forceinline__ void cmpx(uint32_t &a, uint32_t &b) {
  const uint32_t lo = __vminu2(a, b), hi = __vmaxu2(a, b);
  a = lo;
  b = hi;
}
// K is a RUN-TIME argument (5 or 6) and the function is kept out of line: one copy of the two unrolled networks serves
// both rungs (instruction-cache footprint).  rank_lut: the rung's class-rank table in shared memory; n_classes: 208 / 700.
__device__ __noinline__ uint32_t lane_count_sorted(const uint32_t *rd, const uint16_t *rank_lut, int L, int K, int n_classes) {
  const int W = L / K;
  const uint32_t top = 32u - 2u * (uint32_t)K;
  uint32_t v[16];
#pragma unroll
  for (int i = 0; i < 16; i++) {
    uint32_t e[2];
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int j = i + 16 * h;                                        // window index
      e[h] = (uint32_t)((n_classes + j) * 32 + 31);                    // padding: a class of its own, ranked after every real one
      if (j < W) {
        const uint32_t bit = (uint32_t)(2 * j) * (uint32_t)K;
        const uint32_t *p = rd + (bit >> 5) * 32u;
        const uint32_t code = __funnelshift_l(p[32], p[0], bit & 31u) >> top;
        e[h] = (uint32_t)rank_lut[code] * 32u + (uint32_t)j;
      }
    }
    v[i] = e[1] * 65536u + e[0];
  }
  // merge-exchange sort (Batcher) of 16 keys, both halves at once
#pragma unroll
  for (int p = 1; p < 16; p *= 2)
#pragma unroll
    for (int k = p; k >= 1; k /= 2)
#pragma unroll
      for (int j = k % p; j <= 15 - k; j += 2 * k)
#pragma unroll
        for (int i = 0; i <= (k - 1 < 15 - j - k ? k - 1 : 15 - j - k); i++)
          if ((i + j) / (2 * p) == (i + j + k) / (2 * p)) cmpx(v[i + j], v[i + j + k]);
  // bitonic merge of the two sorted halves: low key i against high key 15 - i, then half-cleaners inside each half
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint32_t a = v[i], t = __byte_perm(v[15 - i], 0, 0x1032);
    const uint32_t mn = __vminu2(a, t), mx = __vmaxu2(a, t);
    v[i] = __byte_perm(mn, mx, 0x7610);
    v[15 - i] = __byte_perm(mn, mx, 0x5432);
  }
#pragma unroll
  for (int st = 8; st >= 1; st /= 2)
#pragma unroll
    for (int i = 0; i < 16; i++)
      if ((i & st) == 0) cmpx(v[i], v[i + st]);
  // sweep: longest run of equal classes, earliest completion first.  best = (run length so far * 32 + 31 - window index) << 16 | key:
  // the window index is unique, so the appended key never decides a comparison
  uint32_t prev = 0xffffffffu, cur = 0, best = 0;
#pragma unroll
  for (int h = 0; h < 2; h++)
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const uint32_t e = h == 0 ? (v[i] & 0xffffu) : __umulhi(v[i], 65536u);
      cur = ((e ^ prev) < 32u) ? cur + 1u : 1u;
      best = max(best, (cur * 32u + ((e & 31u) ^ 31u)) * 65536u + e);
      prev = e;
    }
  return W == 0 ? 0u : best;   // run length: best >> 21, class rank: (best & 0xffff) >> 5
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
constexpr int kPreCsaDefault = 6;

__device__ __forceinline__ void prefilter_masks(int L, uint32_t (&V)[5]) {
#pragma unroll
  for (int j = 0; j < 5; j++) {
    const int v0 = min(max(L - 1 - 16 * j, 0), 16), v1 = min(max(L - 1 - 16 * (j + 5), 0), 16);
    const uint32_t m0 = v0 >= 16 ? kOdd : (kOdd & ~(kFull >> (2 * v0)));
    const uint32_t m1 = v1 >= 16 ? kEven : (kEven & ~(kFull >> (2 * v1)));
    V[j] = m0 | m1;
  }
}

// Upper bounds s1 >= s2 of the two largest of the sixteen 2-mer occurrence counts c[a][b] (positions 0 .. L - 2).
// Only NINE of the sixteen cells are counted (a, b < 3) plus three column sums f[b] = positions 0 .. L - 2 whose SUCCESSOR is b;
// the other seven follow from the margins of the 4 x 4 table:
//   f[3]    = (L - 1) - f[0] - f[1] - f[2]                        c[3][b] = f[b] - c[0][b] - c[1][b] - c[2][b]     (exact)
//   row sum n[a] = occurrences of a at positions 0 .. L - 2 = f[a] + [base 0 == a] - [base L-1 == a], so with
//   n+[a] = f[a] + [base 0 == a] = n[a] + [base L-1 == a]:        c[a][3] <= d[a] = n+[a] - (c[a][0] + c[a][1] + c[a][2])   (over by [base L-1 == a])
//                                                                 c[3][3] <= f[3] - (d[0] + d[1] + d[2]) + 1             (over by 1 - [base L-1 < 3])
// -- the last four are over-estimates by at most one, which keeps the filter sound (a segment it finishes really has the
// empty result; at worst a borderline segment more reaches the ladder kernels, which are exact).  12 popcount streams
// instead of 16 and 45 + 15 + 3 mask LOP3 instead of 60 + 20: this kernel is bound by the ALU / popcount pipes.
// `one` is the runtime constant 1 (a kernel argument): a * one + b compiles to IMAD on the FMA pipe, which is idle here, instead of
// IADD3 on the ALU pipe, which is the bottleneck (every ALU instruction costs two issue cycles).
// The top-2 selection runs on packed 16x2 values (VIMNMX.U16x2 / VIMNMX3: half the min/max instructions).
template <int CSA>
__device__ __forceinline__ int popc5(const uint32_t (&m)[5], int idx, int one) {
  if (idx >= CSA) return (((__popc(m[0]) * one + __popc(m[1])) * one + __popc(m[2])) * one + __popc(m[3])) * one + __popc(m[4]);
  const uint32_t x1 = m[0] ^ m[1] ^ m[2], c1 = (m[0] & m[1]) | (m[2] & (m[0] | m[1]));
  const uint32_t x2 = x1 ^ m[3] ^ m[4], c2 = (x1 & m[3]) | (m[4] & (x1 | m[3]));
  return (__popc(c1) * one + __popc(c2)) * (one + one) + __popc(x2);
}

// (a & mask) | (b & ~mask) as ONE LOP3 (the compiler emits two for the spelled-out form with the constants kOdd / kEven)
__device__ __forceinline__ uint32_t bitselect(uint32_t a, uint32_t b, uint32_t mask) {
  uint32_t d;
  asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(mask));
  return d;
}

template <int CSA>
__device__ __forceinline__ void prefilter_top2(const uint32_t (&w)[10], const uint32_t (&V)[5], int n_valid, int &s1, int &s2, int one) {
  uint32_t Dh[6], Dl[6];
#pragma unroll
  for (int j = 0; j < 5; j++) {
    Dh[j] = bitselect(w[j], __umulhi(w[j + 5], 0x80000000u), kOdd);   // x >> 1 as a high multiply: FMA pipe, not ALU
    Dl[j] = bitselect(w[j] << 1, w[j + 5], kOdd);
  }
  Dh[5] = w[5];        // only its top slot pair is used: base 0 of word 5 (the even bit feeds a slot that is never valid)
  Dl[5] = w[5] << 1;
  uint32_t G[5][3];   // positions whose successor is b (and that start a 2-mer of the segment)
#pragma unroll
  for (int j = 0; j < 5; j++) {
    // (the same shift as IMAD + IMAD.HI on the FMA pipe was measured: 236 us against 223 us -- at 70 % issue utilisation two
    // instructions for one no longer pay)
    const uint32_t Qh = __funnelshift_l(Dh[j + 1], Dh[j], 2), Ql = __funnelshift_l(Dl[j + 1], Dl[j], 2);
    G[j][0] = ~Qh & ~Ql & V[j];
    G[j][1] = ~Qh & Ql & V[j];
    G[j][2] = Qh & ~Ql & V[j];
  }
  int c[4][4], f[4];
#pragma unroll
  for (int b = 0; b < 3; b++) {
    int col = 0;
#pragma unroll
    for (int a = 0; a < 3; a++) {
      uint32_t m[5];
#pragma unroll
      for (int j = 0; j < 5; j++) {
        const uint32_t g = G[j][b];
        m[j] = a == 0 ? (g & ~Dh[j] & ~Dl[j]) : (a == 1 ? (g & ~Dh[j] & Dl[j]) : (g & Dh[j] & ~Dl[j]));
      }
      c[a][b] = popc5<CSA>(m, b * 4 + a, one);
      col = a == 0 ? c[a][b] : col * one + c[a][b];
    }
    uint32_t m[5];
#pragma unroll
    for (int j = 0; j < 5; j++) m[j] = G[j][b];
    f[b] = popc5<CSA>(m, b * 4 + 3, one);
    c[3][b] = col * (-one) + f[b];
  }
  f[3] = ((f[0] * one + f[1]) * one + f[2]) * (-one) + n_valid;
  int rest = f[3];   // becomes the bound of c[3][3]
#pragma unroll
  for (int a = 0; a < 3; a++) {
    const int row = (c[a][0] * one + c[a][1]) * one + c[a][2];
    // [base 0 == a]: position 0 is bit 31 of plane 0
    const uint32_t e0 = a == 0 ? (V[0] & ~Dh[0] & ~Dl[0]) : (a == 1 ? (V[0] & ~Dh[0] & Dl[0]) : (V[0] & Dh[0] & ~Dl[0]));
    const int nplus = (int)__umulhi(e0, 2u) * one + f[a];
    const int d = row * (-one) + nplus;     // = c[a][3] + [base L-1 == a]
    c[a][3] = d;
    rest = d * (-one) + rest;
  }
  c[3][3] = rest * one + one;   // = c[3][3] + 1 - [base L-1 is one of the three counted bases] <= c[3][3] + 1
  // top two of the sixteen: pack cell i with cell i + 8 (all values are in 0 .. 161), select on both halves at once
  const int k16 = one << 16;
  uint32_t P[8];
#pragma unroll
  for (int i = 0; i < 8; i++) P[i] = (uint32_t)(c[2 + (i >> 2)][i & 3] * k16 + c[i >> 2][i & 3]);
  uint32_t hi[4], lo[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    hi[i] = __vmaxu2(P[2 * i], P[2 * i + 1]);
    lo[i] = __vminu2(P[2 * i], P[2 * i + 1]);
  }
  // merge (hi, lo) pairs: first = max of the firsts, second = max(min of the firsts, both seconds)
  const uint32_t a1 = __vmaxu2(hi[0], hi[1]), a2 = __vmaxu2(__vmaxu2(__vminu2(hi[0], hi[1]), lo[0]), lo[1]);
  const uint32_t b1 = __vmaxu2(hi[2], hi[3]), b2 = __vmaxu2(__vmaxu2(__vminu2(hi[2], hi[3]), lo[2]), lo[3]);
  const uint32_t t1 = __vmaxu2(a1, b1), t2 = __vmaxu2(__vmaxu2(__vminu2(a1, b1), a2), b2);
  const int t1l = (int)(t1 & 0xffffu), t1h = (int)__umulhi(t1, 65536u), t2l = (int)(t2 & 0xffffu), t2h = (int)__umulhi(t2, 65536u);
  s1 = max(t1l, t1h);
  s2 = max(max(min(t1l, t1h), t2l), t2h);
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
  int nv;     // valid 2-mer start positions: max(L - 1, 0)
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
    fc.nv = L > 0 ? L - 1 : 0;
  }
  int s1, s2;
  prefilter_top2<CSA>(w, fc.V, fc.nv, s1, s2, one);
  return s1 >= fc.r[0] || s1 + s2 >= fc.r[1] || s1 + 2 * s2 >= fc.r[2] || s1 + 3 * s2 >= fc.r[3] || s1 + 4 * s2 >= fc.r[4];
}

// Device scratch of one library call (uint32 words; sized by scan_scratch_words()):
//   hdr[0] long lane-path survivors, hdr[1] short ones, hdr[2] warp-path segments,
//   hdr[3 + i] entries appended to stage list i (R3, C4), hdr[10 + j] group cursor of ladder kernel j
//   listA[n_seg]  lane-path survivors of the pre-filter, two-ended: long segments (>= kLongLen bases) fill it upwards from 0,
//                 short ones downwards from n_seg - 1, so that the groups of 32 a warp takes hold segments of similar length
//   listW[n_seg]  segments for the warp-per-segment kernel: non-ACGT bases, more than 160 bases, stage-list overflow
//   2 stage lists of `cap` entries each, structure of arrays: segment index, packed ScanState, M | leader << 8
struct StageLists {
  uint32_t *hdr, *listA, *listW, *stage;
  uint32_t cap, n_seg;
};
constexpr int kLongLen = 96;
enum { kListR3 = 0, kListC4, kNumStageLists };
static_assert(kNumStageLists == kScanStageLists, "stage list count");

__device__ __forceinline__ StageLists stage_lists(uint32_t *scratch, uint32_t n_seg, uint32_t cap) {
  StageLists sl;
  sl.hdr = scratch;
  sl.listA = scratch + kScanScratchHdr;
  sl.listW = sl.listA + n_seg;
  sl.stage = sl.listW + n_seg;
  sl.cap = cap;
  sl.n_seg = n_seg;
  return sl;
}

// the pre-filter's survivors (one atomic per warp and list)
__device__ __forceinline__ void survivors_push(const StageLists &sl, bool keep_lane, bool is_short, bool keep_warp, uint32_t s, int lane) {
  const uint32_t km = __ballot_sync(kFull, keep_lane);
  const uint32_t wm = __ballot_sync(kFull, keep_warp);
  if ((km | wm) == 0u) return;
  const uint32_t sm = __ballot_sync(kFull, keep_lane && is_short);
  const uint32_t lm = km & ~sm;
  const uint32_t below = (1u << lane) - 1u;
  uint32_t base_l = 0, base_s = 0, base_w = 0;
  if (lane == 0) {
    if (lm) base_l = atomicAdd(sl.hdr, (uint32_t)__popc(lm));
    if (sm) base_s = atomicAdd(sl.hdr + 1, (uint32_t)__popc(sm));
    if (wm) base_w = atomicAdd(sl.hdr + 2, (uint32_t)__popc(wm));
  }
  base_l = __shfl_sync(kFull, base_l, 0);
  base_s = __shfl_sync(kFull, base_s, 0);
  base_w = __shfl_sync(kFull, base_w, 0);
  if (keep_lane) {
    if (is_short) sl.listA[sl.n_seg - 1u - (base_s + __popc(sm & below))] = s;
    else sl.listA[base_l + __popc(lm & below)] = s;
  }
  if (keep_warp) sl.listW[base_w + __popc(wm & below)] = s;
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
                                                                   strgpu_repeat *__restrict__ out, uint32_t *__restrict__ scratch,
                                                                   uint32_t stage_cap, uint32_t n_tma_groups, int one) {
  __shared__ __align__(128) unsigned char stage_buf[kPreWarps][STAGES][kStageBytes];
  __shared__ __align__(8) uint64_t stage_bar[kPreWarps][STAGES];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint16_t *tfilt = thr + kThrFiltOff;
  const StageLists sl = stage_lists(scratch, n_seg, stage_cap);
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
      survivors_push(sl, keep && !has_n, L < kLongLen, has_n, s, lane);
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
    survivors_push(sl, keep && lane_path, L < kLongLen, active && !lane_path, s, lane);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// The ladder for the pre-filter's survivors (utils.nim:242-265), one LANE per segment, two kernels over DENSE lists:
//   ladder_stage<2>  counts k = 2 and 3 (time-stamped counters in shared memory), decides the k = 2 rung (every lane
//                    recounts: best is still -1) and the k = 3 rung when it needs no recount; appends the segment to R3
//                    (its k = 3 recount read.count(s), utils.nim:254, is pending) or to C4, or stores the result
//   ladder_stage<4>  first the entries of R3 -- the pending recount and the k = 3 decision, 32 lanes wide --, then, for them
//                    and for the entries of C4, the rungs k = 4, 5, 6: k = 4 in packed uint8 counters, k = 5 and 6 by sorting
//                    the windows' class ranks in registers (lane_count_sorted); stores the result
// A group of 32 never mixes R3 and C4 entries, so the k = 3 recount -- the one a third of the repeat-bearing reads need --
// runs without divergence; the recounts of rungs 4..6 are rarer and run in place.  State travels between the two kernels
// as 12-byte list entries in HBM / L2.  A list that overflows its capacity spills the segment to listW: the warp-per-segment
// kernel, which runs last, redoes it from rung 2.
constexpr int kStageThreads = 256;
constexpr int kStageWarps = kStageThreads / 32;
template <int K> struct StageCfg;
template <> struct StageCfg<2> { static constexpr int tab = kCls2 + kCls3, list_r = -1, list_c = -1, out_r = kListR3, out_c = kListC4; };
template <> struct StageCfg<4> { static constexpr int tab = kCls4Words, list_r = kListR3, list_c = kListC4, out_r = -1, out_c = -1; };
template <int K> struct StageSize {
  static constexpr int warp_words = (StageCfg<K>::tab + kLaneWords) * 32;
  static constexpr int lut_words = kLutTotal / 2 + (K == 4 ? kRankWords : 0);   // K = 4 also holds the 5- / 6-mer rank tables
  static constexpr int smem_bytes = kStageWarps * warp_words * 4 + lut_words * 4 + 16;
};

__device__ __forceinline__ uint32_t pack_state(const ScanState &st) {
  // written after the k = 2 rung: 0 <= best <= 160, repeat_count <= 80
  return ((uint32_t)(st.best & 0xff) << 24) | ((uint32_t)(st.rc & 0xff) << 16) | ((uint32_t)st.unit_k << 12) | (st.unit_code & 0xfffu);
}
__device__ __forceinline__ ScanState unpack_state(uint32_t a) {
  ScanState st;
  st.best = (int)(a >> 24);
  st.rc = (int)((a >> 16) & 0xffu);
  st.unit_k = (int)((a >> 12) & 0xfu);
  st.unit_code = a & 0xfffu;
  return st;
}

// Warp-aggregated append of this group's lanes to the recount list `which_r` (lanes with what == 2) and / or the count list
// `which_c` (what == 1); every lane of the warp calls this.  Lane 0 / lane 1 issue the two atomics at the same time, so the
// warp pays one L2 round trip for both lists.
__device__ __forceinline__ void stage_push2(const StageLists &sl, int which_r, int which_c, int what, int lane, uint32_t s, uint32_t st,
                                            uint32_t extra) {
  const uint32_t mr = which_r >= 0 ? __ballot_sync(kFull, what == 2) : 0u;
  const uint32_t mc = which_c >= 0 ? __ballot_sync(kFull, what == 1) : 0u;
  if ((mr | mc) == 0u) return;
  uint32_t base = 0;
  if (lane == 0 && mr) base = atomicAdd(sl.hdr + 3 + which_r, (uint32_t)__popc(mr));
  if (lane == 1 && mc) base = atomicAdd(sl.hdr + 3 + which_c, (uint32_t)__popc(mc));
  const uint32_t base_r = __shfl_sync(kFull, base, 0), base_c = __shfl_sync(kFull, base, 1);
  const uint32_t below = (1u << lane) - 1u;
  if (what == 1 || what == 2) {
    const bool r = what == 2;
    const uint32_t i = r ? base_r + __popc(mr & below) : base_c + __popc(mc & below);
    if (i < sl.cap) {
      uint32_t *q = sl.stage + (size_t)(r ? which_r : which_c) * 3u * sl.cap;
      q[i] = s;
      q[sl.cap + i] = st;
      q[2u * sl.cap + i] = r ? extra : 0u;
    } else {
      sl.listW[atomicAdd(sl.hdr + 2, 1u)] = s;   // no room: the warp kernel redoes this segment from rung 2
    }
  }
}

// one lane's work item of a ladder kernel: a list entry and the segment it names
struct StageItem {
  uint32_t s, st, extra;
  strgpu_segment sg;
  bool active;
};

// the twelve raw words that hold the lane's segment (issued together; nothing waits on them here)
__device__ __forceinline__ void lane_fetch(const uint32_t *__restrict__ seq, const strgpu_segment &sg, bool active, uint32_t (&raw)[kLaneWords + 1]) {
  const uint32_t g = sg.base_off >> 4;
  const int n_words = active ? (2 * (int)sg.len + 31) >> 5 : -1;
#pragma unroll
  for (int j = 0; j < kLaneWords + 1; j++) raw[j] = (j <= n_words) ? seq[g + j] : 0u;
}
// re-align (base 0 at bit 31 of word 0) and store as the lane's column of shared memory
__device__ __forceinline__ void lane_put(const uint32_t (&raw)[kLaneWords + 1], const strgpu_segment &sg, uint32_t *rd) {
  const uint32_t sh = 2u * (sg.base_off & 15u);
  uint32_t be[kLaneWords + 1];
#pragma unroll
  for (int j = 0; j < kLaneWords + 1; j++) be[j] = __byte_perm(raw[j], 0, 0x0123);
#pragma unroll
  for (int j = 0; j < kLaneWords; j++) rd[j * 32] = __funnelshift_l(be[j + 1], be[j], sh);
}

template <int K>
__global__ void __launch_bounds__(kStageThreads, 4)
ladder_stage(const uint32_t *__restrict__ seq, const strgpu_segment *__restrict__ segs, const UniformReads u,
             const uint16_t *__restrict__ thr, const uint16_t *__restrict__ luts, strgpu_repeat *__restrict__ out,
             uint32_t *__restrict__ scratch, uint32_t n_seg, uint32_t stage_cap) {
  using Cfg = StageCfg<K>;
  extern __shared__ __align__(16) uint32_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const StageLists sl = stage_lists(scratch, n_seg, stage_cap);
  uint32_t n_r, n_c;   // entries of the first / second input list
  if (K == 2) {
    n_r = sl.hdr[0];   // long survivors
    n_c = sl.hdr[1];   // short survivors
  } else {
    n_r = min(sl.hdr[3 + (Cfg::list_r >= 0 ? Cfg::list_r : 0)], sl.cap);
    n_c = min(sl.hdr[3 + (Cfg::list_c >= 0 ? Cfg::list_c : 0)], sl.cap);
  }
  const uint32_t groups_r = (n_r + 31u) / 32u, n_groups = groups_r + (n_c + 31u) / 32u;
  if (blockIdx.x * kStageWarps >= n_groups) return;   // nothing for this CTA (the grid is sized for the worst case)
  uint16_t *lut = reinterpret_cast<uint16_t *>(smem + kStageWarps * StageSize<K>::warp_words);
  {
    uint32_t *dst = reinterpret_cast<uint32_t *>(lut);
    for (int i = tid; i < kLutTotal / 2; i += kStageThreads) dst[i] = reinterpret_cast<const uint32_t *>(luts)[i];
    if (K == 4) {   // rank tables: 6-mers, then 5-mers
      for (int i = tid; i < 4096 / 2; i += kStageThreads) dst[kLutTotal / 2 + i] = reinterpret_cast<const uint32_t *>(luts + kLut6)[i];
      for (int i = tid; i < 1024 / 2; i += kStageThreads) dst[kLutTotal / 2 + 2048 + i] = reinterpret_cast<const uint32_t *>(luts + kLut5R)[i];
    }
    __syncthreads();
  }
  const uint16_t *rank6 = lut + kLutTotal, *rank5 = rank6 + 4096;
  uint32_t *warp_base = smem + warp * StageSize<K>::warp_words;
  uint32_t *scr = warp_base + lane;                                   // [word][lane]: recount scratch, counter column
  uint32_t *tab = K == 4 ? scr - (kCls2 + kCls3) * 32 : scr;          // the 4-mer LUT addresses words 34..51 of a full column
  uint32_t *rd = warp_base + Cfg::tab * 32 + lane;                   // [word][lane] read column
  const uint16_t *tg = thr + (size_t)(STRGPU_MAX_PCLASS * 5) * kThrLen;
  const uint32_t *q_r = Cfg::list_r >= 0 ? sl.stage + (size_t)(Cfg::list_r >= 0 ? Cfg::list_r : 0) * 3u * sl.cap : nullptr;
  const uint32_t *q_c = Cfg::list_c >= 0 ? sl.stage + (size_t)(Cfg::list_c >= 0 ? Cfg::list_c : 0) * 3u * sl.cap : nullptr;

  // groups of 32 entries are dealt round-robin over the grid's warps (no atomic, so the next group is known in advance and
  // its loads are issued early: its list entry + descriptor while the current group is being counted, its read words while
  // the current group's list appends wait for their atomics)
  auto load_item = [&](uint32_t grp) {
    StageItem it;
    it.s = 0; it.st = 0; it.extra = 0; it.sg = strgpu_segment{0, 0, 0, 0};
    const bool first_list = grp < groups_r;
    const uint32_t item = (first_list ? grp : grp - groups_r) * 32u + (uint32_t)lane;
    it.active = grp < n_groups && item < (first_list ? n_r : n_c);
    if (it.active) {
      if (K == 2) {
        it.s = first_list ? sl.listA[item] : sl.listA[n_seg - 1u - item];
      } else {
        const uint32_t *q = first_list ? q_r : q_c;
        it.s = q[item];
        it.st = q[sl.cap + item];
        it.extra = q[2u * sl.cap + item];
      }
      it.sg = load_segment(segs, nullptr, u, it.s);    // lane-path segments hold no non-ACGT base
    }
    return it;
  };
  const uint32_t warps_total = gridDim.x * kStageWarps;
  uint32_t grp = blockIdx.x * kStageWarps + warp;
  if (grp >= n_groups) return;
  StageItem cur = load_item(grp);
  uint32_t raw[kLaneWords + 1];
  lane_fetch(seq, cur.sg, cur.active, raw);
  while (true) {
    const bool first_list = grp < groups_r;                            // warp-uniform
    const bool active = cur.active;
    const uint32_t act_mask = __ballot_sync(kFull, active);
    const uint32_t s = cur.s;
    uint32_t extra = cur.extra;
    ScanState st = K == 2 ? ScanState{-1, 0u, 0, 0} : unpack_state(cur.st);
    int what = 0;   // 0: finished, 1: next count list, 2: recount list
    __syncwarp();
    if (active) lane_put(raw, cur.sg, rd);
    const uint32_t next = grp + warps_total;
    StageItem nxt = load_item(next);                                   // in flight while this group is counted
    if (active) {
      const strgpu_segment sg = cur.sg;
      const int L = sg.len;
      const int pclass = sg.pclass < STRGPU_MAX_PCLASS ? sg.pclass : STRGPU_MAX_PCLASS - 1;
      const uint16_t *tp = thr + (size_t)(pclass * 5) * kThrLen + L;
      const bool wide = __any_sync(act_mask, L > 64);
      if (K == 2) {
        uint32_t best2, best3;
        lane_count23(rd, tab, lut, L, best2, best3);
        const int M2 = (int)(best2 >> 13), M3 = (int)(best3 >> 13);
        const uint32_t lead2 = M2 ? (uint32_t)lut[kRev234 + (best2 & 31u)] : 0xfu;
        const uint32_t lead3 = M3 ? (uint32_t)lut[kRev234 + kCls2 + (best3 & 31u)] : 0x3fu;
        bool go = lane_decide(rd, scr, L, 2, M2, lead2, tp[0], tg[L], st, wide);
        if (go) {
          if (3 * M3 > st.best) {                        // the k = 3 rung needs its recount
            what = 2;
            extra = (uint32_t)M3 | (lead3 << 8);
          } else {
            what = (M3 < (int)tg[kThrLen + L]) ? 0 : 1;   // break / continue (utils.nim:250-253)
          }
        }
      } else {
        if (first_list) {
          // the recount the k = 3 rung was waiting for (utils.nim:254-262)
          const uint32_t leader = (extra >> 8) & 0xfffu;
          const int c = wide ? lane_recount<10>(rd, scr, L, leader, 3, 65536u / 3u + 1u) : lane_recount<4>(rd, scr, L, leader, 3, 65536u / 3u + 1u);
          const int score = c * 3;
          if (score >= st.best) {
            st.best = score;
            if (c > (int)tp[kThrLen]) {
              st.unit_code = leader;
              st.unit_k = 3;
              st.rc = c;
            }
          }
        }
        // rungs k = 4, 5, 6 (utils.nim:246-265); a false return is the ladder's `break`
        int M;
        uint32_t leader;
        lane_count4(rd, tab, lut, L, M, leader);
        bool go = lane_decide(rd, scr, L, 4, M, leader, tp[2 * kThrLen], tg[2 * kThrLen + L], st, wide);
#pragma unroll 1
        for (int k = 5; k <= 6 && go; k++) {
          const uint32_t b = lane_count_sorted(rd, k == 5 ? rank5 : rank6, L, k, k == 5 ? kCls5 : kCls6);
          M = (int)(b >> 21);
          leader = M ? (uint32_t)luts[(k == 5 ? kRev5 : kRev6) + ((b & 0xffffu) >> 5)] : (1u << (2 * k)) - 1u;   // no window: all ones
          go = lane_decide(rd, scr, L, k, M, leader, tp[(k - 2) * kThrLen], tg[(k - 2) * kThrLen + L], st, wide);
        }
        what = 0;
      }
      if (what == 0) emit_result(out, s, st);
    }
    __syncwarp();
    if (next < n_groups) lane_fetch(seq, nxt.sg, nxt.active, raw);     // in flight while the appends wait for their atomics
    if (Cfg::out_r >= 0 || Cfg::out_c >= 0) stage_push2(sl, Cfg::out_r, Cfg::out_c, what, lane, s, pack_state(st), extra);
    if (next >= n_groups) break;
    cur = nxt;
    grp = next;
  }
}

// the segments of listW (non-ACGT bases, more than 160 bases, stage-list overflow): one warp per segment, from rung 2
__global__ void __launch_bounds__(kStageThreads) ladder_warp_list(const uint32_t *__restrict__ seq, const uint32_t *__restrict__ nmask,
                                                                  const uint32_t *__restrict__ xmask,
                                                                  const strgpu_segment *__restrict__ segs, const UniformReads u,
                                                                  const uint16_t *__restrict__ thr, strgpu_repeat *__restrict__ out,
                                                                  int *status, uint32_t *__restrict__ scratch, uint32_t n_seg,
                                                                  uint32_t stage_cap) {
  __shared__ WarpScratch<512> scratch_w[kStageWarps];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const StageLists sl = stage_lists(scratch, n_seg, stage_cap);
  const uint32_t n = min(sl.hdr[2], n_seg);
  if (blockIdx.x * kStageWarps >= n) return;
  WarpScratch<512> &ws = scratch_w[warp];
  for (int i = lane; i < WarpScratch<512>::kTab / 4; i += 32) reinterpret_cast<uint32_t *>(ws.tab)[i] = 0;
  __syncwarp();
  uint32_t *cursor = sl.hdr + 10 + 5;
  while (true) {
    uint32_t i = 0;
    if (lane == 0) i = atomicAdd(cursor, 1u);
    i = __shfl_sync(kFull, i, 0);
    if (i >= n) break;
    const uint32_t s = sl.listW[i];
    warp_scan_compact(ws, seq, nmask, xmask, load_segment(segs, nmask, u, s), s, thr, lane, 2, ScanState{-1, 0u, 0, 0}, out, status);
  }
}

}  // namespace

// Class LUTs of the lane kernels (uint16 each, kLutTotal entries):
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
  // k = 6: rank of every code's min-rotation class (canonical codes ascend with the rank) and the canonical code of every rank
  {
    auto canon6 = [](uint32_t code) {
      uint32_t m = code, x = code;
      for (int j = 1; j < 6; j++) {
        x = ((x << 2) | (x >> 10)) & 0xfffu;
        if (x < m) m = x;
      }
      return m;
    };
    int n_classes = 0;
    uint16_t rank_of[4096];
    for (uint32_t code = 0; code < 4096; code++)
      if (canon6(code) == code) {
        rank_of[code] = (uint16_t)n_classes;
        dst[kRev6 + n_classes++] = (uint16_t)code;
      }
    for (uint32_t code = 0; code < 4096; code++) dst[kLut6 + code] = rank_of[canon6(code)];
  }
  // k = 5: class rank of every code, recovered from the packed-counter address table (class = word * 4 + byte)
  for (int code = 0; code < 1024; code++) {
    const uint16_t e = dst[kLut5 + code];
    dst[kLut5R + code] = (uint16_t)((e >> 7) * 4 + ((e & 31) >> 3));
  }
}

uint32_t scan_stage_cap(uint32_t n_seg) {
  uint64_t cap = (uint64_t)n_seg / 8u;
  if (cap < 4096u) cap = 4096u;
  return (uint32_t)((cap + 31u) & ~(uint64_t)31u);
}

size_t scan_scratch_words(uint32_t n_seg) {
  return (size_t)kScanScratchHdr + 2u * (size_t)n_seg + (size_t)kScanStageLists * 3u * scan_stage_cap(n_seg) + 4u;
}

namespace {
std::once_flag g_attr_once[64];
cudaError_t g_attr_err[64];

template <typename Kern>
cudaError_t set_smem(Kern kern, int bytes) {
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
}

// function attributes are per device: set once per device, whichever host thread comes first
cudaError_t configure_device(int dev) {
  if (dev < 0 || dev >= 64) dev = 0;
  std::call_once(g_attr_once[dev], [dev]() {
    cudaError_t e = set_smem(ladder_stage<2>, StageSize<2>::smem_bytes);
    if (e == cudaSuccess) e = set_smem(ladder_stage<4>, StageSize<4>::smem_bytes);
    g_attr_err[dev] = e;
  });
  return g_attr_err[dev];
}
}  // namespace

int scan_launches(uint32_t max_len, int variant) { return (max_len <= (uint32_t)kShortMaxLen && variant != 1) ? 4 : 1; }

cudaError_t launch_repeat_scan(const uint32_t *d_seq_words, const uint32_t *d_nmask, const uint32_t *d_xmask,
                               const strgpu_segment *d_segs, uint32_t n_seg, uint32_t max_len, const uint16_t *d_thr,
                               const uint16_t *d_luts, strgpu_repeat *d_out, int *d_status, int sm_count, int variant,
                               cudaStream_t stream, const UniformReads *uniform, uint32_t *d_scratch) {
  if (n_seg == 0) return cudaSuccess;
  UniformReads u{0, 0, 0, 0};
  if (uniform) {
    u = *uniform;
    if (u.read_len > (uint32_t)kShortMaxLen) return cudaErrorInvalidValue;  // callers expand long uniform reads into descriptors
    if (variant == 1) variant = 0;
  }
  constexpr int kWarps = 8;
  const uint32_t blocks_needed = (n_seg + kWarps - 1) / kWarps;
  if (max_len <= (uint32_t)kShortMaxLen && variant != 1 && d_scratch != nullptr) {
    // variants (A/B runs and tests): 0 default; 5 / 7: 0 / 12 of the 12 popcount streams through carry-save adders;
    // 8: stage lists of 64 entries (forces the overflow path)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) dev = 0;
    cudaError_t e = configure_device(dev);
    if (e != cudaSuccess) return e;
    const uint32_t cap = variant == 8 ? 64u : scan_stage_cap(n_seg);
    e = cudaMemsetAsync(d_scratch, 0, kScanScratchHdr * sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    static const bool no_tma = getenv("STRGPU_NO_TMA") != nullptr;   // A/B: per-lane loads for uniform reads too
    static const int stages = getenv("STRGPU_STAGES") ? atoi(getenv("STRGPU_STAGES")) : 4;   // A/B: staging depth
    const uint32_t groups = (n_seg + 31) / 32;
    // resident pre-filter CTAs per SM (8 warps, 64 registers each: 4 fill the register file).  With 3 the ladder kernels of the
    // previous call (other stream) find room next to them and run in the issue slots the pre-filter leaves idle.
    static const int pre_ctas = getenv("STRGPU_PRE_CTAS") ? std::max(1, std::min(4, atoi(getenv("STRGPU_PRE_CTAS")))) : 4;
    uint32_t pre_grid = (uint32_t)sm_count * (uint32_t)pre_ctas;   // grid-stride over groups of 32
    const uint32_t pre_need = (groups + kPreWarps - 1) / kPreWarps;
    if (pre_grid > pre_need) pre_grid = pre_need;
    auto pre = variant == 7 ? (stages == 2 ? repeat_prefilter<12, 2> : repeat_prefilter<12, 4>)
               : variant == 5 ? repeat_prefilter<0, 4>
                              : (stages == 2 ? repeat_prefilter<kPreCsaDefault, 2> : repeat_prefilter<kPreCsaDefault, 4>);
    // uniform reads go through the TMA-staged part when their 32-read spans fit the staging buffers (all groups but the
    // batch's last one: the copy reads 16 bytes past its span)
    uint32_t n_tma = 0;
    if (u.n_reads >= 64u && u.read_len >= 1u && 8u * u.stride + 16u <= (uint32_t)kStageBytes && ((uintptr_t)d_seq_words & 15u) == 0 && !no_tma)
      n_tma = u.n_reads / 32u - 1u;
    pre<<<pre_grid, kPreThreads, 0, stream>>>(d_seq_words, d_nmask, d_segs, n_seg, u, d_thr, d_out, d_scratch, cap, n_tma, 1);
    if ((e = cudaGetLastError()) != cudaSuccess) return e;
    static const int max_stage = getenv("STRGPU_MAX_STAGE") ? atoi(getenv("STRGPU_MAX_STAGE")) : 99;   // profiling only: stop early
    if (max_stage < 2) return cudaSuccess;
    // the ladder kernels: grids sized for the worst case (every segment survives), CTAs without work exit at once
    const uint32_t stage_need = (groups + kStageWarps - 1) / kStageWarps + 1u;
    auto grid_for = [&](int ctas_per_sm) { return std::min((uint32_t)(sm_count * ctas_per_sm), stage_need); };
    ladder_stage<2><<<grid_for(4), kStageThreads, StageSize<2>::smem_bytes, stream>>>(d_seq_words, d_segs, u, d_thr, d_luts, d_out, d_scratch, n_seg, cap);
    if (max_stage < 4) return cudaGetLastError();
    ladder_stage<4><<<grid_for(4), kStageThreads, StageSize<4>::smem_bytes, stream>>>(d_seq_words, d_segs, u, d_thr, d_luts, d_out, d_scratch, n_seg, cap);
    ladder_warp_list<<<grid_for(4), kStageThreads, 0, stream>>>(d_seq_words, d_nmask, d_xmask, d_segs, u, d_thr, d_out, d_status, d_scratch, n_seg, cap);
  } else if (max_len <= (uint32_t)kShortMaxLen) {
    uint32_t grid = (uint32_t)sm_count * 8u;  // 8 resident CTAs of 256 threads per SM
    if (grid > blocks_needed) grid = blocks_needed;
    repeat_scan_warp<kShortMaxLen, kWarps><<<grid, kWarps * 32, 0, stream>>>(d_seq_words, d_nmask, d_xmask, d_segs, n_seg, d_thr,
                                                                              d_out, d_status);
  } else {
    uint32_t grid = (uint32_t)sm_count * 4u;
    if (grid > blocks_needed) grid = blocks_needed;
    repeat_scan_warp<512, kWarps><<<grid, kWarps * 32, 0, stream>>>(d_seq_words, d_nmask, d_xmask, d_segs, n_seg, d_thr, d_out,
                                                                     d_status);
  }
  return cudaGetLastError();
}

}  // namespace strgpu
