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
// K1 v2: one LANE per segment (segments of <= 160 bases without non-ACGT bases).
//   * each thread keeps its read as a private column of shared memory words ([word][thread]: bank == lane, conflict free),
//   * k = 2, 3, 4: min-rotation class of every window through a small shared LUT, counted in a private column of
//     uint32 counters ([class][thread], conflict free) with the running leader updated exactly like Seq.inc (strict >),
//   * recount: bit-parallel pattern match over the ten words; popcount when no two matches can overlap, otherwise a
//     run-wise greedy walk,
//   * segments that survive to k = 5, 6 (a few per cent), segments with N, and nothing else, are compacted into a
//     per-CTA queue and finished by the warp-per-segment code above on the same CTA.
// ================================================================================================
constexpr int kLaneThreads = 128;
constexpr int kLaneWords = 11;         // ten words hold 160 bases; one more absorbs the re-alignment shift
constexpr int kLaneClasses = 70;       // min-rotation classes of 4-mers (24 for 3-mers, 10 for 2-mers)
constexpr int kLutEntries = 16 + 64 + 256;       // class byte-offset LUTs for k = 2, 3, 4
constexpr int kRevEntries = 10 + 24 + 70;        // class -> canonical code
constexpr int kLaneSmemBytes = (kLaneThreads / 32) * (kLaneClasses * 32 + kLaneWords * 32) * 4 + (kLutEntries + kRevEntries) * 2 + 16;

template <int K> struct LaneK;
template <> struct LaneK<2> { static constexpr int wpw = 8, bits = 32, classes = 10, lut = 0, rev = 0; };
template <> struct LaneK<3> { static constexpr int wpw = 5, bits = 30, classes = 24, lut = 16, rev = 10; };
template <> struct LaneK<4> { static constexpr int wpw = 4, bits = 32, classes = 70, lut = 80, rev = 34; };

// word `wi` of the k-specific window stream: k = 2, 4 use the aligned words as they are (8 / 4 windows each);
// k = 3 re-cuts the bit stream into 30-bit pieces (5 windows each)
template <int K>
__device__ __forceinline__ uint32_t lane_word(const uint32_t *rd, int wi) {
  if (K == 3) {
    const int bit = 30 * wi;
    const int a = bit >> 5;
    return __funnelshift_l(rd[(a + 1) * 32], rd[a * 32], bit & 31) >> 2;
  }
  return rd[wi * 32];
}

template <int K>
__device__ __forceinline__ void lane_count(const uint32_t *rd, uint32_t *tab, const uint16_t *lut, const uint16_t *rev, int L,
                                           int &M, uint32_t &leader) {
  using P = LaneK<K>;
  constexpr uint32_t kMask = (1u << (2 * K)) - 1u;
  const int W = L / K;
  const int nfull = W / P::wpw;
  const int rem = W - nfull * P::wpw;
#pragma unroll
  for (int c = 0; c < P::classes; c++) tab[c * 32] = 0;
  M = 0;
  uint32_t lead_off = 0xffffffffu;
  const uint16_t *l = lut + P::lut;
  for (int wi = 0; wi < nfull; wi++) {
    const uint32_t x = lane_word<K>(rd, wi);
#pragma unroll
    for (int t = 0; t < P::wpw; t++) {
      const uint32_t code = (x >> (P::bits - 2 * K * (t + 1))) & kMask;
      const uint32_t off = l[code];
      uint32_t *slot = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(tab) + off);
      const int cnt = (int)*slot + 1;
      *slot = (uint32_t)cnt;
      if (cnt > M) { M = cnt; lead_off = off; }   // Seq.inc: strict >, earlier leader keeps ties (utils.nim:192-195)
    }
  }
  if (rem > 0) {
    const uint32_t x = lane_word<K>(rd, nfull);
    for (int t = 0; t < rem; t++) {
      const uint32_t code = (x >> (P::bits - 2 * K * (t + 1))) & kMask;
      const uint32_t off = l[code];
      uint32_t *slot = reinterpret_cast<uint32_t *>(reinterpret_cast<char *>(tab) + off);
      const int cnt = (int)*slot + 1;
      *slot = (uint32_t)cnt;
      if (cnt > M) { M = cnt; lead_off = off; }
    }
  }
  leader = (lead_off == 0xffffffffu) ? kMask : (uint32_t)rev[P::rev + lead_off / 128u];
}

// read.count(s) for one lane: greedy leftmost non-overlapping matches of the K-base pattern (utils.nim:254).
// One copy for all k (kept out of line: instruction-cache footprint matters more than the call).
__device__ __noinline__ int lane_recount(const uint32_t *rd, int L, uint32_t pat, int K) {
  const int npos = L - K + 1;
  if (npos <= 0) return 0;
  constexpr uint32_t kLow = 0x55555555u;
  uint32_t w[11], m[10];
#pragma unroll
  for (int i = 0; i < 10; i++) w[i] = rd[i * 32];
  w[10] = 0;
#pragma unroll
  for (int i = 0; i < 10; i++) {  // keep positions < npos (position p of word i sits at bit 30 - 2p)
    const int n = npos - 16 * i;
    m[i] = n >= 16 ? kLow : (n <= 0 ? 0u : (kLow & ~((1u << (32 - 2 * n)) - 1u)));
  }
#pragma unroll 1
  for (int j = 0; j < K; j++) {
    const uint32_t rep = ((pat >> (2 * (K - 1 - j))) & 3u) * kLow;
    uint32_t e[11];
#pragma unroll
    for (int i = 0; i < 11; i++) {
      const uint32_t t = w[i] ^ rep;
      e[i] = ~(t | (t >> 1)) & kLow;          // slot LSB set <=> that base equals pattern base j
    }
#pragma unroll
    for (int i = 0; i < 10; i++) m[i] &= __funnelshift_l(e[i + 1], e[i], 2 * j);
  }
  uint32_t conflict = 0;
#pragma unroll 1
  for (int d = 1; d < K; d++) {
#pragma unroll
    for (int i = 0; i < 10; i++) conflict |= m[i] & __funnelshift_l(i < 9 ? m[i + 1] : 0u, m[i], 2 * d);
  }
  int c = 0;
  if (conflict == 0) {  // no two matches closer than K: every match counts
#pragma unroll
    for (int i = 0; i < 10; i++) c += __popc(m[i]);
    return c;
  }
  int next = 0;  // first position the greedy walk may use
#pragma unroll
  for (int i = 0; i < 10; i++) {
    uint32_t mm = m[i];
    while (mm) {
      const int hb = 31 - __clz(mm);                       // earliest remaining match of this word
      const uint32_t gap = ~mm & kLow & ((1u << hb) - 1u);  // first non-match slot after it
      const int hb2 = gap ? 31 - __clz(gap) : -2;
      const int p = 16 * i + ((30 - hb) >> 1);
      const int e = 16 * i + ((30 - hb2) >> 1);            // run of consecutive matches [p, e)
      const int s0 = p > next ? p : next;
      if (s0 < e) {
        const int n = (e - s0 + K - 1) / K;
        c += n;
        next = s0 + n * K;
      }
      mm = gap ? (mm & ((1u << hb2) - 1u)) : 0u;
    }
  }
  return c;
}

template <int K>
__device__ __forceinline__ bool lane_step(const uint32_t *rd, uint32_t *tab, const uint16_t *lut, const uint16_t *rev, int L,
                                          int thr_p, int thr_giveup, ScanState &st) {
  int M;
  uint32_t leader;
  lane_count<K>(rd, tab, lut, rev, L, M, leader);
  int score = M * K;
  if (score <= st.best) return !(M < thr_giveup);
  const int c = lane_recount(rd, L, leader, K);
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

constexpr int kLaneWarps = kLaneThreads / 32;
constexpr int kWarpTabWords = kLaneClasses * 32;
constexpr int kWarpRdWords = kLaneWords * 32;
static_assert(sizeof(WarpScratch<512>) <= (size_t)kWarpTabWords * 4, "warp scratch must fit in the warp's counter region");

// Every warp works on its own groups of 32 segments with its own shared-memory region: no block-level barrier.
__global__ void __launch_bounds__(kLaneThreads) repeat_scan_lane(const uint32_t *__restrict__ seq, const uint32_t *__restrict__ nmask,
                                                                 const strgpu_segment *__restrict__ segs, uint32_t n_seg,
                                                                 const uint16_t *__restrict__ thr, const uint16_t *__restrict__ luts,
                                                                 strgpu_repeat *__restrict__ out, int *status) {
  extern __shared__ __align__(16) uint32_t smem[];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  uint16_t *lut = reinterpret_cast<uint16_t *>(smem + kLaneWarps * (kWarpTabWords + kWarpRdWords));
  uint16_t *rev = lut + kLutEntries;
  for (int i = tid; i < kLutEntries + kRevEntries; i += kLaneThreads) lut[i] = luts[i];
  __syncthreads();
  uint32_t *tab_warp = smem + warp * (kWarpTabWords + kWarpRdWords);   // [class][lane] counters; warp scratch for hand-offs
  uint32_t *tab = tab_warp + lane;
  uint32_t *rd = tab_warp + kWarpTabWords + lane;                        // [word][lane] read columns
  const uint32_t n_groups = (n_seg + 31) / 32;
  const uint32_t warps_total = gridDim.x * kLaneWarps;
  for (uint32_t grp = blockIdx.x * kLaneWarps + warp; grp < n_groups; grp += warps_total) {
    const uint32_t s = grp * 32 + lane;
    const bool active = s < n_seg;
    strgpu_segment sg{0, 0, 0, 0};
    if (active) sg = segs[s];
    const int L = sg.len;
    const bool lane_path = active && L <= kShortMaxLen && !(sg.flags & STRGPU_SEG_HAS_N);
    ScanState st{-1, 0u, 0, 0};
    int handoff_k = (active && !lane_path) ? 2 : 0;       // 0: finished here
    __syncwarp();
    if (lane_path) {
      // ---- stage: eleven words, re-aligned so that base 0 sits at bit 31 of word 0
      const uint32_t g = sg.base_off >> 4;
      const uint32_t sh = 2u * (sg.base_off & 15u);
      const int n_words = (2 * L + 31) >> 5;
      uint32_t raw[kLaneWords + 1];
#pragma unroll
      for (int j = 0; j < kLaneWords + 1; j++) raw[j] = (j <= n_words) ? __byte_perm(seq[g + j], 0, 0x0123) : 0u;
#pragma unroll
      for (int j = 0; j < kLaneWords; j++) rd[j * 32] = __funnelshift_l(raw[j + 1], raw[j], sh);
      const int pclass = sg.pclass < STRGPU_MAX_PCLASS ? sg.pclass : STRGPU_MAX_PCLASS - 1;
      const uint16_t *tp = thr + (size_t)(pclass * 5) * kThrLen + L;
      const uint16_t *tg = thr + (size_t)(STRGPU_MAX_PCLASS * 5) * kThrLen + L;
      bool go = lane_step<2>(rd, tab, lut, rev, L, tp[0], tg[0], st);
      if (go) go = lane_step<3>(rd, tab, lut, rev, L, tp[kThrLen], tg[kThrLen], st);
      if (go) go = lane_step<4>(rd, tab, lut, rev, L, tp[2 * kThrLen], tg[2 * kThrLen], st);
      if (go) handoff_k = 5;
      else emit_result(out, s, st);
    }
    __syncwarp();
    uint32_t pending = __ballot_sync(kFull, handoff_k != 0);
    if (pending) {  // warp-uniform: finish the handed-off segments one at a time on the whole warp
      WarpScratch<512> &ws = *reinterpret_cast<WarpScratch<512> *>(tab_warp);
      for (int i = lane; i < WarpScratch<512>::kTab / 4; i += 32) reinterpret_cast<uint32_t *>(ws.tab)[i] = 0;
      __syncwarp();
      while (pending) {
        const int src = __ffs(pending) - 1;
        pending &= pending - 1;
        ScanState qst;
        qst.best = __shfl_sync(kFull, st.best, src);
        qst.unit_code = __shfl_sync(kFull, st.unit_code, src);
        qst.unit_k = __shfl_sync(kFull, st.unit_k, src);
        qst.rc = __shfl_sync(kFull, st.rc, src);
        const int start_k = __shfl_sync(kFull, handoff_k, src);
        strgpu_segment qsg;
        qsg.base_off = __shfl_sync(kFull, sg.base_off, src);
        const uint32_t packed = __shfl_sync(kFull, (uint32_t)sg.len | ((uint32_t)sg.pclass << 16) | ((uint32_t)sg.flags << 24), src);
        qsg.len = (uint16_t)(packed & 0xffffu);
        qsg.pclass = (uint8_t)((packed >> 16) & 0xffu);
        qsg.flags = (uint8_t)(packed >> 24);
        warp_scan_compact(ws, seq, nmask, qsg, grp * 32 + (uint32_t)src, thr, lane, start_k, qst, out, status);
      }
    }
  }
}

}  // namespace

// class LUTs of the lane kernel: for k = 2, 3, 4 the byte offset (class * 4 * kLaneThreads) of every window code's
// min-rotation class, then the canonical (minimal) code of every class
void build_lane_luts(uint16_t *dst) {
  int lut_off = 0, rev_off = kLutEntries;
  for (int k = 2; k <= 4; k++) {
    const int n = 1 << (2 * k);
    const uint32_t mask = (uint32_t)n - 1u;
    int n_classes = 0;
    for (int code = 0; code < n; code++) {
      uint32_t m = (uint32_t)code, x = (uint32_t)code;
      for (int j = 1; j < k; j++) {
        x = ((x << 2) | (x >> (2 * k - 2))) & mask;
        if (x < m) m = x;
      }
      if (m == (uint32_t)code) dst[rev_off + n_classes++] = (uint16_t)code;  // canonical codes ascend, so class ids do too
    }
    for (int code = 0; code < n; code++) {
      uint32_t m = (uint32_t)code, x = (uint32_t)code;
      for (int j = 1; j < k; j++) {
        x = ((x << 2) | (x >> (2 * k - 2))) & mask;
        if (x < m) m = x;
      }
      int cls = 0;
      while (dst[rev_off + cls] != (uint16_t)m) cls++;
      dst[lut_off + code] = (uint16_t)(cls * 4 * 32);
    }
    lut_off += n;
    rev_off += n_classes;
  }
}

cudaError_t launch_repeat_scan(const uint32_t *d_seq_words, const uint32_t *d_nmask, const strgpu_segment *d_segs,
                               uint32_t n_seg, uint32_t max_len, const uint16_t *d_thr, const uint16_t *d_luts,
                               strgpu_repeat *d_out, int *d_status, int sm_count, int variant, cudaStream_t stream) {
  if (n_seg == 0) return cudaSuccess;
  constexpr int kWarps = 8;
  const uint32_t blocks_needed = (n_seg + kWarps - 1) / kWarps;
  if (max_len <= (uint32_t)kShortMaxLen && variant != 1) {
    static bool configured = false;
    if (!configured) {
      cudaError_t e = cudaFuncSetAttribute(repeat_scan_lane, cudaFuncAttributeMaxDynamicSharedMemorySize, kLaneSmemBytes);
      if (e != cudaSuccess) return e;
      e = cudaFuncSetAttribute(repeat_scan_lane, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
      if (e != cudaSuccess) return e;
      configured = true;
    }
    const uint32_t tiles = (n_seg + kLaneThreads - 1) / kLaneThreads;
    uint32_t grid = (uint32_t)sm_count * 5u;  // 5 resident CTAs of 128 threads per SM (shared-memory bound)
    if (grid > tiles) grid = tiles;
    repeat_scan_lane<<<grid, kLaneThreads, kLaneSmemBytes, stream>>>(d_seq_words, d_nmask, d_segs, n_seg, d_thr, d_luts, d_out,
                                                                     d_status);
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
