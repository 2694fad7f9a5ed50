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
};

template <int MAXLEN, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) repeat_scan_warp(const uint32_t *__restrict__ seq,
                                                               const uint32_t *__restrict__ nmask,
                                                               const strgpu_segment *__restrict__ segs, uint32_t n_seg,
                                                               const uint16_t *__restrict__ thr,
                                                               strgpu_repeat *__restrict__ out, int *status) {
  constexpr int MAXR2 = (MAXLEN / 2 + 31) / 32, MAXR3 = (MAXLEN / 3 + 31) / 32, MAXR4 = (MAXLEN / 4 + 31) / 32,
                MAXR5 = (MAXLEN / 5 + 31) / 32, MAXR6 = (MAXLEN / 6 + 31) / 32;
  constexpr int MAXP = (MAXLEN + 31) / 32;
  __shared__ WarpScratch<MAXLEN> scratch[WARPS];
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  WarpScratch<MAXLEN> &ws = scratch[warp];
  for (int i = lane; i < WarpScratch<MAXLEN>::kTab; i += 32) ws.tab[i] = 0;
  __syncwarp();

  const uint32_t warps_total = gridDim.x * WARPS;
  for (uint32_t s = blockIdx.x * WARPS + warp; s < n_seg; s += warps_total) {
    const strgpu_segment sg = segs[s];
    const int L = sg.len;
    strgpu_repeat res;
#pragma unroll
    for (int i = 0; i < 6; i++) res.unit[i] = 0;
    res.repeat_count = 0;
    if (L > MAXLEN || L > STRGPU_MAX_SEGMENT_LEN) {  // warp-uniform
      if (lane == 0) {
        atomicExch(status, (int)STRGPU_ERR_TOO_LONG);
        store_result(out, s, res);
      }
      continue;
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

    ScanState st{-1, 0u, 0, 0};
    if (n_count <= 20) {  // utils.nim:238
      const int pclass = sg.pclass < STRGPU_MAX_PCLASS ? sg.pclass : STRGPU_MAX_PCLASS - 1;
      const uint16_t *tp = thr + (size_t)(pclass * 5) * kThrLen + L;
      const uint16_t *tg = thr + (size_t)(STRGPU_MAX_PCLASS * 5) * kThrLen + L;
      bool go = ladder_step<2, MAXR2, MAXP>(ws.sw, ws.nm, has_n, ws.tab, L, lane, tp[0], tg[0], st);
      if (go) go = ladder_step<3, MAXR3, MAXP>(ws.sw, ws.nm, has_n, ws.tab, L, lane, tp[kThrLen], tg[kThrLen], st);
      if (go) go = ladder_step<4, MAXR4, MAXP>(ws.sw, ws.nm, has_n, ws.tab, L, lane, tp[2 * kThrLen], tg[2 * kThrLen], st);
      if (go) go = ladder_step<5, MAXR5, MAXP>(ws.sw, ws.nm, has_n, ws.tab, L, lane, tp[3 * kThrLen], tg[3 * kThrLen], st);
      if (go) ladder_step<6, MAXR6, MAXP>(ws.sw, ws.nm, has_n, ws.tab, L, lane, tp[4 * kThrLen], tg[4 * kThrLen], st);
    }
    if (lane == 0) {
      if (st.unit_k > 0) {
        // decode (kmer.decode with alphabet "CATG") + reduce_repeat (utils.nim:220-233,271)
        const uint32_t alpha = 0x47544143u;  // 'C','A','T','G' little-endian
        bool homo = true;
        const uint32_t first = (st.unit_code >> (2 * (st.unit_k - 1))) & 3u;
        for (int j = 0; j < st.unit_k; j++) {
          const uint32_t b = (st.unit_code >> (2 * (st.unit_k - 1 - j))) & 3u;
          res.unit[j] = (char)((alpha >> (8 * b)) & 0xffu);
          homo = homo && (b == first);
        }
        int rc = st.rc;
        if (homo) {
          for (int j = 1; j < st.unit_k; j++) res.unit[j] = 0;
          rc *= st.unit_k;
        }
        res.repeat_count = (uint16_t)rc;
      }
      store_result(out, s, res);
    }
  }
}

}  // namespace

cudaError_t launch_repeat_scan(const uint32_t *d_seq_words, const uint32_t *d_nmask, const strgpu_segment *d_segs,
                               uint32_t n_seg, uint32_t max_len, const uint16_t *d_thr, strgpu_repeat *d_out,
                               int *d_status, int sm_count, cudaStream_t stream) {
  if (n_seg == 0) return cudaSuccess;
  constexpr int kWarps = 8;
  const uint32_t blocks_needed = (n_seg + kWarps - 1) / kWarps;
  if (max_len <= (uint32_t)kShortMaxLen) {
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
