// Launch interface of the repeat-unit scan kernels (K1).  Internal to libstrgpu.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "strgpu.h"

namespace strgpu {

constexpr int kThrLen = 512;                    // thresholds tabulated for len 0..511
constexpr int kThrClasses = STRGPU_MAX_PCLASS + 1;  // + the 0.12 "give up" class (utils.nim:251)
constexpr int kThrMinOff = kThrClasses * 5 * kThrLen;       // then min over k = 2..6 per (proportion class, len): the fused pre-filter's bound
constexpr int kThrFiltOff = kThrMinOff + STRGPU_MAX_PCLASS * kThrLen;   // then 8 uint16 per (class, len): the pre-filter kernel's bounds
constexpr int kThrEntries = kThrFiltOff + STRGPU_MAX_PCLASS * kThrLen * 8;
constexpr int kShortMaxLen = 160;               // kernel variant with the read in <= 10 words

// thr[(cls * 5 + (k - 2)) * kThrLen + len] = int(len * p_cls / k); cls == STRGPU_MAX_PCLASS holds int(len * 0.12 / k);
// thr[kThrMinOff + cls * kThrLen + len] = min over k of the class's thresholds
// thr[kThrFiltOff + (cls * kThrLen + len) * 8 + (k - 2)] = (int(len * p_cls / k) + 1) * (k - 1), k = 2..6 (entries 5..7 unused)
// Implicit whole-read segments: read i = bases [i * stride, i * stride + read_len), proportion class pclass.
struct UniformReads {
  uint32_t n_reads, read_len, stride, pclass;
};

// variant: 0 = default (batches of <= 160-base segments: repeat_prefilter, then one ladder kernel per rung over dense
//              survivor lists, then the warp-per-segment kernel over the segments with non-ACGT bases; batches with longer
//              segments: the warp-per-segment kernel),
//          1 = force the warp-per-segment kernel (kept for A/B measurements and as the long-segment path),
//          5 / 7 = like 0 with 0 / 12 of the 12 popcount streams through carry-save adders (A/B),
//          8 = like 0 with stage lists of 64 entries (exercises the overflow path)
// d_xmask: optional plane of the non-ACGT bases that are NOT the literal 'N' (they stay out of the N > 20 gate)
// d_scratch: device scratch of scan_scratch_words(n_seg) uint32; without it the warp-per-segment kernel runs
cudaError_t launch_repeat_scan(const uint32_t *d_seq_words, const uint32_t *d_nmask, const uint32_t *d_xmask,
                               const strgpu_segment *d_segs, uint32_t n_seg, uint32_t max_len, const uint16_t *d_thr,
                               const uint16_t *d_luts, strgpu_repeat *d_out, int *d_status, int sm_count, int variant,
                               cudaStream_t stream, const UniformReads *uniform = nullptr, uint32_t *d_scratch = nullptr);

constexpr uint32_t kScanScratchHdr = 32;   // header words of the scratch buffer (list counts, group cursors)
constexpr int kScanStageLists = 2;
uint32_t scan_stage_cap(uint32_t n_seg);
size_t scan_scratch_words(uint32_t n_seg);

// kernels one launch_repeat_scan call issues (for the library's launch counter)
int scan_launches(uint32_t max_len, int variant);

constexpr int kLaneLutEntries = 1672 + 4096 + 700 + 1024;  // see build_lane_luts
void build_lane_luts(uint16_t *dst);  // host: fills kLaneLutEntries uint16

}  // namespace strgpu
