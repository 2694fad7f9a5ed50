// Host-side sequence packers of libstrgpu (include/strgpu.h): ASCII or BAM 4-bit SEQ -> 2-bit + N mask.
// Base codes follow the `kmer` nimble package the reference scans with (utils.nim:14): C=0 A=1 T=2 G=3,
// anything else is stored as 1 ('A') and flagged in the N mask; a non-ACGT base that is not the literal 'N' (an IUPAC
// ambiguity code, '=') is flagged in the second plane too, because the reference's N > 20 gate counts 'N' only (utils.nim:238).
#include <cstdint>
#include <cstring>
#if defined(__x86_64__)
#include <immintrin.h>
#endif

#include "strgpu.h"

namespace {

struct AsciiLut {
  uint8_t code[256];
  uint8_t other[256];
  AsciiLut() {
    for (int i = 0; i < 256; i++) { code[i] = 1; other[i] = 1; }
    const char *acgt = "CATG";
    for (int j = 0; j < 4; j++) {
      code[(unsigned char)acgt[j]] = (uint8_t)j;
      other[(unsigned char)acgt[j]] = 0;
      // hts-nim's aln.sequence() only yields upper case; lower case is treated the same way the kmer
      // package's table does
      code[(unsigned char)(acgt[j] + 32)] = (uint8_t)j;
      other[(unsigned char)(acgt[j] + 32)] = 0;
    }
  }
};
const AsciiLut kAscii;

// BAM nibble: "=ACMGRSVTWYHKDBN"; A=1 C=2 G=4 T=8
struct Bam4Lut {
  uint8_t code[16];
  uint8_t other[16];
  uint8_t pair_code[256];   // two nibbles -> 4 bits (first base in the high pair)
  uint8_t pair_other[256];  // bit1 = first base non-ACGT, bit0 = second
  Bam4Lut() {
    for (int i = 0; i < 16; i++) { code[i] = 1; other[i] = 1; }
    code[1] = 1; other[1] = 0;  // A
    code[2] = 0; other[2] = 0;  // C
    code[4] = 3; other[4] = 0;  // G
    code[8] = 2; other[8] = 0;  // T
    for (int b = 0; b < 256; b++) {
      pair_code[b] = (uint8_t)((code[b >> 4] << 2) | code[b & 15]);
      pair_other[b] = (uint8_t)((other[b >> 4] << 1) | other[b & 15]);
    }
  }
};
const Bam4Lut kBam4;

inline void set_n(uint32_t *nmask, uint64_t b) { nmask[b >> 5] |= 1u << (b & 31); }
inline void flag(uint32_t *nmask, uint32_t *xmask, uint64_t b, bool literal_n) {
  if (nmask) set_n(nmask, b);
  if (xmask && !literal_n) set_n(xmask, b);
}

#if defined(__x86_64__)
// Sixteen bases (eight BAM bytes) per step: nibble -> code is two OR-shifts (bit 0 = A | G, bit 1 = G | T with A=1 C=2 G=4
// T=8), the 2-bit codes are gathered with PEXT and the two nibbles of every output byte swapped (BAM puts the first base of
// a byte in the high nibble, PEXT gathers from the low end).  Stops at the first group that holds anything but A, C, G, T
// and returns the number of bases done; the table loop takes over from there.
__attribute__((target("bmi2"))) uint32_t pack_bam4_bmi2(const uint8_t *bam_seq, uint32_t len, uint8_t *dst) {
  uint32_t i = 0;
  for (; i + 16 <= len; i += 16) {
    uint64_t w;
    std::memcpy(&w, bam_seq + (i >> 1), 8);
    const uint64_t pairs = (w & 0x5555555555555555ull) + ((w >> 1) & 0x5555555555555555ull);
    const uint64_t pop = (pairs & 0x3333333333333333ull) + ((pairs >> 2) & 0x3333333333333333ull);
    if (pop != 0x1111111111111111ull) break;  // a nibble that is not one-hot: '=', N or an IUPAC code
    const uint64_t lo = (w | (w >> 2)) & 0x1111111111111111ull;
    const uint64_t hi = ((w >> 2) | (w >> 3)) & 0x1111111111111111ull;
    const uint32_t x = (uint32_t)_pext_u64(lo | (hi << 1), 0x3333333333333333ull);
    const uint32_t y = ((x & 0x0f0f0f0fu) << 4) | ((x >> 4) & 0x0f0f0f0fu);
    std::memcpy(dst + (i >> 2), &y, 4);
  }
  return i;
}
const bool kHaveBmi2 = __builtin_cpu_supports("bmi2");
#endif

}  // namespace

extern "C" {

int strgpu_pack_ascii(const char *seq, uint32_t len, uint8_t *seq2, uint32_t *nmask, uint32_t *xmask, uint64_t base_off) {
  if (!seq2 || (len && !seq) || (base_off & 3)) return STRGPU_ERR_INVALID;
  uint8_t *dst = seq2 + (base_off >> 2);
  int n_other = 0;
  uint32_t i = 0;
  for (; i + 4 <= len; i += 4) {
    const unsigned char a = seq[i], b = seq[i + 1], c = seq[i + 2], d = seq[i + 3];
    dst[i >> 2] = (uint8_t)((kAscii.code[a] << 6) | (kAscii.code[b] << 4) | (kAscii.code[c] << 2) | kAscii.code[d]);
    const int o = kAscii.other[a] | kAscii.other[b] | kAscii.other[c] | kAscii.other[d];
    if (o) {
      for (int j = 0; j < 4; j++)
        if (kAscii.other[(unsigned char)seq[i + j]]) { n_other++; flag(nmask, xmask, base_off + i + j, seq[i + j] == 'N'); }
    }
  }
  if (i < len) {
    uint8_t v = 0;
    for (uint32_t j = 0; j < 4; j++) {
      uint8_t c = 0;
      if (i + j < len) {
        const unsigned char ch = seq[i + j];
        c = kAscii.code[ch];
        if (kAscii.other[ch]) { n_other++; flag(nmask, xmask, base_off + i + j, ch == 'N'); }
      }
      v = (uint8_t)((v << 2) | c);
    }
    dst[i >> 2] = v;
  }
  return n_other;
}

int strgpu_pack_bam4(const uint8_t *bam_seq, uint32_t len, uint8_t *seq2, uint32_t *nmask, uint32_t *xmask, uint64_t base_off) {
  if (!seq2 || (len && !bam_seq) || (base_off & 3)) return STRGPU_ERR_INVALID;
  uint8_t *dst = seq2 + (base_off >> 2);
  int n_other = 0;
  uint32_t i = 0;  // base index
#if defined(__x86_64__)
  if (kHaveBmi2) i = pack_bam4_bmi2(bam_seq, len, dst);
#endif
  for (; i + 4 <= len; i += 4) {
    const uint8_t b0 = bam_seq[i >> 1], b1 = bam_seq[(i >> 1) + 1];
    dst[i >> 2] = (uint8_t)((kBam4.pair_code[b0] << 4) | kBam4.pair_code[b1]);
    const int o = (kBam4.pair_other[b0] << 2) | kBam4.pair_other[b1];
    if (o) {
      for (int j = 0; j < 4; j++)
        if (o & (8 >> j)) {
          const uint32_t b = i + j;
          const uint8_t nib = (b & 1) ? (bam_seq[b >> 1] & 15) : (bam_seq[b >> 1] >> 4);
          n_other++;
          flag(nmask, xmask, base_off + b, nib == 15);
        }
    }
  }
  if (i < len) {
    uint8_t v = 0;
    for (uint32_t j = 0; j < 4; j++) {
      uint8_t c = 0;
      if (i + j < len) {
        const uint32_t b = i + j;
        const uint8_t nib = (b & 1) ? (bam_seq[b >> 1] & 15) : (bam_seq[b >> 1] >> 4);
        c = kBam4.code[nib];
        if (kBam4.other[nib]) { n_other++; flag(nmask, xmask, base_off + b, nib == 15); }
      }
      v = (uint8_t)((v << 2) | c);
    }
    dst[i >> 2] = v;
  }
  return n_other;
}

}  // extern "C"
