// Raw DEFLATE (RFC 1951) decoder for whole BGZF blocks (SAM specification 4.1: every block is one complete deflate
// stream of at most 64 KiB, its compressed and decompressed sizes known before decoding).  Stands in for the htslib /
// zlib inflate the reference reaches through hts-nim (extract.nim:275-329) and is written for exactly that situation: whole-
// buffer in, whole-buffer out, a 64-bit bit buffer refilled with one unaligned load, multi-bit decode tables (11 bits
// litlen / 8 bits distance + sub-tables), up to three literals per refill, word-wide match copies.  Plain C++ without any
// library call so that the same code compiles as a CUDA device function (csrc/decode_kernels.cu, one BGZF block per thread).
//
// Contract of inflate_block(): `in` must be readable for 8 bytes past in + in_len (a BGZF block always has its 8-byte
// CRC32 / ISIZE footer there); exactly out_len bytes are written, nothing beyond out + out_len is touched.
#pragma once
#include <cstdint>
#include <cstring>

#if defined(__CUDACC__)
#define STRLING_HD __host__ __device__
#define STRLING_INFLATE_NO_PAIRS 1   // the device build keeps its tables small (shared memory): no literal-pair entries
#define STRLING_INFLATE_INLINE inline
#else
#define STRLING_HD
// on the host the decoder is instantiated per ISA level (inflate_host.cpp): it has to be inlined into those instances
#define STRLING_INFLATE_INLINE inline __attribute__((always_inline))
#endif

namespace strling {
namespace infl {

constexpr int kLitBits = 11, kDistBits = 8, kPreBits = 7;
constexpr int kLitCap = 2400, kDistCap = 512, kPreCap = 128;
// table entry: value << 16 | flags | codeword bits << 8 | bits to consume (codeword + extra)
// a literal entry with kLiteral2 carries TWO literals (value = first | second << 8) whose codewords both fit the main index
constexpr uint32_t kLiteral = 0x8000u, kSubtable = 0x4000u, kEndOfBlock = 0x2000u, kLiteral2 = 0x1000u;
constexpr uint32_t kInvalid = 0;  // consumes 0 bits and has no flag: only ever seen for incomplete codes / corrupt data

struct Tables {
  uint32_t lit[kLitCap];
  uint32_t dist[kDistCap];
  uint32_t pre[kPreCap];
  uint8_t lens[320];       // litlen + distance code lengths of the current dynamic block
  uint16_t sorted[288];    // scratch of build()
  uint8_t sub_bits[1 << kLitBits];
#if !defined(STRLING_INFLATE_NO_PAIRS)
  uint32_t single[1 << kLitBits];  // scratch of build(): the main table before literal pairs are merged
#endif
  bool fixed_built;
};

enum Status { kOk = 0, kBadBlockType = -1, kBadStored = -2, kBadCode = -3, kBadDistance = -4, kOutputOverrun = -5, kInputOverrun = -6, kOutputShort = -7 };

STRLING_HD inline uint32_t reverse_bits(uint32_t v, int n) {
  uint32_t r = 0;
  for (int i = 0; i < n; i++) { r = (r << 1) | (v & 1u); v >>= 1; }
  return r;
}

// kind 0: literal/length alphabet, 1: distance alphabet, 2: code-length alphabet
STRLING_HD inline bool build(Tables &T, uint32_t *table, int table_bits, int cap, const uint8_t *lens, int n_sym, int kind) {
  const uint16_t len_base[29] = {3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258};
  const uint8_t len_extra[29] = {0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0};
  const uint16_t dist_base[30] = {1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145, 8193, 12289, 16385, 24577};
  const uint8_t dist_extra[30] = {0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13};
  int count[16];
  for (int i = 0; i < 16; i++) count[i] = 0;
  for (int s = 0; s < n_sym; s++) count[lens[s]]++;
  count[0] = 0;
  int offs[16];
  offs[0] = offs[1] = 0;
  int left = 1;
  for (int l = 1; l <= 15; l++) {
    left = (left << 1) - count[l];
    if (left < 0) return false;  // over-subscribed
    if (l < 15) offs[l + 1] = offs[l] + count[l];
  }
  for (int s = 0; s < n_sym; s++)
    if (lens[s]) T.sorted[offs[lens[s]]++] = (uint16_t)s;
  int n_codes = 0;
  for (int l = 1; l <= 15; l++) n_codes += count[l];
  const int main_size = 1 << table_bits;
  for (int i = 0; i < main_size; i++) table[i] = kInvalid;
  auto entry_of = [&](int sym, int code_bits) -> uint32_t {  // code_bits: the bits of the codeword this table level consumes
    if (kind == 2) return ((uint32_t)sym << 16) | kLiteral | ((uint32_t)code_bits << 8) | (uint32_t)code_bits;
    if (kind == 0) {
      if (sym < 256) return ((uint32_t)sym << 16) | kLiteral | ((uint32_t)code_bits << 8) | (uint32_t)code_bits;
      if (sym == 256) return kEndOfBlock | ((uint32_t)code_bits << 8) | (uint32_t)code_bits;
      if (sym > 285) return kInvalid;
      const int i = sym - 257;
      return ((uint32_t)len_base[i] << 16) | ((uint32_t)code_bits << 8) | (uint32_t)(code_bits + len_extra[i]);
    }
    if (sym > 29) return kInvalid;
    return ((uint32_t)dist_base[sym] << 16) | ((uint32_t)code_bits << 8) | (uint32_t)(code_bits + dist_extra[sym]);
  };
  // pass 1: short codes into the main table, and for long codes the deepest length below every main-table prefix
  uint32_t code = 0;
  int idx = 0;
  bool any_long = false;
  for (int l = 1; l <= 15; l++) {
    for (int c = 0; c < count[l]; c++, idx++, code++) {
      const uint32_t rev = reverse_bits(code, l);
      if (l <= table_bits) {
        const uint32_t e = entry_of(T.sorted[idx], l);
        for (uint32_t i = rev; i < (uint32_t)main_size; i += 1u << l) table[i] = e;
      } else {
        const uint32_t prefix = rev & (uint32_t)(main_size - 1);
        if (!any_long) {
          for (int i = 0; i < main_size; i++) T.sub_bits[i] = 0;
          any_long = true;
        }
        if ((int)T.sub_bits[prefix] < l - table_bits) T.sub_bits[prefix] = (uint8_t)(l - table_bits);
      }
    }
    code <<= 1;
  }
  if (any_long) {
  // pass 2: sub-tables
  int next = main_size;
  for (int p = 0; p < main_size; p++) {
    if (!T.sub_bits[p]) continue;
    const int sz = 1 << T.sub_bits[p];
    if (next + sz > cap) return false;
    table[p] = ((uint32_t)next << 16) | kSubtable | ((uint32_t)T.sub_bits[p] << 8) | (uint32_t)table_bits;
    for (int i = 0; i < sz; i++) table[next + i] = kInvalid;
    next += sz;
  }
  code = 0;
  idx = 0;
  for (int l = 1; l <= 15; l++) {
    for (int c = 0; c < count[l]; c++, idx++, code++) {
      if (l <= table_bits) continue;
      const uint32_t rev = reverse_bits(code, l);
      const uint32_t prefix = rev & (uint32_t)(main_size - 1);
      const uint32_t base = table[prefix] >> 16;
      const int sb = (int)T.sub_bits[prefix];
      const uint32_t e = entry_of(T.sorted[idx], l - table_bits);
      for (uint32_t i = rev >> table_bits; i < (1u << sb); i += 1u << (l - table_bits)) table[base + i] = e;
    }
    code <<= 1;
  }
  }
  (void)n_codes;
#if !defined(STRLING_INFLATE_NO_PAIRS)
  if (kind == 0) {  // literal pairs: when the bits behind a short literal code decode to another literal inside the same index
    for (int i = 0; i < main_size; i++) T.single[i] = table[i];
    for (int i = 0; i < main_size; i++) {
      const uint32_t e = T.single[i];
      if (!(e & kLiteral)) continue;
      const int l1 = (int)(e & 0xff), rem = table_bits - l1;
      if (rem < 1) continue;
      const uint32_t e2 = T.single[(i >> l1) & ((1 << rem) - 1)];
      if (!(e2 & kLiteral) || (int)(e2 & 0xff) > rem) continue;
      table[i] = ((e >> 16) << 16) | ((e2 >> 16) << 24) | kLiteral | kLiteral2 | (uint32_t)(l1 + (int)(e2 & 0xff));
    }
  }
#endif
  return true;
}

STRLING_HD inline uint64_t load64(const uint8_t *p) {
#if defined(__CUDA_ARCH__)
  uint64_t v = 0;
  for (int i = 7; i >= 0; i--) v = (v << 8) | p[i];
  return v;
#else
  uint64_t v;
  memcpy(&v, p, 8);
  return v;
#endif
}

struct Bits {
  const uint8_t *in, *in_end;
  uint64_t buf;
  uint32_t cnt;
  STRLING_HD inline bool refill() {
    if (in <= in_end) {
      buf |= load64(in) << cnt;
    } else {  // the last bytes of the stream: never read further than the 8 bytes the caller guarantees behind in_end
      if ((uint64_t)(in - in_end) * 8 > cnt) return false;  // more bits consumed than the stream has
      uint64_t v = 0;
      for (int i = 7; i >= 0; i--) v = (v << 8) | (in + i < in_end + 8 ? in[i] : 0);
      buf |= v << cnt;
    }
    in += (63 - cnt) >> 3;
    cnt |= 56;
    return true;
  }
  STRLING_HD inline void consume(uint32_t n) { buf >>= n; cnt -= n; }
  STRLING_HD inline uint32_t peek(uint32_t n) const { return (uint32_t)(buf & ((1ull << n) - 1)); }
};

STRLING_HD STRLING_INFLATE_INLINE void copy_match(uint8_t *dst, uint32_t dist, uint32_t len, bool room) {
  const uint8_t *src = dst - dist;
#if !defined(__CUDA_ARCH__)
  if (room) {  // at least 16 bytes may be written past dst + len
    if (dist >= 8) {
      uint64_t a, b2;
      memcpy(&a, src, 8);
      memcpy(dst, &a, 8);
      memcpy(&b2, src + 8, 8);
      memcpy(dst + 8, &b2, 8);
      if (len > 16) {
        uint8_t *const end = dst + len;
        dst += 16; src += 16;
        do { memcpy(dst, src, 8); dst += 8; src += 8; } while (dst < end);
      }
      return;
    }
    if (dist == 1) {
      uint64_t v = src[0];
      v *= 0x0101010101010101ull;
      uint8_t *const end = dst + len;
      do { memcpy(dst, &v, 8); dst += 8; } while (dst < end);
      return;
    }
  }
#else
  (void)room;
#endif
  for (uint32_t i = 0; i < len; i++) dst[i] = src[i];
}

// Decodes one complete deflate stream.  Returns kOk when exactly out_len bytes were produced by a stream that ends inside `in`.
STRLING_HD STRLING_INFLATE_INLINE int inflate_block(Tables &T, const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len) {
  Bits b{in, in + in_len, 0, 0};
  uint8_t *o = out;
  uint8_t *const o_end = out + out_len;
  constexpr uint32_t kFastMargin = 8 + 258 + 16;  // three literal pairs (two-byte stores), the longest match, the overshoot of a word-wide copy
  bool last = false;
  while (!last) {
    if (!b.refill()) return kInputOverrun;
    last = b.peek(1) != 0;
    const uint32_t type = (b.peek(3) >> 1);
    b.consume(3);
    if (type == 0) {  // stored
      b.consume(b.cnt & 7);
      const uint8_t *p = b.in - (b.cnt >> 3);
      b.buf = 0;
      b.cnt = 0;
      if (p + 4 > b.in_end) return kBadStored;
      const uint32_t len = (uint32_t)p[0] | ((uint32_t)p[1] << 8), nlen = (uint32_t)p[2] | ((uint32_t)p[3] << 8);
      if ((len ^ nlen) != 0xffffu) return kBadStored;
      p += 4;
      if (p + len > b.in_end) return kInputOverrun;
      if (len > (uint32_t)(o_end - o)) return kOutputOverrun;
      for (uint32_t i = 0; i < len; i++) o[i] = p[i];
      o += len;
      b.in = p + len;
      continue;
    }
    if (type == 3) return kBadBlockType;
    if (type == 1) {
      if (!T.fixed_built) {
        for (int i = 0; i < 144; i++) T.lens[i] = 8;
        for (int i = 144; i < 256; i++) T.lens[i] = 9;
        for (int i = 256; i < 280; i++) T.lens[i] = 7;
        for (int i = 280; i < 288; i++) T.lens[i] = 8;
        for (int i = 0; i < 32; i++) T.lens[288 + i] = 5;
        if (!build(T, T.lit, kLitBits, kLitCap, T.lens, 288, 0) || !build(T, T.dist, kDistBits, kDistCap, T.lens + 288, 32, 1)) return kBadCode;
        T.fixed_built = true;
      }
    } else {
      T.fixed_built = false;
      const uint32_t hlit = b.peek(5) + 257, hdist = (b.peek(10) >> 5) + 1, hclen = (b.peek(14) >> 10) + 4;
      b.consume(14);
      if (hlit > 286 || hdist > 30) return kBadCode;
      const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
      uint8_t pl[19];
      for (int i = 0; i < 19; i++) pl[i] = 0;
      if (!b.refill()) return kInputOverrun;
      for (uint32_t i = 0; i < hclen; i++) {
        if (b.cnt < 3 && !b.refill()) return kInputOverrun;
        pl[order[i]] = (uint8_t)b.peek(3);
        b.consume(3);
        if ((i & 7) == 7 && !b.refill()) return kInputOverrun;
      }
      if (!build(T, T.pre, kPreBits, kPreCap, pl, 19, 2)) return kBadCode;
      uint32_t n = 0;
      while (n < hlit + hdist) {
        if (!b.refill()) return kInputOverrun;
        const uint32_t e = T.pre[b.peek(kPreBits)];
        if (!(e & kLiteral)) return kBadCode;
        b.consume(e & 0xff);
        const uint32_t sym = e >> 16;
        if (sym < 16) { T.lens[n++] = (uint8_t)sym; continue; }
        uint32_t rep, val = 0;
        if (sym == 16) {
          if (n == 0) return kBadCode;
          val = T.lens[n - 1];
          rep = 3 + b.peek(2);
          b.consume(2);
        } else if (sym == 17) {
          rep = 3 + b.peek(3);
          b.consume(3);
        } else {
          rep = 11 + b.peek(7);
          b.consume(7);
        }
        if (n + rep > hlit + hdist) return kBadCode;
        for (uint32_t i = 0; i < rep; i++) T.lens[n++] = (uint8_t)val;
      }
      if (T.lens[256] == 0) return kBadCode;  // no end-of-block code
      // the two alphabets are built from separate arrays: move the distance lengths behind a fixed offset
      uint8_t dl[32];
      for (uint32_t i = 0; i < 32; i++) dl[i] = i < hdist ? T.lens[hlit + i] : 0;
      for (uint32_t i = hlit; i < 288; i++) T.lens[i] = 0;
      if (!build(T, T.lit, kLitBits, kLitCap, T.lens, 288, 0) || !build(T, T.dist, kDistBits, kDistCap, dl, 32, 1)) return kBadCode;
    }
    // ---- symbols of this block
    bool eob = false;
#if !defined(__CUDA_ARCH__)
    {  // fast loop: far enough from both ends that neither the refills nor the word-wide copies need a bounds check
      const uint8_t *ip = b.in;
      uint64_t buf = b.buf;
      uint32_t cnt = b.cnt;
      constexpr uint64_t kLitMask = (1u << kLitBits) - 1, kDistMask = (1u << kDistBits) - 1;
      while (ip + 8 <= b.in_end && (uint32_t)(o_end - o) >= kFastMargin) {
        buf |= load64(ip) << cnt;
        ip += (63 - cnt) >> 3;
        cnt |= 56;
        uint32_t e = T.lit[buf & kLitMask];
#define STRLING_EMIT_LITERALS(e)                                   \
  do {                                                             \
    buf >>= ((e) & 0xff);                                          \
    cnt -= ((e) & 0xff);                                           \
    const uint32_t two = (e) >> 16;                                \
    memcpy(o, &two, 4);                                            \
    o += 1 + (((e) >> 12) & 1);                                    \
  } while (0)
        if (e & kLiteral) {
          STRLING_EMIT_LITERALS(e);
          e = T.lit[buf & kLitMask];
          if (e & kLiteral) {
            STRLING_EMIT_LITERALS(e);
            e = T.lit[buf & kLitMask];
            if (e & kLiteral) {
              STRLING_EMIT_LITERALS(e);
              continue;
            }
          }
          buf |= load64(ip) << cnt;
          ip += (63 - cnt) >> 3;
          cnt |= 56;
        }
        if (__builtin_expect((e & kSubtable) != 0, 0)) {
          buf >>= kLitBits; cnt -= kLitBits;
          e = T.lit[(e >> 16) + (uint32_t)(buf & ((1u << ((e >> 8) & 0x1f)) - 1))];
          if (e & kLiteral) {
            STRLING_EMIT_LITERALS(e);
            continue;
          }
        }
#undef STRLING_EMIT_LITERALS
        const uint32_t total = e & 0xff, cw = (e >> 8) & 0x1f;
        if (__builtin_expect((e & kEndOfBlock) != 0, 0)) { buf >>= total; cnt -= total; eob = true; break; }
        if (total == 0) return kBadCode;
        const uint32_t len = (e >> 16) + ((uint32_t)(buf & ((1ull << total) - 1)) >> cw);
        buf >>= total; cnt -= total;
        uint32_t d = T.dist[buf & kDistMask];
        if (__builtin_expect((d & kSubtable) != 0, 0)) {
          buf >>= kDistBits; cnt -= kDistBits;
          d = T.dist[(d >> 16) + (uint32_t)(buf & ((1u << ((d >> 8) & 0x1f)) - 1))];
        }
        const uint32_t dtotal = d & 0xff, dcw = (d >> 8) & 0x1f;
        if (dtotal == 0) return kBadCode;
        const uint32_t dist = (d >> 16) + ((uint32_t)(buf & ((1ull << dtotal) - 1)) >> dcw);
        buf >>= dtotal; cnt -= dtotal;
        if (dist > (uint32_t)(o - out)) return kBadDistance;
        copy_match(o, dist, len, true);
        o += len;
      }
      b.in = ip;
      b.buf = buf;
      b.cnt = cnt;
    }
#endif
    while (!eob) {
      if (!b.refill()) return kInputOverrun;
      const bool room = (uint32_t)(o_end - o) >= kFastMargin;
      uint32_t e = T.lit[b.peek(kLitBits)];
      if (e & kSubtable) { b.consume(kLitBits); e = T.lit[(e >> 16) + b.peek((e >> 8) & 0x1f)]; }
      if (e & kLiteral) {  // one literal or a pair; the tail of a block is decoded one entry per refill
        const uint32_t n_lit = 1 + ((e >> 12) & 1);
        if ((uint32_t)(o_end - o) < n_lit) return kOutputOverrun;
        b.consume(e & 0xff);
        *o++ = (uint8_t)(e >> 16);
        if (n_lit == 2) *o++ = (uint8_t)(e >> 24);
        continue;
      }
      if (e & kEndOfBlock) { b.consume(e & 0xff); break; }
      const uint32_t total = e & 0xff, cw = (e >> 8) & 0x1f;
      if (total == 0) return kBadCode;
      const uint32_t len = (e >> 16) + (b.peek(total) >> cw);
      b.consume(total);
      uint32_t d = T.dist[b.peek(kDistBits)];
      if (d & kSubtable) { b.consume(kDistBits); d = T.dist[(d >> 16) + b.peek((d >> 8) & 0x1f)]; }
      const uint32_t dtotal = d & 0xff, dcw = (d >> 8) & 0x1f;
      if (dtotal == 0) return kBadCode;
      const uint32_t dist = (d >> 16) + (b.peek(dtotal) >> dcw);
      b.consume(dtotal);
      if (dist > (uint32_t)(o - out)) return kBadDistance;
      if (len > (uint32_t)(o_end - o)) return kOutputOverrun;
      copy_match(o, dist, len, room);
      o += len;
    }
  }
  if (o != o_end) return kOutputShort;
  // bytes actually consumed: b.in minus the whole bytes still in the bit buffer
  if (b.in - (b.cnt >> 3) > b.in_end) return kInputOverrun;
  return kOk;
}

// ---- The same decoder as a pull-style state machine: next() decodes up to the next COPY (a match or a stored block), writing the
// literals on the way straight into the output window, and returns that copy as a command.  This is the form the
// warp-per-block CUDA kernel (csrc/decode_kernels.cu, kernel v2) runs: lane 0 calls next(), the command is broadcast and all 32
// lanes execute the copy.  inflate_block_stream() below executes the commands serially -- the CPU check of this decoder
// (`strling debug inflate-selftest` runs it beside inflate_block on every case).
enum CommandType { kCmdEnd = 0, kCmdMatch = 1, kCmdStored = 2 };   // negative: a Status
struct Command {
  int type;
  uint32_t o;   // where the copy starts in the output
  uint32_t a;   // match: distance; stored: offset of the bytes in the input
  uint32_t b;   // length
};

struct Stream {
  Bits b;
  const uint8_t *in0;
  uint32_t o, o_end;
  bool last, in_block;

  STRLING_HD void init(const uint8_t *in, uint32_t in_len, uint32_t out_len) {
    b.in = in;
    b.in_end = in + in_len;
    b.buf = 0;
    b.cnt = 0;
    in0 = in;
    o = 0;
    o_end = out_len;
    last = false;
    in_block = false;
  }

  STRLING_HD static Command status(int st) { return Command{st, 0, 0, 0}; }

  // block header of a fixed / dynamic block (the 3 header bits are consumed already): builds T.lit / T.dist
  STRLING_HD int read_tables(Tables &T, uint32_t type) {
    if (type == 1) {
      if (!T.fixed_built) {
        for (int i = 0; i < 144; i++) T.lens[i] = 8;
        for (int i = 144; i < 256; i++) T.lens[i] = 9;
        for (int i = 256; i < 280; i++) T.lens[i] = 7;
        for (int i = 280; i < 288; i++) T.lens[i] = 8;
        for (int i = 0; i < 32; i++) T.lens[288 + i] = 5;
        if (!build(T, T.lit, kLitBits, kLitCap, T.lens, 288, 0) || !build(T, T.dist, kDistBits, kDistCap, T.lens + 288, 32, 1)) return kBadCode;
        T.fixed_built = true;
      }
      return kOk;
    }
    T.fixed_built = false;
    const uint32_t hlit = b.peek(5) + 257, hdist = (b.peek(10) >> 5) + 1, hclen = (b.peek(14) >> 10) + 4;
    b.consume(14);
    if (hlit > 286 || hdist > 30) return kBadCode;
    const uint8_t order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    uint8_t pl[19];
    for (int i = 0; i < 19; i++) pl[i] = 0;
    for (uint32_t i = 0; i < hclen; i++) {
      if ((i & 7) == 0 && !b.refill()) return kInputOverrun;
      pl[order[i]] = (uint8_t)b.peek(3);
      b.consume(3);
    }
    if (!build(T, T.pre, kPreBits, kPreCap, pl, 19, 2)) return kBadCode;
    uint32_t n = 0;
    while (n < hlit + hdist) {
      if (!b.refill()) return kInputOverrun;
      const uint32_t e = T.pre[b.peek(kPreBits)];
      if (!(e & kLiteral)) return kBadCode;
      b.consume(e & 0xff);
      const uint32_t sym = e >> 16;
      if (sym < 16) { T.lens[n++] = (uint8_t)sym; continue; }
      uint32_t rep, val = 0;
      if (sym == 16) {
        if (n == 0) return kBadCode;
        val = T.lens[n - 1];
        rep = 3 + b.peek(2);
        b.consume(2);
      } else if (sym == 17) {
        rep = 3 + b.peek(3);
        b.consume(3);
      } else {
        rep = 11 + b.peek(7);
        b.consume(7);
      }
      if (n + rep > hlit + hdist) return kBadCode;
      for (uint32_t i = 0; i < rep; i++) T.lens[n++] = (uint8_t)val;
    }
    if (T.lens[256] == 0) return kBadCode;
    uint8_t dl[32];
    for (uint32_t i = 0; i < 32; i++) dl[i] = i < hdist ? T.lens[hlit + i] : 0;
    for (uint32_t i = hlit; i < 288; i++) T.lens[i] = 0;
    if (!build(T, T.lit, kLitBits, kLitCap, T.lens, 288, 0) || !build(T, T.dist, kDistBits, kDistCap, dl, 32, 1)) return kBadCode;
    return kOk;
  }

  STRLING_HD Command next(Tables &T, uint8_t *window) {
    while (true) {
      if (!in_block) {
        if (last) {
          if (o != o_end) return status(kOutputShort);
          if (b.in - (b.cnt >> 3) > b.in_end) return status(kInputOverrun);
          return Command{kCmdEnd, o, 0, 0};
        }
        if (!b.refill()) return status(kInputOverrun);
        last = b.peek(1) != 0;
        const uint32_t type = b.peek(3) >> 1;
        b.consume(3);
        if (type == 0) {
          b.consume(b.cnt & 7);
          const uint8_t *p = b.in - (b.cnt >> 3);
          b.buf = 0;
          b.cnt = 0;
          if (p + 4 > b.in_end) return status(kBadStored);
          const uint32_t len = (uint32_t)p[0] | ((uint32_t)p[1] << 8), nlen = (uint32_t)p[2] | ((uint32_t)p[3] << 8);
          if ((len ^ nlen) != 0xffffu) return status(kBadStored);
          p += 4;
          if (p + len > b.in_end) return status(kInputOverrun);
          if (len > o_end - o) return status(kOutputOverrun);
          b.in = p + len;
          if (len == 0) continue;
          const Command c{kCmdStored, o, (uint32_t)(p - in0), len};
          o += len;
          return c;
        }
        if (type == 3) return status(kBadBlockType);
        const int rc = read_tables(T, type);
        if (rc != kOk) return status(rc);
        in_block = true;
      }
      // a symbol takes at most 15 + 5 + 15 + 13 = 48 bits: the bit buffer is topped up only when it holds fewer
      if (b.cnt < 48 && !b.refill()) return status(kInputOverrun);
      uint32_t e = T.lit[b.peek(kLitBits)];
      if (e & kSubtable) { b.consume(kLitBits); e = T.lit[(e >> 16) + b.peek((e >> 8) & 0x1f)]; }
      if (e & kLiteral) {
        const uint32_t n_lit = 1 + ((e >> 12) & 1);
        if (o_end - o < n_lit) return status(kOutputOverrun);
        b.consume(e & 0xff);
        window[o++] = (uint8_t)(e >> 16);
        if (n_lit == 2) window[o++] = (uint8_t)(e >> 24);
        continue;
      }
      if (e & kEndOfBlock) { b.consume(e & 0xff); in_block = false; continue; }
      const uint32_t total = e & 0xff, cw = (e >> 8) & 0x1f;
      if (total == 0) return status(kBadCode);
      const uint32_t len = (e >> 16) + (b.peek(total) >> cw);
      b.consume(total);
      uint32_t d = T.dist[b.peek(kDistBits)];
      if (d & kSubtable) { b.consume(kDistBits); d = T.dist[(d >> 16) + b.peek((d >> 8) & 0x1f)]; }
      const uint32_t dtotal = d & 0xff, dcw = (d >> 8) & 0x1f;
      if (dtotal == 0) return status(kBadCode);
      const uint32_t dist = (d >> 16) + (b.peek(dtotal) >> dcw);
      b.consume(dtotal);
      if (dist > o) return status(kBadDistance);
      if (len > o_end - o) return status(kOutputOverrun);
      const Command c{kCmdMatch, o, dist, len};
      o += len;
      return c;
    }
  }
};

// byte i of a match copy, for an executor that copies the bytes of one command in any order (the source of byte i lies
// before the command's first output byte even when the match overlaps itself)
STRLING_HD inline uint32_t match_source(const Command &c, uint32_t i) { return c.o - c.a + (i < c.a ? i : i % c.a); }

STRLING_HD inline int inflate_block_stream(Tables &T, const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len) {
  Stream s;
  s.init(in, in_len, out_len);
  while (true) {
    const Command c = s.next(T, out);
    if (c.type == kCmdEnd) return kOk;
    if (c.type < 0) return c.type;
    if (c.type == kCmdMatch) {
      for (uint32_t i = c.b; i-- > 0;) out[c.o + i] = out[match_source(c, i)];   // back to front: any order must do
    } else {
      for (uint32_t i = 0; i < c.b; i++) out[c.o + i] = in[c.a + i];
    }
  }
}

}  // namespace infl
}  // namespace strling
