// `strling debug synth-bam`: a coordinate-sorted synthetic BAM in the shape of BASELINE.json configs[1] (SURVEY 8d, config 2:
// 150-bp pairs, positions uniform over 24 contigs, insert ~N(400, 80) clipped to [150, 4095]; per read 90 % plain 150M,
// 7 % messy (one I / D or a soft clip <= 16), 2 % clipped-STR (an S > 16 clip that is a repeat), 1 % STR (a repeat of a random
// 1..6-mer with 1 % substitutions, mapped poorly), 1 % of the pairs without coordinates at the end), written at several
// hundred MB per second so that the command-line measurements (`bench.py` cli leg, tools/bench_cli.py) run on BAMs of 10^6-10^7
// reads.  Every field of a pair is a pure function of (seed, pair index), so the records are generated where they are
// needed (twice: once per mate) and in parallel.  Measurement input only: no parity claim rests on this file.
#include <zlib.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

#include "commands.hpp"

namespace strling {

namespace {

struct Rng {
  uint64_t s;
  explicit Rng(uint64_t seed) : s(seed) {}
  uint64_t next() {  // splitmix64
    uint64_t z = (s += 0x9e3779b97f4a7c15ull);
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
  }
  uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
  double unit() { return (double)(next() >> 11) * (1.0 / 9007199254740992.0); }
};

constexpr int kContigs = 24;
constexpr uint32_t kContigLen = 125000000u;
constexpr int kReadLen = 150;

struct Mate {
  int32_t tid, pos, mate_tid, mate_pos, isize;
  uint16_t flag;
  uint8_t mapq;
  uint32_t cigar[3];
  int n_cigar;
  char seq[kReadLen];
};

void random_bases(Rng &g, char *dst, int n) {
  for (int i = 0; i < n;) {
    uint64_t r = g.next();
    for (int j = 0; j < 32 && i < n; j++, i++, r >>= 2) dst[i] = "ACGT"[r & 3];
  }
}
void repeat_bases(Rng &g, char *dst, int n, const char *unit, int k) {
  const int phase = (int)g.below((uint32_t)k);
  for (int i = 0; i < n; i++) dst[i] = unit[(i + phase) % k];
  for (int i = 0; i < n; i++)
    if (g.below(100) == 0) dst[i] = "ACGT"[g.below(4)];
}

// both mates of pair `i` (mate 0 forward at `pos`, mate 1 reverse at pos + insert - 150)
void make_pair(uint64_t seed, uint64_t i, bool unplaced, Mate m[2]) {
  Rng g(seed * 0x100000001b3ull + i * 0x9e3779b97f4a7c15ull + 12345);
  double z = 0;
  for (int k = 0; k < 12; k++) z += g.unit();
  int insert = (int)std::lround(400.0 + 80.0 * (z - 6.0));
  insert = std::max(150, std::min(4095, insert));
  const int tid = (int)g.below(kContigs);
  const int pos = (int)g.below(kContigLen - 5000u) + 100;
  for (int w = 0; w < 2; w++) {
    Mate &a = m[w];
    a.tid = unplaced ? -1 : tid;
    a.pos = unplaced ? -1 : (w == 0 ? pos : pos + insert - kReadLen);
    a.mapq = 60;
    a.n_cigar = 1;
    a.cigar[0] = (uint32_t)kReadLen << 4;  // 150M
    const uint32_t cls = g.below(100);
    if (cls < 90) {
      random_bases(g, a.seq, kReadLen);
    } else if (cls < 97) {  // messy: one insertion / deletion or a short clip
      random_bases(g, a.seq, kReadLen);
      const uint32_t kind = g.below(3);
      const uint32_t at = 20 + g.below(100);
      const uint32_t len = 1 + g.below(12);
      if (kind == 0) { a.n_cigar = 3; a.cigar[0] = at << 4; a.cigar[1] = (len << 4) | 1u; a.cigar[2] = ((uint32_t)kReadLen - at - len) << 4; }
      else if (kind == 1) { a.n_cigar = 3; a.cigar[0] = at << 4; a.cigar[1] = (len << 4) | 2u; a.cigar[2] = ((uint32_t)kReadLen - at) << 4; }
      else { a.n_cigar = 2; a.cigar[0] = ((uint32_t)kReadLen - len) << 4; a.cigar[1] = (len << 4) | 4u; }
    } else {
      char unit[6];
      const int k = 1 + (int)g.below(6);
      random_bases(g, unit, k);
      if (cls < 99) {  // clipped-STR: the clipped part is the repeat
        const int clip = 30 + (int)g.below(60);
        a.n_cigar = 2;
        if (g.below(2)) {
          repeat_bases(g, a.seq, clip, unit, k);
          random_bases(g, a.seq + clip, kReadLen - clip);
          a.cigar[0] = ((uint32_t)clip << 4) | 4u;
          a.cigar[1] = (uint32_t)(kReadLen - clip) << 4;
        } else {
          random_bases(g, a.seq, kReadLen - clip);
          repeat_bases(g, a.seq + kReadLen - clip, clip, unit, k);
          a.cigar[0] = (uint32_t)(kReadLen - clip) << 4;
          a.cigar[1] = ((uint32_t)clip << 4) | 4u;
        }
      } else {  // STR read, placed poorly
        repeat_bases(g, a.seq, kReadLen, unit, k);
        a.mapq = (uint8_t)g.below(20);
      }
    }
    if (unplaced) { a.n_cigar = 0; a.mapq = 0; }
  }
  for (int w = 0; w < 2; w++) {
    Mate &a = m[w];
    const Mate &b = m[1 - w];
    a.mate_tid = b.tid;
    a.mate_pos = b.pos;
    a.isize = unplaced ? 0 : (w == 0 ? insert : -insert);
    a.flag = (uint16_t)(0x1 | (w == 0 ? 0x40 | 0x20 : 0x80 | 0x10));
    if (unplaced) a.flag = (uint16_t)(0x1 | 0x4 | 0x8 | (w == 0 ? 0x40 : 0x80));
    else a.flag |= 0x2;
  }
}

int ref_length(const Mate &a) {
  int rl = 0;
  for (int i = 0; i < a.n_cigar; i++) {
    const int op = (int)(a.cigar[i] & 15u);
    if (op == 0 || op == 2) rl += (int)(a.cigar[i] >> 4);
  }
  return rl ? rl : 1;
}
int reg2bin(int beg, int end) {  // SAM specification 5.3
  --end;
  if (beg >> 14 == end >> 14) return ((1 << 15) - 1) / 7 + (beg >> 14);
  if (beg >> 17 == end >> 17) return ((1 << 12) - 1) / 7 + (beg >> 17);
  if (beg >> 20 == end >> 20) return ((1 << 9) - 1) / 7 + (beg >> 20);
  if (beg >> 23 == end >> 23) return ((1 << 6) - 1) / 7 + (beg >> 23);
  if (beg >> 26 == end >> 26) return ((1 << 3) - 1) / 7 + (beg >> 26);
  return 0;
}

constexpr int kNameLen = 24;  // "SYN1:0000:00000000000/p" stand-in for an instrument read name, NUL included

size_t record_bytes(const Mate &a) { return 4 + 32 + kNameLen + 4 * (size_t)a.n_cigar + (kReadLen + 1) / 2 + kReadLen; }

void write_record(uint8_t *p, const Mate &a, uint64_t pair, uint64_t seed) {
  const int32_t block_size = (int32_t)record_bytes(a) - 4;
  auto put32 = [&](size_t off, int32_t v) { std::memcpy(p + off, &v, 4); };
  auto put16 = [&](size_t off, uint16_t v) { std::memcpy(p + off, &v, 2); };
  put32(0, block_size);
  put32(4, a.tid);
  put32(8, a.pos);
  p[12] = (uint8_t)kNameLen;
  p[13] = a.mapq;
  put16(14, (uint16_t)(a.tid < 0 ? 4680 : reg2bin(a.pos, a.pos + ref_length(a))));
  put16(16, (uint16_t)a.n_cigar);
  put16(18, a.flag);
  put32(20, kReadLen);
  put32(24, a.mate_tid);
  put32(28, a.mate_pos);
  put32(32, a.isize);
  char name[64];
  std::snprintf(name, sizeof(name), "SYN%llu:%04llu:%013llu", (unsigned long long)(seed % 10), (unsigned long long)(pair % 9973), (unsigned long long)pair);
  std::memset(p + 36, 0, kNameLen);
  std::memcpy(p + 36, name, std::min<size_t>(std::strlen(name), kNameLen - 1));
  uint8_t *q = p + 36 + kNameLen;
  std::memcpy(q, a.cigar, 4 * (size_t)a.n_cigar);
  q += 4 * (size_t)a.n_cigar;
  auto nib = [](char c) -> uint8_t { return c == 'A' ? 1 : c == 'C' ? 2 : c == 'G' ? 4 : c == 'T' ? 8 : 15; };
  for (int i = 0; i < kReadLen; i += 2) q[i / 2] = (uint8_t)((nib(a.seq[i]) << 4) | (i + 1 < kReadLen ? nib(a.seq[i + 1]) : 0));
  q += (kReadLen + 1) / 2;
  // binned qualities in runs (what current instruments emit): compresses about like a real file
  Rng g(pair * 2 + (uint64_t)(a.flag & 0x80) + seed);
  static const uint8_t bins[4] = {37, 37, 25, 11};
  for (int i = 0; i < kReadLen;) {
    const uint8_t v = bins[g.below(4)];
    const int run = 1 + (int)g.below(24);
    for (int j = 0; j < run && i < kReadLen; j++, i++) q[i] = v;
  }
}

template <typename F>
void run_threads(int nt, size_t n, F f) {
  std::vector<std::thread> th;
  const size_t per = (n + (size_t)nt - 1) / (size_t)nt;
  for (int t = 0; t < nt; t++) {
    const size_t a = (size_t)t * per, e = std::min(n, a + per);
    if (a >= e) break;
    th.emplace_back([=]() { f(a, e); });
  }
  for (auto &x : th) x.join();
}

}  // namespace

int synth_bam(const std::string &path, uint64_t n_pairs, uint64_t seed, int level, int threads) {
  const int nt = threads > 0 ? threads : (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  const uint64_t n_unplaced = n_pairs / 100;
  const uint64_t n_placed_pairs = n_pairs - n_unplaced;
  // sort keys of the placed records: tid << 40 | pos << 8 | ..., then (pair, mate)
  struct Key { uint64_t key; uint64_t who; };
  std::vector<Key> keys((size_t)n_placed_pairs * 2);
  run_threads(nt, (size_t)n_placed_pairs, [&](size_t lo, size_t hi) {
    Mate m[2];
    for (size_t i = lo; i < hi; i++) {
      make_pair(seed, i, false, m);
      for (int w = 0; w < 2; w++) keys[2 * i + (size_t)w] = Key{((uint64_t)m[w].tid << 32) | (uint32_t)m[w].pos, (uint64_t)i * 2 + (uint64_t)w};
    }
  });
  std::sort(keys.begin(), keys.end(), [](const Key &a, const Key &b) { return a.key != b.key ? a.key < b.key : a.who < b.who; });
  const size_t n_rec = keys.size() + (size_t)n_unplaced * 2;
  std::vector<uint64_t> off(n_rec + 1);
  // header
  std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
  for (int c = 0; c < kContigs; c++) text += "@SQ\tSN:chr" + std::to_string(c + 1) + "\tLN:" + std::to_string(kContigLen) + "\n";
  std::string head("BAM\1", 4);
  auto app32 = [&](int32_t v) { head.append(reinterpret_cast<const char *>(&v), 4); };
  app32((int32_t)text.size());
  head += text;
  app32(kContigs);
  for (int c = 0; c < kContigs; c++) {
    const std::string nm = "chr" + std::to_string(c + 1);
    app32((int32_t)nm.size() + 1);
    head.append(nm.c_str(), nm.size() + 1);
    app32((int32_t)kContigLen);
  }
  // record sizes depend on the CIGAR only; recompute per record (cheap relative to deflate)
  std::vector<uint32_t> sizes(n_rec);
  run_threads(nt, n_rec, [&](size_t lo, size_t hi) {
    Mate m[2];
    for (size_t r = lo; r < hi; r++) {
      const bool un = r >= keys.size();
      const uint64_t who = un ? (uint64_t)(r - keys.size()) : keys[r].who;
      const uint64_t pair = un ? n_placed_pairs + who / 2 : who / 2;
      make_pair(seed, pair, un, m);
      sizes[r] = (uint32_t)record_bytes(m[who & 1]);
    }
  });
  off[0] = head.size();
  for (size_t r = 0; r < n_rec; r++) off[r + 1] = off[r] + sizes[r];
  const uint64_t total = off[n_rec];
  std::vector<uint8_t> raw((size_t)total);
  std::memcpy(raw.data(), head.data(), head.size());
  run_threads(nt, n_rec, [&](size_t lo, size_t hi) {
    Mate m[2];
    for (size_t r = lo; r < hi; r++) {
      const bool un = r >= keys.size();
      const uint64_t who = un ? (uint64_t)(r - keys.size()) : keys[r].who;
      const uint64_t pair = un ? n_placed_pairs + who / 2 : who / 2;
      make_pair(seed, pair, un, m);
      write_record(raw.data() + off[r], m[who & 1], pair, seed);
    }
  });
  // BGZF: blocks of 0xff00 input bytes, deflated in parallel
  constexpr size_t kBlock = 0xff00;
  const size_t n_blocks = (size_t)((total + kBlock - 1) / kBlock);
  std::vector<std::vector<uint8_t>> comp(n_blocks);
  std::string err;
  run_threads(nt, n_blocks, [&](size_t lo, size_t hi) {
    std::vector<uint8_t> buf(kBlock + 1024);
    for (size_t b = lo; b < hi; b++) {
      const size_t from = b * kBlock, len = (size_t)std::min<uint64_t>(kBlock, total - from);
      z_stream zs;
      std::memset(&zs, 0, sizeof(zs));
      if (deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY) != Z_OK) { err = "deflateInit2"; return; }
      zs.next_in = raw.data() + from;
      zs.avail_in = (uInt)len;
      zs.next_out = buf.data();
      zs.avail_out = (uInt)buf.size();
      const int rc = deflate(&zs, Z_FINISH);
      const size_t clen = buf.size() - zs.avail_out;
      deflateEnd(&zs);
      if (rc != Z_STREAM_END || clen + 26 > 65536) { err = "deflate"; return; }
      std::vector<uint8_t> &o = comp[b];
      o.resize(18 + clen + 8);
      static const uint8_t hdr[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 'B', 'C', 2, 0};
      std::memcpy(o.data(), hdr, 16);
      const uint16_t bsize = (uint16_t)(o.size() - 1);
      std::memcpy(o.data() + 16, &bsize, 2);
      std::memcpy(o.data() + 18, buf.data(), clen);
      const uint32_t crc = (uint32_t)crc32(crc32(0L, Z_NULL, 0), raw.data() + from, (uInt)len), isize = (uint32_t)len;
      std::memcpy(o.data() + 18 + clen, &crc, 4);
      std::memcpy(o.data() + 22 + clen, &isize, 4);
    }
  });
  if (!err.empty()) throw std::runtime_error("[strling] synth-bam: " + err + " failed");
  FILE *fh = std::fopen(path.c_str(), "wb");
  if (!fh) throw std::runtime_error("[strling] synth-bam: cannot write " + path);
  for (const auto &o : comp) std::fwrite(o.data(), 1, o.size(), fh);
  static const uint8_t eof_block[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  std::fwrite(eof_block, 1, 28, fh);
  std::fclose(fh);
  std::printf("{\"records\": %zu, \"raw_bytes\": %llu, \"blocks\": %zu}\n", n_rec, (unsigned long long)total, n_blocks);
  return 0;
}

}  // namespace strling
