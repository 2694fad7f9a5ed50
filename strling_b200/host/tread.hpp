// Host-side record type and the small per-pair arithmetic of `strling extract` that stays on the CPU because it is
// order dependent (extract.nim:51-61,134-190; utils.nim:37-83,139-146,291-310), plus the `.bin` codec
// (cluster.nim:38-50, unpack.nim:36-133, version.nim).  The per-read scan itself is NOT here: it only exists as CUDA.
#pragma once
#include <algorithm>
#include <array>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

namespace strling {

enum Soft : uint8_t { kLeft = 0, kRight = 1, kBoth = 2, kNone = 3, kNoneRight = 4, kNoneLeft = 5 };  // cluster.nim:14-20

struct TreadCore {  // cluster.nim:23-32 without the qname: what the pair arithmetic works on (20 bytes, trivially copyable)
  int32_t tid = 0;
  uint32_t position = 0;
  std::array<char, 6> repeat{{0, 0, 0, 0, 0, 0}};
  uint16_t flag = 0;
  uint8_t split = kNone;
  uint8_t mapping_quality = 0;
  uint8_t repeat_count = 0;
  uint8_t align_length = 0;
};
struct Tread : TreadCore {
  std::string qname;
  Tread() = default;
  Tread(const TreadCore &c, const char *name, size_t len) : TreadCore(c), qname(name, len) {}
};

struct Options {  // utils.nim:119-127 (fields used on this path)
  int median_fragment_length = 0;
  double proportion_repeat = 0.8;
  uint8_t min_mapq = 40;
};

inline int unit_length(const std::array<char, 6> &u) {
  int n = 0;
  while (n < 6 && u[(size_t)n] != 0) n++;
  return n;
}

// extract.nim:56-58 : the uint8 product wraps (checks are off in the release build)
inline double p_repeat(const TreadCore &t) {
  const uint8_t prod = (uint8_t)(t.repeat_count * (uint8_t)unit_length(t.repeat));
  return (double)prod / (double)std::max<uint8_t>(1, t.align_length);
}

// 2-bit order of the `kmer` package: C < A < T < G
inline int base_rank(char c) {
  switch (c) {
    case 'C': return 0;
    case 'T': return 2;
    case 'G': return 3;
    default: return 1;
  }
}
inline char complement(char c) {
  switch (c) {
    case 'A': return 'T';
    case 'T': return 'A';
    case 'C': return 'G';
    case 'G': return 'C';
    default: return c;
  }
}

// utils.nim:61-80 : reverse complement the unit, then take the rotation with the smallest 2-bit code
inline void min_rev_complement(std::array<char, 6> &u) {
  const int k = unit_length(u);
  if (k == 0) return;
  char rc[6];
  for (int i = 0; i < k; i++) rc[k - 1 - i] = complement(u[(size_t)i]);
  int best_rot = 0;
  uint32_t best = 0xffffffffu;
  for (int r = 0; r < k; r++) {
    uint32_t v = 0;
    for (int i = 0; i < k; i++) v = v * 4u + (uint32_t)base_rank(rc[(r + i) % k]);
    if (v < best) { best = v; best_rot = r; }
  }
  static const char alpha[4] = {'C', 'A', 'T', 'G'};
  for (int i = 0; i < k; i++) u[(size_t)i] = alpha[base_rank(rc[(best_rot + i) % k])];  // decode() spells CATG only
}

// utils.nim:304-310
inline std::array<char, 6> canonical_repeat(const std::array<char, 6> &u) {
  std::array<char, 6> r = u;
  min_rev_complement(r);
  for (size_t i = 0; i < 6; i++)
    if (i == 5 || r[i] != u[i]) return ((unsigned char)r[i] < (unsigned char)u[i]) ? r : u;
  return u;
}

inline uint32_t half_length(const TreadCore &t) { return (uint32_t)((double)t.align_length / 2.0 + 0.5); }

// extract.nim:141-179
inline bool adjust_by(TreadCore &A, const TreadCore &B, const Options &o, uint32_t B_position) {
  if (A.repeat_count == 0) return false;
  const bool a_proper = (A.flag & 0x2) != 0;
  if (B.mapping_quality > o.min_mapq &&
      ((p_repeat(A) > o.proportion_repeat && p_repeat(B) < 0.2) || (!a_proper && A.mapping_quality < o.min_mapq))) {
    if (B.flag & 0x10) {
      A.position = B_position - (uint32_t)o.median_fragment_length + (uint32_t)B.align_length + half_length(A);
      if (B.split == kNoneLeft) A.position = B_position;
    } else {
      A.position = B_position + (uint32_t)o.median_fragment_length - half_length(A);
      if (B.split == kNoneRight) A.position = B_position + (uint32_t)B.align_length;
    }
    A.split = kNone;
    A.tid = B.tid;
    A.mapping_quality = std::max(A.mapping_quality, B.mapping_quality);
    bool rev = !(A.flag & 0x20);          // should_reverse, extract.nim:134-139
    if (A.flag & 0x10) rev = !rev;
    if (rev) min_rev_complement(A.repeat);
  } else if (A.mapping_quality >= o.min_mapq || a_proper) {
    A.position += half_length(A);
    A.mapping_quality = std::max(A.mapping_quality, B.mapping_quality);
  }
  return true;
}

// extract.nim:182-190
inline bool unplaced_pair(const TreadCore &A, const TreadCore &B, const Options &o) {
  const double pa = p_repeat(A), pb = p_repeat(B);
  if (pa > o.proportion_repeat && pb > o.proportion_repeat) return true;
  if (pa > o.proportion_repeat && B.mapping_quality < o.min_mapq) return true;
  if (pb > o.proportion_repeat && A.mapping_quality < o.min_mapq) return true;
  return false;
}

// utils.nim:139-146
inline int frag_median(const std::array<uint32_t, 4096> &f, double pct = 0.5) {
  uint32_t n = 0;
  for (uint32_t c : f) n += c;
  const uint32_t target = (uint32_t)(0.5 + (double)n / (1.0 / pct));
  uint32_t count = 0;
  for (int i = 0; i < 4096; i++) {
    count += f[(size_t)i];
    if (count >= target) return i;
  }
  return 4096;
}

// ------------------------------------------------------------------------------------------------ .bin codec
constexpr const char *kStrlingVersion = "0.6.0";  // version.nim:1
constexpr int16_t kFmtVersion = 0;                // version.nim:4

struct MsgpackOut {  // the msgpack4nim subset pack_type uses: smallest-form ints, fixarray, str
  std::vector<uint8_t> &b;
  void u8(uint8_t v) { b.push_back(v); }
  void be(uint64_t v, int n) { for (int i = n - 1; i >= 0; i--) b.push_back((uint8_t)(v >> (8 * i))); }
  void uint_(uint64_t v) {
    if (v < 128) u8((uint8_t)v);
    else if (v <= 0xff) { u8(0xcc); be(v, 1); }
    else if (v <= 0xffff) { u8(0xcd); be(v, 2); }
    else if (v <= 0xffffffffull) { u8(0xce); be(v, 4); }
    else { u8(0xcf); be(v, 8); }
  }
  void int_(int64_t v) {
    if (v >= 0) { uint_((uint64_t)v); return; }
    if (v >= -32) u8((uint8_t)(int8_t)v);
    else if (v >= -128) { u8(0xd0); be((uint64_t)v, 1); }
    else if (v >= -32768) { u8(0xd1); be((uint64_t)v, 2); }
    else if (v >= -2147483648ll) { u8(0xd2); be((uint64_t)v, 4); }
    else { u8(0xd3); be((uint64_t)v, 8); }
  }
  void str(const std::string &s) {
    const size_t n = s.size();
    if (n < 32) u8((uint8_t)(0xa0 | n));
    else if (n <= 0xff) { u8(0xd9); be(n, 1); }
    else if (n <= 0xffff) { u8(0xda); be(n, 2); }
    else { u8(0xdb); be(n, 4); }
    b.insert(b.end(), s.begin(), s.end());
  }
};

inline void pack_tread(std::vector<uint8_t> &out, const Tread &t) {  // cluster.nim:38-50
  MsgpackOut m{out};
  m.int_(t.tid);
  m.uint_(t.position);
  m.u8(0x96);
  for (char c : t.repeat) m.uint_((uint8_t)c);
  m.uint_(t.flag);
  m.uint_(t.split);
  m.uint_(t.mapping_quality);
  m.uint_(t.repeat_count);
  m.uint_(t.align_length);
  m.uint_((uint32_t)t.qname.size());
  m.str(t.qname);
}

struct BinFile {
  float proportion_repeat = 0;
  uint8_t min_mapq = 0;
  std::array<uint32_t, 4096> frag_dist{};
  std::string header;
  std::vector<Tread> reads;
};

inline void write_bin(const std::string &path, const BinFile &bf) {  // extract.nim:331-348
  std::vector<uint8_t> out;
  out.insert(out.end(), {'S', 'T', 'R'});
  out.push_back((uint8_t)(kFmtVersion & 0xff));
  out.push_back((uint8_t)(kFmtVersion >> 8));
  char ver[9] = {0};
  std::strncpy(ver, kStrlingVersion, 9);
  out.insert(out.end(), ver, ver + 9);
  uint8_t f4[4];
  std::memcpy(f4, &bf.proportion_repeat, 4);
  out.insert(out.end(), f4, f4 + 4);
  out.push_back(bf.min_mapq);
  const uint8_t *fd = reinterpret_cast<const uint8_t *>(bf.frag_dist.data());
  out.insert(out.end(), fd, fd + 4096 * 4);
  const int32_t hl = (int32_t)bf.header.size();
  const uint8_t *hp = reinterpret_cast<const uint8_t *>(&hl);
  out.insert(out.end(), hp, hp + 4);
  out.insert(out.end(), bf.header.begin(), bf.header.end());
  const int32_t n = (int32_t)bf.reads.size();
  const uint8_t *np = reinterpret_cast<const uint8_t *>(&n);
  out.insert(out.end(), np, np + 4);
  for (const Tread &t : bf.reads) pack_tread(out, t);
  FILE *fh = std::fopen(path.c_str(), "wb");
  if (!fh) throw std::runtime_error("[strling] couldnt open binary output file");
  const bool ok = std::fwrite(out.data(), 1, out.size(), fh) == out.size();
  std::fclose(fh);
  if (!ok) throw std::runtime_error("[strling] short write on binary output file");
}

struct MsgpackIn {
  const uint8_t *p, *end;
  uint64_t be(int n) {
    if (end - p < n) throw std::runtime_error("[strling] truncated bin file");
    uint64_t v = 0;
    for (int i = 0; i < n; i++) v = (v << 8) | *p++;
    return v;
  }
  int64_t integer() {
    const uint8_t t = (uint8_t)be(1);
    if (t < 0x80) return t;
    if (t >= 0xe0) return (int8_t)t;
    switch (t) {
      case 0xcc: return (int64_t)be(1);
      case 0xcd: return (int64_t)be(2);
      case 0xce: return (int64_t)be(4);
      case 0xcf: return (int64_t)be(8);
      case 0xd0: return (int8_t)be(1);
      case 0xd1: return (int16_t)be(2);
      case 0xd2: return (int32_t)be(4);
      case 0xd3: return (int64_t)be(8);
      default: throw std::runtime_error("[strling] bin file: expected an integer");
    }
  }
  std::string bytes_like() {  // str / bin
    const uint8_t t = (uint8_t)be(1);
    size_t n;
    if ((t & 0xe0) == 0xa0) n = t & 0x1f;
    else if (t == 0xd9 || t == 0xc4) n = (size_t)be(1);
    else if (t == 0xda || t == 0xc5) n = (size_t)be(2);
    else if (t == 0xdb || t == 0xc6) n = (size_t)be(4);
    else throw std::runtime_error("[strling] bin file: expected a string");
    if ((size_t)(end - p) < n) throw std::runtime_error("[strling] truncated bin file");
    std::string s(reinterpret_cast<const char *>(p), n);
    p += n;
    return s;
  }
  std::array<char, 6> unit() {  // fixarray of six uint8; also accept str/bin of length 6
    std::array<char, 6> u{{0, 0, 0, 0, 0, 0}};
    if (p < end && *p == 0x96) {
      p++;
      for (auto &c : u) c = (char)integer();
    } else {
      const std::string s = bytes_like();
      if (s.size() != 6) throw std::runtime_error("[strling] bin file: bad repeat field");
      std::memcpy(u.data(), s.data(), 6);
    }
    return u;
  }
};

inline BinFile read_bin(const std::string &path) {  // unpack.nim:58-133
  FILE *fh = std::fopen(path.c_str(), "rb");
  if (!fh) throw std::runtime_error("[strling] unable to open " + path + " for reading. please check path");
  std::vector<uint8_t> d;
  uint8_t buf[1 << 16];
  size_t got;
  while ((got = std::fread(buf, 1, sizeof(buf), fh)) > 0) d.insert(d.end(), buf, buf + got);
  std::fclose(fh);
  BinFile bf;
  const size_t fixed = 3 + 2 + 9 + 4 + 1 + 16384 + 4;
  if (d.size() < fixed || std::memcmp(d.data(), "STR", 3) != 0)
    throw std::runtime_error("[strling] expected bin file to start with \"STR\". This may indicate that this bin file was generated by an old version of STRling. Please re-run the extract step with this version.");
  int16_t fmt;
  std::memcpy(&fmt, d.data() + 3, 2);
  if (fmt != kFmtVersion) throw std::runtime_error("[strling] this bin file was generated using a different format. Please re-run the extract step with the same version of STRling.");
  std::memcpy(&bf.proportion_repeat, d.data() + 14, 4);
  bf.min_mapq = d[18];
  std::memcpy(bf.frag_dist.data(), d.data() + 19, 16384);
  int32_t hl;
  std::memcpy(&hl, d.data() + 19 + 16384, 4);
  size_t off = fixed;
  if (hl < 0 || d.size() < off + (size_t)hl + 4) throw std::runtime_error("[strling] truncated bin file");
  bf.header.assign(reinterpret_cast<const char *>(d.data() + off), (size_t)hl);
  off += (size_t)hl;
  int32_t n;
  std::memcpy(&n, d.data() + off, 4);
  off += 4;
  MsgpackIn in{d.data() + off, d.data() + d.size()};
  while (in.p < in.end) {
    Tread t;
    t.tid = (int32_t)in.integer();
    t.position = (uint32_t)in.integer();
    t.repeat = in.unit();
    t.flag = (uint16_t)in.integer();
    t.split = (uint8_t)in.integer();
    t.mapping_quality = (uint8_t)in.integer();
    t.repeat_count = (uint8_t)in.integer();
    t.align_length = (uint8_t)in.integer();
    const uint32_t L = (uint32_t)in.integer();
    if (L > 0) t.qname = in.bytes_like();   // unpack.nim:123-124 reads the string only when L > 0
    bf.reads.push_back(std::move(t));
  }
  if ((int64_t)bf.reads.size() != (int64_t)n)
    throw std::runtime_error("[strling] expected " + std::to_string(n) + " got " + std::to_string(bf.reads.size()));
  return bf;
}

inline std::vector<std::pair<std::string, uint32_t>> targets_from_header(const std::string &h) {
  std::vector<std::pair<std::string, uint32_t>> t;
  size_t i = 0;
  while (i < h.size()) {
    size_t e = h.find('\n', i);
    if (e == std::string::npos) e = h.size();
    const std::string line = h.substr(i, e - i);
    if (line.rfind("@SQ", 0) == 0) {
      std::string name;
      uint32_t len = 0;
      size_t j = 0;
      while (j < line.size()) {
        size_t f = line.find('\t', j);
        if (f == std::string::npos) f = line.size();
        const std::string fld = line.substr(j, f - j);
        if (fld.rfind("SN:", 0) == 0) name = fld.substr(3);
        if (fld.rfind("LN:", 0) == 0) len = (uint32_t)std::stoul(fld.substr(3));
        j = f + 1;
      }
      t.emplace_back(name, len);
    }
    i = e + 1;
  }
  return t;
}

}  // namespace strling
