// BGZF + BAM reader for the host side of `strling extract` / `strling call` (stands in for htslib, which the
// reference reaches through hts-nim: extract.nim:275-329).  Blocks are inflated in parallel batches with the repo's own
// whole-block DEFLATE decoder (inflate_fast.hpp; STRLING_ZLIB=1 switches back to zlib's inflate for A/B runs).
#pragma once
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <zlib.h>

#include <chrono>
#include <cstdlib>
#include <functional>
#include <memory>

#include "inflate_fast.hpp"
#include "pool.hpp"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace strling {

struct Target {
  std::string name;
  uint32_t length = 0;
};

// A view into the reader's decode buffer; valid until the next call to next().
struct BamRecord {
  int32_t tid, pos, mate_tid, mate_pos, isize;
  uint16_t flag, n_cigar;
  uint8_t mapq;
  int32_t l_seq;
  const char *qname;       // NUL terminated
  uint32_t l_qname;        // without the NUL
  const uint32_t *cigar;   // len << 4 | op  (may be unaligned: use cigar_at)
  const uint8_t *seq;      // 4-bit packed
  uint64_t voffset;        // BGZF virtual offset of the record start (coffset << 16 | uoffset)

  uint32_t cigar_at(int i) const {
    uint32_t v;
    std::memcpy(&v, reinterpret_cast<const uint8_t *>(cigar) + 4 * i, 4);
    return v;
  }
  static int op(uint32_t c) { return (int)(c & 15u); }
  static uint32_t oplen(uint32_t c) { return c >> 4; }
  // htslib bam_endpos: pos + reference length; unmapped or zero-length alignments count as length 1
  int32_t stop() const {
    int64_t rl = 0;
    if (!(flag & 4))
      for (int i = 0; i < n_cigar; i++) {
        const uint32_t c = cigar_at(i);
        const int o = op(c);
        if (o == 0 || o == 2 || o == 3 || o == 7 || o == 8) rl += oplen(c);
      }
    if (rl == 0) rl = 1;
    return (int32_t)(pos + rl);
  }
};

// inflate_host.cpp: the decoder compiled for the best ISA level this CPU has
int inflate_block_host(infl::Tables &T, const uint8_t *in, uint32_t in_len, uint8_t *out, uint32_t out_len);

// one BGZF block: `in` is followed by the block's 8-byte footer (the decoder may read into it)
inline void inflate_bgzf_block(const uint8_t *in, uint32_t csize, uint8_t *out, uint32_t isize) {
  static const bool use_zlib = std::getenv("STRLING_ZLIB") != nullptr;
  if (isize == 0) return;
  if (!use_zlib) {
    static thread_local std::unique_ptr<infl::Tables> tables;
    if (!tables) { tables.reset(new infl::Tables); tables->fixed_built = false; }
    const int rc = inflate_block_host(*tables, in, csize, out, isize);
    if (rc != infl::kOk) throw std::runtime_error("BGZF: inflate failed (" + std::to_string(rc) + ")");
    return;
  }
  z_stream zs;
  std::memset(&zs, 0, sizeof(zs));
  if (inflateInit2(&zs, -15) != Z_OK) throw std::runtime_error("BGZF: inflateInit2");
  zs.next_in = const_cast<uint8_t *>(in);
  zs.avail_in = csize;
  zs.next_out = out;
  zs.avail_out = isize;
  const int rc = inflate(&zs, Z_FINISH);
  inflateEnd(&zs);
  if (rc != Z_STREAM_END || zs.avail_out != 0) throw std::runtime_error("BGZF: inflate failed");
}

class BamReader {
 public:
  explicit BamReader(const std::string &path, int threads = 0) : path_(path) {
    fh_ = std::fopen(path.c_str(), "rb");
    if (!fh_) throw std::runtime_error("couldn't open bam");
    threads_ = threads > 0 ? threads : (int)std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    read_header();
  }
  ~BamReader() {
    if (fh_) std::fclose(fh_);
  }
  BamReader(const BamReader &) = delete;
  BamReader &operator=(const BamReader &) = delete;

  const std::string &header_text() const { return header_text_; }
  const std::vector<Target> &targets() const { return targets_; }

  // next record, or false at EOF
  bool next(BamRecord &r) {
    if (!ensure(4)) return false;
    int32_t block_size;
    std::memcpy(&block_size, data_.data() + cur_, 4);
    if (block_size < 32) throw std::runtime_error("corrupt BAM record");
    if (!ensure(4 + (size_t)block_size)) throw std::runtime_error("truncated BAM record");
    r.voffset = voffset_at(cur_);
    const uint8_t *p = data_.data() + cur_ + 4;
    uint8_t l_read_name;
    uint16_t bin;
    std::memcpy(&r.tid, p, 4);
    std::memcpy(&r.pos, p + 4, 4);
    l_read_name = p[8];
    r.mapq = p[9];
    std::memcpy(&bin, p + 10, 2);
    std::memcpy(&r.n_cigar, p + 12, 2);
    std::memcpy(&r.flag, p + 14, 2);
    std::memcpy(&r.l_seq, p + 16, 4);
    std::memcpy(&r.mate_tid, p + 20, 4);
    std::memcpy(&r.mate_pos, p + 24, 4);
    std::memcpy(&r.isize, p + 28, 4);
    r.qname = reinterpret_cast<const char *>(p + 32);
    r.l_qname = l_read_name ? (uint32_t)l_read_name - 1 : 0;
    r.cigar = reinterpret_cast<const uint32_t *>(p + 32 + l_read_name);
    r.seq = p + 32 + l_read_name + 4 * (size_t)r.n_cigar;
    if (32 + (size_t)l_read_name + 4 * (size_t)r.n_cigar + (size_t)(r.l_seq + 1) / 2 > (size_t)block_size)
      throw std::runtime_error("corrupt BAM record (field lengths)");
    cur_ += 4 + (size_t)block_size;
    return true;
  }

  // BGZF virtual offset of the next record (valid right after construction: the first record)
  uint64_t tell() {
    if (!ensure(1)) return (fpos_ << 16);
    return voffset_at(cur_);
  }

  // restart iteration at a BGZF virtual offset previously taken from BamRecord::voffset
  void seek(uint64_t voffset) {
    data_.clear();
    spans_.clear();
    cur_ = 0;
    eof_ = false;
    fill_blocks_ = 16;
    if (std::fseek(fh_, (long)(voffset >> 16), SEEK_SET) != 0) throw std::runtime_error("seek failed");
    fpos_ = voffset >> 16;
    fill();
    cur_ = (size_t)(voffset & 0xffff);
  }

 private:
  struct Span {       // one inflated BGZF block inside data_
    size_t begin;     // offset in data_
    uint64_t coffset; // file offset of the compressed block
  };

  uint64_t voffset_at(size_t pos) const {
    // binary search the block that holds pos
    size_t lo = 0, hi = spans_.size();
    while (hi - lo > 1) {
      const size_t mid = (lo + hi) / 2;
      if (spans_[mid].begin <= pos) lo = mid; else hi = mid;
    }
    return (spans_[lo].coffset << 16) | (uint64_t)(pos - spans_[lo].begin);
  }

  bool ensure(size_t need) {
    while (data_.size() - cur_ < need) {
      if (eof_) return false;
      fill();
    }
    return true;
  }

  // read up to kBatch compressed blocks and inflate them in parallel, appending to data_
  void fill() {
    // drop consumed bytes
    if (cur_ > 0) {
      size_t keep_from = cur_;
      // keep whole spans so voffsets stay right: find the span containing cur_
      size_t si = 0;
      while (si + 1 < spans_.size() && spans_[si + 1].begin <= keep_from) si++;
      const size_t cut = spans_.empty() ? 0 : spans_[si].begin;
      if (cut > 0) {
        data_.erase(data_.begin(), data_.begin() + (long)cut);
        cur_ -= cut;
        std::vector<Span> ns;
        for (size_t i = si; i < spans_.size(); i++) ns.push_back(Span{spans_[i].begin - cut, spans_[i].coffset});
        spans_.swap(ns);
      }
    }
    struct Blk { size_t coff; uint32_t csize, isize; uint64_t fpos; };
    std::vector<Blk> blks;
    comp_.clear();
    size_t total_out = 0;
    const size_t kBatch = fill_blocks_;   // small at first (the header is usually all a caller wants), then up to 1024
    fill_blocks_ = std::min<size_t>(fill_blocks_ * 4, 1024);
    while (blks.size() < kBatch) {
      uint8_t hdr[18];
      const size_t got = std::fread(hdr, 1, 18, fh_);
      if (got == 0) { eof_ = true; break; }
      if (got < 18 || hdr[0] != 31 || hdr[1] != 139 || hdr[2] != 8 || !(hdr[3] & 4)) throw std::runtime_error("not a BGZF file");
      uint16_t xlen;
      std::memcpy(&xlen, hdr + 10, 2);
      // locate the BC subfield (usually first)
      std::vector<uint8_t> extra(xlen);
      std::memcpy(extra.data(), hdr + 12, std::min<size_t>(6, xlen));
      if (xlen > 6 && std::fread(extra.data() + 6, 1, xlen - 6, fh_) != (size_t)xlen - 6) throw std::runtime_error("truncated BGZF header");
      uint32_t bsize = 0;
      for (size_t o = 0; o + 4 <= extra.size();) {
        uint16_t slen;
        std::memcpy(&slen, extra.data() + o + 2, 2);
        if (extra[o] == 'B' && extra[o + 1] == 'C' && slen == 2) {
          uint16_t bs;
          std::memcpy(&bs, extra.data() + o + 4, 2);
          bsize = (uint32_t)bs + 1;
          break;
        }
        o += 4 + slen;
      }
      if (!bsize) throw std::runtime_error("BGZF block without BC field");
      const uint32_t csize = bsize - xlen - 12 - 8;  // deflate payload
      const size_t coff = comp_.size();
      comp_.resize(coff + csize + 8);
      if (std::fread(comp_.data() + coff, 1, csize + 8, fh_) != csize + 8) throw std::runtime_error("truncated BGZF block");
      uint32_t isize;
      std::memcpy(&isize, comp_.data() + coff + csize + 4, 4);
      blks.push_back(Blk{coff, csize, isize, fpos_});
      fpos_ += bsize;
      total_out += isize;
    }
    if (blks.empty()) return;
    const size_t base = data_.size();
    data_.resize(base + total_out);
    std::vector<size_t> outoff(blks.size());
    size_t o = base;
    for (size_t i = 0; i < blks.size(); i++) {
      outoff[i] = o;
      if (blks[i].isize) spans_.push_back(Span{o, blks[i].fpos});
      o += blks[i].isize;
    }
    if (spans_.empty()) spans_.push_back(Span{base, blks[0].fpos});
    auto work = [&](size_t from, size_t to, std::string *err) {
      try {
        for (size_t i = from; i < to; i++) inflate_bgzf_block(comp_.data() + blks[i].coff, blks[i].csize, data_.data() + outoff[i], blks[i].isize);
      } catch (const std::exception &e) {
        *err = e.what();
      }
    };
    const int nt = (int)std::min<size_t>((size_t)threads_, (blks.size() + 15) / 16);
    std::vector<std::string> errs((size_t)std::max(nt, 1));
    if (nt <= 1) {
      work(0, blks.size(), &errs[0]);
    } else {
      std::vector<std::thread> th;
      const size_t per = (blks.size() + (size_t)nt - 1) / (size_t)nt;
      for (int t = 0; t < nt; t++) {
        const size_t a = (size_t)t * per, b = std::min(blks.size(), a + per);
        if (a >= b) break;
        th.emplace_back(work, a, b, &errs[(size_t)t]);
      }
      for (auto &x : th) x.join();
    }
    for (auto &e : errs)
      if (!e.empty()) throw std::runtime_error(e);
  }

  void read_raw(void *dst, size_t n) {
    if (!ensure(n)) throw std::runtime_error("truncated BAM header");
    std::memcpy(dst, data_.data() + cur_, n);
    cur_ += n;
  }

  void read_header() {
    fpos_ = 0;
    char magic[4];
    read_raw(magic, 4);
    if (std::memcmp(magic, "BAM\1", 4) != 0) throw std::runtime_error("not a BAM file (CRAM/SAM input is not supported)");
    int32_t l_text;
    read_raw(&l_text, 4);
    header_text_.resize((size_t)l_text);
    if (l_text) read_raw(&header_text_[0], (size_t)l_text);
    // htslib keeps the text up to its first NUL
    const size_t z = header_text_.find('\0');
    if (z != std::string::npos) header_text_.resize(z);
    int32_t n_ref;
    read_raw(&n_ref, 4);
    targets_.resize((size_t)n_ref);
    for (auto &t : targets_) {
      int32_t l_name;
      read_raw(&l_name, 4);
      std::string nm((size_t)l_name, '\0');
      read_raw(&nm[0], (size_t)l_name);
      if (!nm.empty() && nm.back() == '\0') nm.pop_back();
      t.name = nm;
      int32_t ln;
      read_raw(&ln, 4);
      t.length = (uint32_t)ln;
    }
  }

  std::string path_;
  FILE *fh_ = nullptr;
  int threads_ = 1;
  uint64_t fpos_ = 0;
  bool eof_ = false;
  std::vector<uint8_t> data_, comp_;
  size_t fill_blocks_ = 16;
  std::vector<Span> spans_;
  size_t cur_ = 0;
  std::string header_text_;
  std::vector<Target> targets_;
};


// ------------------------------------------------------------------------------------------------
// Chunked reader for the extract pipeline: hands out large buffers of WHOLE decoded records (the file is memory
// mapped, BGZF blocks are inflated in parallel straight from the mapping) together with the record offsets, so that
// staging can be spread over threads and the buffer's ownership can travel with the batch (record fields such as
// qname are then used in place, never copied).
struct RawBuffer {  // uninitialised, growable byte buffer (std::vector would zero-fill hundreds of MB per chunk)
  uint8_t *p = nullptr;
  size_t size = 0, cap = 0;
  // optional allocator (pinned host memory when the GPU inflates into the buffer); contents are NOT kept across a growth
  void *(*alloc_fn)(size_t) = nullptr;
  void (*free_fn)(void *) = nullptr;
  RawBuffer() = default;
  RawBuffer(const RawBuffer &) = delete;
  RawBuffer &operator=(const RawBuffer &) = delete;
  ~RawBuffer() { release(); }
  void release() {
    if (p) { if (free_fn) free_fn(p); else std::free(p); }
    p = nullptr;
    size = cap = 0;
  }
  void set_allocator(void *(*a)(size_t), void (*f)(void *)) { release(); alloc_fn = a; free_fn = f; }
  void resize(size_t n) {
    if (n > cap) {
      const size_t nc = n + n / 8 + 4096;
      uint8_t *q = static_cast<uint8_t *>(alloc_fn ? alloc_fn(nc) : std::malloc(nc));
      if (!q) throw std::runtime_error("out of memory");
      if (p) { if (free_fn) free_fn(p); else std::free(p); }
      p = q;
      cap = nc;
    }
    size = n;
  }
  uint8_t *data() { return p; }
  const uint8_t *data() const { return p; }
};

// a BGZF block of a chunk: the same 24-byte layout as strgpu_bgzf_block (include/strgpu.h)
struct BgzfBlockRef {
  uint64_t in_off;   // offset of the raw DEFLATE payload in the file
  uint32_t csize, isize;
  uint64_t out_off;  // offset of the inflated bytes in the chunk's buffer
};

struct BamChunk {
  RawBuffer data;                   // decoded bytes; records [rec_off[i], rec_off[i+1])
  std::vector<uint32_t> rec_off;    // n_records + 1 entries
  struct Span { uint32_t begin; uint64_t coffset; uint32_t skip; };  // data[begin..] came from block coffset, starting `skip` bytes into it
  std::vector<Span> spans;
  size_t n_records() const { return rec_off.empty() ? 0 : rec_off.size() - 1; }
  uint64_t voffset_of(size_t rec) const {
    const uint32_t pos = rec_off[rec];
    size_t lo = 0, hi = spans.size();
    while (hi - lo > 1) {
      const size_t mid = (lo + hi) / 2;
      if (spans[mid].begin <= pos) lo = mid; else hi = mid;
    }
    return (spans[lo].coffset << 16) | (uint64_t)(pos - spans[lo].begin + spans[lo].skip);
  }
  static BamRecord view(const uint8_t *rec) {
    BamRecord r;
    const uint8_t *p = rec + 4;
    std::memcpy(&r.tid, p, 4);
    std::memcpy(&r.pos, p + 4, 4);
    const uint8_t l_read_name = p[8];
    r.mapq = p[9];
    std::memcpy(&r.n_cigar, p + 12, 2);
    std::memcpy(&r.flag, p + 14, 2);
    std::memcpy(&r.l_seq, p + 16, 4);
    std::memcpy(&r.mate_tid, p + 20, 4);
    std::memcpy(&r.mate_pos, p + 24, 4);
    std::memcpy(&r.isize, p + 28, 4);
    r.qname = reinterpret_cast<const char *>(p + 32);
    r.l_qname = l_read_name ? (uint32_t)l_read_name - 1 : 0;
    r.cigar = reinterpret_cast<const uint32_t *>(p + 32 + l_read_name);
    r.seq = p + 32 + l_read_name + 4 * (size_t)r.n_cigar;
    r.voffset = 0;
    return r;
  }
};

class BamChunkReader {
 public:
  // `pool`: the threads that inflate and walk (nullptr: a private pool of `threads`); `n_ref`: number of reference
  // sequences in the header (only sharpens the record-start guesses of the parallel walk; any value is correct)
  BamChunkReader(const std::string &path, uint64_t start_voffset, int threads, Pool *pool = nullptr, int32_t n_ref = INT32_MAX) : n_ref_(n_ref) {
    fd_ = ::open(path.c_str(), O_RDONLY);
    if (fd_ < 0) throw std::runtime_error("couldn't open bam");
    struct stat st;
    if (::fstat(fd_, &st) != 0) throw std::runtime_error("couldn't stat bam");
    file_size_ = (size_t)st.st_size;
    if (file_size_) {
      map_ = static_cast<const uint8_t *>(::mmap(nullptr, file_size_, PROT_READ, MAP_PRIVATE, fd_, 0));
      if (map_ == MAP_FAILED) throw std::runtime_error("couldn't mmap bam");
      ::madvise(const_cast<uint8_t *>(map_), file_size_, MADV_SEQUENTIAL);
    }
    if (pool) {
      pool_ = pool;
    } else {
      own_pool_.reset(new Pool(threads > 0 ? threads : (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()))));
      pool_ = own_pool_.get();
    }
    fpos_ = start_voffset >> 16;
    first_skip_ = (uint32_t)(start_voffset & 0xffff);
  }
  ~BamChunkReader() {
    if (map_ && map_ != MAP_FAILED) ::munmap(const_cast<uint8_t *>(map_), file_size_);
    if (fd_ >= 0) ::close(fd_);
  }
  BamChunkReader(const BamChunkReader &) = delete;
  BamChunkReader &operator=(const BamChunkReader &) = delete;

  // When set, inflates the blocks of a chunk instead of the host threads (`strling extract --gpu-inflate`): file = the
  // mapped BAM, every block's isize bytes go to out + out_off; must throw on failure.
  std::function<void(const uint8_t *file, size_t file_size, const BgzfBlockRef *blocks, size_t n_blocks, uint8_t *out, size_t out_bytes)> inflate_hook;

  double t_read = 0, t_alloc = 0, t_inflate = 0, t_walk = 0;  // seconds spent per phase (diagnostics)
  size_t n_rewalked = 0;                                      // parts of the parallel walk whose guessed start was wrong

  // Fills `c` with the next whole records (about max_blocks BGZF blocks); false at EOF.
  bool next(BamChunk &c, size_t max_blocks) {
    using clk = std::chrono::steady_clock;
    const auto q0 = clk::now();
    c.rec_off.clear();
    c.spans.clear();
    struct Blk { uint64_t in_off; uint32_t csize, isize; uint64_t fpos; };
    std::vector<Blk> blks;
    blks.reserve(max_blocks);
    size_t total_out = 0;
    while (blks.size() < max_blocks && fpos_ < file_size_) {
      if (fpos_ + 18 > file_size_) throw std::runtime_error("truncated BGZF header");
      const uint8_t *hdr = map_ + fpos_;
      if (hdr[0] != 31 || hdr[1] != 139 || hdr[2] != 8 || !(hdr[3] & 4)) throw std::runtime_error("not a BGZF file");
      uint16_t xlen;
      std::memcpy(&xlen, hdr + 10, 2);
      if (fpos_ + 12 + xlen > file_size_) throw std::runtime_error("truncated BGZF header");
      uint32_t bsize = 0;
      for (size_t o = 0; o + 4 <= xlen;) {
        const uint8_t *e = hdr + 12 + o;
        uint16_t slen;
        std::memcpy(&slen, e + 2, 2);
        if (e[0] == 'B' && e[1] == 'C' && slen == 2) {
          uint16_t bs;
          std::memcpy(&bs, e + 4, 2);
          bsize = (uint32_t)bs + 1;
          break;
        }
        o += 4 + slen;
      }
      if (!bsize || bsize < (uint32_t)xlen + 20) throw std::runtime_error("BGZF block without BC field");
      if (fpos_ + bsize > file_size_) throw std::runtime_error("truncated BGZF block");
      const uint32_t csize = bsize - xlen - 12 - 8;
      uint32_t isize;
      std::memcpy(&isize, hdr + bsize - 4, 4);
      blks.push_back(Blk{fpos_ + 12 + xlen, csize, isize, fpos_});
      fpos_ += bsize;
      total_out += isize;
    }
    const bool eof = fpos_ >= file_size_;
    const auto q1 = clk::now();
    // bytes of a record that straddled the previous chunk's end come first
    const size_t base = carry_.size();
    if (base + total_out > 0xfffffff0ull) throw std::runtime_error("BAM chunk too large");
    c.data.resize(base + total_out);
    if (base) std::memcpy(c.data.data(), carry_.data(), base);
    for (const auto &sp : carry_spans_) c.spans.push_back(sp);
    carry_.clear();
    carry_spans_.clear();
    std::vector<size_t> outoff(blks.size());
    size_t o = base;
    for (size_t i = 0; i < blks.size(); i++) {
      outoff[i] = o;
      o += blks[i].isize;
    }
    const auto q2 = clk::now();
    uint8_t *out = c.data.data();
    if (inflate_hook && !blks.empty()) {
      std::vector<BgzfBlockRef> refs(blks.size());
      for (size_t i = 0; i < blks.size(); i++) refs[i] = BgzfBlockRef{blks[i].in_off, blks[i].csize, blks[i].isize, (uint64_t)outoff[i]};
      inflate_hook(map_, file_size_, refs.data(), refs.size(), out, c.data.size);
    } else {
      pool_->run(blks.size(), [&](size_t i) { inflate_bgzf_block(map_ + blks[i].in_off, blks[i].csize, out + outoff[i], blks[i].isize); });
    }
    const auto q3 = clk::now();
    for (size_t i = 0; i < blks.size(); i++)
      if (blks[i].isize) c.spans.push_back(BamChunk::Span{(uint32_t)outoff[i], blks[i].fpos, 0});
    size_t n = c.data.size;
    // a stream opened at a virtual offset starts mid-block: drop the bytes before it
    if (first_skip_ && !blks.empty()) {
      const size_t drop = first_skip_;
      if (drop > blks[0].isize) throw std::runtime_error("bad virtual offset");
      std::memmove(out + base, out + base + drop, n - base - drop);
      n -= drop;
      bool first = true;
      for (auto &sp : c.spans) {
        if (sp.begin < base) continue;
        if (first) { sp.skip = (uint32_t)drop; first = false; }
        else sp.begin -= (uint32_t)drop;
      }
      first_skip_ = 0;
    }
    const size_t pos = walk_records(out, n, c.rec_off);
    if (pos < n) {
      if (eof && blks.empty()) throw std::runtime_error("truncated BAM record at end of file");
      carry_.assign(out + pos, out + n);
      for (size_t i = 0; i < c.spans.size(); i++) {  // spans that cover the carried bytes, rebased to the carry buffer
        const size_t sb = c.spans[i].begin;
        const size_t se = (i + 1 < c.spans.size()) ? c.spans[i + 1].begin : n;
        if (se <= pos) continue;
        BamChunk::Span sp = c.spans[i];
        if (sb < pos) { sp.skip += (uint32_t)(pos - sb); sp.begin = 0; }
        else sp.begin = (uint32_t)(sb - pos);
        carry_spans_.push_back(sp);
      }
    }
    c.data.size = pos;
    const auto q4 = clk::now();
    t_read += std::chrono::duration<double>(q1 - q0).count();
    t_alloc += std::chrono::duration<double>(q2 - q1).count();
    t_inflate += std::chrono::duration<double>(q3 - q2).count();
    t_walk += std::chrono::duration<double>(q4 - q3).count();
    return c.n_records() > 0 || !eof || !carry_.empty();
  }

  // ---- record boundaries.  A BAM record only says where the NEXT one starts, so the walk is a chain through memory the
  // inflate threads have just written.  It is cut into parts: every part but the first GUESSES its first record start
  // (the first offset at or after the part's beginning from which kChain records in a row look like BAM records) and walks
  // from there; afterwards the parts are checked in order -- part p is accepted only if it started exactly where the
  // (already verified) chain through part p-1 arrived, otherwise it is walked again from that position.  data[0] is a
  // record start by construction, so by induction the result is the one chain from data[0]: exact, whatever the guesses were.
  struct WalkPart {
    size_t start = 0, landing = 0;     // first record walked / first record start at or after the part's end (or the incomplete tail)
    bool ok = false;
    std::vector<uint32_t> offs;
  };
  static constexpr int kChain = 4;

  bool plausible(const uint8_t *d, size_t n, size_t c, size_t *next) const {
    if (c + 36 > n) return false;
    int32_t bs, tid, pos, l_seq, mtid, mpos;
    uint16_t n_cig;
    std::memcpy(&bs, d + c, 4);
    std::memcpy(&tid, d + c + 4, 4);
    std::memcpy(&pos, d + c + 8, 4);
    const uint32_t l_name = d[c + 12];
    std::memcpy(&n_cig, d + c + 16, 2);
    std::memcpy(&l_seq, d + c + 20, 4);
    std::memcpy(&mtid, d + c + 24, 4);
    std::memcpy(&mpos, d + c + 28, 4);
    if (bs < 32 || bs > (1 << 24)) return false;
    if (tid < -1 || tid >= n_ref_ || mtid < -1 || mtid >= n_ref_ || pos < -1 || mpos < -1 || l_name < 1 || l_seq < 0) return false;
    const uint64_t need = 32ull + l_name + 4ull * n_cig + ((uint64_t)l_seq + 1) / 2 + (uint64_t)l_seq;
    if (need > (uint64_t)bs) return false;
    const size_t nul = c + 36 + l_name - 1;
    if (nul < n && d[nul] != 0) return false;
    *next = c + 4 + (size_t)bs;
    return true;
  }

  // appends the complete records that start in [pos, limit); returns the first record start >= limit, or the position of the
  // incomplete tail / of the first implausible length (then *bad is set, when bad is given; otherwise that is an error)
  static size_t walk_range(const uint8_t *d, size_t n, size_t pos, size_t limit, std::vector<uint32_t> &offs, bool *bad) {
    while (pos < limit && pos + 4 <= n) {
      int32_t block_size;
      std::memcpy(&block_size, d + pos, 4);
      if (block_size < 32) {
        if (bad) { *bad = true; return pos; }
        throw std::runtime_error("corrupt BAM record");
      }
      if (pos + 4 + (size_t)block_size > n) break;
      offs.push_back((uint32_t)pos);
      pos += 4 + (size_t)block_size;
      __builtin_prefetch(d + pos + 1024);
    }
    return pos;
  }

  // fills rec_off (n_records + 1 entries) and returns the end of the last complete record
  size_t walk_records(const uint8_t *d, size_t n, std::vector<uint32_t> &rec_off) {
    size_t parts = (size_t)pool_->size() * 2;
    if (n < (1u << 22) || parts < 2) parts = 1;
    std::vector<WalkPart> wp(parts);
    const size_t per = (n + parts - 1) / parts;
    pool_->run(parts, [&](size_t p) {
      WalkPart &w = wp[p];
      const size_t begin = p * per, limit = std::min(n, begin + per);
      w.offs.reserve(per / 160 + 16);
      if (p == 0) {
        w.start = 0;
        w.landing = walk_range(d, n, 0, limit, w.offs, nullptr);
        w.ok = true;
        return;
      }
      for (size_t c = begin; c < limit; c++) {
        size_t q = c, nx = 0;
        int k = 0;
        while (k < kChain && plausible(d, n, q, &nx)) { q = nx; k++; }
        if (k < kChain && !(k > 0 && q + 36 > n)) continue;  // a chain that runs into the end of the data counts as well
        bool bad = false;
        w.start = c;
        w.landing = walk_range(d, n, c, limit, w.offs, &bad);
        w.ok = !bad;
        return;
      }
      w.ok = false;
    });
    // verification in chain order (serial, one comparison per part; a wrong guess costs one re-walk of that part)
    size_t at = wp[0].landing;
    for (size_t p = 1; p < parts; p++) {
      WalkPart &w = wp[p];
      const size_t begin = p * per, limit = std::min(n, begin + per);
      (void)begin;
      if (at >= limit) { w.offs.clear(); w.landing = at; continue; }   // the chain jumps over this part entirely
      if (!(w.ok && w.start == at)) {
        n_rewalked++;
        w.offs.clear();
        w.landing = walk_range(d, n, at, limit, w.offs, nullptr);
      }
      at = w.landing;
    }
    size_t total = 0;
    std::vector<size_t> first(parts);
    for (size_t p = 0; p < parts; p++) { first[p] = total; total += wp[p].offs.size(); }
    rec_off.resize(total + 1);
    pool_->run(parts, [&](size_t p) {
      if (!wp[p].offs.empty()) std::memcpy(rec_off.data() + first[p], wp[p].offs.data(), wp[p].offs.size() * sizeof(uint32_t));
    });
    rec_off[total] = (uint32_t)at;
    return at;
  }

 private:
  int fd_ = -1;
  const uint8_t *map_ = nullptr;
  size_t file_size_ = 0;
  Pool *pool_ = nullptr;
  std::unique_ptr<Pool> own_pool_;
  int32_t n_ref_ = INT32_MAX;
  uint64_t fpos_ = 0;
  uint32_t first_skip_ = 0;
  std::vector<uint8_t> carry_;
  std::vector<BamChunk::Span> carry_spans_;
};

}  // namespace strling
