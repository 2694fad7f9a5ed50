// `strling extract` (extract.nim:250-350) with the per-read repeat-unit scan on the GPU.
//
// The scan (get_repeat, utils.nim:236) is a pure function of (bases, length, proportion_repeat), so every
// candidate segment of a batch of records -- the whole read (unless the genome-STR filter of extract.nim:30-34 skips
// it) and each soft-clipped end under both proportion classes add() may use (extract.nim:207-211,241-244) -- is
// submitted to libstrgpu up front.  The order-dependent part (mate table, add_soft conditions, unplaced / adjust_by,
// append order of cache.cache == `.bin` record order; extract.nim:93-132,192-248) is then replayed on the host in
// file order with the scan results.  Inflate, decode and staging are spread over the host threads (the decode buffer
// travels with the batch, so qnames are never copied); a consumer thread waits for the GPU and replays batch n-1 while
// the producer stages batch n+1.
#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <deque>
#include <mutex>
#include <thread>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "bam.hpp"
#include "commands.hpp"
#include "mate_table.hpp"
#include "pool.hpp"
#include "strgpu.h"
#include "tread.hpp"

namespace strling {

namespace {

struct Intervals {  // stands in for Lapper[region] (read_bed.nim:30-50); find = any overlap with [start, stop)
  std::vector<int64_t> start, max_stop;
  bool find(int64_t s, int64_t e) const {
    const size_t idx = (size_t)(std::lower_bound(start.begin(), start.end(), e) - start.begin());  // starts < e
    return idx > 0 && max_stop[idx - 1] > s;
  }
};

std::unordered_map<std::string, Intervals> read_bed(const std::string &path) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("[strling] couldn't open genome repeats file: " + path);
  std::unordered_map<std::string, std::vector<std::pair<int64_t, int64_t>>> raw;
  std::string line;
  while (std::getline(in, line)) {
    if (line.rfind("track ", 0) == 0 || (!line.empty() && line[0] == '#')) continue;
    while (!line.empty() && (line.back() == '\r' || line.back() == ' ' || line.back() == '\t')) line.pop_back();
    const size_t a = line.find('\t');
    if (a == std::string::npos) continue;
    const size_t b = line.find('\t', a + 1);
    if (b == std::string::npos) continue;
    size_t c = line.find('\t', b + 1);
    if (c == std::string::npos) c = line.size();
    raw[line.substr(0, a)].emplace_back(std::stoll(line.substr(a + 1, b - a - 1)), std::stoll(line.substr(b + 1, c - b - 1)));
  }
  std::unordered_map<std::string, Intervals> out;
  for (auto &kv : raw) {
    std::sort(kv.second.begin(), kv.second.end());
    Intervals iv;
    int64_t mx = INT64_MIN;
    for (auto &p : kv.second) {
      iv.start.push_back(p.first);
      mx = std::max(mx, p.second);
      iv.max_stop.push_back(mx);
    }
    out.emplace(kv.first, std::move(iv));
  }
  return out;
}

constexpr int kClsRead = 0, kClsFirstSeen = 1, kClsSecondSeen = 2;
constexpr uint8_t kNoOwner = 255;   // a record the replay does not look at (secondary / supplementary, placed records in the tail pass)
constexpr int kMaxShards = 16;

struct Pending {  // what the replay needs from one BAM record (qname stays in the batch's decode buffer)
  int32_t tid, pos, stop, mate_tid, mate_pos;
  uint16_t flag, n_cigar;
  uint8_t mapq;
  uint8_t n_seg;           // segments this record contributes (0..5)
  uint32_t first_cig, last_cig;
  int32_t l_seq, m_len;
  int32_t seg_full;        // segment index or -1 (filtered by the genome-STR rule)
  int32_t seg_clip[2][2];  // [0 left / 1 right][0 first-seen class / 1 second-seen class], -1 = none
  const char *qname;
  uint32_t qname_len;
  uint32_t aligned_bases;  // bases reserved in seq2 (0: the record's SEQ is not needed)
  uint64_t hash;           // of the qname: picks the replay shard and the slot in its mate table
  bool skip, clip_l, clip_r, primary;
};

template <typename T>
struct PodVec {  // uninitialised growable array (std::vector would value-initialise ~50 MB per batch on one thread)
  T *p = nullptr;
  size_t n = 0, cap = 0;
  PodVec() = default;
  PodVec(const PodVec &) = delete;
  PodVec &operator=(const PodVec &) = delete;
  ~PodVec() { std::free(p); }
  void resize(size_t m) {
    if (m > cap) {
      std::free(p);
      cap = m + m / 8 + 64;
      p = static_cast<T *>(std::malloc(cap * sizeof(T)));
      if (!p) throw std::runtime_error("[strling] out of memory");
    }
    n = m;
  }
  size_t size() const { return n; }
  T &operator[](size_t i) { return p[i]; }
  const T &operator[](size_t i) const { return p[i]; }
};

struct Batch {
  BamChunk chunk;                // owns the decoded records (qname / SEQ are used in place)
  PodVec<Pending> recs;          // one per record of the chunk, file order
  PodVec<uint8_t> owner;         // replay shard of every record, kNoOwner: not replayed
  std::vector<std::pair<uint32_t, int32_t>> tid_changes;  // (record, tid) wherever the reference id changes (progress messages)
  uint64_t n_replayed = 0;
  uint8_t *seq2 = nullptr;       // pinned
  uint32_t *nmask = nullptr;     // pinned
  uint32_t *xmask = nullptr;     // pinned: non-ACGT bases other than the literal N (stay out of the N > 20 gate, utils.nim:238)
  strgpu_segment *segs = nullptr;
  strgpu_repeat *out = nullptr;
  uint64_t n_bases = 0, cap_bases = 0;
  uint32_t n_seg = 0, cap_seg = 0;
  uint32_t max_len = 0;
  bool any_n = false;
  int ticket = -1;
};

// One replay shard: the reads whose qname hashes to it, in file order.  Everything add() (extract.nim:192-248) does is keyed
// by the qname, so shards never interact; what they append to cache.cache is tagged with (record, ordinal within the record)
// and merged back into file order once per batch.
struct Shard {
  MateTable tbl;
  std::vector<std::pair<uint64_t, Tread>> out;
  uint64_t tag = 0;
};

// `--gpu-inflate`: the compressed bytes of a chunk's blocks (one contiguous stretch of the file) are copied into a pinned
// staging buffer, the block descriptors rebased to it, and strgpu_inflate_bgzf inflates them on the device into the chunk's
// (pinned) buffer.  gpu == nullptr (`strling debug extract` with STRLING_DEBUG_STAGED_INFLATE): the same staging and
// descriptors, decoded by the host threads -- the CPU check of this plumbing.
struct StagedInflater {
  strgpu_ctx *gpu = nullptr;
  Pool *pool = nullptr;
  uint8_t *staging = nullptr;
  size_t cap = 0;
  ~StagedInflater() { release(); }
  void release() {
    if (staging) { if (gpu) strgpu_host_free(staging); else std::free(staging); }
    staging = nullptr;
    cap = 0;
  }
  void operator()(const uint8_t *file, size_t file_size, const BgzfBlockRef *blocks, size_t n, uint8_t *out, size_t out_bytes) {
    const uint64_t lo = blocks[0].in_off;
    const uint64_t hi = std::min<uint64_t>(file_size, blocks[n - 1].in_off + blocks[n - 1].csize + 8);  // + the last block's footer
    const size_t bytes = (size_t)(hi - lo);
    if (bytes + 16 > cap) {
      release();
      cap = bytes + bytes / 4 + 4096;
      if (gpu) {
        void *p = nullptr;
        if (strgpu_host_alloc(&p, cap) != STRGPU_OK) throw std::runtime_error("[strling] gpu: host_alloc failed");
        staging = static_cast<uint8_t *>(p);
      } else {
        staging = static_cast<uint8_t *>(std::malloc(cap));
        if (!staging) throw std::runtime_error("[strling] out of memory");
      }
    }
    pool->ranges(bytes, (bytes + (1u << 20) - 1) >> 20, [&](size_t a, size_t e, size_t) { std::memcpy(staging + a, file + lo + a, e - a); });
    std::memset(staging + bytes, 0, 16);
    std::vector<strgpu_bgzf_block> d(n);
    for (size_t i = 0; i < n; i++) d[i] = strgpu_bgzf_block{blocks[i].in_off - lo, blocks[i].csize, blocks[i].isize, blocks[i].out_off};
    if (gpu) {
      if (strgpu_inflate_bgzf(gpu, staging, bytes, d.data(), (uint32_t)n, out, out_bytes) != STRGPU_OK)
        throw std::runtime_error(std::string("[strling] gpu: inflate_bgzf: ") + strgpu_last_error(gpu));
    } else {
      pool->run(n, [&](size_t i) { inflate_bgzf_block(staging + d[i].in_off, d[i].csize, out + d[i].out_off, d[i].isize); });
    }
  }
};

struct Extractor {
  strgpu_ctx *gpu = nullptr;
  Options opts;
  double p = 0.8;
  const std::vector<Target> *targets = nullptr;
  std::vector<const Intervals *> genome_str_by_tid;  // nullptr: chrom not in genome_str
  Shard shards[kMaxShards];                          // Cache.tbl, split by qname hash
  int n_shards = 1;
  std::vector<Tread> cache;                          // Cache.cache
  uint64_t n_reads = 0, n_scanned = 0;
  std::atomic<uint64_t> n_warned{0};
  double t_wait = 0, t_replay = 0, t_submit = 0, t_decode = 0, t_stage = 0;  // host stage timers (seconds)
  double t_stage_parse = 0, t_stage_grow = 0, t_stage_pack = 0, t_replay_shards = 0;
  bool verbose = false;
  int threads = 1;
  Pool *pool = nullptr;
  // `strling debug extract`: the scan is replaced by a segment dump / by results read from a file (commands.hpp); no GPU then
  FILE *dump = nullptr, *results = nullptr;
  bool debug_mode() const { return dump != nullptr || results != nullptr; }

  void gpu_check(int rc, const char *what) {
    if (rc != STRGPU_OK) throw std::runtime_error(std::string("[strling] gpu: ") + what + ": " + strgpu_last_error(gpu));
  }

  void alloc_batch(Batch &b, uint64_t cap_bases, uint32_t cap_seg) {
    b.cap_bases = cap_bases;
    b.cap_seg = cap_seg;
    void *p1, *p2, *p3, *p4, *p5;
    auto get = [&](void **p, size_t bytes) {
      if (debug_mode()) {   // no CUDA runtime on the machines the CPU-side checks run on: plain memory
        *p = std::malloc(bytes ? bytes : 1);
        if (!*p) throw std::runtime_error("[strling] out of memory");
      } else {
        gpu_check(strgpu_host_alloc(p, bytes), "host_alloc");
      }
    };
    get(&p1, strgpu_seq2_bytes(cap_bases));
    get(&p2, strgpu_nmask_bytes(cap_bases));
    get(&p3, (size_t)cap_seg * sizeof(strgpu_segment));
    get(&p4, (size_t)cap_seg * sizeof(strgpu_repeat));
    get(&p5, strgpu_nmask_bytes(cap_bases));
    b.seq2 = (uint8_t *)p1; b.nmask = (uint32_t *)p2; b.segs = (strgpu_segment *)p3; b.out = (strgpu_repeat *)p4;
    b.xmask = (uint32_t *)p5;
    std::memset(b.nmask, 0, strgpu_nmask_bytes(cap_bases));
    std::memset(b.xmask, 0, strgpu_nmask_bytes(cap_bases));
  }
  void free_batch(Batch &b) {
    if (debug_mode()) { std::free(b.seq2); std::free(b.nmask); std::free(b.xmask); std::free(b.segs); std::free(b.out); return; }
    strgpu_host_free(b.seq2); strgpu_host_free(b.nmask); strgpu_host_free(b.xmask); strgpu_host_free(b.segs); strgpu_host_free(b.out);
  }
  void grow_batch(Batch &b, uint64_t need_bases, uint32_t need_seg) {
    if (need_bases <= b.cap_bases && need_seg <= b.cap_seg) return;
    free_batch(b);
    alloc_batch(b, std::max<uint64_t>(need_bases + need_bases / 8, b.cap_bases), std::max<uint32_t>(need_seg + need_seg / 8, b.cap_seg));
  }

  // Decode every record of the chunk and lay the batch out: phase 1 (parallel) parses the fields, hashes the qname and
  // decides which segments a record contributes; a prefix sum over the parts fixes every part's place in seq2 / segs; phase 2
  // (parallel, same parts) packs the SEQ fields and writes the descriptors.
  void stage(Batch &b, bool pass1) {
    const size_t n = b.chunk.n_records();
    const auto s0 = std::chrono::steady_clock::now();
    b.recs.resize(n);
    b.owner.resize(n);
    b.tid_changes.clear();
    const uint8_t *data = b.chunk.data.data();  // RawBuffer
    const size_t parts = std::max<size_t>(1, std::min<size_t>((size_t)pool->size() * 4, (n + 2047) / 2048));
    struct PartSum { uint64_t bases = 0; uint32_t segs = 0; uint64_t replayed = 0; std::vector<std::pair<uint32_t, int32_t>> tids; };
    std::vector<PartSum> sums(parts);
    const uint32_t n_sh = (uint32_t)n_shards;
    pool->ranges(n, parts, [&](size_t lo, size_t hi, size_t part) {
      PartSum ps;
      int32_t last_tid = INT32_MIN;
      for (size_t i = lo; i < hi; i++) {
        if (i + 6 < hi) {
          __builtin_prefetch(data + b.chunk.rec_off[i + 6]);
          __builtin_prefetch(data + b.chunk.rec_off[i + 6] + 64);
        }
        const BamRecord r = BamChunk::view(data + b.chunk.rec_off[i]);
        Pending &pr = b.recs[i];
        pr.tid = r.tid; pr.pos = r.pos; pr.mate_tid = r.mate_tid; pr.mate_pos = r.mate_pos;
        pr.flag = r.flag; pr.n_cigar = r.n_cigar; pr.mapq = r.mapq; pr.l_seq = r.l_seq;
        pr.first_cig = r.n_cigar ? r.cigar_at(0) : 0;
        pr.last_cig = r.n_cigar ? r.cigar_at(r.n_cigar - 1) : 0;
        pr.m_len = 0;
        pr.seg_full = -1;
        pr.seg_clip[0][0] = pr.seg_clip[0][1] = pr.seg_clip[1][0] = pr.seg_clip[1][1] = -1;
        pr.qname = r.qname;
        pr.qname_len = r.l_qname;
        pr.primary = !(r.flag & (0x100 | 0x800));
        pr.skip = pr.clip_l = pr.clip_r = false;
        pr.n_seg = 0;
        pr.aligned_bases = 0;
        pr.stop = 0;
        pr.hash = 0;
        b.owner[i] = kNoOwner;
        if (r.tid != last_tid) { ps.tids.emplace_back((uint32_t)i, r.tid); last_tid = r.tid; }
        if (!pr.primary) continue;
        if (!pass1 && r.tid >= 0) continue;   // ibam.query("*") returns no-coordinate records only (extract.nim:326-329)
        if (r.l_seq > STRGPU_MAX_SEGMENT_LEN)
          throw std::runtime_error("[strling] read longer than " + std::to_string(STRGPU_MAX_SEGMENT_LEN) + " bp: " + std::string(r.qname));
        pr.stop = r.stop();
        pr.hash = hash_name(r.qname, r.l_qname);
        b.owner[i] = (uint8_t)((pr.hash >> 40) % n_sh);
        ps.replayed++;
        // extract.nim:30-34 : exact single-M match outside every genome STR region -> no scan
        if (r.n_cigar == 1 && BamRecord::op(pr.first_cig) == 0 && r.tid >= 0 && (size_t)r.tid < genome_str_by_tid.size() &&
            genome_str_by_tid[(size_t)r.tid] != nullptr && !genome_str_by_tid[(size_t)r.tid]->find(r.pos, pr.stop)) {
          pr.skip = true;
          pr.m_len = (int32_t)BamRecord::oplen(pr.first_cig);
        }
        // add_soft preconditions that do not depend on scan results (extract.nim:97-104)
        if (r.mapq >= opts.min_mapq && r.n_cigar > 0) {
          pr.clip_l = BamRecord::op(pr.first_cig) == 4;
          pr.clip_r = BamRecord::op(pr.last_cig) == 4 && r.n_cigar > 1;  // a single op is handled as "left" twice (extract.nim:102-112)
        }
        pr.n_seg = (uint8_t)((pr.skip ? 0 : 1) + (pr.clip_l ? 2 : 0) + (pr.clip_r ? 2 : 0));
        if (pr.n_seg) pr.aligned_bases = ((uint32_t)r.l_seq + 31u) & ~31u;  // 32-base alignment: no two records share an nmask word
        ps.bases += pr.aligned_bases;
        ps.segs += pr.n_seg;
      }
      sums[part] = std::move(ps);
    }, 1);
    const auto s1 = std::chrono::steady_clock::now();
    uint64_t nb = 0;
    uint32_t ns = 0;
    uint64_t n_rep = 0;
    std::vector<uint64_t> part_base(parts);
    std::vector<uint32_t> part_seg(parts);
    int32_t last_tid = INT32_MIN;
    for (size_t q = 0; q < parts; q++) {
      part_base[q] = nb;
      part_seg[q] = ns;
      nb += sums[q].bases;
      if ((uint64_t)ns + sums[q].segs > 0xfffffff0ull) throw std::runtime_error("[strling] batch too large: lower --batch-reads");
      ns += sums[q].segs;
      n_rep += sums[q].replayed;
      for (const auto &tc : sums[q].tids)
        if (tc.second != last_tid) { b.tid_changes.push_back(tc); last_tid = tc.second; }
    }
    if (nb > 0xffffff00ull) throw std::runtime_error("[strling] batch too large: lower --batch-reads");
    grow_batch(b, nb + 64, ns + 8);
    b.n_bases = nb;
    b.n_seg = ns;
    b.n_replayed = n_rep;
    const auto s2 = std::chrono::steady_clock::now();
    std::vector<uint32_t> tmax(parts, 0);
    std::vector<uint8_t> tany(parts, 0);
    pool->ranges(n, parts, [&](size_t lo, size_t hi, size_t part) {
      uint32_t mx = 0;
      bool any = false;
      uint64_t base = part_base[part];
      uint32_t si = part_seg[part];
      for (size_t i = lo; i < hi; i++) {
        if (i + 6 < hi) {
          __builtin_prefetch(data + b.chunk.rec_off[i + 6]);
          __builtin_prefetch(data + b.chunk.rec_off[i + 6] + 64);
          __builtin_prefetch(data + b.chunk.rec_off[i + 6] + 128);
        }
        Pending &pr = b.recs[i];
        if (!pr.n_seg) continue;
        const BamRecord r = BamChunk::view(data + b.chunk.rec_off[i]);
        const int n_other = strgpu_pack_bam4(r.seq, (uint32_t)r.l_seq, b.seq2, b.nmask, b.xmask, base);
        if (n_other < 0) throw std::runtime_error("[strling] pack_bam4 failed");
        const bool has_n = n_other > 0;
        any = any || has_n;
        auto put = [&](uint64_t off, uint32_t len, int cls) {
          b.segs[si] = strgpu_segment{(uint32_t)off, (uint16_t)len, (uint8_t)cls, (uint8_t)(has_n ? STRGPU_SEG_HAS_N : 0)};
          mx = std::max(mx, len);
          return (int32_t)si++;
        };
        if (!pr.skip) pr.seg_full = put(base, (uint32_t)r.l_seq, kClsRead);
        if (pr.clip_l) {
          const uint32_t len = std::min<uint32_t>(BamRecord::oplen(pr.first_cig), (uint32_t)r.l_seq);
          pr.seg_clip[0][0] = put(base, len, kClsFirstSeen);
          pr.seg_clip[0][1] = put(base, len, kClsSecondSeen);
        }
        if (pr.clip_r) {
          const uint32_t len = std::min<uint32_t>(BamRecord::oplen(pr.last_cig), (uint32_t)r.l_seq);
          pr.seg_clip[1][0] = put(base + (uint64_t)r.l_seq - len, len, kClsFirstSeen);
          pr.seg_clip[1][1] = put(base + (uint64_t)r.l_seq - len, len, kClsSecondSeen);
        }
        base += pr.aligned_bases;
      }
      tmax[part] = mx;
      tany[part] = any;
    }, 1);
    b.max_len = *std::max_element(tmax.begin(), tmax.end());
    b.any_n = std::any_of(tany.begin(), tany.end(), [](uint8_t v) { return v != 0; });
    const auto s3 = std::chrono::steady_clock::now();
    t_stage_parse += std::chrono::duration<double>(s1 - s0).count();
    t_stage_grow += std::chrono::duration<double>(s2 - s1).count();
    t_stage_pack += std::chrono::duration<double>(s3 - s2).count();
  }

  void submit(Batch &b) {
    n_scanned += b.n_seg;
    if (dump) {   // one line per staged segment, in the order the results are consumed in
      std::string line;
      for (uint32_t i = 0; i < b.n_seg; i++) {
        const strgpu_segment &sg = b.segs[i];
        line.assign(std::to_string((int)sg.pclass));
        line.push_back('\t');
        for (uint32_t k = 0; k < sg.len; k++) {
          const uint64_t base = (uint64_t)sg.base_off + k;
          char c = "CATG"[(b.seq2[base >> 2] >> (6 - 2 * (base & 3))) & 3];
          if (b.any_n && ((b.nmask[base >> 5] >> (base & 31)) & 1u)) c = ((b.xmask[base >> 5] >> (base & 31)) & 1u) ? 'R' : 'N';
          line.push_back(c);
        }
        line.push_back('\n');
        std::fputs(line.c_str(), dump);
      }
      std::memset(b.out, 0, (size_t)b.n_seg * sizeof(strgpu_repeat));
      return;
    }
    if (results) {
      if (b.n_seg && std::fread(b.out, sizeof(strgpu_repeat), b.n_seg, results) != b.n_seg)
        throw std::runtime_error("[strling] debug extract: the scan-results file is shorter than the staged segments");
      return;
    }
    gpu_check(strgpu_scan_submit(gpu, b.seq2, b.n_bases, b.any_n ? b.nmask : nullptr, b.any_n ? b.xmask : nullptr, b.segs, b.n_seg,
                                 b.max_len, b.out, &b.ticket),
              "scan_submit");
  }
  void wait(Batch &b) {
    if (debug_mode()) return;
    gpu_check(strgpu_scan_wait(gpu, b.ticket), "scan_wait");
  }
  void recycle(Batch &b) {
    if (b.any_n) {
      std::memset(b.nmask, 0, (size_t)((b.n_bases + 31) / 32) * 4 + 8);
      std::memset(b.xmask, 0, (size_t)((b.n_bases + 31) / 32) * 4 + 8);
    }
    b.n_bases = 0; b.n_seg = 0; b.max_len = 0; b.any_n = false; b.ticket = -1;
  }

  // `--gpu-inflate` only: the submit slots get their device buffers (and the scan kernels are loaded) before the pipeline starts.
  // cudaMalloc waits for every kernel in flight, so a first-use allocation inside strgpu_scan_submit would otherwise stall the
  // staging thread behind a running inflate kernel, once per slot and buffer.  Empty segments: nothing is scanned.
  void warm_up(Batch *bs, int n) {
    if (debug_mode()) return;
    int tickets[STRGPU_SLOTS];
    const int m = std::min(n, (int)STRGPU_SLOTS);
    for (int i = 0; i < m; i++) {
      Batch &b = bs[i];
      const uint32_t n_seg = b.cap_seg > 16 ? b.cap_seg - 16 : 0;
      std::memset(b.segs, 0, (size_t)n_seg * sizeof(strgpu_segment));
      gpu_check(strgpu_scan_submit(gpu, b.seq2, b.cap_bases > 64 ? b.cap_bases - 64 : 0, nullptr, nullptr, b.segs, n_seg, 160, b.out, &tickets[i]), "scan_submit (warm-up)");
    }
    for (int i = 0; i < m; i++) gpu_check(strgpu_scan_wait(gpu, tickets[i]), "scan_wait (warm-up)");
  }

  // ---- replay: extract.nim:63-132,192-248 with scan results looked up instead of computed.  The pair arithmetic runs on
  // records without the qname (every read of a pair carries the same one); emit() attaches it.
  TreadCore to_tread(const Batch &b, const Pending &r) {
    TreadCore t;
    t.tid = r.tid;
    t.position = (uint32_t)std::max(0, r.pos);
    t.flag = r.flag;
    t.split = kNone;
    t.mapping_quality = r.mapq;
    int align_length = r.m_len, repeat_count = 0;
    if (r.seg_full >= 0) {
      const strgpu_repeat &res = b.out[r.seg_full];
      std::memcpy(t.repeat.data(), res.unit, 6);
      repeat_count = res.repeat_count;
      align_length = r.l_seq;
    }
    if (repeat_count >= 256)  // doAssert, extract.nim:72
      throw std::runtime_error("[strling] repeat_count >= 256 for read " + std::string(r.qname, r.qname_len));
    t.repeat_count = (uint8_t)repeat_count;
    t.align_length = (uint8_t)align_length;
    if (r.n_cigar > 1 && BamRecord::op(r.first_cig) == 4 && BamRecord::oplen(r.first_cig) > 16) t.split = kNoneLeft;
    if (r.n_cigar > 1 && BamRecord::op(r.last_cig) == 4 && BamRecord::oplen(r.last_cig) > 16) t.split = kNoneRight;
    return t;
  }

  static void emit(Shard &sh, const TreadCore &t, const Pending &r) { sh.out.emplace_back(sh.tag++, Tread(t, r.qname, r.qname_len)); }

  void add_soft(Shard &sh, const Batch &b, const Pending &r, int cls_idx, const std::array<char, 6> &read_repeat) {
    if (r.mapq < opts.min_mapq) return;
    if (r.n_cigar == 0 || (BamRecord::op(r.first_cig) != 4 && BamRecord::op(r.last_cig) != 4)) return;
    const int idxs[2] = {0, (int)r.n_cigar - 1};
    for (int cig_index : idxs) {
      const uint32_t c = cig_index == 0 ? r.first_cig : r.last_cig;
      if (BamRecord::op(c) != 4) continue;
      const uint32_t clen = BamRecord::oplen(c);
      if (read_repeat[0] == 0 && clen <= 16) continue;
      const bool is_left = cig_index == 0;
      const int32_t si = r.seg_clip[is_left ? 0 : 1][cls_idx];
      if (si < 0) continue;
      const strgpu_repeat &res = b.out[si];
      if (res.repeat_count == 0) continue;
      TreadCore tr;
      tr.tid = r.tid;
      tr.position = (uint32_t)(is_left ? std::max(0, r.pos) : std::max(0, r.stop));
      tr.flag = r.flag;
      std::memcpy(tr.repeat.data(), res.unit, 6);
      tr.repeat_count = (uint8_t)res.repeat_count;
      tr.align_length = (uint8_t)std::min<uint32_t>(clen, (uint32_t)r.l_seq);
      tr.split = is_left ? kLeft : kRight;
      tr.mapping_quality = r.mapq;
      if (p_repeat(tr) < 0.9) continue;
      emit(sh, tr, r);
    }
  }

  void add(Shard &sh, const Batch &b, const Pending &r) {
    MateTable &tbl = sh.tbl;
    size_t slot = SIZE_MAX;
    bool looked = false;
    bool after_mate = r.tid > r.mate_tid;
    if (!after_mate && r.tid == r.mate_tid) {
      if (r.pos > r.mate_pos) after_mate = true;
      else if (r.pos == r.mate_pos) { slot = tbl.find(r.qname, r.qname_len, r.hash); looked = true; after_mate = slot != SIZE_MAX; }
    }
    if (after_mate) {
      if (!looked) slot = tbl.find(r.qname, r.qname_len, r.hash);
      if (slot == SIZE_MAX) return;
      TreadCore mate = tbl.at(slot).t;
      tbl.erase(slot);
      TreadCore self = to_tread(b, r);
      add_soft(sh, b, r, 1, self.repeat);  // opts.proportion_repeat = min(p, 0.6)
      if (mate.repeat_count == 0 && self.repeat_count == 0) return;
      if (unplaced_pair(self, mate, opts)) {
        if (self.repeat[0] == 0 || mate.repeat[0] == 0) return;
        self.repeat = canonical_repeat(self.repeat);
        self.position = 0;
        self.tid = -1;
        mate.repeat = canonical_repeat(mate.repeat);
        mate.position = 0;
        mate.tid = -1;
        emit(sh, self, r);
        emit(sh, mate, r);
        return;
      }
      const uint32_t mp = mate.position;
      if (adjust_by(mate, self, opts, self.position)) emit(sh, mate, r);
      if (adjust_by(self, mate, opts, mp)) emit(sh, self, r);
    } else {
      const TreadCore tr = to_tread(b, r);
      add_soft(sh, b, r, 0, tr.repeat);  // opts.proportion_repeat = p - 0.07
      if (!looked) slot = tbl.find(r.qname, r.qname_len, r.hash);
      if (slot != SIZE_MAX) {  // hasKeyOrPut found the key: warn and drop it (extract.nim:245-248)
        if (n_warned.fetch_add(1) < 20)
          std::fprintf(stderr, "[strling] warning. bad read (this happens with bwa-kit alignments):%.*s already in table\n", (int)r.qname_len, r.qname);
        tbl.erase(slot);
      } else {
        tbl.insert(r.qname, r.qname_len, r.hash, tr);
      }
    }
  }

  // One batch: every shard walks the batch's records in file order and handles its own; what they emitted is merged into
  // cache.cache by (record, ordinal) -- the order a single add() loop would have appended in.
  void replay(const Batch &b) {
    const size_t n = b.recs.size();
    const uint8_t *own = b.owner.p;
    auto shard_work = [&](size_t s) {
      Shard &sh = shards[s];
      sh.out.clear();
      for (size_t i = 0; i < n; i++) {
        if (own[i] != (uint8_t)s) continue;
        sh.tag = (uint64_t)i << 4;
        add(sh, b, b.recs[i]);
      }
    };
    const auto r0 = std::chrono::steady_clock::now();
    if (n_shards == 1) shard_work(0);
    else pool->run((size_t)n_shards, shard_work, 2);
    t_replay_shards += std::chrono::duration<double>(std::chrono::steady_clock::now() - r0).count();
    size_t total = 0;
    for (int s = 0; s < n_shards; s++) total += shards[s].out.size();
    if (n_shards == 1) {
      for (auto &e : shards[0].out) cache.push_back(std::move(e.second));
      return;
    }
    cache.reserve(cache.size() + total);
    size_t head[kMaxShards] = {0};
    for (size_t k = 0; k < total; k++) {
      int best = -1;
      uint64_t best_tag = UINT64_MAX;
      for (int s = 0; s < n_shards; s++)
        if (head[s] < shards[s].out.size() && shards[s].out[head[s]].first < best_tag) { best_tag = shards[s].out[head[s]].first; best = s; }
      cache.push_back(std::move(shards[best].out[head[best]++].second));
    }
  }

  size_t table_size() const {
    size_t n = 0;
    for (int s = 0; s < n_shards; s++) n += shards[s].tbl.size();
    return n;
  }
};

}  // namespace

// utils.nim:86-111 with n_reads = 2_000_000, skip_reads = 100_000 (extract.nim:273, call.nim:76)
std::array<uint32_t, 4096> fragment_length_distribution(const std::string &bam, int threads) {
  std::array<uint32_t, 4096> frag{};
  uint64_t first;
  int32_t n_ref;
  {
    BamReader hdr(bam, 1);
    first = hdr.tell();
    n_ref = (int32_t)hdr.targets().size();
  }
  BamChunkReader rd(bam, first, threads, nullptr, n_ref);  // blocks inflated in parallel; the records are looked at in file order
  BamChunk c;
  int64_t i = -1, counted = 0;
  std::vector<int32_t> skipped;
  bool done = false;
  size_t blocks = 64;
  while (!done && rd.next(c, blocks)) {
    blocks = std::min<size_t>(blocks * 4, 2048);
    const uint8_t *data = c.data.data();
    const size_t n = c.n_records();
    for (size_t k = 0; k < n; k++) {
      if (k + 8 < n) __builtin_prefetch(data + c.rec_off[k + 8]);
      const uint8_t *p = data + c.rec_off[k] + 4;
      uint16_t flag;
      int32_t isize;
      std::memcpy(&flag, p + 14, 2);
      std::memcpy(&isize, p + 28, 4);
      i++;
      if (!(flag & 0x2)) continue;
      if (flag & (0x800 | 0x100)) continue;
      if (isize < 0 || isize > 4095) continue;
      if (i < 100000) { skipped.push_back(isize); continue; }
      skipped.clear();
      frag[(size_t)isize]++;
      if (++counted > 2000000) { done = true; break; }
    }
  }
  uint64_t sum = 0;
  for (uint32_t v : frag) sum += v;
  if (sum == 0) {
    std::fprintf(stderr, "using first reads in fragment_length_distribution calculation as there were not enough\n");
    for (int32_t v : skipped) frag[(size_t)v]++;
  }
  return frag;
}


int extract_run(const ExtractArgs &a) {
  using clk = std::chrono::steady_clock;
  const auto t_start = clk::now();
  // ---- pass 1: fragment length distribution (utils.nim:86-111), skip_reads = 100000 (extract.nim:273)
  std::array<uint32_t, 4096> frag = fragment_length_distribution(a.bam, a.threads);
  Extractor ex;
  ex.verbose = a.verbose;
  ex.p = a.proportion_repeat;
  ex.opts.median_fragment_length = frag_median(frag);
  ex.opts.proportion_repeat = a.proportion_repeat;
  ex.opts.min_mapq = (uint8_t)a.min_mapq;
  if (a.verbose) {
    std::fprintf(stderr, "Calculated median fragment length:%d\n", ex.opts.median_fragment_length);
    std::fprintf(stderr, "10th, 90th percentile of fragment length:%d %d\n", frag_median(frag, 0.1), frag_median(frag, 0.9));
  }

  BamReader rd(a.bam, a.threads);
  ex.targets = &rd.targets();
  std::unordered_map<std::string, Intervals> genome_str;
  // genome_repeats (genome_strs.nim:107-141): use the -g bed if it exists, otherwise build it from the fasta
  // (into -g when given, else only in memory -- the reference uses a temporary file)
  bool have_bed = false;
  if (!a.genome_repeats.empty()) {
    std::ifstream probe(a.genome_repeats);
    have_bed = (bool)probe;
  }
  if (have_bed) {
    std::fprintf(stderr, "[strling] using existing file %s for genome repeats\n", a.genome_repeats.c_str());
    genome_str = read_bed(a.genome_repeats);
  } else if (!a.fasta.empty()) {
    if (!a.debug_dump_segments.empty() || !a.debug_scan_results.empty())
      throw std::runtime_error("[strling] debug extract: building the genome-repeats index needs the GPU; pass an existing -g file");
    const std::vector<std::string> lines = genome_repeat_lines(a.fasta, a.proportion_repeat, a.device);
    std::fprintf(stderr, "[strling] found %zu STR-like regions in the genome\n", lines.size());
    std::string path = a.genome_repeats;
    const bool tmp = path.empty();
    if (tmp) path = a.bin + ".genome_repeats.tmp";
    {
      std::ofstream out(path);
      if (!out) throw std::runtime_error("[strling] couldn't open bed file: " + path + " for writing");
      for (const auto &l : lines) out << l << "\n";
    }
    genome_str = read_bed(path);
    if (tmp) std::remove(path.c_str());
  } else {
    if (!a.genome_repeats.empty())
      throw std::runtime_error("[strling] genome repeats file " + a.genome_repeats + " does not exist and no -f fasta was given to build it");
    std::fprintf(stderr, "[strling] no -f fasta / -g genome repeats: every read is scanned (the reference requires -f)\n");
  }
  std::fprintf(stderr, "[strling] got STR repeats from genome into an interval tree\n");
  ex.genome_str_by_tid.assign(rd.targets().size(), nullptr);
  for (size_t t = 0; t < rd.targets().size(); t++) {
    auto it = genome_str.find(rd.targets()[t].name);
    if (it != genome_str.end()) ex.genome_str_by_tid[t] = &it->second;
  }

  if (!a.debug_dump_segments.empty()) {
    ex.dump = std::fopen(a.debug_dump_segments.c_str(), "w");
    if (!ex.dump) throw std::runtime_error("[strling] debug extract: cannot write " + a.debug_dump_segments);
  } else if (!a.debug_scan_results.empty()) {
    ex.results = std::fopen(a.debug_scan_results.c_str(), "rb");
    if (!ex.results) throw std::runtime_error("[strling] debug extract: cannot read " + a.debug_scan_results);
  } else {
    int rc = strgpu_create(&ex.gpu, a.device);
    if (rc != STRGPU_OK) throw std::runtime_error(std::string("[strling] gpu: ") + strgpu_error_string(rc));
    const double classes[3] = {a.proportion_repeat, a.proportion_repeat - 0.07, std::min(a.proportion_repeat, 0.6)};
    ex.gpu_check(strgpu_set_proportions(ex.gpu, classes, 3), "set_proportions");
  }

  ex.threads = a.threads > 0 ? a.threads : (int)std::max(1u, std::min(64u, std::thread::hardware_concurrency()));
  Pool pool(ex.threads);
  ex.pool = &pool;
  // replay shards (by qname hash): the mate-table work of one batch spread over a quarter of the threads
  ex.n_shards = std::max(1, std::min(kMaxShards, a.replay_shards > 0 ? a.replay_shards : ex.threads / 4));
  const uint64_t first_voffset = rd.tell();
  constexpr int kBatches = STRGPU_SLOTS + 1;  // one being inflated, one being staged, up to two on the GPU / in replay
  Batch batches[kBatches];
  StagedInflater staged;
  staged.gpu = ex.gpu;
  staged.pool = &pool;
  if (a.gpu_inflate && ex.gpu)  // the device copies the inflated records straight into the chunk buffers: pinned memory
    for (auto &b : batches)
      b.chunk.data.set_allocator([](size_t n) -> void * { void *p = nullptr; return strgpu_host_alloc(&p, n) == STRGPU_OK ? p : nullptr; },
                                 [](void *p) { strgpu_host_free(p); });
  // a chunk is a number of ~64 KiB BGZF blocks (~220 records of a 150-bp library each); the pinned buffers start out with
  // room for a quarter more than that and grow should a file pack more reads into its blocks
  const size_t blocks_per_chunk = std::max<size_t>(16, (size_t)a.batch_reads / 220);
  for (auto &b : batches) ex.alloc_batch(b, (uint64_t)a.batch_reads * 200 + 4096, (uint32_t)std::min<uint64_t>((uint64_t)a.batch_reads * 2 + 64, 0xfffffff0u));

  if (a.gpu_inflate) ex.warm_up(batches, kBatches);
  std::fprintf(stderr, "[strling] collecting str-like reads\n");
  const auto t0 = clk::now();
  // producer (this thread): inflate + decode + stage (all parallel) + submit.  consumer: wait for the GPU, replay in
  // file order (the mate table makes that part sequential), recycle the batch.
  std::mutex mu;
  std::condition_variable cv;
  std::deque<int> submitted;
  int on_gpu = 0;   // batches submitted and not yet waited for: never more than the library's STRGPU_SLOTS submit slots
  bool is_free[kBatches];
  for (bool &f : is_free) f = true;
  bool done = false;
  std::string consumer_error;
  bool first_pass = true;
  int32_t tid_seen = -1;
  std::thread consumer([&]() {
    try {
      while (true) {
        int bi;
        {
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&]() { return !submitted.empty() || done; });
          if (submitted.empty()) return;
          bi = submitted.front();
          submitted.pop_front();
        }
        Batch &b = batches[bi];
        const auto w0 = clk::now();
        ex.wait(b);
        {
          std::lock_guard<std::mutex> lk(mu);
          on_gpu--;
        }
        cv.notify_all();
        const auto w1 = clk::now();
        if (first_pass)
          for (const auto &tc : b.tid_changes)
            if (tc.second != tid_seen && tc.second >= 0) {
              if (rd.targets()[(size_t)tc.second].length > 2000000u)
                std::fprintf(stderr, "[strling] extracting chromosome:%s\n", rd.targets()[(size_t)tc.second].name.c_str());
              tid_seen = tc.second;
            }
        ex.replay(b);
        if (ex.verbose && (ex.n_reads + b.n_replayed) / 10000000 != ex.n_reads / 10000000) {
          const double dt = std::chrono::duration<double>(clk::now() - t0).count();
          std::fprintf(stderr, "%llu %.1f reads/sec tbl len: %zu cache len: %zu\n", (unsigned long long)(ex.n_reads + b.n_replayed),
                       (ex.n_reads + b.n_replayed) / dt, ex.table_size(), ex.cache.size());
        }
        ex.n_reads += b.n_replayed;
        ex.recycle(b);
        ex.t_wait += std::chrono::duration<double>(w1 - w0).count();
        ex.t_replay += std::chrono::duration<double>(clk::now() - w1).count();
        {
          std::lock_guard<std::mutex> lk(mu);
          is_free[bi] = true;
        }
        cv.notify_all();
      }
    } catch (const std::exception &e) {
      std::lock_guard<std::mutex> lk(mu);
      consumer_error = e.what();
      for (bool &f : is_free) f = true;
      on_gpu = 0;
      cv.notify_all();
    }
  });
  uint64_t tail_voffset = 0;
  bool have_tail = false;
  auto drain = [&]() {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&]() { return (submitted.empty() && std::all_of(is_free, is_free + kBatches, [](bool f) { return f; })) || !consumer_error.empty(); });
  };
  auto acquire_free = [&]() -> int {
    std::unique_lock<std::mutex> lk(mu);
    cv.wait(lk, [&]() { return std::any_of(is_free, is_free + kBatches, [](bool f) { return f; }) || !consumer_error.empty(); });
    if (!consumer_error.empty()) return -1;
    for (int i = 0; i < kBatches; i++)
      if (is_free[i]) { is_free[i] = false; return i; }
    return -1;
  };
  auto feed = [&](uint64_t voffset, bool pass1) {
    // reader thread: inflate the next chunk into a free batch while this thread stages the previous one
    std::deque<int> decoded;
    bool reader_done = false;
    std::string reader_error;
    std::thread reader_thread([&]() {
      try {
        BamChunkReader reader(a.bam, voffset, ex.threads, &pool, (int32_t)rd.targets().size());
        if (a.gpu_inflate)
          reader.inflate_hook = [&](const uint8_t *file, size_t file_size, const BgzfBlockRef *blocks, size_t n, uint8_t *out, size_t out_bytes) {
            staged(file, file_size, blocks, n, out, out_bytes);
          };
        while (true) {
          const int bi = acquire_free();
          if (bi < 0) break;
          const auto d0 = clk::now();
          const bool more = reader.next(batches[bi].chunk, blocks_per_chunk);
          ex.t_decode += std::chrono::duration<double>(clk::now() - d0).count();
          std::lock_guard<std::mutex> lk(mu);
          if (more && batches[bi].chunk.n_records() > 0) decoded.push_back(bi);
          else is_free[bi] = true;
          cv.notify_all();
          if (!more) break;
        }
      } catch (const std::exception &e) {
        std::lock_guard<std::mutex> lk(mu);
        reader_error = e.what();
      }
      std::lock_guard<std::mutex> lk(mu);
      reader_done = true;
      cv.notify_all();
    });
    std::string stage_error;
    while (true) {
      int bi = -1;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv.wait(lk, [&]() { return !decoded.empty() || reader_done; });
        if (decoded.empty()) break;
        bi = decoded.front();
        decoded.pop_front();
      }
      Batch &b = batches[bi];
      if (!stage_error.empty() || !consumer_error.empty()) {  // keep draining so the reader can finish
        std::lock_guard<std::mutex> lk(mu);
        is_free[bi] = true;
        cv.notify_all();
        continue;
      }
      try {
        const auto d1 = clk::now();
        ex.stage(b, pass1);
        if (pass1 && !have_tail)
          for (const auto &tc : b.tid_changes)
            if (tc.second < 0) { have_tail = true; tail_voffset = b.chunk.voffset_of(tc.first); break; }
        const auto d2 = clk::now();
        ex.t_stage += std::chrono::duration<double>(d2 - d1).count();
        {
          // back-pressure instead of STRGPU_ERR_BUSY: wait until the consumer has waited for an earlier ticket
          std::unique_lock<std::mutex> lk(mu);
          cv.wait(lk, [&]() { return on_gpu < STRGPU_SLOTS || !consumer_error.empty(); });
          if (!consumer_error.empty()) throw std::runtime_error(consumer_error);
          on_gpu++;
        }
        ex.submit(b);
        ex.t_submit += std::chrono::duration<double>(clk::now() - d2).count();
        std::lock_guard<std::mutex> lk(mu);
        submitted.push_back(bi);
      } catch (const std::exception &e) {
        stage_error = e.what();
        std::lock_guard<std::mutex> lk(mu);
        is_free[bi] = true;
      }
      cv.notify_all();
    }
    reader_thread.join();
    if (!reader_error.empty()) throw std::runtime_error(reader_error);
    if (!stage_error.empty()) throw std::runtime_error(stage_error);
  };
  try {
    feed(first_voffset, true);            // pass 2: every record in file order (extract.nim:308-322)
    drain();
    std::fprintf(stderr, "[strling] extracting unmapped reads\n");
    if (have_tail && consumer_error.empty()) {  // ibam.query("*"): the no-coordinate tail again (extract.nim:326-329)
      first_pass = false;
      feed(tail_voffset, false);
      drain();
    }
  } catch (...) {
    {
      std::lock_guard<std::mutex> lk(mu);
      done = true;
    }
    cv.notify_all();
    consumer.join();
    throw;
  }
  {
    std::lock_guard<std::mutex> lk(mu);
    done = true;
  }
  cv.notify_all();
  consumer.join();
  if (!consumer_error.empty()) throw std::runtime_error(consumer_error);
  const double dt = std::chrono::duration<double>(clk::now() - t0).count();

  if (ex.dump) {   // segments only: the replay ran on empty results, there is nothing to write
    std::fclose(ex.dump);
    for (auto &b : batches) ex.free_batch(b);
    return 0;
  }
  if (ex.results) std::fclose(ex.results);
  std::fprintf(stderr, "[strling] writing binary file:%s\n", a.bin.c_str());
  BinFile bf;
  bf.proportion_repeat = (float)a.proportion_repeat;
  bf.min_mapq = (uint8_t)a.min_mapq;
  bf.frag_dist = frag;
  bf.header = rd.header_text();
  bf.reads = std::move(ex.cache);
  write_bin(a.bin, bf);
  std::fprintf(stderr, "[strling] finished extraction\n");
  if (a.verbose) {
    const double total = std::chrono::duration<double>(clk::now() - t_start).count();
    std::fprintf(stderr, "[strling] perf: {\"reads\": %llu, \"segments_scanned\": %llu, \"str_reads\": %zu, \"scan_pass_s\": %.3f, \"threads\": %d, \"gpu_inflate\": %s, \"inflate_s\": %.3f, \"stage_s\": %.3f, \"submit_s\": %.3f, \"gpu_wait_s\": %.3f, \"replay_s\": %.3f, \"stage_parse_s\": %.3f, \"stage_grow_s\": %.3f, \"stage_pack_s\": %.3f, \"replay_shards\": %d, \"replay_shard_phase_s\": %.3f, \"reads_per_s\": %.1f, \"total_s\": %.3f, \"gpu_launches\": %llu}\n",
                 (unsigned long long)ex.n_reads, (unsigned long long)ex.n_scanned, bf.reads.size(), dt, ex.threads, a.gpu_inflate ? "true" : "false", ex.t_decode, ex.t_stage, ex.t_submit,
                 ex.t_wait, ex.t_replay, ex.t_stage_parse, ex.t_stage_grow, ex.t_stage_pack, ex.n_shards, ex.t_replay_shards, ex.n_reads / std::max(dt, 1e-9), total,
                 (unsigned long long)(ex.gpu ? strgpu_launch_count(ex.gpu) : 0));
  }
  for (auto &b : batches) ex.free_batch(b);
  if (ex.gpu) strgpu_destroy(ex.gpu);
  return 0;
}

}  // namespace strling
