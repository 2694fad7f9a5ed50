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
  bool skip, clip_l, clip_r, primary;
};

struct Batch {
  BamChunk chunk;                // owns the decoded records (qname / SEQ are used in place)
  std::vector<Pending> recs;     // one per record of the chunk, file order
  std::vector<uint64_t> base_off;
  std::vector<uint32_t> seg_off;
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

struct Extractor {
  strgpu_ctx *gpu = nullptr;
  Options opts;
  double p = 0.8;
  const std::vector<Target> *targets = nullptr;
  std::vector<const Intervals *> genome_str_by_tid;  // nullptr: chrom not in genome_str
  std::unordered_map<std::string, Tread> tbl;        // Cache.tbl (extract.nim:89-91)
  std::vector<Tread> cache;                          // Cache.cache
  uint64_t n_reads = 0, n_scanned = 0, n_warned = 0;
  double t_wait = 0, t_replay = 0, t_submit = 0, t_decode = 0, t_stage = 0;  // host stage timers (seconds)
  bool verbose = false;
  int threads = 1;
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

  template <typename F>
  void parallel_for(size_t n, F f) {  // f(begin, end, thread)
    const int nt = (int)std::min<size_t>((size_t)threads, (n + 4095) / 4096);
    if (nt <= 1) { f((size_t)0, n, 0); return; }
    std::vector<std::thread> th;
    std::vector<std::string> errs((size_t)nt);
    const size_t per = (n + (size_t)nt - 1) / (size_t)nt;
    for (int t = 0; t < nt; t++) {
      const size_t a = (size_t)t * per, e = std::min(n, a + per);
      if (a >= e) break;
      th.emplace_back([&, a, e, t]() {
        try { f(a, e, t); } catch (const std::exception &ex) { errs[(size_t)t] = ex.what(); }
      });
    }
    for (auto &x : th) x.join();
    for (auto &e : errs)
      if (!e.empty()) throw std::runtime_error(e);
  }

  // Decode every record of the chunk and lay the batch out: phase 1 (parallel) parses the fields and decides which
  // segments a record contributes; a prefix sum fixes every record's place in seq2 / segs; phase 2 (parallel) packs the
  // SEQ fields and writes the descriptors.
  void stage(Batch &b) {
    const size_t n = b.chunk.n_records();
    b.recs.resize(n);
    b.base_off.resize(n + 1);
    b.seg_off.resize(n + 1);
    const uint8_t *data = b.chunk.data.data();  // RawBuffer
    parallel_for(n, [&](size_t lo, size_t hi, int) {
      for (size_t i = lo; i < hi; i++) {
        const BamRecord r = BamChunk::view(data + b.chunk.rec_off[i]);
        Pending &pr = b.recs[i];
        pr.tid = r.tid; pr.pos = r.pos; pr.stop = r.stop(); pr.mate_tid = r.mate_tid; pr.mate_pos = r.mate_pos;
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
        if (!pr.primary) continue;
        if (r.l_seq > STRGPU_MAX_SEGMENT_LEN)
          throw std::runtime_error("[strling] read longer than " + std::to_string(STRGPU_MAX_SEGMENT_LEN) + " bp: " + std::string(r.qname));
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
      }
    });
    uint64_t nb = 0;
    uint32_t ns = 0;
    for (size_t i = 0; i < n; i++) {
      b.base_off[i] = nb;
      b.seg_off[i] = ns;
      nb += b.recs[i].aligned_bases;
      ns += b.recs[i].n_seg;
    }
    b.base_off[n] = nb;
    b.seg_off[n] = ns;
    if (nb > 0xffffff00ull) throw std::runtime_error("[strling] batch too large: lower --batch-reads");
    grow_batch(b, nb + 64, ns + 8);
    b.n_bases = nb;
    b.n_seg = ns;
    std::vector<uint32_t> tmax((size_t)threads + 1, 0);
    std::vector<uint8_t> tany((size_t)threads + 1, 0);
    parallel_for(n, [&](size_t lo, size_t hi, int t) {
      uint32_t mx = 0;
      bool any = false;
      for (size_t i = lo; i < hi; i++) {
        Pending &pr = b.recs[i];
        if (!pr.n_seg) continue;
        const BamRecord r = BamChunk::view(data + b.chunk.rec_off[i]);
        const uint64_t base = b.base_off[i];
        const int n_other = strgpu_pack_bam4(r.seq, (uint32_t)r.l_seq, b.seq2, b.nmask, b.xmask, base);
        if (n_other < 0) throw std::runtime_error("[strling] pack_bam4 failed");
        const bool has_n = n_other > 0;
        any = any || has_n;
        uint32_t si = b.seg_off[i];
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
      }
      tmax[(size_t)t] = mx;
      tany[(size_t)t] = any;
    });
    b.max_len = *std::max_element(tmax.begin(), tmax.end());
    b.any_n = std::any_of(tany.begin(), tany.end(), [](uint8_t v) { return v != 0; });
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

  // ---- replay: extract.nim:63-132,192-248 with scan results looked up instead of computed
  Tread to_tread(const Batch &b, const Pending &r) {
    Tread t;
    t.tid = r.tid;
    t.position = (uint32_t)std::max(0, r.pos);
    t.flag = r.flag;
    t.split = kNone;
    t.mapping_quality = r.mapq;
    t.qname.assign(r.qname, r.qname_len);
    int align_length = r.m_len, repeat_count = 0;
    if (r.seg_full >= 0) {
      const strgpu_repeat &res = b.out[r.seg_full];
      std::memcpy(t.repeat.data(), res.unit, 6);
      repeat_count = res.repeat_count;
      align_length = r.l_seq;
    }
    if (repeat_count >= 256) throw std::runtime_error("[strling] repeat_count >= 256 for read " + t.qname);  // doAssert, extract.nim:72
    t.repeat_count = (uint8_t)repeat_count;
    t.align_length = (uint8_t)align_length;
    if (r.n_cigar > 1 && BamRecord::op(r.first_cig) == 4 && BamRecord::oplen(r.first_cig) > 16) t.split = kNoneLeft;
    if (r.n_cigar > 1 && BamRecord::op(r.last_cig) == 4 && BamRecord::oplen(r.last_cig) > 16) t.split = kNoneRight;
    return t;
  }

  void add_soft(const Batch &b, const Pending &r, int cls_idx, const std::array<char, 6> &read_repeat, const std::string &qname) {
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
      Tread tr;
      tr.tid = r.tid;
      tr.position = (uint32_t)(is_left ? std::max(0, r.pos) : std::max(0, r.stop));
      tr.flag = r.flag;
      std::memcpy(tr.repeat.data(), res.unit, 6);
      tr.repeat_count = (uint8_t)res.repeat_count;
      tr.align_length = (uint8_t)std::min<uint32_t>(clen, (uint32_t)r.l_seq);
      tr.split = is_left ? kLeft : kRight;
      tr.mapping_quality = r.mapq;
      tr.qname = qname;
      if (p_repeat(tr) < 0.9) continue;
      cache.push_back(std::move(tr));
    }
  }

  void add(const Batch &b, const Pending &r) {
    std::string qname(r.qname, r.qname_len);
    auto it = tbl.end();
    bool after_mate = r.tid > r.mate_tid;
    if (!after_mate && r.tid == r.mate_tid) {
      if (r.pos > r.mate_pos) after_mate = true;
      else if (r.pos == r.mate_pos) { it = tbl.find(qname); after_mate = it != tbl.end(); }
    }
    if (after_mate) {
      if (it == tbl.end()) it = tbl.find(qname);
      if (it == tbl.end()) return;
      Tread mate = std::move(it->second);
      tbl.erase(it);
      Tread self = to_tread(b, r);
      add_soft(b, r, 1, self.repeat, qname);  // opts.proportion_repeat = min(p, 0.6)
      if (mate.repeat_count == 0 && self.repeat_count == 0) return;
      if (unplaced_pair(self, mate, opts)) {
        if (self.repeat[0] == 0 || mate.repeat[0] == 0) return;
        self.repeat = canonical_repeat(self.repeat);
        self.position = 0;
        self.tid = -1;
        mate.repeat = canonical_repeat(mate.repeat);
        mate.position = 0;
        mate.tid = -1;
        cache.push_back(std::move(self));
        cache.push_back(std::move(mate));
        return;
      }
      const uint32_t mp = mate.position;
      if (adjust_by(mate, self, opts, self.position)) cache.push_back(mate);
      if (adjust_by(self, mate, opts, mp)) cache.push_back(std::move(self));
    } else {
      Tread tr = to_tread(b, r);
      add_soft(b, r, 0, tr.repeat, qname);  // opts.proportion_repeat = p - 0.07
      auto ins = tbl.emplace(qname, std::move(tr));
      if (!ins.second) {  // hasKeyOrPut found the key: warn and drop it (extract.nim:245-248)
        if (n_warned++ < 20)
          std::fprintf(stderr, "[strling] warning. bad read (this happens with bwa-kit alignments):%s already in table\n", qname.c_str());
        tbl.erase(ins.first);
      }
    }
  }

  void replay(const Batch &b) {
    for (const Pending &r : b.recs)
      if (r.primary) add(b, r);
  }
};

}  // namespace

// utils.nim:86-111 with n_reads = 2_000_000, skip_reads = 100_000 (extract.nim:273, call.nim:76)
std::array<uint32_t, 4096> fragment_length_distribution(const std::string &bam, int threads) {
  std::array<uint32_t, 4096> frag{};
  BamReader rd(bam, threads);
  BamRecord r;
  int64_t i = -1, counted = 0;
  std::vector<int32_t> skipped;
  while (rd.next(r)) {
    i++;
    if (!(r.flag & 0x2)) continue;
    if (r.flag & (0x800 | 0x100)) continue;
    if (r.isize < 0 || r.isize > 4095) continue;
    if (i < 100000) { skipped.push_back(r.isize); continue; }
    skipped.clear();
    frag[(size_t)r.isize]++;
    if (++counted > 2000000) break;
  }
  uint64_t sum = 0;
  for (uint32_t c : frag) sum += c;
  if (sum == 0) {
    std::fprintf(stderr, "using first reads in fragment_length_distribution calculation as there were not enough\n");
    for (int32_t s : skipped) frag[(size_t)s]++;
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

  ex.threads = a.threads > 0 ? a.threads : (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
  const uint64_t first_voffset = rd.tell();
  constexpr int kBatches = STRGPU_SLOTS + 1;  // one being inflated, one being staged, up to two on the GPU / in replay
  Batch batches[kBatches];
  for (auto &b : batches) ex.alloc_batch(b, (uint64_t)a.batch_reads * 160 + 4096, (uint32_t)std::min<uint64_t>((uint64_t)a.batch_reads * 2 + 64, 0xfffffff0u));
  const size_t blocks_per_chunk = std::max<size_t>(16, (size_t)a.batch_reads / 200);  // ~64 KiB blocks of ~300-byte records

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
        for (const Pending &r : b.recs) {
          if (!r.primary) continue;
          if (!first_pass && r.tid >= 0) continue;   // ibam.query("*") returns no-coordinate records only
          if (first_pass && r.tid != tid_seen && r.tid >= 0) {
            if (rd.targets()[(size_t)r.tid].length > 2000000u)
              std::fprintf(stderr, "[strling] extracting chromosome:%s\n", rd.targets()[(size_t)r.tid].name.c_str());
            tid_seen = r.tid;
          }
          ex.n_reads++;
          if (ex.verbose && ex.n_reads % 10000000 == 0) {
            const double dt = std::chrono::duration<double>(clk::now() - t0).count();
            std::fprintf(stderr, "%llu %.1f reads/sec tbl len: %zu cache len: %zu\n", (unsigned long long)ex.n_reads, ex.n_reads / dt,
                         ex.tbl.size(), ex.cache.size());
          }
          ex.add(b, r);
        }
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
        BamChunkReader reader(a.bam, voffset, ex.threads);
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
        ex.stage(b);
        if (pass1 && !have_tail)
          for (size_t i = 0; i < b.recs.size(); i++)
            if (b.recs[i].tid < 0) { have_tail = true; tail_voffset = b.chunk.voffset_of(i); break; }
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
    std::fprintf(stderr, "[strling] perf: {\"reads\": %llu, \"segments_scanned\": %llu, \"str_reads\": %zu, \"scan_pass_s\": %.3f, \"threads\": %d, \"inflate_s\": %.3f, \"stage_s\": %.3f, \"submit_s\": %.3f, \"gpu_wait_s\": %.3f, \"replay_s\": %.3f, \"reads_per_s\": %.1f, \"total_s\": %.3f, \"gpu_launches\": %llu}\n",
                 (unsigned long long)ex.n_reads, (unsigned long long)ex.n_scanned, bf.reads.size(), dt, ex.threads, ex.t_decode, ex.t_stage, ex.t_submit,
                 ex.t_wait, ex.t_replay, ex.n_reads / std::max(dt, 1e-9), total,
                 (unsigned long long)(ex.gpu ? strgpu_launch_count(ex.gpu) : 0));
  }
  for (auto &b : batches) ex.free_batch(b);
  if (ex.gpu) strgpu_destroy(ex.gpu);
  return 0;
}

}  // namespace strling
