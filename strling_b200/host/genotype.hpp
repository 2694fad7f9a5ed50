// The rest of `strling call` downstream of the cluster kernels: spanning-read / spanning-pair evidence
// (collect.nim), the smoothed fragment distribution (spanning.nim), the genotyper (genotyper.nim) and the formatting of
// -genotype.txt (genotyper.nim:50-53).  Host code: float modelling and BAM traversal, nothing here belongs on the GPU.
//
// B200-first difference from the reference: the reference runs one indexed BAM query per locus (collect.nim:141); here ALL loci
// are resolved in ONE streaming pass over the BAM (parallel BGZF inflate, loci sorted by window start, binary search per record),
// so `call` needs no .bai and reads the file once.  A locus sees exactly the records its query would return (pos < window stop
// and end > window start, file order), so the evidence is identical.
#pragma once
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "bam.hpp"
#include "pool.hpp"
#include "tread.hpp"

namespace strling {

enum SupportType : uint8_t { kSpanningFragment = 0, kSpanningRead = 1, kOverlappingRead = 2 };  // collect.nim:10-13

struct Support {  // collect.nim:15-31 (the fields the genotyper reads)
  uint8_t type = kOverlappingRead;
  uint32_t frag_len = 0;
  double frag_pct = 0;
  uint8_t rc = 0, ins = 0, dele = 0;
};

// ---- spanning.nim:8-19 : proportion of fragments shorter than the index, smoothed over +-11 bp; float32 throughout
inline std::vector<float> cumulative(const std::array<uint32_t, 4096> &f) {
  std::vector<float> r(4096, 0.0f);
  for (int i = 0; i < 4096; i++) {
    float acc = 0.0f;
    for (int j = std::max(0, i - 11); j <= std::min(i + 11, 4095); j++) acc += (float)f[(size_t)j];
    r[(size_t)i] = acc;
  }
  for (int i = 1; i < 4096; i++) r[(size_t)i] = r[(size_t)i - 1] + r[(size_t)i];
  const float fmax = r[4095];
  for (auto &v : r) v = v / fmax;
  return r;
}

// spanning.nim:21-52
inline double expected_spanning_probability(const std::vector<float> &cd, int start, int stop, bool reverse, int event_start, int event_stop,
                                            int min_spanning_bases = 20) {
  int dist;
  if (start < event_stop - min_spanning_bases) {
    if (reverse) return 0.0;
    dist = event_start - start;
    if (dist < 0) return 0.0;
    if (dist + (event_stop - event_start) < min_spanning_bases) return 0.0;
  } else {
    if (!reverse) return 0.0;
    dist = stop - event_stop;
    if (dist < 0) return 0.0;
    if (dist + (event_stop - event_start) < min_spanning_bases) return 0.0;
  }
  dist += min_spanning_bases;
  dist += event_stop - event_start;
  if (dist < 0 || dist > 4095) return 0.0;
  const float v = 1.0f - cd[(size_t)dist];
  return (double)v;
}

// utils.nim:129-137
inline double frag_percentile(const std::array<uint32_t, 4096> &f, int fragment_length) {
  uint32_t total = 0;
  for (uint32_t c : f) total += c;
  long s = 0;
  for (int i = 0; i < 4096; i++) {
    s += (long)f[(size_t)i];
    if (i >= fragment_length) break;
  }
  return (double)s / (double)std::max<uint32_t>(1u, total);
}

// utils.nim:148-158
inline int median_depth(const std::vector<int> &D) {
  std::vector<long> H(1048, 0);
  for (int d : D) H[(size_t)std::min(d, 1047)]++;
  long s = 0;
  for (int i = 0; i < 1048; i++) {
    s += H[(size_t)i];
    if ((double)s > (double)D.size() / 2.0) return i;
  }
  return 0;
}

// ---- Nim 1.6 CountTable: which key wins `largest` / leads `most_frequent` (genotyper.nim:76-92): the maximal count, ties to the
// lowest slot of the open-addressing table (hashWangYi1, 64 initial slots, linear probing, x2 growth re-inserting in slot order)
inline uint64_t nim_hash_wangyi1(uint64_t x) {
  auto hi_xor_lo = [](uint64_t a, uint64_t b) {
    const __uint128_t p = (__uint128_t)a * (__uint128_t)b;
    return (uint64_t)(p >> 64) ^ (uint64_t)p;
  };
  const uint64_t P0 = 0xa0761d6478bd642fULL, P1 = 0xe7037ed1a0b428dbULL, P58 = 0xeb44accab455d165ULL ^ 8ULL;
  return hi_xor_lo(hi_xor_lo(P0, x ^ P1), P58);
}

// returns false when `keys` is empty
inline bool counttable_top(const std::vector<int64_t> &keys, int64_t &top) {
  size_t cap = 64, counter = 0;
  std::vector<int64_t> sk(cap, 0);
  std::vector<long> sv(cap, 0);
  auto raw_insert = [](std::vector<int64_t> &k, std::vector<long> &v, size_t c, int64_t key, long val) {
    size_t h = (size_t)(nim_hash_wangyi1((uint64_t)key) & (uint64_t)(c - 1));
    while (v[h] != 0) h = (h + 1) & (c - 1);
    k[h] = key;
    v[h] = val;
  };
  for (int64_t key : keys) {
    size_t h = (size_t)(nim_hash_wangyi1((uint64_t)key) & (uint64_t)(cap - 1));
    bool found = false;
    while (sv[h] != 0) {
      if (sk[h] == key) { sv[h]++; found = true; break; }
      h = (h + 1) & (cap - 1);
    }
    if (found) continue;
    if (cap * 2 < counter * 3 || cap - counter < 4) {
      const size_t ncap = cap * 2;
      std::vector<int64_t> nk(ncap, 0);
      std::vector<long> nv(ncap, 0);
      for (size_t i = 0; i < cap; i++)
        if (sv[i] != 0) raw_insert(nk, nv, ncap, sk[i], sv[i]);
      sk.swap(nk);
      sv.swap(nv);
      cap = ncap;
    }
    raw_insert(sk, sv, cap, key, 1);
    counter++;
  }
  if (counter == 0) return false;
  size_t mi = 0;
  for (size_t h = 1; h < cap; h++)
    if (sv[mi] < sv[h]) mi = h;
  top = sk[mi];
  return true;
}

// ---- collect.nim
inline bool cigar_query(int op) { return op == 0 || op == 1 || op == 4 || op == 7 || op == 8; }      // M I S = X
inline bool cigar_reference(int op) { return op == 0 || op == 2 || op == 3 || op == 7 || op == 8; }  // M D N = X

// collect.nim:50-72
inline int find_read_position(const BamRecord &a, int position) {
  int r_off = a.pos, q_off = 0;
  for (int i = 0; i < a.n_cigar; i++) {
    if (r_off > position) return -1;
    const uint32_t c = a.cigar_at(i);
    const int op = BamRecord::op(c), len = (int)BamRecord::oplen(c);
    if (cigar_query(op)) q_off += len;
    if (cigar_reference(op)) r_off += len;
    if (r_off < position) continue;
    const int over = r_off - position;
    if (over > q_off) return -1;
    if (!cigar_query(op)) return -1;
    return q_off - over;
  }
  return -1;
}

// strutils.count(s, sub) with overlapping = false
inline int count_nonoverlap(const std::string &s, const std::string &sub) {
  if (sub.empty()) return 0;
  int n = 0;
  size_t i = 0;
  while ((i = s.find(sub, i)) != std::string::npos) {
    n++;
    i += sub.size();
  }
  return n;
}

inline std::string decode_seq(const BamRecord &a) {
  static const char tbl[] = "=ACMGRSVTWYHKDBN";
  std::string s((size_t)a.l_seq, 'N');
  for (int i = 0; i < a.l_seq; i++) s[(size_t)i] = tbl[(a.seq[i >> 1] >> ((~i & 1) << 2)) & 15];
  return s;
}

// collect.nim:75-95
inline int count_in_bounds(const BamRecord &a, int left, int right, const std::string &repeat) {
  if (right < left) return 0;
  const std::string dna = decode_seq(a);
  int rl = find_read_position(a, left), rr = find_read_position(a, right);
  if (rl >= 0 && rr < 0) rr = (int)dna.size();
  if (rl < 0 && rr < 0) return 0;
  if (rl < 0) rl = 0;
  const std::string S = rr >= rl ? dna.substr((size_t)rl, (size_t)(rr - rl)) : std::string();
  int res = count_nonoverlap(S, repeat);
  if (res < (int)((double)S.size() * 0.7 / (double)repeat.size())) res = 0;
  return res;
}

inline int bounds_slop(int left, int right, const std::string &repeat) {
  int slop = (int)repeat.size() - 1;
  if (right - left < 5) slop += 5 - (right - left);
  return slop;
}

// A Bounds being genotyped, with the evidence its BAM window yields (collect.nim:130-183)
struct GLocus {
  int32_t tid = 0;
  uint32_t left = 0, right = 0;
  std::string repeat;
  uint16_t n_left = 0, n_right = 0;
  // evidence
  int wl = 0, wr = 0, qbeg = 0;
  std::vector<int> depths;
  std::vector<Support> support;
  struct PairRec { int pos, stop, isize; };
  std::unordered_map<std::string, std::vector<PairRec>> pairs;
  std::unordered_map<std::string, size_t> exp_idx;
  std::vector<double> exp_vals;      // insertion order (the reference sums a Table's values: order unpinned)
  bool dead = false;                 // > 20000 read names in the window: the reference gives up on the locus (collect.nim:166-169)
  bool finished = false;             // the stream has passed the window: evidence reduced to support / median depth / expectation
  // results
  int median_depth_v = 0;
  float expected = 0.0f;
};

inline void locus_add_record(GLocus &L, const BamRecord &a, int stop, const std::vector<float> &cd, uint8_t min_mapq, int max_size = 5000) {
  if (L.dead) return;
  if (L.depths.empty()) L.depths.assign((size_t)(L.wr - L.wl), 0);   // allocated when the first record of the window arrives
  if (a.flag & (0x100 | 0x800 | 0x400)) return;
  if (a.mapq < min_mapq) return;
  const int left = (int)L.left, right = (int)L.right;
  const double prob = expected_spanning_probability(cd, a.pos, stop, (a.flag & 0x10) != 0, left, right);
  const std::string qname(a.qname, a.l_qname);
  if (prob > 0) {
    auto it = L.exp_idx.find(qname);
    if (it != L.exp_idx.end()) L.exp_vals[it->second] = 0.5 * (L.exp_vals[it->second] + prob);
    else { L.exp_idx.emplace(qname, L.exp_vals.size()); L.exp_vals.push_back(prob); }
  }
  const int hi = (int)L.depths.size() - 1;
  L.depths[(size_t)std::max(0, a.pos - L.wl - 1)] += 1;
  L.depths[(size_t)std::min(hi, stop - L.wl - 1)] -= 1;
  // overlapping_read (collect.nim:99-121; Record/Bounds overlap: cluster.nim:104-108)
  if (std::max(a.pos, left) <= std::min(stop, right)) {
    Support s;
    s.type = kOverlappingRead;
    s.rc = (uint8_t)count_in_bounds(a, left, right, L.repeat);
    const int slop = bounds_slop(left, right, L.repeat);
    if (a.pos < left - slop && stop > right + slop) {
      s.type = kSpanningRead;
      for (int i = 0; i < a.n_cigar; i++) {
        const uint32_t c = a.cigar_at(i);
        if (BamRecord::op(c) == 1) s.ins = (uint8_t)(s.ins + (uint8_t)BamRecord::oplen(c));
        if (BamRecord::op(c) == 2) s.dele = (uint8_t)(s.dele + (uint8_t)BamRecord::oplen(c));
      }
    }
    L.support.push_back(s);
  }
  if (a.tid != a.mate_tid) return;
  if (std::abs(a.isize) > max_size) return;
  L.pairs[qname].push_back(GLocus::PairRec{a.pos, stop, a.isize});
  if (L.pairs.size() > 20000) {
    L.dead = true;
    L.support.clear();
  }
}

inline void locus_finish(GLocus &L, const std::array<uint32_t, 4096> &frag) {
  if (L.finished) return;
  L.finished = true;
  if (L.dead) {
    L.median_depth_v = -1;
    L.expected = 0.0f;
    L.support.clear();
    std::unordered_map<std::string, std::vector<GLocus::PairRec>>().swap(L.pairs);
    std::unordered_map<std::string, size_t>().swap(L.exp_idx);
    std::vector<int>().swap(L.depths);
    return;
  }
  if (L.depths.empty()) L.depths.assign((size_t)(L.wr - L.wl), 0);   // no record at all: depth 0 everywhere
  float e = 0.0f;
  for (double v : L.exp_vals) e += (float)v;
  L.expected = e;
  const int left = (int)L.left, right = (int)L.right;
  const int slop = bounds_slop(left, right, L.repeat);
  for (const auto &kv : L.pairs) {
    if (kv.second.size() != 2) continue;
    const auto &A = kv.second[0], &B = kv.second[1];   // file order: A.pos <= B.pos (collect.nim:36)
    if (A.pos < left - slop && B.stop > right + slop) {  // spanning_fragment, collect.nim:35-48
      Support s;
      s.type = kSpanningFragment;
      s.frag_len = std::max<uint32_t>(1u, (uint32_t)std::abs(A.isize));
      s.frag_pct = frag_percentile(frag, (int)s.frag_len);
      L.support.push_back(s);
    }
  }
  long run = 0;
  for (auto &d : L.depths) { run += d; d = (int)run; }
  L.median_depth_v = median_depth(L.depths);
  // the window's working set (read names, per-base depth) is released as soon as the locus is done: a 30x `call` with 10^5 loci
  // would otherwise hold every window of the genome until the end of the BAM
  std::unordered_map<std::string, std::vector<GLocus::PairRec>>().swap(L.pairs);
  std::unordered_map<std::string, size_t>().swap(L.exp_idx);
  std::vector<double>().swap(L.exp_vals);
  std::vector<int>().swap(L.depths);
}

// One pass over the BAM for every locus at once.  The reference queries an index per locus (collect.nim:130-146), which only
// works on a coordinate-sorted BAM; this pass needs the same order (it is what lets a locus be finished, and its memory
// released, as soon as the stream has passed its window) and says so when the file is not sorted.
// The pass is chunked and parallel (round 2): a chunk of BGZF blocks is inflated and walked on the thread pool
// (BamChunkReader), one parallel sweep takes (tid, pos, stop) of every record and checks the order, and then every locus whose
// window meets the chunk takes ITS records -- a binary search on the sorted positions gives the range -- in file order.  Loci
// only ever touch their own state, so they are worked off in parallel; a locus sees exactly the records, in exactly the
// order, the record-at-a-time loop gave it.  (A chunk in which a placed record follows a no-coordinate record is not sorted by
// the SAM specification's rule but was tolerated by that loop: such a chunk goes through the loop's logic record by record.)
inline void collect_evidence(const std::string &bam, std::vector<GLocus> &loci, int window, const std::array<uint32_t, 4096> &frag, uint8_t min_mapq,
                             int threads = 0) {
  const std::vector<float> cd = cumulative(frag);
  int n_tid = 0;
  for (auto &L : loci) {
    L.wl = (int)L.left - window;
    L.wr = (int)L.right + window;
    L.qbeg = std::max(0, L.wl);
    n_tid = std::max(n_tid, L.tid + 1);
  }
  std::vector<std::vector<uint32_t>> by_tid((size_t)n_tid);
  for (uint32_t i = 0; i < loci.size(); i++) by_tid[(size_t)loci[i].tid].push_back(i);
  std::vector<int> max_span((size_t)n_tid, 0);
  for (int t = 0; t < n_tid; t++) {
    auto &v = by_tid[(size_t)t];
    std::stable_sort(v.begin(), v.end(), [&](uint32_t a, uint32_t b) { return loci[a].qbeg < loci[b].qbeg; });
    for (uint32_t i : v) max_span[(size_t)t] = std::max(max_span[(size_t)t], loci[i].wr - loci[i].qbeg);
  }
  Pool pool(threads > 0 ? threads : (int)std::max(1u, std::min(64u, std::thread::hardware_concurrency())));
  std::vector<size_t> first_active((size_t)n_tid, 0);   // loci before this index (in qbeg order) are finished
  std::vector<uint32_t> to_finish;
  auto finish_upto = [&](int tid, int pos) {            // every window of `tid` that ends at or before pos
    auto &v = by_tid[(size_t)tid];
    size_t &f = first_active[(size_t)tid];
    while (f < v.size() && (loci[v[f]].finished || loci[v[f]].wr <= pos)) to_finish.push_back(v[f++]);
  };
  auto run_finish = [&]() {
    pool.run(to_finish.size(), [&](size_t k) { locus_finish(loci[to_finish[k]], frag); });
    to_finish.clear();
  };
  uint64_t first_voffset;
  int32_t n_ref;
  {
    BamReader hdr(bam, 1);
    first_voffset = hdr.tell();
    n_ref = (int32_t)hdr.targets().size();
  }
  BamChunkReader rd(bam, first_voffset, pool.size(), &pool, n_ref);
  BamChunk c;
  std::vector<int32_t> tid, pos, stop;
  int last_tid = -1, last_pos = -1;
  auto unsorted = [&](const uint8_t *rec) {
    const BamRecord a = BamChunk::view(rec);
    return std::runtime_error("[strling] call: " + bam + " is not coordinate-sorted (record " + std::string(a.qname, a.l_qname) +
                              "); the reference needs a sorted, indexed BAM here too");
  };
  // test hooks: STRLING_CALL_BLOCKS = BGZF blocks per chunk (many small chunks), STRLING_CALL_SERIAL = every chunk through the
  // record-by-record logic
  static const size_t env_blocks = std::getenv("STRLING_CALL_BLOCKS") ? (size_t)std::atol(std::getenv("STRLING_CALL_BLOCKS")) : 0;
  static const bool force_serial = std::getenv("STRLING_CALL_SERIAL") != nullptr;
  size_t blocks = env_blocks ? env_blocks : 256;
  while (rd.next(c, blocks)) {
    blocks = env_blocks ? env_blocks : 1024;
    const size_t n = c.n_records();
    if (n == 0) continue;
    const uint8_t *data = c.data.data();
    tid.resize(n); pos.resize(n); stop.resize(n);
    const size_t parts = std::max<size_t>(1, std::min<size_t>((size_t)pool.size() * 4, (n + 4095) / 4096));
    struct PartInfo { int64_t bad = -1; bool placed_after_unplaced = false; bool any_unplaced = false; int32_t first_tid = -2, first_pos = 0, last_tid = -2, last_pos = 0; int max_len = 0; };
    std::vector<PartInfo> info(parts);
    pool.ranges(n, parts, [&](size_t lo, size_t hi, size_t part) {
      PartInfo pi;
      for (size_t i = lo; i < hi; i++) {
        if (i + 6 < hi) __builtin_prefetch(data + c.rec_off[i + 6]);
        const BamRecord a = BamChunk::view(data + c.rec_off[i]);
        tid[i] = a.tid;
        pos[i] = a.pos;
        if (a.tid < 0) { stop[i] = a.pos; pi.any_unplaced = true; continue; }
        stop[i] = a.stop();
        pi.max_len = std::max(pi.max_len, stop[i] - a.pos);
        if (pi.any_unplaced) pi.placed_after_unplaced = true;
        if (pi.first_tid == -2) { pi.first_tid = a.tid; pi.first_pos = a.pos; }
        else if (pi.bad < 0 && (a.tid < pi.last_tid || (a.tid == pi.last_tid && a.pos < pi.last_pos))) pi.bad = (int64_t)i;
        pi.last_tid = a.tid;
        pi.last_pos = a.pos;
      }
      info[part] = pi;
    });
    // order across the parts and against the previous chunk; the earliest offending record is the one reported
    bool mixed = force_serial, seen_unplaced = false;
    int max_len = 0;
    int ct = last_tid, cp = last_pos;
    const size_t per = (n + parts - 1) / parts;
    for (size_t q = 0; q < parts; q++) {
      const PartInfo &pi = info[q];
      max_len = std::max(max_len, pi.max_len);
      if (pi.first_tid != -2) {
        if (seen_unplaced) mixed = true;
        if (pi.first_tid < ct || (pi.first_tid == ct && pi.first_pos < cp)) {
          size_t i = q * per;   // the part's first placed record
          while (tid[i] < 0) i++;
          throw unsorted(data + c.rec_off[i]);
        }
        if (pi.bad >= 0) throw unsorted(data + c.rec_off[(size_t)pi.bad]);
        ct = pi.last_tid;
        cp = pi.last_pos;
      }
      if (pi.placed_after_unplaced) mixed = true;
      if (pi.any_unplaced) seen_unplaced = true;
    }
    if (mixed) {   // record by record, as before
      for (size_t i = 0; i < n; i++) {
        if (tid[i] < 0) continue;
        const BamRecord a = BamChunk::view(data + c.rec_off[i]);
        if (a.tid != last_tid)
          for (int t = std::max(0, last_tid); t < std::min(a.tid, n_tid); t++) finish_upto(t, INT32_MAX);
        last_tid = a.tid;
        last_pos = a.pos;
        if (a.tid >= n_tid) continue;
        const auto &v = by_tid[(size_t)a.tid];
        if (v.empty()) continue;
        finish_upto(a.tid, a.pos);
        run_finish();
        const int lo_key = a.pos - max_span[(size_t)a.tid];
        auto it = std::upper_bound(v.begin(), v.end(), lo_key, [&](int key, uint32_t k) { return key < loci[k].qbeg; });
        for (; it != v.end() && loci[*it].qbeg < stop[i]; ++it) {
          GLocus &L = loci[*it];
          if (!L.finished && a.pos < L.wr && stop[i] > L.qbeg) locus_add_record(L, a, stop[i], cd, min_mapq);
        }
      }
      run_finish();
      continue;
    }
    size_t n_placed = n;
    while (n_placed > 0 && tid[n_placed - 1] < 0) n_placed--;   // no-coordinate records only ever trail here
    if (n_placed == 0) continue;
    // one stretch of records per reference sequence
    size_t a0 = 0;
    while (a0 < n_placed) {
      const int t = tid[a0];
      const size_t a1 = (size_t)(std::upper_bound(tid.begin() + (long)a0, tid.begin() + (long)n_placed, t) - tid.begin());
      if (t != last_tid)
        for (int u = std::max(0, last_tid); u < std::min(t, n_tid); u++) finish_upto(u, INT32_MAX);
      last_tid = t;
      last_pos = pos[a1 - 1];
      if (t < n_tid && !by_tid[(size_t)t].empty()) {
        const auto &v = by_tid[(size_t)t];
        finish_upto(t, pos[a0]);   // windows that ended before this stretch began
        // loci whose window can meet a record of the stretch: qbeg < (largest stop) and wr > first pos
        const int reach = pos[a1 - 1] + max_len;
        std::vector<uint32_t> todo;
        for (size_t k = first_active[(size_t)t]; k < v.size() && loci[v[k]].qbeg < reach; k++)
          if (!loci[v[k]].finished && loci[v[k]].wr > pos[a0]) todo.push_back(v[k]);
        pool.run(todo.size(), [&](size_t k) {
          GLocus &L = loci[todo[k]];
          // records with pos < wr and stop > qbeg; stop - pos <= max_len bounds the search from below
          const auto lo_it = std::upper_bound(pos.begin() + (long)a0, pos.begin() + (long)a1, L.qbeg - max_len - 1);
          const auto hi_it = std::lower_bound(pos.begin() + (long)a0, pos.begin() + (long)a1, L.wr);
          for (size_t i = (size_t)(lo_it - pos.begin()); i < (size_t)(hi_it - pos.begin()); i++)
            if (stop[i] > L.qbeg) locus_add_record(L, BamChunk::view(data + c.rec_off[i]), stop[i], cd, min_mapq);
        });
        finish_upto(t, pos[a1 - 1]);
      }
      a0 = a1;
    }
    run_finish();
  }
  to_finish.clear();
  for (uint32_t i = 0; i < loci.size(); i++) to_finish.push_back(i);
  run_finish();   // locus_finish is a no-op for the ones already done
}

// ---- genotyper.nim
struct Call {  // genotyper.nim:25-47
  std::string chrom, repeat;
  uint32_t start = 0, stop = 0;
  double allele1 = 0.0, allele2 = 0.0;
  uint32_t anchored_reads = 0, spanning_reads = 0, spanning_pairs = 0, left_clips = 0, right_clips = 0, sum_str_counts = 0;
  float expected_spanning_fragments = 0.0f, oe_percentile = 0.0f;
  int32_t unplaced_reads = 0;
  double depth = 0.0;
  bool is_large = false;
};

inline double anchored_lm(unsigned long sum_str_counts, double depth) {  // genotyper.nim:113-120
  if (sum_str_counts == 0) return std::nan("");
  const double y = std::log2((double)sum_str_counts / std::max(1.0, depth) + 1.0) * 0.7565329 + 4.3558142;
  return std::pow(2.0, y);
}
inline double unplaced_est(int unplaced_count, double depth) {  // genotyper.nim:131-136
  const double y = std::log2((double)unplaced_count / depth + 1.0) * 0.7595562 + 8.9199168;
  return std::pow(2.0, y);
}

struct GenotypeOpts { uint16_t min_clip = 0, min_clip_total = 0; int min_support = 5, median_fragment_length = 0; };

// tandems: the cluster's STR reads (repeat_count, split, qname)
inline Call genotype(const GLocus &L, const std::string &chrom, const std::vector<const Tread *> &tandems, const GenotypeOpts &o) {  // genotyper.nim:142-196
  Call c;
  c.chrom = chrom;
  c.start = L.left;
  c.stop = L.right;
  c.left_clips = L.n_left;
  c.right_clips = L.n_right;
  c.repeat = L.repeat;
  c.depth = (double)L.median_depth_v;
  const int ru = (int)L.repeat.size();
  if (L.support.empty()) {
    c.allele1 = std::nan("");
  } else {
    std::vector<int64_t> indels;
    uint32_t n_span = 0, n_pairs = 0;
    for (const Support &s : L.support) {
      if (s.type == kSpanningRead) { indels.push_back((int64_t)s.ins - (int64_t)s.dele); n_span++; }
      if (s.type == kSpanningFragment) n_pairs++;
    }
    int64_t top;
    if (counttable_top(indels, top)) c.allele1 = (double)top / (double)std::max(1, ru);   // allele1_bp (genotyper.nim:86-92)
    c.spanning_reads = n_span;
    c.spanning_pairs = n_pairs;
  }
  // evaluated while allele2 is still 0.0, exactly like genotyper.nim:169
  c.is_large = L.n_left >= o.min_clip && L.n_right >= o.min_clip && (uint16_t)(L.n_left + L.n_right) >= o.min_clip_total &&
               (int)tandems.size() >= o.min_support && c.allele2 > (double)o.median_fragment_length;
  unsigned long sum = 0;
  for (const Tread *t : tandems) sum += t->repeat_count;
  c.sum_str_counts = (uint32_t)sum;
  c.allele2 = anchored_lm(sum, c.depth) / (double)std::max(1, ru);
  std::vector<const std::string *> names;
  for (const Tread *t : tandems)
    if (t->split == kNone) names.push_back(&t->qname);
  std::sort(names.begin(), names.end(), [](const std::string *a, const std::string *b) { return *a < *b; });
  c.anchored_reads = (uint32_t)(std::unique(names.begin(), names.end(), [](const std::string *a, const std::string *b) { return *a == *b; }) - names.begin());
  return c;
}

inline std::string fmt2(double x) {  // strformat "{x:.2f}"
  if (std::isnan(x)) return "nan";
  if (std::isinf(x)) return x > 0 ? "inf" : "-inf";
  char b[64];
  std::snprintf(b, sizeof(b), "%.2f", x);
  return b;
}

inline std::string nim_float(double x) {  // Nim `$`(float) for the integral values that reach it here (the depth)
  if (std::isnan(x)) return "nan";
  if (std::isinf(x)) return x > 0 ? "inf" : "-inf";
  char b[64];
  if (x == std::floor(x) && std::fabs(x) < 1e15) std::snprintf(b, sizeof(b), "%.1f", x);
  else std::snprintf(b, sizeof(b), "%.17g", x);
  return b;
}

constexpr const char *kGtHeader =
    "#chrom\tleft\tright\trepeatunit\tallele1_est\tallele2_est\tanchored_reads\tspanning_reads\tspanning_pairs\texpected_spanning_pairs\t"
    "spanning_pairs_pctl\tleft_clips\tright_clips\tunplaced_pairs\tdepth\tsum_str_counts";   // genotyper.nim:50

inline std::string call_line(const Call &c) {  // genotyper.nim:52-53
  std::string s = c.chrom + "\t" + std::to_string(c.start) + "\t" + std::to_string(c.stop) + "\t" + c.repeat + "\t" + fmt2(c.allele1) + "\t" +
                  fmt2(c.allele2) + "\t" + std::to_string(c.anchored_reads) + "\t" + std::to_string(c.spanning_reads) + "\t" +
                  std::to_string(c.spanning_pairs) + "\t" + fmt2((double)c.expected_spanning_fragments) + "\t" + fmt2((double)c.oe_percentile) + "\t" +
                  std::to_string(c.left_clips) + "\t" + std::to_string(c.right_clips) + "\t" + std::to_string(c.unplaced_reads) + "\t" +
                  nim_float(c.depth) + "\t" + std::to_string(c.sum_str_counts);
  return s;
}

inline float oe_ratio(const Call &c) {  // call.nim:31-34
  const float obs = (float)c.spanning_pairs, ex = c.expected_spanning_fragments;
  return (1.0f + obs - ex) / (ex + 1.0f);
}

inline void add_percentile(std::vector<Call> &calls) {  // call.nim:37-47
  std::vector<float> oes;
  oes.reserve(calls.size());
  for (const Call &c : calls) oes.push_back(oe_ratio(c));
  std::sort(oes.begin(), oes.end());
  for (Call &c : calls) {
    const float lb = (float)(std::lower_bound(oes.begin(), oes.end(), oe_ratio(c)) - oes.begin());
    c.oe_percentile = lb / (float)((long)oes.size() - 1);
  }
}

}  // namespace strling
