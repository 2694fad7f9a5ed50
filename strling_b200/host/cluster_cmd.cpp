// `strling merge` (merge.nim:47-187) and the cluster loop of `strling call` (call.nim:50-130,223-235,280-281), with
// grouping, sorting, clustering and bounds on the GPU (strgpu_cluster).  Host work: `.bin` decoding, the fragment
// distribution medians that parameterise the kernels, and writing `-bounds.txt` / `-unplaced.txt`.
// `-l` / `-b` loci take their reads first (assign_reads_locus, callclusters.nim:14-50; `merge -l` on the GPU through
// strgpu_cluster_loci, `call -l / -b` on the host because its genotyper needs the reads themselves).
// `call` then gathers the spanning-read / spanning-pair / depth evidence and genotypes every locus on the host
// (genotype.hpp: collect.nim, spanning.nim, genotyper.nim, call.nim:158-281) and writes `-bounds.txt` with the trailing
// median-depth column (call.nim:255), `-unplaced.txt` and `-genotype.txt`.
#include <algorithm>
#include <cctype>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <numeric>
#include <sstream>
#include <string>
#include <vector>

#include "bam.hpp"
#include "commands.hpp"
#include "genotype.hpp"
#include "strgpu.h"
#include "tread.hpp"

namespace strling {

namespace {

const char *kBoundsHeader = "#chrom\tleft\tright\trepeat\tname\tleft_most\tright_most\tcenter_mass\tn_left\tn_right\tn_total";  // cluster.nim:89

struct Args {
  std::map<std::string, std::string> opt;
  std::vector<std::string> pos;
  bool has(const std::string &k) const { return opt.count(k) > 0; }
  std::string get(const std::string &k, const std::string &d) const { auto it = opt.find(k); return it == opt.end() ? d : it->second; }
};

// options: {"-x", "--long", takes_value}
struct OptSpec { const char *s, *l; bool value; };

Args parse(int argc, char **argv, const std::vector<OptSpec> &specs, const char *usage) {
  Args a;
  for (int i = 0; i < argc; i++) {
    std::string t = argv[i];
    if (t == "-h" || t == "--help") { std::fputs(usage, stdout); std::exit(0); }
    bool matched = false;
    for (const auto &sp : specs) {
      std::string val;
      bool hit = false;
      if (t == sp.s || t == sp.l) {
        hit = true;
        if (sp.value) {
          if (i + 1 >= argc) { std::fprintf(stderr, "option %s needs a value\n", t.c_str()); std::exit(1); }
          val = argv[++i];
        }
      } else if (sp.value && t.rfind(std::string(sp.l) + "=", 0) == 0) {
        hit = true;
        val = t.substr(std::strlen(sp.l) + 1);
      }
      if (hit) { a.opt[sp.l] = sp.value ? val : "1"; matched = true; break; }
    }
    if (!matched) {
      if (t.size() > 1 && t[0] == '-' && !(t[1] >= '0' && t[1] <= '9')) { std::fprintf(stderr, "unknown option %s\n%s", t.c_str(), usage); std::exit(1); }
      a.pos.push_back(t);
    }
  }
  return a;
}

strgpu_tread to_pod(const Tread &t, int32_t sample) {
  strgpu_tread p;
  p.tid = t.tid; p.position = t.position;
  std::memcpy(p.repeat, t.repeat.data(), 6);
  p.flag = t.flag; p.split = t.split; p.mapping_quality = t.mapping_quality; p.repeat_count = t.repeat_count;
  p.align_length = t.align_length; p.sample = sample;
  return p;
}

std::string bounds_line(const strgpu_bounds &b, const std::vector<std::pair<std::string, uint32_t>> &targets) {  // cluster.nim:262-266
  char unit[7] = {0};
  std::memcpy(unit, b.repeat, 6);
  char buf[512];
  std::snprintf(buf, sizeof(buf), "%s\t%u\t%u\t%s\t\t%u\t%u\t%u\t%u\t%u\t%u", targets[(size_t)b.tid].first.c_str(), b.left, b.right, unit,
                b.left_most, b.right_most, b.center_mass, (unsigned)b.n_left, (unsigned)b.n_right, (unsigned)b.n_total);
  return buf;
}

struct LocusLine {  // a Bounds read from a bed / bounds file (cluster.nim:111-163)
  strgpu_locus key;
  uint32_t left = 0, right = 0, center_mass = 0;
  std::string repeat, name;
};

std::vector<std::string> split_ws(const std::string &l) {
  std::vector<std::string> f;
  size_t i = 0;
  while (i < l.size()) {
    while (i < l.size() && std::isspace((unsigned char)l[i])) i++;
    size_t j = i;
    while (j < l.size() && !std::isspace((unsigned char)l[j])) j++;
    if (j > i) f.push_back(l.substr(i, j - i));
    i = j;
  }
  return f;
}

int tid_of(const std::string &name, const std::vector<std::pair<std::string, uint32_t>> &targets) {
  for (size_t t = 0; t < targets.size(); t++)
    if (targets[t].first == name) return (int)t;
  return -1;
}

void check_dna(const std::string &rep, const std::string &line) {
  for (char c : rep)
    if (c != 'A' && c != 'T' && c != 'C' && c != 'G')
      throw std::runtime_error("Error reading loci bed file. Expected DNA (ATCG only) in the 4th field, and got an unexpected character on line: " + line);
}

// parse_bed (cluster.nim:111-141)
std::vector<LocusLine> parse_bed(const std::string &path, const std::vector<std::pair<std::string, uint32_t>> &targets, uint32_t window, int32_t only_tid) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("couldn't open bed file");
  std::vector<LocusLine> out;
  std::string line;
  while (std::getline(in, line)) {
    const auto f = split_ws(line);
    if (f.size() != 4 && f.size() != 5)
      throw std::runtime_error("Error reading loci bed file. Expected 4 or 5 fields and got " + std::to_string(f.size()) + " on line: " + line);
    LocusLine L;
    if (f.size() == 5) L.name = f[4];
    const int tid = tid_of(f[0], targets);
    if (tid < 0) throw std::runtime_error("Error reading loci bed file. Unknown chromosome on line: " + line);
    L.left = (uint32_t)std::stoul(f[1]);
    L.right = (uint32_t)std::stoul(f[2]);
    L.repeat = f[3];
    if (L.repeat.size() > 6)
      throw std::runtime_error("ERROR: STRling currently only supports 1-6 bp repeat units. Input bed contains repeat unit length " + std::to_string(L.repeat.size()) + "\n" + line);
    check_dna(L.repeat, line);
    std::memset(&L.key, 0, sizeof(L.key));
    L.key.tid = tid;
    L.key.left_most = (uint32_t)std::max((int32_t)L.left - (int32_t)window, 0);
    L.key.right_most = std::min(L.right + window, targets[(size_t)tid].second);
    std::memcpy(L.key.repeat, L.repeat.data(), L.repeat.size());
    if (L.left > L.right || L.key.left_most > L.key.right_most) throw std::runtime_error("bad locus: " + line);
    if (only_tid != INT32_MIN && tid != only_tid) continue;
    out.push_back(L);
  }
  return out;
}

// parse_bounds (cluster.nim:143-169)
std::vector<LocusLine> parse_bounds(const std::string &path, const std::vector<std::pair<std::string, uint32_t>> &targets) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("couldn't open bounds file");
  std::vector<LocusLine> out;
  std::string line;
  while (std::getline(in, line)) {
    if (!line.empty() && line[0] == '#') continue;
    std::vector<std::string> f;
    size_t i = 0;
    while (true) {
      const size_t j = line.find('\t', i);
      f.push_back(line.substr(i, j == std::string::npos ? std::string::npos : j - i));
      if (j == std::string::npos) break;
      i = j + 1;
    }
    if (f.size() != 11) throw std::runtime_error("Error reading loci bed file. Expected 11 fields and got " + std::to_string(f.size()) + " on line: " + line);
    LocusLine L;
    const int tid = tid_of(f[0], targets);
    if (tid < 0) throw std::runtime_error("Error reading bounds file. Unknown chromosome on line: " + line);
    L.left = (uint32_t)std::stoul(f[1]);
    L.right = (uint32_t)std::stoul(f[2]);
    L.repeat = f[3];
    L.name = f[4];
    check_dna(L.repeat, line);
    std::memset(&L.key, 0, sizeof(L.key));
    L.key.tid = tid;
    L.key.left_most = (uint32_t)std::stoul(f[5]);
    L.key.right_most = (uint32_t)std::stoul(f[6]);
    L.center_mass = (uint32_t)std::stoul(f[7]);
    std::memcpy(L.key.repeat, L.repeat.data(), std::min<size_t>(6, L.repeat.size()));
    if (L.left > L.right || L.key.left_most > L.key.right_most) throw std::runtime_error("bad bounds line: " + line);
    out.push_back(L);
  }
  return out;
}

std::string locus_line(const LocusLine &L, const std::vector<std::pair<std::string, uint32_t>> &targets) {  // Bounds.tostring
  char buf[640];
  std::snprintf(buf, sizeof(buf), "%s\t%u\t%u\t%s\t%s\t%u\t%u\t%u\t%u\t%u\t%u", targets[(size_t)L.key.tid].first.c_str(), L.left, L.right,
                L.repeat.c_str(), L.name.c_str(), L.key.left_most, L.key.right_most, L.center_mass, (unsigned)L.key.n_left,
                (unsigned)L.key.n_right, (unsigned)L.key.n_total);
  return buf;
}

std::vector<strgpu_bounds> run_cluster(const std::vector<strgpu_tread> &treads, const strgpu_cluster_params &p, int device, bool verbose,
                                       std::vector<LocusLine> *loci = nullptr) {
  strgpu_ctx *gpu = nullptr;
  int rc = strgpu_create(&gpu, device);
  if (rc != STRGPU_OK) throw std::runtime_error(std::string("[strling] gpu: ") + strgpu_error_string(rc));
  std::vector<strgpu_bounds> out(std::max<size_t>(1024, treads.size() / 4));
  uint32_t n_out = 0;
  std::vector<strgpu_locus> keys;
  if (loci)
    for (const auto &l : *loci) keys.push_back(l.key);
  rc = strgpu_cluster_loci(gpu, treads.data(), (uint32_t)treads.size(), &p, keys.data(), (uint32_t)keys.size(), out.data(), (uint32_t)out.size(), &n_out);
  if (rc == STRGPU_ERR_OVERFLOW) {
    out.resize(n_out);
    rc = strgpu_cluster_loci(gpu, treads.data(), (uint32_t)treads.size(), &p, keys.data(), (uint32_t)keys.size(), out.data(), (uint32_t)out.size(), &n_out);
  }
  if (loci)
    for (size_t i = 0; i < keys.size(); i++) (*loci)[i].key = keys[i];
  if (rc != STRGPU_OK) {
    const std::string msg = strgpu_last_error(gpu);
    strgpu_destroy(gpu);
    throw std::runtime_error("[strling] gpu: cluster: " + msg);
  }
  if (verbose) std::fprintf(stderr, "[strling] gpu clustered %zu STR reads into %u records (%llu kernel launches)\n", treads.size(), n_out,
                            (unsigned long long)strgpu_launch_count(gpu));
  strgpu_destroy(gpu);
  out.resize(n_out);
  return out;
}

bool same_targets(const std::vector<std::pair<std::string, uint32_t>> &a, const std::vector<std::pair<std::string, uint32_t>> &b) {
  return a == b;  // unpack.nim:46-56
}

}  // namespace

int merge_main(int argc, char **argv) {
  static const char *usage =
      "strling merge [-f fasta] [-w window] [-m min-support] [--chromosome C] [-c min-clip] [-t min-clip-total] [-q min-mapq]\n"
      "              [-o output-prefix] [-d] [-v] [--device N] <bin>...\n";
  Args a = parse(argc, argv, {{"-f", "--fasta", true}, {"-w", "--window", true}, {"-m", "--min-support", true}, {"", "--chromosome", true},
                              {"-c", "--min-clip", true}, {"-t", "--min-clip-total", true}, {"-q", "--min-mapq", true}, {"-l", "--bed", true},
                              {"-o", "--output-prefix", true}, {"-d", "--diff-refs", false}, {"-v", "--verbose", false}, {"", "--device", true}},
                 usage);
  if (a.pos.empty()) { std::fputs(usage, stdout); return 0; }
  if (a.has("--diff-refs")) throw std::runtime_error("[strling merge] -d/--diff-refs is not part of this build");
  int window = std::stoi(a.get("--window", "-1"));
  const int min_support = std::stoi(a.get("--min-support", "5"));
  const uint16_t min_clip = (uint16_t)std::stoi(a.get("--min-clip", "0"));
  const uint16_t min_clip_total = (uint16_t)std::stoi(a.get("--min-clip-total", "0"));
  const std::string prefix = a.get("--output-prefix", "strling");
  const bool verbose = a.has("--verbose");
  const std::string chromosome = a.get("--chromosome", "-2");

  std::array<uint32_t, 4096> frag{};
  std::vector<std::pair<std::string, uint32_t>> targets;
  std::vector<strgpu_tread> treads;
  int32_t requested_tid = INT32_MIN;
  for (size_t si = 0; si < a.pos.size(); si++) {
    if (verbose) std::fprintf(stderr, "[strling] reading bin file: %s\n", a.pos[si].c_str());
    BinFile bf = read_bin(a.pos[si]);
    auto tg = targets_from_header(bf.header);
    if (targets.empty()) {
      targets = tg;
      if (chromosome != "-2") {  // the reference resolves the name against the fasta index (merge.nim:37-45); same names, same order
        for (size_t t = 0; t < targets.size(); t++)
          if (targets[t].first == chromosome) requested_tid = (int32_t)t;
        if (requested_tid == INT32_MIN) throw std::runtime_error("[strling merge] chromosome: " + chromosome + " not found, check name and 'chr' prefix");
      }
    } else if (!same_targets(tg, targets)) {
      throw std::runtime_error("[strling] Error: inconsistent bam header for " + a.pos[si] + ". Were all samples run on the same reference genome?");
    }
    for (size_t i = 0; i < 4096; i++) {
      const uint32_t before = frag[i];
      frag[i] += bf.frag_dist[i];
      if (frag[i] < before) throw std::runtime_error("overflow");  // merge.nim:112-115
    }
    size_t kept = 0;
    for (const Tread &t : bf.reads) {
      if (requested_tid != INT32_MIN && t.tid != requested_tid) continue;
      if (t.tid < 0) continue;  // drop_unplaced=true (merge.nim:101)
      treads.push_back(to_pod(t, (int32_t)si));  // qname := sample index (merge.nim:121-124)
      kept++;
    }
    std::fprintf(stderr, "[strling] read %zu STR reads from file: %s\n", kept, a.pos[si].c_str());
  }
  if (verbose) {
    std::fprintf(stderr, "[strling] read %zu STR reads across all samples.\n", treads.size());
    std::fprintf(stderr, "[strling] Calculated median fragment length accross all samples:%d\n", frag_median(frag));
  }
  if (window < 0) window = frag_median(frag, 0.98);
  strgpu_cluster_params p;
  p.window = (uint32_t)window;
  p.min_support = min_support;
  p.min_clip = min_clip;
  p.min_clip_total = min_clip_total;
  p.max_clip_dist = (uint16_t)(0.5 * (double)frag_median(frag, 0.5));  // merge.nim:181
  p.merge_mode = 1;
  std::vector<LocusLine> loci;
  if (a.has("--bed")) loci = parse_bed(a.get("--bed", ""), targets, (uint32_t)window, requested_tid);  // merge.nim:154-157
  std::vector<strgpu_bounds> bounds = run_cluster(treads, p, std::stoi(a.get("--device", "0")), verbose, loci.empty() ? nullptr : &loci);

  std::ofstream out(prefix + "-bounds.txt");
  if (!out) throw std::runtime_error("couldn't open output file");
  out << kBoundsHeader << "\n";
  for (const auto &l : loci) out << locus_line(l, targets) << "\n";  // merge.nim:166-168
  for (const auto &b : bounds)
    if (b.tid >= 0) out << bounds_line(b, targets) << "\n";
  out.close();
  if (verbose) std::fprintf(stderr, "[strling] Wrote merged str bounds to %s-bounds.txt\n", prefix.c_str());
  return 0;
}

// The -b / -l part of call_main before clustering (call.nim:158-218 minus the genotyping): fills `loci` (output order), the reads
// each locus takes, and -- when there are loci -- `remaining`, the records left for the cluster loop in `.bin` order.
void prepare_call_loci(const BinFile &bf, const std::vector<std::pair<std::string, uint32_t>> &targets, uint32_t window, const std::string &bounds_path,
                       const std::string &bed_path, std::vector<LocusLine> &loci, std::vector<std::vector<const Tread *>> &loci_reads,
                       std::vector<Tread> &remaining) {
  // call.nim:158-187: -b bounds, then -l loci; a locus that overlaps a bound (same contig and unit) overwrites its name and
  // interval and is consumed (seq.del: the last locus moves into its place); the remaining loci are appended
  if (!bounds_path.empty()) loci = parse_bounds(bounds_path, targets);
  if (!bed_path.empty()) {
    std::vector<LocusLine> bed = parse_bed(bed_path, targets, window, INT32_MIN);
    for (LocusLine &bound : loci)
      for (size_t i = 0; i < bed.size(); i++) {
        const LocusLine &l = bed[i];
        if (l.key.tid == bound.key.tid && l.repeat == bound.repeat && std::max(l.left, bound.left) <= std::min(l.right, bound.right)) {  // cluster.nim:96-100
          bound.name = l.name;
          bound.left = l.left;
          bound.right = l.right;
          bed[i] = bed.back();
          bed.pop_back();
          break;
        }
      }
    loci.insert(loci.end(), bed.begin(), bed.end());
  }
  // assign_reads_locus (callclusters.nim:14-50) on the host, in list order: a locus takes the reads of its (tid, repeat) bucket with
  // left_most - 1 <= position <= right_most; the first read after that window is dropped from the bucket as well (the reference's
  // `ri + 1`).  The genotyper needs the reads themselves, so `call` does this here rather than through strgpu_cluster_loci.
  loci_reads.assign(loci.size(), {});
  if (!loci.empty()) {
    std::map<std::pair<int32_t, std::array<char, 6>>, std::vector<uint32_t>> buckets;
    for (uint32_t i = 0; i < bf.reads.size(); i++) buckets[{bf.reads[i].tid, bf.reads[i].repeat}].push_back(i);
    for (auto &kv : buckets)
      std::stable_sort(kv.second.begin(), kv.second.end(), [&](uint32_t x, uint32_t y) { return bf.reads[x].position < bf.reads[y].position; });
    for (size_t li = 0; li < loci.size(); li++) {
      LocusLine &L = loci[li];
      std::array<char, 6> ru{{0, 0, 0, 0, 0, 0}};
      std::memcpy(ru.data(), L.repeat.data(), std::min<size_t>(6, L.repeat.size()));
      L.key.n_left = L.key.n_right = L.key.n_total = 0;
      auto it = buckets.find({L.key.tid, ru});
      if (it == buckets.end() || it->second.empty()) continue;
      std::vector<uint32_t> &trs = it->second;
      const uint32_t lm = L.key.left_most == 0 ? 0u : L.key.left_most - 1u;
      const size_t lo = (size_t)(std::lower_bound(trs.begin(), trs.end(), lm, [&](uint32_t x, uint32_t v) { return bf.reads[x].position < v; }) - trs.begin());
      const size_t hi = (size_t)(std::upper_bound(trs.begin(), trs.end(), L.key.right_most, [&](uint32_t v, uint32_t x) { return v < bf.reads[x].position; }) - trs.begin());
      for (size_t j = lo; j < hi; j++) {
        const Tread &r = bf.reads[trs[j]];
        loci_reads[li].push_back(&r);
        L.key.n_total++;
        if (r.split == kRight) L.key.n_right++;
        else if (r.split == kLeft) L.key.n_left++;
      }
      std::vector<uint32_t> keep(trs.begin(), trs.begin() + (long)lo);
      if (hi + 1 < trs.size()) keep.insert(keep.end(), trs.begin() + (long)hi + 1, trs.end());   // `if ri < trs.high: add trs[ri + 1 ..]`
      trs.swap(keep);
    }
    std::vector<uint32_t> left_over;
    for (auto &kv : buckets) left_over.insert(left_over.end(), kv.second.begin(), kv.second.end());
    std::sort(left_over.begin(), left_over.end());    // back to .bin order
    remaining.reserve(left_over.size());
    for (uint32_t i : left_over) remaining.push_back(bf.reads[i]);
  }
}

// call.nim:223-281 downstream of the cluster kernels: spanning evidence, genotypes, the three output files.  `bounds` is what
// strgpu_cluster returned for bf.reads (ascending tid, repeat, position; unplaced buckets as tid == -1 records).
// `loci` (may be null): the -b / -l Bounds with the reads assign_reads_locus gave them (call.nim:189-218); they are genotyped and
// written first.  `reads` is the record array the cluster records index (all of the .bin, or what the loci left over).
void write_call_outputs(const std::vector<strgpu_bounds> &bounds, const std::vector<Tread> &reads, const std::vector<std::pair<std::string, uint32_t>> &targets,
                        const std::array<uint32_t, 4096> &frag, const strgpu_cluster_params &p, uint8_t min_mapq, const std::string &bam,
                        std::ostream &bo, std::ostream &un, std::ostream &gt, bool verbose,
                        const std::vector<LocusLine> *loci = nullptr, const std::vector<std::vector<const Tread *>> *loci_reads = nullptr) {
  bo << kBoundsHeader << "\tdepth\n";   // call.nim:145

  // The cluster's reads (call.nim:225 `c.reads`) = n_reads records from first_read of the (tid, repeat, position)-sorted order.
  std::vector<uint32_t> order(reads.size());
  std::iota(order.begin(), order.end(), 0u);
  std::stable_sort(order.begin(), order.end(), [&](uint32_t x, uint32_t y) {
    const Tread &A = reads[x], &B = reads[y];
    if (A.tid != B.tid) return A.tid < B.tid;
    const int c = std::memcmp(A.repeat.data(), B.repeat.data(), 6);
    if (c != 0) return c < 0;
    return A.position < B.position;
  });
  // spanning evidence for every discovered locus in one pass over the BAM (collect.nim:130-183)
  std::vector<GLocus> gl;
  std::vector<size_t> gl_bound;      // index into `bounds`, or into `loci` when gl_locus is set
  std::vector<char> gl_locus;
  std::map<std::string, uint32_t> unplaced_counts;
  if (loci)
    for (size_t i = 0; i < loci->size(); i++) {
      const LocusLine &ll = (*loci)[i];
      if (ll.right - ll.left > 1000u) {   // call.nim:192-194
        std::fprintf(stderr, "large bounds:%s skipping\n", locus_line(ll, targets).c_str());
        continue;
      }
      GLocus L;
      L.tid = ll.key.tid; L.left = ll.left; L.right = ll.right; L.repeat = ll.repeat; L.n_left = ll.key.n_left; L.n_right = ll.key.n_right;
      gl.push_back(std::move(L));
      gl_bound.push_back(i);
      gl_locus.push_back(1);
    }
  for (size_t i = 0; i < bounds.size(); i++) {
    const strgpu_bounds &b = bounds[i];
    char unit[7] = {0};
    std::memcpy(unit, b.repeat, 6);
    if (b.tid < 0) {
      un << unit << "\t" << b.n_reads << "\n";  // call.nim:226-228,280-281
      unplaced_counts[unit] = b.n_reads;
      continue;
    }
    GLocus L;
    L.tid = b.tid; L.left = b.left; L.right = b.right; L.repeat = unit; L.n_left = b.n_left; L.n_right = b.n_right;
    gl.push_back(std::move(L));
    gl_bound.push_back(i);
    gl_locus.push_back(0);
  }
  const auto ev0 = std::chrono::steady_clock::now();
  collect_evidence(bam, gl, (int)p.window, frag, min_mapq);
  if (std::getenv("STRLING_CALL_TIMING"))
    std::fprintf(stderr, "[strling] call: evidence pass over the BAM for %zu loci: %.3f s\n", gl.size(),
                 std::chrono::duration<double>(std::chrono::steady_clock::now() - ev0).count());
  GenotypeOpts go;
  go.min_clip = p.min_clip; go.min_clip_total = p.min_clip_total; go.min_support = p.min_support; go.median_fragment_length = frag_median(frag);
  std::vector<Call> calls;
  std::vector<std::string> canon;
  for (size_t k = 0; k < gl.size(); k++) {
    const GLocus &L = gl[k];
    if (L.support.size() > 5000 || L.median_depth_v == -1) continue;   // call.nim:196-202,236-241
    const bool is_locus = gl_locus[k] != 0;
    std::vector<const Tread *> tandems;
    if (is_locus) {
      tandems = (*loci_reads)[gl_bound[k]];
    } else {
      const strgpu_bounds &b = bounds[gl_bound[k]];
      tandems.reserve(b.n_reads);
      for (uint32_t j = 0; j < b.n_reads; j++) tandems.push_back(&reads[order[(size_t)b.first_read + j]]);
    }
    Call c = genotype(L, targets[(size_t)L.tid].first, tandems, go);
    c.expected_spanning_fragments = L.expected;
    std::array<char, 6> ru{{0, 0, 0, 0, 0, 0}};
    std::memcpy(ru.data(), L.repeat.data(), std::min<size_t>(6, L.repeat.size()));
    const std::array<char, 6> cr = canonical_repeat(ru);
    canon.emplace_back(cr.data(), (size_t)unit_length(cr));
    calls.push_back(std::move(c));
    if (is_locus) bo << locus_line((*loci)[gl_bound[k]], targets) << "\t" << L.median_depth_v << "\n";   // call.nim:214
    else bo << bounds_line(bounds[gl_bound[k]], targets) << "\t" << L.median_depth_v << "\n";                   // call.nim:255
  }
  add_percentile(calls);
  // call.nim:266-278: per canonical repeat unit; a lone "large" genotype takes the unit's unplaced-read count.  (The reference
  // walks a Nim Table here, i.e. in hash order; we write ascending by canonical unit, discovery order inside a unit.)
  std::map<std::string, std::vector<size_t>> by_rep;
  for (size_t i = 0; i < calls.size(); i++) by_rep[canon[i]].push_back(i);
  for (auto &kv : by_rep) {
    std::vector<size_t> large;
    for (size_t i : kv.second)
      if (calls[i].is_large) { large.push_back(i); if (large.size() > 1) break; }
    if (large.size() == 1) {
      Call &c = calls[large[0]];
      const auto it = unplaced_counts.find(kv.first);
      const int n_un = it == unplaced_counts.end() ? 0 : (int)it->second;
      c.unplaced_reads = n_un;                                                   // genotyper.nim:190-196
      if (n_un > 2) c.allele2 = unplaced_est(n_un, c.depth) / (double)c.repeat.size();
    }
    for (size_t i : kv.second) gt << call_line(calls[i]) << "\n";
  }
  if (verbose) std::fprintf(stderr, "[strling] genotyped %zu of %zu loci\n", calls.size(), gl.size());
}

int call_main(int argc, char **argv) {
  static const char *usage =
      "strling call [-f fasta] [-m min-support] [-c min-clip] [-t min-clip-total] [-q min-mapq] [-o output-prefix] [-v] [--device N] <bam> <bin>\n"
      "  writes <prefix>-bounds.txt, <prefix>-unplaced.txt and <prefix>-genotype.txt (genotypes only without -l / -b)\n";
  Args a = parse(argc, argv, {{"-f", "--fasta", true}, {"-m", "--min-support", true}, {"-c", "--min-clip", true}, {"-t", "--min-clip-total", true},
                              {"-q", "--min-mapq", true}, {"-l", "--loci", true}, {"-b", "--bounds", true}, {"-o", "--output-prefix", true},
                              {"-v", "--verbose", false}, {"", "--device", true}},
                 usage);
  if (a.pos.size() != 2) { std::fputs(usage, stdout); return a.pos.empty() ? 0 : 1; }
  const std::string prefix = a.get("--output-prefix", "strling");
  const bool verbose = a.has("--verbose");
  // call.nim:96-114 : the fragment distribution is re-derived from the BAM, window = its 0.99 quantile
  const std::array<uint32_t, 4096> frag = fragment_length_distribution(a.pos[0], 0);
  if (verbose) std::fprintf(stderr, "Calculated median fragment length:%d\n", frag_median(frag));
  BinFile bf = read_bin(a.pos[1]);
  auto targets = targets_from_header(bf.header);
  {
    BamReader rd(a.pos[0]);
    std::vector<std::pair<std::string, uint32_t>> bt;
    for (const auto &t : rd.targets()) bt.emplace_back(t.name, t.length);
    if (!same_targets(bt, targets)) throw std::runtime_error("[strling] bin file and bam have different targets (call.nim:122)");
  }
  std::vector<strgpu_tread> treads;
  treads.reserve(bf.reads.size());
  for (const Tread &t : bf.reads) treads.push_back(to_pod(t, 0));
  strgpu_cluster_params p;
  p.window = (uint32_t)frag_median(frag, 0.99);
  p.min_support = std::stoi(a.get("--min-support", "5"));
  p.min_clip = (uint16_t)std::stoi(a.get("--min-clip", "0"));
  p.min_clip_total = (uint16_t)std::stoi(a.get("--min-clip-total", "0"));
  p.max_clip_dist = (uint16_t)(0.5 * (double)frag_median(frag, 0.5));  // call.nim:232
  p.merge_mode = 0;
  std::vector<LocusLine> loci;
  std::vector<std::vector<const Tread *>> loci_reads;
  std::vector<Tread> remaining;
  prepare_call_loci(bf, targets, p.window, a.get("--bounds", ""), a.get("--loci", ""), loci, loci_reads, remaining);
  const std::vector<Tread> *cluster_reads = &bf.reads;
  if (!loci.empty()) {
    cluster_reads = &remaining;
    treads.clear();
    for (const Tread &t : remaining) treads.push_back(to_pod(t, 0));
  }
  std::vector<strgpu_bounds> bounds = run_cluster(treads, p, std::stoi(a.get("--device", "0")), verbose, nullptr);

  std::ofstream bo(prefix + "-bounds.txt"), un(prefix + "-unplaced.txt"), gt(prefix + "-genotype.txt");
  if (!bo || !un || !gt) throw std::runtime_error("couldn't open output file");
  gt << kGtHeader << "\n";
  write_call_outputs(bounds, *cluster_reads, targets, frag, p, (uint8_t)std::stoi(a.get("--min-mapq", "40")), a.pos[0], bo, un, gt, verbose,
                     loci.empty() ? nullptr : &loci, loci.empty() ? nullptr : &loci_reads);
  return 0;
}

// `strling debug genotype <bam> <bin> <clusters.tsv> <prefix> <window> <min_support> <min_mapq>`: the host half of `call` with the
// cluster records read from a file (tid left right repeat left_most right_most center_mass n_left n_right n_total first_read
// n_reads per line) instead of coming from the GPU, so that it can be compared with the oracle on a machine without one.
int debug_genotype(int argc, char **argv) {
  if (argc != 7 && argc != 9) { std::fprintf(stderr, "usage: debug genotype bam bin clusters.tsv prefix window min_support min_mapq [bounds|- bed|-]\n"); return 1; }
  const std::array<uint32_t, 4096> frag = fragment_length_distribution(argv[0], 0);
  BinFile bf = read_bin(argv[1]);
  auto targets = targets_from_header(bf.header);
  std::vector<strgpu_bounds> bounds;
  std::ifstream in(argv[2]);
  std::string line;
  while (std::getline(in, line)) {
    if (line.empty()) continue;
    std::istringstream ss(line);
    strgpu_bounds b;
    std::memset(&b, 0, sizeof(b));
    std::string rep;
    unsigned nl, nr, nt;
    ss >> b.tid >> b.left >> b.right >> rep >> b.left_most >> b.right_most >> b.center_mass >> nl >> nr >> nt >> b.first_read >> b.n_reads;
    b.n_left = (uint16_t)nl; b.n_right = (uint16_t)nr; b.n_total = (uint16_t)nt;
    std::memcpy(b.repeat, rep.data(), std::min<size_t>(6, rep.size()));
    bounds.push_back(b);
  }
  strgpu_cluster_params p;
  std::memset(&p, 0, sizeof(p));
  p.window = (uint32_t)std::stoi(argv[4]);
  p.min_support = std::stoi(argv[5]);
  const std::string prefix = argv[3];
  std::ofstream bo(prefix + "-bounds.txt"), un(prefix + "-unplaced.txt"), gt(prefix + "-genotype.txt");
  if (!bo || !un || !gt) throw std::runtime_error("couldn't open output file");
  gt << kGtHeader << "\n";
  std::vector<LocusLine> loci;
  std::vector<std::vector<const Tread *>> loci_reads;
  std::vector<Tread> remaining;
  if (argc == 9)
    prepare_call_loci(bf, targets, p.window, std::string(argv[7]) == "-" ? "" : argv[7], std::string(argv[8]) == "-" ? "" : argv[8], loci, loci_reads, remaining);
  write_call_outputs(bounds, loci.empty() ? bf.reads : remaining, targets, frag, p, (uint8_t)std::stoi(argv[6]), argv[0], bo, un, gt, true,
                     loci.empty() ? nullptr : &loci, loci.empty() ? nullptr : &loci_reads);
  return 0;
}

int extract_main(int argc, char **argv) {
  static const char *usage =
      "strling extract [-f fasta] [-g genome-repeats] [-p proportion-repeat=0.8] [-q min-mapq=40] [-v] [--device N] [--threads N]\n"
      "                [--batch-reads N] [--replay-shards N] [--gpu-inflate] <bam> <bin>\n"
      "  --gpu-inflate: the BGZF blocks are inflated on the GPU instead of the host threads (same .bin)\n"
      "  <bam> must be coordinate-sorted with the no-coordinate reads at the end (as `samtools sort` writes it).  Reads longer\n"
      "  than 510 bases are refused: beyond that the reference's uint8 k-mer count tables wrap (utils.nim:113-117).\n";
  Args a = parse(argc, argv, {{"-f", "--fasta", true}, {"-g", "--genome-repeats", true}, {"-p", "--proportion-repeat", true}, {"-q", "--min-mapq", true},
                              {"-v", "--verbose", false}, {"", "--device", true}, {"", "--threads", true}, {"", "--batch-reads", true}, {"", "--replay-shards", true}, {"", "--gpu-inflate", false}},
                 usage);
  if (a.pos.size() != 2) { std::fputs(usage, stdout); return a.pos.empty() ? 0 : 1; }
  ExtractArgs e;
  e.fasta = a.get("--fasta", "");
  e.genome_repeats = a.get("--genome-repeats", "");
  e.proportion_repeat = std::stod(a.get("--proportion-repeat", "0.8"));
  e.min_mapq = std::stoi(a.get("--min-mapq", "40"));
  e.verbose = a.has("--verbose");
  e.device = std::stoi(a.get("--device", "0"));
  e.threads = std::stoi(a.get("--threads", "0"));
  e.batch_reads = (uint32_t)std::stoul(a.get("--batch-reads", "262144"));
  e.replay_shards = std::stoi(a.get("--replay-shards", "0"));
  e.gpu_inflate = a.has("--gpu-inflate");
  e.bam = a.pos[0];
  e.bin = a.pos[1];
  return extract_run(e);
}

}  // namespace strling
