// `strling index` (genome_strs.nim:61-131,169-199): STR-like regions of a reference genome.  Every 100-bp window
// (step 60) of every chromosome is one segment of a libstrgpu scan batch -- segments overlap and start at any base --
// and the host then merges adjacent same-unit windows and trims them exactly like the reference (window merge
// :75-86, trim :22-59).  Also used by `strling extract` when the -g file does not exist yet (genome_strs.nim:117-128).
#include <algorithm>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <string>
#include <vector>

#include "commands.hpp"
#include "strgpu.h"
#include "tread.hpp"

namespace strling {

namespace {

struct Chrom {
  std::string name, seq;
};

std::vector<Chrom> read_fasta(const std::string &path) {
  std::ifstream in(path);
  if (!in) throw std::runtime_error("[strling] couldn't open fasta " + path + " make sure file is present and has a .fai index");
  std::vector<Chrom> out;
  std::string line;
  while (std::getline(in, line)) {
    if (!line.empty() && line.back() == '\r') line.pop_back();
    if (!line.empty() && line[0] == '>') {
      size_t e = 1;
      while (e < line.size() && !std::isspace((unsigned char)line[e])) e++;
      out.push_back(Chrom{line.substr(1, e - 1), ""});
    } else if (!out.empty()) {
      for (char c : line) out.back().seq.push_back((char)std::toupper((unsigned char)c));  // fai.get(chrom).toUpperAscii
    }
  }
  return out;
}

// min-rotation 2-bit code of the k bases at s (slide_by's per-window value, utils.nim:10-34)
uint64_t min_rot_code(const char *s, int k) {
  uint64_t best = ~0ull;
  for (int r = 0; r < k; r++) {
    uint64_t v = 0;
    for (int i = 0; i < k; i++) v = v * 4u + (uint64_t)base_rank(s[(r + i) % k]);
    best = std::min(best, v);
  }
  return best;
}

struct Window {
  int64_t start = 0, stop = -1;
  std::string repeat;
};

// genome_strs.nim:22-59
void trim(Window &w, const std::string &dna) {
  const int k = (int)w.repeat.size();
  const int64_t n = (int64_t)dna.size();
  const uint64_t expected = min_rot_code(w.repeat.data(), k);
  for (int64_t i = 0; i + k <= n; i += k) {
    if (min_rot_code(dna.data() + i, k) != expected) w.start += k;
    else break;
  }
  if (!(w.start < w.stop)) throw std::runtime_error("repeat " + w.repeat + " not found in expected region");
  std::string rrep(w.repeat.rbegin(), w.repeat.rend());
  const uint64_t rexpected = min_rot_code(rrep.data(), k);
  char buf[8];
  for (int64_t i = 0; i + k <= n; i += k) {  // windows of the reversed sequence
    for (int j = 0; j < k; j++) buf[j] = dna[(size_t)(n - 1 - i - j)];
    if (min_rot_code(buf, k) != rexpected) w.stop -= k;
    else break;
  }
  if (!(w.start < w.stop)) throw std::runtime_error("repeat " + w.repeat + " not found in expected region");
}

}  // namespace

// Returns the bed lines (chrom, start, stop, unit) of genome_repeats (genome_strs.nim:107-125)
std::vector<std::string> genome_repeat_lines(const std::string &fasta, double proportion_repeat, int device) {
  const std::vector<Chrom> chroms = read_fasta(fasta);
  strgpu_ctx *gpu = nullptr;
  int rc = strgpu_create(&gpu, device);
  if (rc != STRGPU_OK) throw std::runtime_error(std::string("[strling] gpu: ") + strgpu_error_string(rc));
  auto check = [&](int r, const char *what) {
    if (r != STRGPU_OK) {
      const std::string msg = strgpu_last_error(gpu);
      strgpu_destroy(gpu);
      throw std::runtime_error(std::string("[strling] gpu: ") + what + ": " + msg);
    }
  };
  check(strgpu_set_proportions(gpu, &proportion_repeat, 1), "set_proportions");
  constexpr int kWindow = 100, kStep = 60;
  std::vector<std::string> lines;
  for (const Chrom &c : chroms) {
    const int64_t L = (int64_t)c.seq.size();
    if (L > 2000000) std::fprintf(stderr, "[strling] finding STR regions on reference chromosome: %s\n", c.name.c_str());
    if (L == 0) continue;
    if (L > 0xfffffff0ll) throw std::runtime_error("[strling] chromosome too long: " + c.name);
    std::vector<uint8_t> seq2(strgpu_seq2_bytes((uint64_t)L), 0);
    std::vector<uint32_t> nmask(strgpu_nmask_bytes((uint64_t)L) / 4, 0), xmask(strgpu_nmask_bytes((uint64_t)L) / 4, 0);
    const int n_other = strgpu_pack_ascii(c.seq.data(), (uint32_t)L, seq2.data(), nmask.data(), xmask.data(), 0);
    if (n_other < 0) throw std::runtime_error("[strling] pack_ascii failed");
    const int64_t n_win = (L + kStep - 1) / kStep;
    std::vector<strgpu_segment> segs((size_t)n_win);
    for (int64_t i = 0; i < n_win; i++) {
      const int64_t start = i * kStep, len = std::min<int64_t>(kWindow, L - start);
      bool has_n = false;
      if (n_other)
        for (int64_t b = start; b < start + len && !has_n; b++) has_n = (nmask[(size_t)(b >> 5)] >> (b & 31)) & 1u;
      segs[(size_t)i] = strgpu_segment{(uint32_t)start, (uint16_t)len, 0, (uint8_t)(has_n ? STRGPU_SEG_HAS_N : 0)};
    }
    std::vector<strgpu_repeat> res((size_t)n_win);
    const int64_t kChunk = 1 << 22;
    for (int64_t a = 0; a < n_win; a += kChunk) {
      const uint32_t n = (uint32_t)std::min<int64_t>(kChunk, n_win - a);
      check(strgpu_scan(gpu, seq2.data(), (uint64_t)L, n_other ? nmask.data() : nullptr, n_other ? xmask.data() : nullptr, segs.data() + a, n, kWindow,
                        res.data() + a), "scan");
    }
    // genome_strs.nim:70-92 : merge adjacent windows of the same unit (one window may be skipped), pad, trim
    Window last;
    auto flush = [&]() {
      if (last.stop != -1 && last.stop - last.start >= (kWindow - kStep)) {
        last.start = std::max<int64_t>(0, last.start - kWindow);
        last.stop = std::min<int64_t>(last.stop + kWindow, L);
        trim(last, c.seq.substr((size_t)last.start, (size_t)(last.stop - last.start)));
        lines.push_back(c.name + "\t" + std::to_string(last.start) + "\t" + std::to_string(last.stop) + "\t" + last.repeat);
      }
    };
    for (int64_t i = 0; i < n_win; i++) {
      const strgpu_repeat &r = res[(size_t)i];
      if (r.repeat_count == 0) continue;
      Window w;
      w.start = i * kStep;
      w.stop = w.start + segs[(size_t)i].len;
      for (char ch : r.unit)
        if (ch) w.repeat.push_back(ch);
      if (last.repeat != w.repeat || w.start > last.stop + (kWindow - kStep)) {
        flush();
        last = w;
      } else {
        last.stop = w.stop;
      }
    }
    flush();
  }
  strgpu_destroy(gpu);
  return lines;
}

int index_main(int argc, char **argv) {
  std::string genome_repeats, fasta;
  double p = 0.8;
  int device = 0;
  for (int i = 0; i < argc; i++) {
    const std::string t = argv[i];
    if ((t == "-g" || t == "--genome-repeats") && i + 1 < argc) genome_repeats = argv[++i];
    else if ((t == "-p" || t == "--proportion-repeat") && i + 1 < argc) p = std::stod(argv[++i]);
    else if (t == "--device" && i + 1 < argc) device = std::stoi(argv[++i]);
    else if (t == "-h" || t == "--help") { std::puts("strling index [-g genome-repeats] [-p proportion-repeat=0.8] [--device N] <fasta>"); return 0; }
    else fasta = t;
  }
  if (fasta.empty()) { std::puts("strling index [-g genome-repeats] [-p proportion-repeat=0.8] [--device N] <fasta>"); return 0; }
  if (genome_repeats.empty()) {  // lastPathPart(fasta) & ".str" (genome_strs.nim:186-187)
    const size_t slash = fasta.find_last_of('/');
    genome_repeats = (slash == std::string::npos ? fasta : fasta.substr(slash + 1)) + ".str";
  }
  std::fprintf(stderr, "Writing genome str index to: %s\n", genome_repeats.c_str());
  {
    std::ifstream probe(genome_repeats);
    if (probe) {  // genome_strs.nim:129-130
      std::fprintf(stderr, "[strling] using existing file %s for genome repeats\n", genome_repeats.c_str());
      return 0;
    }
  }
  const auto lines = genome_repeat_lines(fasta, p, device);
  std::ofstream out(genome_repeats);
  if (!out) throw std::runtime_error("[strling] couldn't open bed file: " + genome_repeats + " for writing");
  for (const auto &l : lines) out << l << "\n";
  std::fprintf(stderr, "[strling] found %zu STR-like regions in the genome\n", lines.size());
  return 0;
}

}  // namespace strling
