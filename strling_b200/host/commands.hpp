// Entry points of the `strling` command line (src/strling.nim:18-41 dispatcher) implemented by this build.
#pragma once
#include <array>
#include <cstdint>
#include <string>
#include <vector>

namespace strling {

struct ExtractArgs {
  std::string fasta, genome_repeats, bam, bin;
  double proportion_repeat = 0.8;
  int min_mapq = 40;
  bool verbose = false;
  int device = 0;
  int threads = 0;
  uint32_t batch_reads = 1u << 18;
  int replay_shards = 0;          // 0: a quarter of the threads
  bool gpu_inflate = false;       // BGZF blocks are inflated on the GPU (strgpu_inflate_bgzf) instead of the host threads
  // `strling debug extract` only (CPU-side checks of the staging and replay logic; never set by `strling extract`):
  // dump: write every staged segment ("pclass<TAB>bases", submission order) instead of scanning it, no .bin is written;
  // results: read the scan results (8-byte strgpu_repeat records in that same order) from a file instead of the GPU
  std::string debug_dump_segments, debug_scan_results;
};
int extract_run(const ExtractArgs &a);
std::array<uint32_t, 4096> fragment_length_distribution(const std::string &bam, int threads);

int extract_main(int argc, char **argv);
int merge_main(int argc, char **argv);
int call_main(int argc, char **argv);
int debug_main(int argc, char **argv);
int debug_genotype(int argc, char **argv);
int index_main(int argc, char **argv);
int synth_bam(const std::string &path, uint64_t n_pairs, uint64_t seed, int level, int threads);
std::vector<std::string> genome_repeat_lines(const std::string &fasta, double proportion_repeat, int device);

}  // namespace strling
