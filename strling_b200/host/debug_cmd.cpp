// `strling debug ...`: CPU-only introspection of the host side (BAM decode, `.bin` codec, pair arithmetic) so that
// tests can compare it with the oracle without a GPU.  Not part of the reference CLI.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>

#include "bam.hpp"
#include "commands.hpp"
#include "tread.hpp"

namespace strling {

namespace {

std::array<char, 6> unit_of(const std::string &s) {
  std::array<char, 6> u{{0, 0, 0, 0, 0, 0}};
  if (s != ".")
    for (size_t i = 0; i < s.size() && i < 6; i++) u[i] = s[i];
  return u;
}
std::string unit_str(const std::array<char, 6> &u) {
  std::string s;
  for (char c : u)
    if (c) s.push_back(c);
  return s.empty() ? "." : s;
}
Tread read_tread(std::istream &in) {
  Tread t;
  std::string unit;
  int flag, split, mapq, rc, alen;
  long long pos;
  in >> t.tid >> pos >> unit >> flag >> split >> mapq >> rc >> alen;
  t.position = (uint32_t)pos;
  t.repeat = unit_of(unit);
  t.flag = (uint16_t)flag; t.split = (uint8_t)split; t.mapping_quality = (uint8_t)mapq; t.repeat_count = (uint8_t)rc; t.align_length = (uint8_t)alen;
  return t;
}

}  // namespace

int debug_main(int argc, char **argv) {
  if (argc < 1) return 1;
  const std::string what = argv[0];
  if (what == "genotype") return debug_genotype(argc - 1, argv + 1);
  if (what == "extract" && argc >= 5) {
    // strling debug extract {dump <segments.tsv> | replay <results.bin>} <bam> <bin> [p] [min_mapq] [batch_reads] [genome-repeats.bed]
    // The staging and the order-dependent replay of `strling extract` WITHOUT the scan: `dump` writes every staged segment, `replay`
    // takes the scan results from a file (the CPU tests compute them with the oracle).  It is a test hook, not a CPU path: the
    // results never come from this program.
    ExtractArgs e;
    const std::string mode = argv[1];
    if (mode == "dump") e.debug_dump_segments = argv[2];
    else if (mode == "replay") e.debug_scan_results = argv[2];
    else return 1;
    e.bam = argv[3];
    e.bin = argv[4];
    if (argc >= 6) e.proportion_repeat = std::atof(argv[5]);
    if (argc >= 7) e.min_mapq = std::atoi(argv[6]);
    if (argc >= 8) e.batch_reads = (uint32_t)std::atol(argv[7]);
    if (argc >= 9) e.genome_repeats = argv[8];
    e.threads = 2;
    return extract_run(e);
  }
  if (what == "bam" && argc >= 2) {
    BamReader rd(argv[1]);
    std::printf("@targets %zu\n", rd.targets().size());
    for (const auto &t : rd.targets()) std::printf("@target\t%s\t%u\n", t.name.c_str(), t.length);
    std::printf("@text_bytes %zu\n", rd.header_text().size());
    BamRecord r;
    static const char *nib = "=ACMGRSVTWYHKDBN";
    while (rd.next(r)) {
      std::string cig, seq;
      for (int i = 0; i < r.n_cigar; i++) cig += std::to_string(BamRecord::oplen(r.cigar_at(i))) + "MIDNSHP=X"[BamRecord::op(r.cigar_at(i))];
      for (int i = 0; i < r.l_seq; i++) seq.push_back(nib[(i & 1) ? (r.seq[i >> 1] & 15) : (r.seq[i >> 1] >> 4)]);
      std::printf("%s\t%u\t%d\t%d\t%u\t%s\t%d\t%d\t%d\t%s\t%d\n", r.qname, (unsigned)r.flag, r.tid, r.pos, (unsigned)r.mapq, cig.empty() ? "*" : cig.c_str(),
                  r.mate_tid, r.mate_pos, r.isize, seq.c_str(), r.stop());
    }
    return 0;
  }
  if (what == "chunks" && argc >= 2) {  // timing of the chunked reader alone
    BamReader hdr(argv[1]);
    const int threads = argc >= 3 ? std::atoi(argv[2]) : 0;
    const size_t blocks = argc >= 4 ? (size_t)std::atol(argv[3]) : 4096;
    BamChunkReader rd(argv[1], hdr.tell(), threads);
    BamChunk c;
    size_t n = 0, chunks = 0;
    const auto t0 = std::chrono::steady_clock::now();
    while (rd.next(c, blocks)) { n += c.n_records(); chunks++; }
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("records\t%zu\tchunks\t%zu\tseconds\t%.3f\tread\t%.3f\talloc\t%.3f\tinflate\t%.3f\twalk\t%.3f\n", n, chunks, dt, rd.t_read, rd.t_alloc,
                rd.t_inflate, rd.t_walk);
    return 0;
  }
  if (what == "fragdist" && argc >= 2) {
    const auto f = fragment_length_distribution(argv[1], 0);
    for (int i = 0; i < 4096; i++)
      if (f[(size_t)i]) std::printf("%d\t%u\n", i, f[(size_t)i]);
    std::printf("median\t%d\t%d\t%d\n", frag_median(f), frag_median(f, 0.98), frag_median(f, 0.99));
    return 0;
  }
  if (what == "bin" && argc >= 2) {
    const BinFile bf = read_bin(argv[1]);
    std::printf("p\t%.9g\nmin_mapq\t%u\nheader_bytes\t%zu\nn\t%zu\n", (double)bf.proportion_repeat, (unsigned)bf.min_mapq, bf.header.size(), bf.reads.size());
    for (const Tread &t : bf.reads)
      std::printf("%d\t%u\t%s\t%u\t%u\t%u\t%u\t%u\t%s\n", t.tid, t.position, unit_str(t.repeat).c_str(), (unsigned)t.flag, (unsigned)t.split,
                  (unsigned)t.mapping_quality, (unsigned)t.repeat_count, (unsigned)t.align_length, t.qname.c_str());
    if (argc >= 3) write_bin(argv[2], bf);  // re-encode: must be byte-identical
    return 0;
  }
  if (what == "logic") {  // line protocol on stdin
    std::string line;
    while (std::getline(std::cin, line)) {
      std::istringstream in(line);
      std::string op;
      in >> op;
      if (op == "canonical") {
        std::string u;
        in >> u;
        std::printf("%s\n", unit_str(canonical_repeat(unit_of(u))).c_str());
      } else if (op == "minrc") {
        std::string u;
        in >> u;
        auto a = unit_of(u);
        min_rev_complement(a);
        std::printf("%s\n", unit_str(a).c_str());
      } else if (op == "adjust") {
        Tread A = read_tread(in), B = read_tread(in);
        Options o;
        int mq;
        long long bpos;
        in >> o.proportion_repeat >> mq >> o.median_fragment_length >> bpos;
        o.min_mapq = (uint8_t)mq;
        const bool r = adjust_by(A, B, o, (uint32_t)bpos);
        std::printf("%d\t%d\t%u\t%s\t%u\t%u\n", (int)r, A.tid, A.position, unit_str(A.repeat).c_str(), (unsigned)A.split, (unsigned)A.mapping_quality);
      } else if (op == "unplaced") {
        Tread A = read_tread(in), B = read_tread(in);
        Options o;
        int mq;
        in >> o.proportion_repeat >> mq;
        o.min_mapq = (uint8_t)mq;
        std::printf("%d\n", (int)unplaced_pair(A, B, o));
      } else if (op == "prepeat") {
        Tread A = read_tread(in);
        std::printf("%.17g\n", p_repeat(A));
      }
    }
    return 0;
  }
  std::fprintf(stderr, "strling debug {bam <bam> | fragdist <bam> | bin <bin> [out.bin] | logic | extract {dump|replay} <file> <bam> <bin> ...}\n");
  return 1;
}

}  // namespace strling
