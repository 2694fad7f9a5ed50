// `strling debug ...`: CPU-only introspection of the host side (BAM decode, `.bin` codec, pair arithmetic) so that
// tests can compare it with the oracle without a GPU.  Not part of the reference CLI.
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <iostream>
#include <sstream>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "bam.hpp"
#include "commands.hpp"
#include "mate_table.hpp"
#include "pool.hpp"
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
    if (const char *t = std::getenv("STRLING_DEBUG_THREADS")) { e.threads = std::atoi(t); e.verbose = true; }
    if (const char *t = std::getenv("STRLING_DEBUG_SHARDS")) e.replay_shards = std::atoi(t);
    // the --gpu-inflate plumbing (compressed bytes staged contiguously, rebased block descriptors) with the blocks decoded on the
    // host from that staging buffer: everything but the CUDA calls of strgpu_inflate_bgzf
    if (std::getenv("STRLING_DEBUG_STAGED_INFLATE")) e.gpu_inflate = true;  // stage timings of the host side
    return extract_run(e);
  }
  if (what == "synth-bam" && argc >= 3)  // strling debug synth-bam <out.bam> <n_pairs> [seed] [deflate level] [threads]
    return synth_bam(argv[1], (uint64_t)std::atoll(argv[2]), argc >= 4 ? (uint64_t)std::atoll(argv[3]) : 2, argc >= 5 ? std::atoi(argv[4]) : 1,
                     argc >= 6 ? std::atoi(argv[5]) : 0);
  if (what == "pool-selftest") {
    // pool.hpp: three threads submit jobs of different priorities at the same time (as the reader, stager and consumer of extract
    // do); every task must run exactly once, run() must not return before its job is done, and a throwing task must surface in
    // the caller of that job only
    const int threads = argc >= 2 ? std::atoi(argv[1]) : 6;
    Pool pool(threads);
    std::atomic<long> bad{0};
    auto submitter = [&](int prio, int rounds, size_t n) {
      for (int r = 0; r < rounds; r++) {
        std::vector<std::atomic<int>> hit(n);
        for (auto &h : hit) h = 0;
        bool threw = false;
        const bool want_throw = (r % 7) == 3;
        try {
          pool.run(n, [&](size_t i) {
            hit[i]++;
            if (want_throw && i == n / 2) throw std::runtime_error("task failed");
            volatile uint64_t x = 0;
            for (int k = 0; k < 200 + (int)(i % 50) * 20; k++) x += (uint64_t)k * i;
          }, prio);
        } catch (const std::exception &) {
          threw = true;
        }
        if (threw != want_throw) bad++;
        for (auto &h : hit)
          if (h != 1) bad++;
      }
    };
    std::thread a(submitter, 0, 300, (size_t)997), b(submitter, 1, 400, (size_t)64), c(submitter, 2, 500, (size_t)5);
    a.join(); b.join(); c.join();
    size_t covered = 0;
    pool.ranges(1000003, 37, [&](size_t lo, size_t hi, size_t) { __atomic_fetch_add(&covered, hi - lo, __ATOMIC_RELAXED); });
    if (covered != 1000003) bad++;
    std::printf(bad == 0 ? "ok\n" : "FAIL %ld\n", bad.load());
    return bad == 0 ? 0 : 1;
  }
  if (what == "matetable-selftest") {
    // MateTable (the mate table of extract's replay) against std::unordered_map under random insert / find / take traffic: names of
    // 1..120 bytes (inline and heap storage), growth from 1024 slots to hundreds of thousands of live entries, and -- with a degraded
    // hash that keeps only a few bits -- long probe runs, which is where backward-shift deletion goes wrong if it is wrong
    const uint64_t seed = argc >= 2 ? (uint64_t)std::atoll(argv[1]) : 1;
    const long n_ops = argc >= 3 ? std::atol(argv[2]) : 2000000;
    uint64_t st = seed * 0x9e3779b97f4a7c15ull + 7;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
    for (int mode = 0; mode < 3; mode++) {
      const uint64_t hash_mask = mode == 0 ? ~0ull : mode == 1 ? 0xfffull : 0x3full;   // full hash, 4096 classes, 64 classes
      const size_t n_names = mode == 0 ? 400000 : mode == 1 ? 20000 : 2000;   // degraded hashes: probe runs as long as the table is full
      MateTable tbl;
      std::unordered_map<std::string, uint32_t> ref;
      std::vector<std::string> names(n_names);
      for (size_t i = 0; i < n_names; i++) {
        const size_t len = 1 + (size_t)(rnd() % (i % 50 == 0 ? 120 : 40));
        names[i].resize(len);
        for (auto &c : names[i]) c = (char)('!' + rnd() % 90);
        names[i] += std::to_string(i);  // unique
      }
      for (long op = 0; op < n_ops / (mode == 0 ? 1 : 10); op++) {
        const std::string &nm = names[(size_t)(rnd() % n_names)];
        const uint64_t h = hash_name(nm.data(), nm.size()) & hash_mask;
        const size_t slot = tbl.find(nm.data(), (uint32_t)nm.size(), h);
        auto it = ref.find(nm);
        if ((slot != SIZE_MAX) != (it != ref.end())) { std::printf("FAIL presence mode %d op %ld\n", mode, op); return 1; }
        if (slot != SIZE_MAX) {
          if (tbl.at(slot).t.position != it->second) { std::printf("FAIL value mode %d op %ld\n", mode, op); return 1; }
          if (rnd() % 3) { tbl.erase(slot); ref.erase(it); }
        } else if (rnd() % 4) {
          TreadCore t;
          t.position = (uint32_t)rnd();
          tbl.insert(nm.data(), (uint32_t)nm.size(), h, t);
          ref.emplace(nm, t.position);
        }
        if (tbl.size() != ref.size()) { std::printf("FAIL size mode %d op %ld\n", mode, op); return 1; }
      }
      // everything that should be there is there
      for (const auto &kv : ref) {
        const uint64_t h = hash_name(kv.first.data(), kv.first.size()) & hash_mask;
        const size_t slot = tbl.find(kv.first.data(), (uint32_t)kv.first.size(), h);
        if (slot == SIZE_MAX || tbl.at(slot).t.position != kv.second) { std::printf("FAIL final mode %d\n", mode); return 1; }
      }
      std::printf("mode %d ok\tlive %zu\n", mode, ref.size());
      std::fflush(stdout);
    }
    return 0;
  }
  if (what == "inflate-selftest") {
    // the repo's DEFLATE decoder against zlib's encoder: every deflate level and strategy (stored, fixed and dynamic Huffman
    // blocks, long sub-table codes, every match distance / length class) over several kinds of data and all small sizes
    const uint64_t seed = argc >= 2 ? (uint64_t)std::atoll(argv[1]) : 1;
    const int rounds = argc >= 3 ? std::atoi(argv[2]) : 200;
    uint64_t st = seed * 0x9e3779b97f4a7c15ull + 1;
    auto rnd = [&]() { st ^= st << 13; st ^= st >> 7; st ^= st << 17; return st; };
    auto same_bytes = [](const uint8_t *a, const uint8_t *b, size_t len) { return len == 0 || std::memcmp(a, b, len) == 0; };
    infl::Tables *T = new infl::Tables;
    T->fixed_built = false;
    size_t cases = 0, bytes = 0;
    for (int r = 0; r < rounds; r++) {
      const size_t n = r < 64 ? (size_t)r : (size_t)(rnd() % 65281);
      std::vector<uint8_t> raw(n);
      const int kind = (int)(rnd() % 6);
      for (size_t i = 0; i < n; i++) {
        switch (kind) {
          case 0: raw[i] = (uint8_t)rnd(); break;                                         // incompressible
          case 1: raw[i] = (uint8_t)"ACGT"[rnd() & 3]; break;                             // 2 bits of entropy per byte
          case 2: raw[i] = (uint8_t)(i % (1 + seed % 7 + (size_t)r % 300)); break;        // periodic: every short distance
          case 3: raw[i] = (uint8_t)((rnd() % 100) < 3 ? rnd() : 'F'); break;             // long runs (distance 1)
          case 4: raw[i] = (uint8_t)(i > 4000 && (rnd() % 8) ? raw[i - 1 - (size_t)(rnd() % 4000)] : rnd()); break;  // far matches
          default: raw[i] = (uint8_t)(rnd() % (2 + (size_t)r % 250)); break;              // skewed alphabets: long codes
        }
      }
      for (int level = 0; level <= 9; level += (level < 2 ? 1 : 4)) {
        for (int strategy : {Z_DEFAULT_STRATEGY, Z_FIXED, Z_HUFFMAN_ONLY, Z_RLE}) {
          std::vector<uint8_t> comp(n + n / 8 + 1024);
          z_stream zs;
          std::memset(&zs, 0, sizeof(zs));
          if (deflateInit2(&zs, level, Z_DEFLATED, -15, 1 + (r % 9), strategy) != Z_OK) return 2;
          zs.next_in = raw.data();
          zs.avail_in = (uInt)n;
          zs.next_out = comp.data();
          zs.avail_out = (uInt)comp.size() - 8;
          // a flush in the middle makes multi-block streams (and empty stored blocks)
          if (n > 100 && (r & 1)) { zs.avail_in = (uInt)(n / 2); deflate(&zs, Z_FULL_FLUSH); zs.avail_in = (uInt)(n - n / 2); }
          if (deflate(&zs, Z_FINISH) != Z_STREAM_END) return 2;
          const size_t clen = comp.size() - 8 - zs.avail_out;
          deflateEnd(&zs);
          std::vector<uint8_t> got(n + 64, 0xAB);
          const int rc = infl::inflate_block(*T, comp.data(), (uint32_t)clen, got.data() + 32, (uint32_t)n);
          bool ok = rc == infl::kOk && same_bytes(got.data() + 32, raw.data(), n);
          for (int g = 0; g < 32; g++) ok = ok && got[(size_t)g] == 0xAB && got[32 + n + (size_t)g] == 0xAB;  // nothing outside [out, out + n)
          if (!ok) { std::printf("FAIL round %d kind %d level %d strategy %d n %zu rc %d\n", r, kind, level, strategy, n, rc); return 1; }
          // the instance the BAM readers call (compiled for BMI2 where the CPU has it)
          std::fill(got.begin(), got.end(), 0xAB);
          const int rc1 = inflate_block_host(*T, comp.data(), (uint32_t)clen, got.data() + 32, (uint32_t)n);
          bool ok1 = rc1 == infl::kOk && same_bytes(got.data() + 32, raw.data(), n);
          for (int g = 0; g < 32; g++) ok1 = ok1 && got[(size_t)g] == 0xAB && got[32 + n + (size_t)g] == 0xAB;
          if (!ok1) { std::printf("FAIL (dispatched) round %d kind %d level %d strategy %d n %zu rc %d\n", r, kind, level, strategy, n, rc1); return 1; }
          // the command-stream form of the decoder (what the CUDA kernel runs), commands executed serially
          std::fill(got.begin(), got.end(), 0xAB);
          const int rc2 = infl::inflate_block_stream(*T, comp.data(), (uint32_t)clen, got.data() + 32, (uint32_t)n);
          bool ok2 = rc2 == infl::kOk && same_bytes(got.data() + 32, raw.data(), n);
          for (int g = 0; g < 32; g++) ok2 = ok2 && got[(size_t)g] == 0xAB && got[32 + n + (size_t)g] == 0xAB;
          if (!ok2) { std::printf("FAIL (stream) round %d kind %d level %d strategy %d n %zu rc %d\n", r, kind, level, strategy, n, rc2); return 1; }
          // a truncated or corrupted stream must be refused or at least stay inside the output buffer
          if (clen > 4) {
            std::vector<uint8_t> bad(comp);
            bad[(size_t)(rnd() % clen)] ^= (uint8_t)(1u << (rnd() % 8));
            std::fill(got.begin(), got.end(), 0xAB);
            const int rb1 = infl::inflate_block(*T, bad.data(), (uint32_t)clen, got.data() + 32, (uint32_t)n);
            for (int g = 0; g < 32; g++)
              if (got[(size_t)g] != 0xAB || got[32 + n + (size_t)g] != 0xAB) { std::printf("FAIL overrun on corrupt input, round %d\n", r); return 1; }
            std::vector<uint8_t> first(got.begin() + 32, got.begin() + 32 + (long)n);
            std::fill(got.begin(), got.end(), 0xAB);
            const int rb2 = infl::inflate_block_stream(*T, bad.data(), (uint32_t)clen, got.data() + 32, (uint32_t)n);
            for (int g = 0; g < 32; g++)
              if (got[(size_t)g] != 0xAB || got[32 + n + (size_t)g] != 0xAB) { std::printf("FAIL (stream) overrun on corrupt input, round %d\n", r); return 1; }
            // both forms must agree on whether the damaged stream is acceptable, and on its bytes when it is
            if ((rb1 == infl::kOk) != (rb2 == infl::kOk) || (rb1 == infl::kOk && !same_bytes(first.data(), got.data() + 32, n))) {
              std::printf("FAIL the two decoders disagree on a corrupt stream, round %d: %d vs %d\n", r, rb1, rb2);
              return 1;
            }
          }
          cases++;
          bytes += n;
        }
      }
    }
    delete T;
    std::printf("ok\t%zu cases\t%zu bytes\n", cases, bytes);
    return 0;
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
    BamChunkReader rd(argv[1], hdr.tell(), threads, nullptr, (int32_t)hdr.targets().size());
    BamChunk c;
    size_t n = 0, chunks = 0;
    uint64_t digest = 1469598103934665603ull;  // FNV-1a over (tid, pos, flag, l_seq, qname) of every record, in order
    auto mix = [&](const void *p, size_t len) {
      const uint8_t *b = static_cast<const uint8_t *>(p);
      for (size_t i = 0; i < len; i++) digest = (digest ^ b[i]) * 1099511628211ull;
    };
    const auto t0 = std::chrono::steady_clock::now();
    while (rd.next(c, blocks)) {
      n += c.n_records();
      chunks++;
      if (argc >= 5)
        for (size_t i = 0; i < c.n_records(); i++) {
          const BamRecord r = BamChunk::view(c.data.data() + c.rec_off[i]);
          mix(&r.tid, 4); mix(&r.pos, 4); mix(&r.flag, 2); mix(&r.l_seq, 4); mix(r.qname, r.l_qname);
        }
    }
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("records\t%zu\tchunks\t%zu\tseconds\t%.3f\tread\t%.3f\talloc\t%.3f\tinflate\t%.3f\twalk\t%.3f\trewalked\t%zu\tdigest\t%016llx\n", n, chunks, dt, rd.t_read,
                rd.t_alloc, rd.t_inflate, rd.t_walk, rd.n_rewalked, (unsigned long long)digest);
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
