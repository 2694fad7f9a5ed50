// `strling` command line: the dispatcher of src/strling.nim:12-41, restricted to the subcommands on the hot path.
#include <cstdio>
#include <cstring>
#include <exception>
#include <string>

#include "commands.hpp"
#include "strgpu.h"

int main(int argc, char **argv) {
  static const char *usage =
      "strling-b200 (STRling 0.6.0 hot path on B200)\n\nCommands:\n"
      "  extract       :   extract informative STR reads from a BAM (repeat-unit scan on the GPU)\n"
      "  merge         :   merge putative STR loci from multiple samples (clustering on the GPU)\n"
      "  index         :   find STR-like regions of a reference genome (the genome-repeats file of extract -g)\n"
      "  call          :   discover and genotype the STR loci of one sample (clustering on the GPU)\n";
  if (argc < 2 || !std::strcmp(argv[1], "-h") || !std::strcmp(argv[1], "--help")) {
    std::fputs(usage, stdout);
    return argc < 2 ? 1 : 0;
  }
  const std::string cmd = argv[1];
  try {
    if (cmd == "extract") return strling::extract_main(argc - 2, argv + 2);
    if (cmd == "merge") return strling::merge_main(argc - 2, argv + 2);
    if (cmd == "call") return strling::call_main(argc - 2, argv + 2);
    if (cmd == "index") return strling::index_main(argc - 2, argv + 2);
    if (cmd == "debug") return strling::debug_main(argc - 2, argv + 2);
    if (cmd == "--version" || cmd == "version") { std::printf("%s\n", strgpu_version()); return 0; }
    std::fprintf(stderr, "unknown program '%s'\n%s", cmd.c_str(), usage);
    return 1;
  } catch (const std::exception &e) {  // the reference `quit`s with a message and a non-zero status
    std::fprintf(stderr, "%s\n", e.what());
    return 1;
  }
}
