#!/usr/bin/env python
"""Times the `strling extract` command line (C++ host + CUDA scan) on a synthetic BAM (`strling debug synth-bam`, the
configs[1] read mix) and prints the stage report the binary emits with -v.
usage: python tools/bench_cli.py [n_pairs] [threads] [--host-only]
--host-only: no GPU needed -- the scan results are taken from a file of zeros (`strling debug extract replay`), which times
inflate + decode + staging + replay exactly as `extract` runs them (every read then looks non-repetitive to the replay)."""
import json
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from strling_b200 import build  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith("--")]
host_only = "--host-only" in sys.argv
n_pairs = int(args[0]) if len(args) > 0 else 1_500_000
threads = args[1] if len(args) > 1 else "0"
cli = build.build_cli()
d = tempfile.mkdtemp(prefix="strcli")
bam, out, zeros = os.path.join(d, "bench.bam"), os.path.join(d, "bench.bin"), os.path.join(d, "zeros.bin")
print(subprocess.run([cli, "debug", "synth-bam", bam, str(n_pairs)], capture_output=True, text=True, check=True).stdout.strip(), file=sys.stderr)
if host_only:
    with open(zeros, "wb") as fh:
        fh.write(bytes(8 * 3 * 2 * n_pairs))
best = None
for rep in range(3):
    if host_only:
        env = dict(os.environ, STRLING_DEBUG_THREADS=threads if threads != "0" else str(os.cpu_count()))
        r = subprocess.run([cli, "debug", "extract", "replay", zeros, bam, out], capture_output=True, text=True, env=env)
    else:
        r = subprocess.run([cli, "extract", "-v", "--threads", threads, bam, out], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    perf = json.loads(re.search(r"perf: (\{.*\})", r.stderr).group(1))
    if best is None or perf["scan_pass_s"] < best["scan_pass_s"]:
        best = perf
best["bam_mb"] = round(os.path.getsize(bam) / 1e6, 1)
best["mode"] = "host only (scan results from a file)" if host_only else "extract"
print(json.dumps(best))
