#!/usr/bin/env python
"""Times the `strling extract` command line (C++ host + CUDA scan) on a synthetic BAM and prints the stage report the
binary emits with -v.  usage: python tools/bench_cli.py [n_pairs] [threads]"""
import json
import os
import re
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from strling_b200 import bamio, build  # noqa: E402

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 300_000
threads = sys.argv[2] if len(sys.argv) > 2 else "0"
cli = build.build_cli()
targets = [(f"chr{i + 1}", 50_000_000) for i in range(8)]
loci = [(i % 8, 1_000_000 + 137_000 * i, 1_000_000 + 137_000 * i + 60, u) for i, u in enumerate(["CAG", "AAAG", "ATTCT", "A", "AC", "CCG", "AAGGG", "CACGAT"] * 20)]
d = tempfile.mkdtemp(prefix="strcli")
bam, out = os.path.join(d, "bench.bam"), os.path.join(d, "bench.bin")
t0 = time.time()
recs = bamio.simulate_alignments(5, n_pairs, targets, loci, str_pair_frac=0.03, unmapped_pairs=n_pairs // 100)
t1 = time.time()
bamio.write_bam(bam, bamio.sam_header(targets), targets, recs, level=1)
t2 = time.time()
print(f"generated {len(recs)} records in {t1 - t0:.1f}s, wrote {os.path.getsize(bam) / 1e6:.0f} MB BAM in {t2 - t1:.1f}s", file=sys.stderr)
best = None
for rep in range(3):
    t0 = time.time()
    r = subprocess.run([cli, "extract", "-v", "--threads", threads, bam, out], capture_output=True, text=True)
    dt = time.time() - t0
    assert r.returncode == 0, r.stderr
    m = re.search(r"perf: (\{.*\})", r.stderr)
    perf = json.loads(m.group(1))
    perf["wall_s"] = dt
    if best is None or perf["scan_pass_s"] < best["scan_pass_s"]:
        best = perf
print(json.dumps(best))
