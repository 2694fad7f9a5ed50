#!/usr/bin/env python
"""Config 5 in miniature on a multi-GPU box: N synthetic samples -> `strling extract` each (GPU) -> `strling merge` on one GPU
and `python -m strling_b200.joint` over all visible GPUs -> the two -bounds.txt files must be identical.
usage: python tools/joint_check.py [n_samples=10] [pairs_per_sample=20000]"""
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from strling_b200 import bamio  # noqa: E402
from strling_b200 import build as sb_build  # noqa: E402


def main():
    n_samples = int(sys.argv[1]) if len(sys.argv) > 1 else 10
    n_pairs = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    cli = sb_build.build_cli()
    n_gpu = torch.cuda.device_count()
    targets = [(f"chr{i + 1}", 1_500_000) for i in range(6)]
    loci = [(i % 6, 100_000 + 61_000 * i, 100_000 + 61_000 * i + 30 + 7 * (i % 9), u)
            for i, u in enumerate(["CAG", "AAAG", "AC", "CCG", "ATTCT", "A", "AAGGG", "CTG", "AAAAT", "GGC", "AT", "CAGG"] * 2)]
    hdr = bamio.sam_header(targets)
    d = tempfile.mkdtemp(prefix="joint_check_")
    bins = []
    for s in range(n_samples):
        recs = bamio.simulate_alignments(900 + s, n_pairs, targets, loci, str_pair_frac=0.3, unmapped_pairs=50)
        bam, out = os.path.join(d, f"s{s}.bam"), os.path.join(d, f"s{s}.bin")
        bamio.write_bam(bam, hdr, targets, recs)
        subprocess.run([cli, "extract", bam, out], check=True, capture_output=True)
        bins.append(out)
    t0 = time.time()
    subprocess.run([cli, "merge", "-m", "5", "-o", os.path.join(d, "one"), *bins], check=True, capture_output=True)
    t1 = time.time()
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n_gpu}", "--master-addr", "127.0.0.1",
           "--master-port", "29544", "-m", "strling_b200.joint", "-m", "5", "-o", os.path.join(d, "many"), *bins]
    r = subprocess.run(cmd, capture_output=True, text=True, cwd=ROOT)
    t2 = time.time()
    if r.returncode != 0:
        print(r.stderr[-2000:])
        raise SystemExit("joint merge failed")
    a = open(os.path.join(d, "one-bounds.txt")).read()
    b = open(os.path.join(d, "many-bounds.txt")).read()
    print(f"samples {n_samples}, GPUs {n_gpu}: one-GPU merge {t1 - t0:.2f}s, joint merge {t2 - t1:.2f}s (incl. process start), "
          f"{a.count(chr(10)) - 1} bounds lines, identical: {a == b}")
    if a != b:
        raise SystemExit("joint merge output differs from strling merge")


if __name__ == "__main__":
    main()
