#!/usr/bin/env python
"""The .bin `strling extract` must write for the synthetic BAM of bench.py's cli leg, computed WITHOUT a GPU: `strling debug synth-bam`
-> `strling debug extract dump` (the staged segments) -> the oracle scans every segment (oracle/liboracle.so, all cores) ->
`strling debug extract replay` -> md5 of the .bin.  bench.py compares the md5 of the .bin the GPU path wrote with the constant
this prints (CLI_BIN_MD5_ORACLE): a parity check of the whole command line against the oracle at 6x10^6 reads.
usage: python tools/cli_oracle_md5.py [n_pairs=3000000] [seed=2] [deflate level=6]"""
import hashlib
import multiprocessing as mp
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import oracle as orc  # noqa: E402
from strling_b200 import build  # noqa: E402

n_pairs = sys.argv[1] if len(sys.argv) > 1 else "3000000"
seed = sys.argv[2] if len(sys.argv) > 2 else "2"
level = sys.argv[3] if len(sys.argv) > 3 else "6"
cli = build.build_cli()
d = tempfile.mkdtemp(prefix="cli_oracle_")
bam, segs, res, out = (os.path.join(d, n) for n in ("x.bam", "segs.tsv", "res.bin", "x.bin"))
env = dict(os.environ, STRLING_DEBUG_THREADS=str(os.cpu_count() or 1))
subprocess.run([cli, "debug", "synth-bam", bam, n_pairs, seed, level], check=True, capture_output=True)
subprocess.run([cli, "debug", "extract", "dump", segs, bam, out], check=True, capture_output=True, env=env)
data = np.fromfile(segs, dtype=np.uint8)
nl = np.flatnonzero(data == 10)
starts = np.concatenate(([0], nl[:-1] + 1))
off = (starts + 2).astype(np.uint64)
length = (nl - starts - 2).astype(np.uint32)
p = np.array([0.8, 0.8 - 0.07, 0.6])[data[starts] - 48]   # the three proportion classes of extract (extract.nim:204-211,240-244)
n = len(off)
cores = os.cpu_count() or 1


def work(rng):
    a, b = rng
    return orc.get_repeat_batch(data, off[a:b], length[a:b], p[a:b])


with mp.get_context("fork").Pool(cores) as pool:
    parts = pool.map(work, [(i * n // cores, (i + 1) * n // cores) for i in range(cores)])
results = np.zeros(n, dtype=[("unit", "S6"), ("repeat_count", "<u2")])
results["unit"] = np.concatenate([r[0] for r in parts])
results["repeat_count"] = np.concatenate([r[1] for r in parts])
results.tofile(res)
r = subprocess.run([cli, "debug", "extract", "replay", res, bam, out], check=True, capture_output=True, text=True, env=env)
md5 = hashlib.md5(open(out, "rb").read()).hexdigest()
print(f"segments {n}  STR segments {int((results['repeat_count'] > 0).sum())}  .bin bytes {os.path.getsize(out)}  md5 {md5}")
for f in (bam, segs, res, out):
    os.remove(f)
os.rmdir(d)
