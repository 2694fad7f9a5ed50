#!/usr/bin/env python
"""One strgpu_cluster_device call on the bench's cluster-leg input (6x10^6 synthetic STR-read records), for ncu captures:
   STRGPU_NO_GRAPH=1 ncu --set full -k regex:"radix|cluster_|piece_|bucket_|make_sort|scan_" -o out python tools/profile_cluster.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

import strling_b200 as sb  # noqa: E402
from strling_b200 import synth  # noqa: E402

n_treads = int(sys.argv[1]) if len(sys.argv) > 1 else 6_000_000
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 1
treads = synth.make_treads(max(10, n_treads // 120), seed=40, noise_reads=n_treads - (n_treads // 120) * 26, unplaced=n_treads // 200)
dev = torch.device("cuda", 0)
d_t = torch.from_numpy(treads.view(np.uint8).reshape(-1).copy()).to(dev)
cap = max(1024, len(treads) // 2)
d_out = torch.zeros(cap * 48, dtype=torch.uint8, device=dev)
d_n = torch.zeros(1, dtype=torch.int32, device=dev)
g = sb.StrGpu(0)
p = g.cluster_params(window=480, min_support=5, max_clip_dist=190)
st = torch.cuda.current_stream().cuda_stream
for _ in range(calls):
    g.cluster_device(d_t.data_ptr(), len(treads), p, d_out.data_ptr(), cap, d_n.data_ptr(), st)
torch.cuda.synchronize()
print(len(treads), "records ->", int(d_n.item()), "cluster records")
