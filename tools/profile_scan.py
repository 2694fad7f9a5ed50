#!/usr/bin/env python
"""Times the scan kernels of one library call with CUDA events (no profiler attached): the bench's shard (12e6 reads of the
config-2 mix, device resident) through strgpu_scan_reads_device, `--calls` times on one stream.  STRGPU_MAX_STAGE=<k> (profiling
knob of launch_repeat_scan) stops the call after rung k, so successive runs give the cumulative cost of each kernel.
Also the input of the `ncu --set full` captures under profiles/ (few calls, -k regex:...)."""
import argparse
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(args):
    import torch

    import strling_b200 as sb
    from strling_b200 import synth

    dev = torch.device("cuda", 0)
    reads, cls, lclip, rclip = synth.make_reads(args.reads, seed=2)
    seq2, nmask, stride = synth.pack_matrix(reads, align_bases=args.align)
    segs, _ = synth.segments_for(reads, lclip, rclip, stride)
    n_seg, n = len(segs), args.reads
    seq_bytes = n * stride // 4
    d_seq = torch.zeros(args.regions * seq_bytes + 64, dtype=torch.uint8, device=dev)
    h = torch.from_numpy(seq2[:seq_bytes].copy()).to(dev)
    for r in range(args.regions):
        d_seq[r * seq_bytes:(r + 1) * seq_bytes].copy_(h)
    d_segs = torch.from_numpy(segs.view(np.uint8).reshape(-1).copy()).to(dev)
    d_out = torch.zeros(args.regions * n_seg * 8, dtype=torch.uint8, device=dev)
    extra_max = int(segs["len"][n:].max()) if n_seg > n else 0
    g = sb.StrGpu(0)
    g.set_proportions([0.8, 0.73, 0.6])
    st = torch.cuda.current_stream().cuda_stream

    def call(i):
        r = i % args.regions
        g.scan_reads_device(d_seq.data_ptr() + r * seq_bytes, n, 150, stride, 0, None, d_segs.data_ptr() + n * 8, n_seg - n, extra_max,
                            d_out.data_ptr() + r * n_seg * 8, st)

    for i in range(3):
        call(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(args.calls):
        call(i)
    e1.record()
    torch.cuda.synchronize()
    g.device_status(st)
    us = e0.elapsed_time(e1) * 1e3 / args.calls
    res = d_out[: n_seg * 8].cpu().numpy().view(sb.REPEAT_DTYPE)
    print(json.dumps({"max_stage": os.environ.get("STRGPU_MAX_STAGE", "all"), "us_per_call": us, "reads": n, "segments": n_seg,
                      "found": int((res["repeat_count"] > 0).sum()), "variant": os.environ.get("STRGPU_SCAN_VARIANT", "0")}))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=12_000_000)
    ap.add_argument("--calls", type=int, default=20)
    ap.add_argument("--regions", type=int, default=4)
    ap.add_argument("--align", type=int, default=16)
    ap.add_argument("--sweep", action="store_true", help="run once per STRGPU_MAX_STAGE value (subprocesses)")
    args = ap.parse_args()
    if args.sweep:
        for k in ("0", "2", "99"):
            env = dict(os.environ, STRGPU_MAX_STAGE=k)
            subprocess.run([sys.executable, os.path.abspath(__file__), "--reads", str(args.reads), "--calls", str(args.calls),
                            "--regions", str(args.regions), "--align", str(args.align)], env=env, check=True)
        return
    run(args)


if __name__ == "__main__":
    main()
