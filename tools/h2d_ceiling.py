#!/usr/bin/env python
"""Host-side ceiling of the end-to-end leg: N processes (one per GPU, like bench.py) each copy the bench's pinned 456 MB batch
host->device and an 98 MB result device->host, back to back, with NO kernels in between.  What this reaches per GPU at N ranks is
the most the box's host memory / PCIe complex gives the e2e leg of bench.py (which moves exactly these buffers around its kernels).
   python -m torch.distributed.run --nnodes=1 --nproc-per-node=N --master-addr 127.0.0.1 --master-port 29581 tools/h2d_ceiling.py"""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    h2d_bytes, d2h_bytes, reps = 456_000_000, 98_000_000, 30
    h_in = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    h_out = torch.empty(d2h_bytes, dtype=torch.uint8).pin_memory()
    d_in = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    d_out = torch.empty(d2h_bytes, dtype=torch.uint8, device=dev)
    s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
    res = {}
    for mode in ("h2d", "h2d+d2h"):
        for _ in range(3):
            d_in.copy_(h_in, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            with torch.cuda.stream(s_in):
                d_in.copy_(h_in, non_blocking=True)
            if mode != "h2d":
                with torch.cuda.stream(s_out):
                    h_out.copy_(d_out, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        t = torch.tensor([dt], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        res[mode] = {"h2d_gb_per_s_per_gpu": h2d_bytes * reps / float(t) / 1e9,
                     "d2h_gb_per_s_per_gpu": (d2h_bytes * reps / float(t) / 1e9) if mode != "h2d" else 0.0}
    if rank == 0:
        print(json.dumps({"n_gpus": world, "copies": res, "note": "pinned host buffers of bench.py's e2e leg, no kernels"}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
