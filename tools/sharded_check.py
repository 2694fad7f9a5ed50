#!/usr/bin/env python
"""strgpu_cluster_sharded (NCCL inside libstrgpu.so) against ONE GPU clustering the concatenated shards, on every visible GPU:
   python -m torch.distributed.run --nnodes=1 --nproc-per-node=N --master-addr 127.0.0.1 --master-port 29555 tools/sharded_check.py [loci_per_rank]
Checks the host-buffer entry point and the device-resident one (default pair capacity), call and merge semantics, and
times the device-resident path with CUDA events (max over ranks)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import strling_b200 as sb  # noqa: E402
from strling_b200 import synth  # noqa: E402

FIELDS = ("tid", "left", "left_most", "right", "right_most", "center_mass", "n_left", "n_right", "n_total", "repeat", "n_reads")


def main():
    n_loci = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    g = sb.StrGpu(local)
    g.comm_init_torch()
    shards = [synth.make_treads(n_loci + 37 * r, seed=500 + r, n_samples=3, noise_reads=20 * n_loci, unplaced=n_loci // 10) for r in range(world)]
    mine = shards[rank]
    max_n = max(len(s) for s in shards)
    cat = np.concatenate(shards)
    ok = True
    for merge_mode in (False, True):
        kw = dict(window=480, min_support=5, max_clip_dist=190, merge_mode=merge_mode)
        one_b, one_u = g.cluster(cat, **kw)
        many_b, many_u = g.cluster_sharded(mine, max_n, **kw)
        same = len(one_b) == len(many_b) and all(np.array_equal(one_b[f], many_b[f]) for f in FIELDS) and one_u == many_u
        # device-resident entry point, default (hash-balanced) pair capacity
        p = g.cluster_params(**kw)
        d_t = torch.from_numpy(mine.view(np.uint8).reshape(-1).copy()).to(dev)
        cap = max(1024, 2 * len(cat))
        d_out = torch.zeros(cap * 48, dtype=torch.uint8, device=dev)
        d_n = torch.zeros(1, dtype=torch.int32, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        g.cluster_sharded_device(d_t.data_ptr(), len(mine), max_n, p, d_out.data_ptr(), cap, d_n.data_ptr(), st, pair_capacity=0)
        try:
            g.comm_status(st)
            n_tot = int(d_n.item())
            dev_b = d_out[: n_tot * 48].cpu().numpy().view(sb.BOUNDS_DTYPE)
            dev_b = dev_b[dev_b["tid"] >= 0] if not merge_mode else dev_b
            same_dev = len(dev_b) == len(one_b) and all(np.array_equal(one_b[f], dev_b[f]) for f in FIELDS)
        except sb.StrGpuError as e:
            same_dev = f"overflow ({e})"
        if rank == 0:
            print(f"merge_mode={merge_mode}: {len(cat)} treads over {world} ranks -> {len(one_b)} bounds, {len(one_u)} unplaced units; "
                  f"host API == one GPU: {same}; device API == one GPU: {same_dev}", flush=True)
        ok = ok and same and same_dev is True
    # timing of the device path
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        g.cluster_sharded_device(d_t.data_ptr(), len(mine), max_n, p, d_out.data_ptr(), cap, d_n.data_ptr(), st)
    torch.cuda.synchronize()
    dist.barrier()
    e0.record()
    for _ in range(10):
        g.cluster_sharded_device(d_t.data_ptr(), len(mine), max_n, p, d_out.data_ptr(), cap, d_n.data_ptr(), st)
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 10], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    # the timed calls were CUDA-graph replays (unchanged arguments): the result must still be the single-GPU one
    g.comm_status(st)
    n_tot = int(d_n.item())
    dev_b = d_out[: n_tot * 48].cpu().numpy().view(sb.BOUNDS_DTYPE)
    replay_ok = len(dev_b) == len(one_b) and all(np.array_equal(one_b[f], dev_b[f]) for f in FIELDS)
    ok = ok and replay_ok
    if rank == 0:
        print(f"sharded cluster (device-resident, {len(mine)} treads/rank): {float(t):.3f} ms per call (max over ranks); "
              f"after 13 replays == one GPU: {replay_ok}", flush=True)
    flag = torch.tensor([0 if ok else 1], device=dev)
    dist.all_reduce(flag)
    g.close()
    dist.destroy_process_group()
    if int(flag.item()):
        raise SystemExit("sharded clustering differs from one GPU")


if __name__ == "__main__":
    main()
