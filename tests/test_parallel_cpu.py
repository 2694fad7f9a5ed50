"""world_size-2 gloo test of the only collective on the path: the all-gather of per-shard cluster records, plus the
batch -> rank ownership rule.  Runs on CPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from strling_b200 import parallel
from strling_b200.binding import BOUNDS_DTYPE


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _records(rank, n):
    b = np.zeros(n, dtype=BOUNDS_DTYPE)
    b["tid"] = rank
    b["left"] = np.arange(n) * 10 + rank
    b["right"] = b["left"] + 1
    b["repeat"] = b"CAG" if rank == 0 else b"AAAG"
    b["n_total"] = 5 + rank
    return b


def _worker(rank, world, port, counts, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = _records(rank, counts[rank])
    buf = torch.zeros(max(counts[rank], 1) * BOUNDS_DTYPE.itemsize + 96, dtype=torch.uint8)  # capacity > payload
    buf[: mine.nbytes] = torch.from_numpy(mine.view(np.uint8).copy())
    allb, got_counts = parallel.allgather_records(buf, counts[rank])
    res = parallel.bounds_from_bytes(allb)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), res)
    assert got_counts == list(counts)
    dist.destroy_process_group()


def test_allgather_of_cluster_records_world2(tmp_path):
    for counts in ((3, 5), (0, 4), (7, 0)):
        port = _free_port()
        mp.spawn(_worker, args=(2, port, counts, str(tmp_path)), nprocs=2, join=True)
        expect = np.concatenate([_records(0, counts[0]), _records(1, counts[1])])
        for r in range(2):
            got = np.load(tmp_path / f"r{r}.npy")
            assert got.dtype == BOUNDS_DTYPE and np.array_equal(got, expect)


def test_batch_ownership_partitions_all_batches():
    for world in (1, 2, 4, 8):
        owned = [parallel.batches_of_rank(37, r, world) for r in range(world)]
        assert sorted(sum(owned, [])) == list(range(37))
        assert all(b % world == r for r, bs in enumerate(owned) for b in bs)


def test_single_process_passthrough():
    mine = _records(0, 4)
    buf = torch.from_numpy(mine.view(np.uint8).copy())
    allb, counts = parallel.allgather_records(buf, 4)
    assert counts == [4] and np.array_equal(parallel.bounds_from_bytes(allb), mine)
