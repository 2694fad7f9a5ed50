"""Multi-GPU plumbing for the hot path (one process per GPU, torch.distributed: NCCL on GPUs, gloo in CPU tests).

The scan shards with no exchange: read batches are independent, batch i goes to rank i % world.  The only collective on
the path is the all-gather of per-shard cluster records (48-byte strgpu_bounds PODs) at the merge step; the payload is
tiny (<= ~1e5 records), so it is latency- not bandwidth-bound."""
from __future__ import annotations

import numpy as np
import torch
import torch.distributed as dist

from .binding import BOUNDS_DTYPE


def batches_of_rank(n_batches: int, rank: int, world: int):
    """Round-robin batch ownership: batch i -> rank i % world."""
    return list(range(rank, n_batches, world))


def allgather_records(local: torch.Tensor, n_local: int, record_bytes: int = BOUNDS_DTYPE.itemsize):
    """local: uint8 tensor holding n_local fixed-size records (capacity may be larger) on the collective's device.
    Returns (uint8 tensor with every rank's records concatenated in rank order, per-rank counts)."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return local[: n_local * record_bytes].clone(), [n_local]
    dev = local.device
    counts = torch.zeros(world, dtype=torch.int64, device=dev)
    mine = torch.tensor([n_local], dtype=torch.int64, device=dev)
    dist.all_gather_into_tensor(counts, mine)
    counts_h = [int(c) for c in counts.cpu()]
    cap = max(max(counts_h), 1) * record_bytes
    send = torch.zeros(cap, dtype=torch.uint8, device=dev)
    send[: n_local * record_bytes] = local[: n_local * record_bytes]
    recv = torch.empty(world * cap, dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(recv, send)
    parts = [recv[r * cap: r * cap + counts_h[r] * record_bytes] for r in range(world)]
    return torch.cat(parts), counts_h


def bounds_from_bytes(buf: torch.Tensor) -> np.ndarray:
    return buf.cpu().numpy().view(BOUNDS_DTYPE)
