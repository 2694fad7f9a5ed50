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


# ------------------------------------------------------------------------------------------------ sharded clustering
# Clustering shards by bucket: reads are grouped by (tid, repeat) before anything else happens to them (call.nim:124-125,
# merge.nim:125), buckets never interact, and the order inside a bucket is the `.bin` / concatenation order.  So every
# bucket gets an owner rank, the STR-read records travel to their owner ONCE (the one real exchange on the path:
# all-to-all over NVLink, rank-major arrival order == concatenation order), every rank clusters what it owns, and the
# 48-byte cluster records are all-gathered.  The merged result equals one GPU clustering the concatenation of all shards.
TREAD_WORDS = 6  # strgpu_tread = 24 bytes = six int32 words: tid, position, repeat[0..3], repeat[4..5] | flag << 16, ...


def bucket_owner(t32: torch.Tensor, world: int) -> torch.Tensor:
    """t32: int32 [n, 6] view of strgpu_tread records.  Owner rank of each record's (tid, repeat) bucket."""
    tid = t32[:, 0].to(torch.int64)
    r0 = t32[:, 2].to(torch.int64) & 0xFFFFFFFF
    r1 = t32[:, 3].to(torch.int64) & 0xFFFF
    h = tid * 0x9E3779B1 + r0 * 0x85EBCA6B + r1 * 0xC2B2AE35
    h = h ^ (h >> 29)
    return torch.remainder(h, world)


def exchange_by_owner(t32: torch.Tensor):
    """Sends every record to the rank that owns its bucket.  Returns int32 [m, 6]: the records this rank owns, ordered by
    source rank and, within a source, in their original order."""
    world = dist.get_world_size() if dist.is_initialized() else 1
    if world == 1:
        return t32
    dev = t32.device
    owner = bucket_owner(t32, world)
    # stable partition by owner: a one-pass radix sort on 8-bit keys (world <= 256) instead of 64-bit ones
    order = torch.sort(owner.to(torch.uint8) if world <= 256 else owner, stable=True).indices
    send = t32[order].contiguous()
    send_counts = torch.bincount(owner, minlength=world).to(torch.int64)
    if dist.get_backend() == "nccl":
        recv_counts = torch.empty_like(send_counts)
        dist.all_to_all_single(recv_counts, send_counts)
        rc, sc = [int(x) for x in recv_counts.cpu()], [int(x) for x in send_counts.cpu()]
        recv = torch.empty((sum(rc), TREAD_WORDS), dtype=torch.int32, device=dev)
        dist.all_to_all_single(recv, send, output_split_sizes=rc, input_split_sizes=sc)
        return recv
    # backends without all-to-all (gloo on CPU, used by the tests): all-gather the partitioned records and take my slices
    rank = dist.get_rank()
    counts = [torch.empty_like(send_counts) for _ in range(world)]
    dist.all_gather(counts, send_counts)
    n_max = max(int(c.sum()) for c in counts)
    padded = torch.zeros((max(n_max, 1), TREAD_WORDS), dtype=torch.int32, device=dev)
    padded[: send.shape[0]] = send
    everyone = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(everyone, padded)
    parts = []
    for src in range(world):
        c = [int(x) for x in counts[src]]
        off = sum(c[:rank])
        parts.append(everyone[src][off: off + c[rank]])
    return torch.cat(parts)


def sort_bounds(b: np.ndarray) -> np.ndarray:
    """Rank-major gathered records -> the order one GPU emits: ascending (tid, repeat), position order kept inside a bucket."""
    key_rep = np.frombuffer(b["repeat"].tobytes(), dtype=np.uint8).reshape(-1, 6) if len(b) else np.zeros((0, 6), dtype=np.uint8)
    keys = [key_rep[:, j] for j in range(5, -1, -1)] + [b["tid"]]
    return b[np.lexsort(keys)]


def cluster_sharded(cluster_fn, t32: torch.Tensor):
    """cluster_fn(int32 [m, 6] records on t32's device) -> (uint8 tensor of 48-byte bounds records, count).
    Returns (numpy bounds of the whole job in single-GPU order, per-rank counts).  `first_read` indexes the owner rank's
    sorted record array."""
    mine = exchange_by_owner(t32)
    local, n_local = cluster_fn(mine)
    allb, counts = allgather_records(local, n_local)
    return sort_bounds(bounds_from_bytes(allb)), counts
