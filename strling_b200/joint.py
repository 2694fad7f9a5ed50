"""`strling merge` over several GPUs (N samples' `.bin` files -> joint -bounds.txt).

    torchrun --nproc-per-node N -m strling_b200.joint -o joint [-m 5] [-c 0] [-t 0] [-w -1] a.bin b.bin ...

One process per GPU.  The input files are split into contiguous blocks, one per rank (so that rank-major order is file order,
the concatenation order of merge.nim:95-125); every rank parses its files, the fragment-length histograms are summed with an
all-reduce (merge.nim:112-115: window = the 0.98 quantile of the sum); then libstrgpu.so's sharded clustering
(strgpu_cluster_sharded: the STR-read records travel to the owner of their (tid, repeat) bucket over NCCL, every rank clusters
what it owns in merge mode (has_per_sample_reads, merge.nim:18-25) on its GPU, the 48-byte cluster records are all-gathered)
and rank 0 writes `<prefix>-bounds.txt` -- the same lines, in the same order, as
`strling merge` on one GPU.  `-l` loci and `--chromosome` are the single-GPU command's business (strling_b200/bin/strling merge).
The record parsing below reads the `.bin` layout of extract.nim:331-348 / cluster.nim:38-50; it contains no clustering logic:
that is the CUDA library's (strgpu_cluster_sharded)."""
from __future__ import annotations

import argparse
import os
import struct
import sys

import numpy as np
import torch
import torch.distributed as dist

from . import parallel
from .binding import BOUNDS_DTYPE, TREAD_DTYPE

BOUNDS_HEADER = "#chrom\tleft\tright\trepeat\tname\tleft_most\tright_most\tcenter_mass\tn_left\tn_right\tn_total"


def read_bin(path: str):
    """-> (frag_dist uint32[4096], header text, strgpu_tread records without the unplaced ones)  (unpack.nim:58-133)"""
    import msgpack

    data = open(path, "rb").read()
    if data[:3] != b"STR":
        raise SystemExit(f"[strling] {path}: not a STRling bin file")
    (fmt,) = struct.unpack_from("<h", data, 3)
    if fmt != 0:
        raise SystemExit(f"[strling] {path}: unsupported bin format version {fmt}")
    frag = np.frombuffer(data, dtype="<u4", count=4096, offset=19).copy()
    off = 19 + 16384
    (hl,) = struct.unpack_from("<i", data, off)
    header = data[off + 4: off + 4 + hl].decode()
    off += 4 + hl
    (n,) = struct.unpack_from("<i", data, off)
    up = msgpack.Unpacker(raw=True)
    up.feed(data[off + 4:])
    vals = list(up)
    if len(vals) != 10 * n:
        raise SystemExit(f"[strling] {path}: expected {n} records, found {len(vals) // 10}")
    t = np.zeros(n, dtype=TREAD_DTYPE)
    t["tid"] = vals[0::10]
    t["position"] = vals[1::10]
    t["repeat"] = [bytes(v).rstrip(b"\0") for v in vals[2::10]]
    t["flag"], t["split"], t["mapq"] = vals[3::10], vals[4::10], vals[5::10]
    t["repeat_count"], t["align_length"] = vals[6::10], vals[7::10]
    return frag, header, t[t["tid"] >= 0]      # drop_unplaced = true (merge.nim:101)


def targets_from_header(header: str):
    out = []
    for line in header.splitlines():
        if line.startswith("@SQ"):
            f = dict(x.split(":", 1) for x in line.split("\t")[1:] if ":" in x)
            out.append((f["SN"], int(f["LN"])))
    return out


def frag_median(frag: np.ndarray, pct: float) -> int:  # utils.nim:139-146
    n = int(frag.sum()) & 0xFFFFFFFF
    target = int(0.5 + float(n) / (1.0 / pct)) & 0xFFFFFFFF
    count = 0
    for i in range(4096):
        count = (count + int(frag[i])) & 0xFFFFFFFF
        if count >= target:
            return i
    return 4096


def bounds_line(b, targets) -> str:  # cluster.nim:262-266
    rep = bytes(b["repeat"]).rstrip(b"\0").decode()
    return (f"{targets[int(b['tid'])][0]}\t{b['left']}\t{b['right']}\t{rep}\t\t{b['left_most']}\t{b['right_most']}\t{b['center_mass']}\t"
            f"{b['n_left']}\t{b['n_right']}\t{b['n_total']}")


def files_of_rank(n_files: int, rank: int, world: int):
    """Contiguous blocks: rank-major order == file order."""
    per = (n_files + world - 1) // world
    return list(range(rank * per, min(n_files, (rank + 1) * per)))


def joint_merge(paths, cluster_fn, device, window=-1, min_support=5, min_clip=0, min_clip_total=0, lib=None):
    """lib: a StrGpu context whose communicator spans the job (comm_init): the exchange, the per-rank clustering and the
    all-gather then all happen inside libstrgpu.so (strgpu_cluster_sharded) -- the GPU path.  Without it (CPU tests of the
    host logic over gloo) cluster_fn(int32 [m, 6] records on `device`, params dict) -> (uint8 tensor of 48-byte records, count)
    is called on what torch.distributed delivered.
    Returns (bounds lines, or None on ranks other than 0; per-rank record counts)."""
    rank = dist.get_rank() if dist.is_initialized() else 0
    world = dist.get_world_size() if dist.is_initialized() else 1
    frag = np.zeros(4096, dtype=np.int64)
    parts, header = [], None
    for fi in files_of_rank(len(paths), rank, world):
        f, h, t = read_bin(paths[fi])
        header = header or h
        frag += f
        t["sample"] = fi                      # qname := sample index (merge.nim:121-124)
        parts.append(t)
    # every rank needs the header of file 0 (targets) and the summed fragment histogram
    hdr0 = [header if rank == 0 else None]
    if world > 1:
        dist.broadcast_object_list(hdr0, src=0)
        fr = torch.from_numpy(frag).to(device)
        dist.all_reduce(fr)
        frag = fr.cpu().numpy()
    if int(frag.max()) >= 2 ** 32:
        raise SystemExit("overflow")          # merge.nim:112-115
    targets = targets_from_header(hdr0[0])
    if header is not None and targets_from_header(header) != targets:
        raise SystemExit("[strling] Error: inconsistent bam header. Were all samples run on the same reference genome?")
    frag = frag.astype(np.uint32)
    if window < 0:
        window = frag_median(frag, 0.98)      # merge.nim:148-152
    params = dict(window=window, min_support=min_support, min_clip=min_clip, min_clip_total=min_clip_total,
                  max_clip_dist=int(0.5 * float(frag_median(frag, 0.5))) & 0xFFFF, merge_mode=True)
    mine = np.concatenate(parts) if parts else np.zeros(0, dtype=TREAD_DTYPE)
    if lib is not None:
        max_n = len(mine)
        if world > 1:
            m = torch.tensor([max_n], dtype=torch.int64, device=device)
            dist.all_reduce(m, op=dist.ReduceOp.MAX)
            max_n = int(m.item())
        if world > 1:
            res, _ = lib.cluster_sharded(mine, max(max_n, 1), **params)
        else:
            res, _ = lib.cluster(mine, **params)
        if rank != 0:
            return None, [len(res)]
        return [bounds_line(b, targets) for b in res if b["tid"] >= 0], [len(res)]
    t32 = torch.from_numpy(mine.view(np.uint8).reshape(-1).view(np.int32).reshape(-1, 6).copy()).to(device)
    res, counts = parallel.cluster_sharded(lambda owned: cluster_fn(owned, params), t32)
    if rank != 0:
        return None, counts
    return [bounds_line(b, targets) for b in res if b["tid"] >= 0], counts


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m strling_b200.joint", description=__doc__.split("\n\n")[0])
    ap.add_argument("-w", "--window", type=int, default=-1)
    ap.add_argument("-m", "--min-support", type=int, default=5)
    ap.add_argument("-c", "--min-clip", type=int, default=0)
    ap.add_argument("-t", "--min-clip-total", type=int, default=0)
    ap.add_argument("-o", "--output-prefix", default="strling")
    ap.add_argument("bins", nargs="+")
    a = ap.parse_args(argv)
    import strling_b200 as sb

    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("[strling] joint merge: no CUDA device (the cluster kernels have no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    with sb.StrGpu(local) as g:
        if world > 1:
            g.comm_init_torch()
        lines, counts = joint_merge(a.bins, None, dev, a.window, a.min_support, a.min_clip, a.min_clip_total, lib=g)
    if lines is not None:
        with open(a.output_prefix + "-bounds.txt", "w") as fh:
            fh.write(BOUNDS_HEADER + "\n" + "".join(l + "\n" for l in lines))
        print(f"[strling] joint merge over {world} GPU(s): {sum(counts)} cluster records -> {a.output_prefix}-bounds.txt", file=sys.stderr)
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
