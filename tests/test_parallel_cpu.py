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


# ---- sharded clustering: exchange by bucket owner + per-rank clustering + all-gather == one process over the concatenation
def _shard_treads(rank):
    from strling_b200 import synth

    return synth.make_treads(60, seed=500 + rank, n_tids=5, n_samples=2, noise_reads=400, unplaced=30, dense=True)


def _oracle_cluster_fn(t32):
    # CPU stand-in for strgpu_cluster_device in this test: the oracle's cluster loop (tests may use the oracle)
    from oracle import oracle as orc

    treads = t32.numpy().view(np.uint8).reshape(-1).view(orc.TREAD_DTYPE) if t32.shape[0] else np.zeros(0, dtype=orc.TREAD_DTYPE)
    b, unplaced = orc.cluster_all(treads, 480, 3, 0, 0, 190)
    out = np.zeros(len(b) + len(unplaced), dtype=BOUNDS_DTYPE)
    for f in ("tid", "left", "left_most", "right", "right_most", "center_mass", "n_left", "n_right", "n_total", "repeat", "n_reads"):
        out[f][: len(b)] = b[f]
    for i, (unit, cnt) in enumerate(sorted(unplaced.items())):
        out["tid"][len(b) + i], out["repeat"][len(b) + i], out["n_reads"][len(b) + i] = -1, unit, cnt
    return torch.from_numpy(out.view(np.uint8).reshape(-1).copy()), len(out)


def _sharded_worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = _shard_treads(rank)
    t32 = torch.from_numpy(mine.view(np.uint8).reshape(-1).view(np.int32).reshape(-1, 6).copy())
    owned = parallel.exchange_by_owner(t32)
    # every record of a bucket lands on one rank, in concatenation order
    own = parallel.bucket_owner(owned, world)
    assert bool((own == rank).all())
    res, counts = parallel.cluster_sharded(_oracle_cluster_fn, t32)
    np.save(os.path.join(out_dir, f"s{rank}.npy"), res)
    np.save(os.path.join(out_dir, f"o{rank}.npy"), owned.numpy())
    dist.destroy_process_group()


def test_sharded_clustering_equals_single_process_world2(tmp_path):
    port = _free_port()
    mp.spawn(_sharded_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    allt = np.concatenate([_shard_treads(0), _shard_treads(1)])
    t32 = torch.from_numpy(allt.view(np.uint8).reshape(-1).view(np.int32).reshape(-1, 6).copy())
    raw, n = _oracle_cluster_fn(t32)
    expect = parallel.sort_bounds(parallel.bounds_from_bytes(raw[: n * BOUNDS_DTYPE.itemsize]))
    # owned records: a partition of the concatenation that keeps its order inside every bucket
    owned = [np.load(tmp_path / f"o{r}.npy") for r in range(2)]
    assert sum(len(o) for o in owned) == len(allt)
    for r in range(2):
        got = np.load(tmp_path / f"s{r}.npy")
        assert len(got) == len(expect) and len(expect) > 20
        for f in ("tid", "left", "left_most", "right", "right_most", "center_mass", "n_left", "n_right", "n_total", "repeat", "n_reads"):
            assert np.array_equal(got[f], expect[f]), f


# ---- config 5: joint merge of several samples' .bin files over ranks == `strling merge` semantics on one process
def _joint_oracle_cluster_fn(t32, params):
    from oracle import oracle as orc

    treads = t32.numpy().view(np.uint8).reshape(-1).view(orc.TREAD_DTYPE) if t32.shape[0] else np.zeros(0, dtype=orc.TREAD_DTYPE)
    b, _ = orc.cluster_all(treads, params["window"], params["min_support"], params["min_clip"], params["min_clip_total"],
                           params["max_clip_dist"], merge_mode=True)
    out = np.zeros(len(b), dtype=BOUNDS_DTYPE)
    for f in ("tid", "left", "left_most", "right", "right_most", "center_mass", "n_left", "n_right", "n_total", "repeat", "n_reads"):
        out[f] = b[f]
    return torch.from_numpy(out.view(np.uint8).reshape(-1).copy()), len(out)


def _joint_worker(rank, world, port, paths, out_dir, kw):
    from strling_b200 import joint

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    lines, counts = joint.joint_merge(paths, _joint_oracle_cluster_fn, torch.device("cpu"), **kw)
    if rank == 0:
        open(os.path.join(out_dir, "joint.txt"), "w").write("\n".join(lines))
    else:
        assert lines is None
    dist.destroy_process_group()


def test_joint_merge_world2_matches_single_process_merge(tmp_path):
    from oracle import extract_oracle as eo
    from strling_b200 import bamio, joint

    targets = [("chr1", 400_000), ("chr2", 300_000)]
    loci = [(0, 50_000, 50_060, "CAG"), (0, 220_000, 220_040, "AAAG"), (1, 80_000, 80_090, "AC"), (1, 150_000, 150_030, "CCG")]
    hdr = bamio.sam_header(targets)
    paths, datas = [], []
    for s in range(5):
        recs = bamio.simulate_alignments(70 + s, 2500, targets, loci, str_pair_frac=0.4, unmapped_pairs=20)
        data, _, _ = eo.extract(recs, targets, hdr)
        p = str(tmp_path / f"s{s}.bin")
        open(p, "wb").write(data)
        paths.append(p)
        datas.append(data)
    assert joint.files_of_rank(5, 0, 2) == [0, 1, 2] and joint.files_of_rank(5, 1, 2) == [3, 4]
    for kw in (dict(min_support=3), dict(min_support=2, min_clip=1, min_clip_total=1), dict(min_support=4, window=300)):
        port = _free_port()
        mp.spawn(_joint_worker, args=(2, port, paths, str(tmp_path), kw), nprocs=2, join=True)
        got = [l for l in open(tmp_path / "joint.txt").read().split("\n") if l]
        exp, _ = eo.merge(datas, **kw)
        assert got == exp and (len(exp) > 3 or "min_clip" in kw)


def test_joint_helpers_agree_with_the_oracle(tmp_path):
    # the product-side .bin reader and quantile of strling_b200/joint.py against the oracle's (unpack.nim:58-133, utils.nim:139-146)
    from oracle import extract_oracle as eo
    from oracle import oracle as orc
    from strling_b200 import bamio, joint

    targets = [("chr1", 200_000)]
    loci = [(0, 50_000, 50_060, "CAG")]
    recs = bamio.simulate_alignments(3, 1500, targets, loci, str_pair_frac=0.5, unmapped_pairs=30)
    data, _, _ = eo.extract(recs, targets, bamio.sam_header(targets))
    p = str(tmp_path / "x.bin")
    open(p, "wb").write(data)
    frag, header, t = joint.read_bin(p)
    u = eo.unpack_bin(data)
    keep = u["treads"]["tid"] >= 0
    assert header == u["header"] and np.array_equal(frag, u["frag_dist"]) and len(t) == int(keep.sum()) and len(t) > 100
    for f in ("tid", "position", "repeat", "flag", "split", "mapq", "repeat_count", "align_length"):
        assert np.array_equal(t[f], u["treads"][f][keep]), f
    assert joint.targets_from_header(header) == targets
    rng = np.random.default_rng(1)
    for _ in range(20):
        fd = rng.integers(0, 5000, size=4096).astype(np.uint32) * (rng.random(4096) < 0.3)
        for pct in (0.5, 0.98, 0.99, 0.1):
            assert joint.frag_median(fd.astype(np.uint32), pct) == orc.median(fd.astype(np.uint32), pct)
