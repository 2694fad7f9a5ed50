"""Parity of the CUDA cluster path (K2 sort, K3 chain, K4 bounds; through the C ABI) against the CPU oracle:
every Bounds field bit-exact, same order, same unplaced counts.  Needs a B200: run with -m gpu."""
import numpy as np
import pytest

import strling_b200 as sb
from oracle import oracle as orc
from strling_b200 import synth

pytestmark = pytest.mark.gpu

FIELDS = ["tid", "left", "left_most", "right", "right_most", "center_mass", "n_left", "n_right", "n_total", "repeat",
          "first_read", "n_reads"]


@pytest.fixture(scope="module")
def gpu():
    g = sb.StrGpu(0)
    yield g
    g.close()


def compare(gpu, treads, **kw):
    exp, exp_unplaced = orc.cluster_all(treads.astype(orc.TREAD_DTYPE), kw["window"], kw["min_support"], kw.get("min_clip", 0),
                                        kw.get("min_clip_total", 0), kw.get("max_clip_dist", 200), kw.get("merge_mode", False))
    got, got_unplaced = gpu.cluster(treads, **kw)
    assert len(got) == len(exp), (len(got), len(exp))
    for f in FIELDS:
        bad = np.nonzero(got[f] != exp[f])[0]
        assert len(bad) == 0, (f, int(bad[0]), got[bad[0]], exp[bad[0]])
    assert got_unplaced == exp_unplaced
    return got


def t(position, split=synth.SOFT_NONE, tid=1, repeat=b"ATG", sample=0):
    r = np.zeros(1, dtype=synth.TREAD_DTYPE)
    r["tid"], r["position"], r["split"], r["repeat"], r["sample"] = tid, position, split, repeat, sample
    return r


def test_reference_cluster_vectors(gpu):
    # tests/test_cluster.nim:38-55 : one cluster of 4 reads chr1 1..200
    reads = np.concatenate([t(p, repeat=b"AAAAAT") for p in (1, 1, 1, 200, 255)])
    got = compare(gpu, reads, window=125, min_support=3)
    assert len(got) == 1 and got[0]["n_reads"] == 4
    # tests/test_cluster.nim:200-228 : split into 6 + 5 reads
    spec = [(370, 3), (391, 1), (391, 1), (391, 1), (403, 3), (503, 3), (850, 0), (850, 0), (850, 0), (850, 0), (880, 3)]
    reads = np.concatenate([t(p, s, tid=0, repeat=b"CAG") for p, s in spec])
    got = compare(gpu, reads, window=500, min_support=1)
    assert [int(x) for x in got["n_reads"]] == [6, 5]
    # tests/test_cluster.nim:58-79 : bounds from clip modes
    spec = [(123, 3), (123, 3)] + [(223, 0)] * 4 + [(253, 1)] * 4 + [(283, 3)]
    reads = np.concatenate([t(p, s) for p, s in spec])
    got = compare(gpu, reads, window=500, min_support=1)
    assert (got[0]["left"], got[0]["right"]) == (223, 253)


@pytest.mark.parametrize("seed,dense", [(1, False), (2, True), (3, True)])
def test_random_loci_call_mode(gpu, seed, dense):
    treads = synth.make_treads(3000, seed=seed, noise_reads=20000, unplaced=500, dense=dense, n_tids=6 if dense else 24)
    for window, ms in ((500, 5), (125, 3), (40, 1)):
        got = compare(gpu, treads, window=window, min_support=ms, max_clip_dist=150)
        assert len(got) > 10


def test_clip_filters_and_merge_mode(gpu):
    treads = synth.make_treads(4000, seed=9, n_samples=7, noise_reads=10000, unplaced=300)
    compare(gpu, treads, window=450, min_support=5, min_clip=1, min_clip_total=3, max_clip_dist=200, merge_mode=True)
    compare(gpu, treads, window=450, min_support=2, min_clip=0, min_clip_total=2, max_clip_dist=30, merge_mode=True)
    compare(gpu, treads, window=450, min_support=5, min_clip=2, min_clip_total=0, max_clip_dist=200, merge_mode=False)


def test_clip_position_ties_follow_nim_counttable_order(gpu):
    # many distinct clip positions with equal counts: `largest` must pick Nim's first-slot winner
    rng = np.random.default_rng(5)
    parts = []
    for locus in range(300):
        base = 10_000 + locus * 5_000
        k = int(rng.integers(2, 40))
        lefts = base + rng.choice(400, size=k, replace=False)
        rights = base + 500 + rng.choice(400, size=k, replace=False)
        reps = int(rng.integers(2, 4))
        parts += [t(int(p), synth.SOFT_LEFT, tid=2, repeat=b"CAG") for p in np.repeat(lefts, reps)]
        parts += [t(int(p), synth.SOFT_RIGHT, tid=2, repeat=b"CAG") for p in np.repeat(rights, reps)]
        parts += [t(base + 450 + int(d), synth.SOFT_NONE, tid=2, repeat=b"CAG") for d in rng.integers(0, 100, size=6)]
    treads = np.concatenate(parts)
    treads = treads[rng.permutation(len(treads))]
    got = compare(gpu, treads, window=600, min_support=2, max_clip_dist=2000)
    assert len(got) >= 250


def test_wraparound_and_edges(gpu):
    # positions near 0 (posmed - max_dist wraps, cluster.nim:344) and near 2^32 (posmed + max_dist + 100 wraps, :336)
    reads = np.concatenate([t(p, tid=0) for p in (0, 0, 3, 5, 5, 9, 40)] +
                           [t(p, tid=3) for p in (4294967000, 4294967100, 4294967200, 4294967290, 4294967295)])
    compare(gpu, reads, window=500, min_support=2)
    # empty input and a single read
    got, unplaced = gpu.cluster(np.zeros(0, dtype=synth.TREAD_DTYPE), window=500, min_support=1)
    assert len(got) == 0 and unplaced == {}
    compare(gpu, t(77), window=500, min_support=1)
    # a cluster above the uint16 limit is skipped (callclusters.nim:53-55)
    big = np.concatenate([np.repeat(t(1000, tid=5, repeat=b"A"), 70000), np.repeat(t(90000, tid=5, repeat=b"A"), 10)])
    got = compare(gpu, big, window=500, min_support=5)
    assert len(got) == 1 and got[0]["n_reads"] == 10


def test_large_scale_properties(gpu):
    # 3e6 treads: beyond what the oracle comparison needs, checked through invariants + oracle on the same data
    treads = synth.make_treads(150_000, seed=12, noise_reads=600_000, unplaced=5_000)
    got = compare(gpu, treads, window=480, min_support=5, max_clip_dist=190)
    assert np.all(got["left"] <= got["right"]) and np.all(got["left_most"] <= got["left"]) and np.all(got["right_most"] >= got["right"])
    key = np.stack([got["tid"].astype(np.int64), got["first_read"].astype(np.int64)], axis=1)
    assert np.all((key[1:, 0] > key[:-1, 0]) | ((key[1:, 0] == key[:-1, 0]) & (key[1:, 1] > key[:-1, 1])))  # output is ordered


def test_assign_reads_locus_then_cluster(gpu):
    # callclusters.nim:14-50 before the cluster loop (merge -l / call -l -b): loci of one bucket are sequential, the
    # read right after each window is dropped too, unknown buckets and empty windows are no-ops
    rng = np.random.default_rng(17)
    for seed, dense in ((5, True), (6, False)):
        treads = synth.make_treads(2000, seed=seed, noise_reads=15000, unplaced=200, dense=dense, n_tids=4 if dense else 24, n_samples=3)
        units = np.unique(treads["repeat"])
        k = 400
        loci = np.zeros(k, dtype=orc.LOCUS_DTYPE)
        loci["tid"] = rng.integers(-1, 5 if dense else 24, size=k)
        loci["repeat"] = rng.choice(np.concatenate([units, np.array([b"GGGGGC"], dtype="S6")]), size=k)
        left = rng.integers(0, 6000, size=k)
        if not dense:  # centre the windows on real reads, and use their buckets
            pick = treads[rng.integers(0, len(treads), size=k)]
            left = np.maximum(pick["position"].astype(np.int64) - rng.integers(0, 700, size=k), 0)
            loci["tid"] = np.where(rng.random(k) < 0.8, pick["tid"], loci["tid"])
            loci["repeat"] = np.where(rng.random(k) < 0.8, pick["repeat"], loci["repeat"])
        loci["left_most"] = np.where(rng.random(k) < 0.1, 0, left)
        loci["right_most"] = loci["left_most"] + rng.integers(0, 1500, size=k)
        # overlapping windows in the same bucket exercise the sequential dependency
        loci[1::7] = loci[0::7][: len(loci[1::7])]
        for merge_mode in (False, True):
            el, eb, eu = orc.cluster_all_loci(treads.astype(orc.TREAD_DTYPE), loci, 450, 3, 0, 0, 180, merge_mode)
            gl, gb, gu = gpu.cluster_loci(treads, loci.astype(sb.binding.LOCUS_DTYPE), 450, 3, 0, 0, 180, merge_mode)
            for f in ("n_left", "n_right", "n_total"):
                assert np.array_equal(gl[f], el[f]), f
            assert el["n_total"].sum() > 100
            assert len(gb) == len(eb) and gu == eu
            for f in FIELDS:
                assert np.array_equal(gb[f], eb[f]), f


def test_sharded_entry_points_with_a_one_rank_communicator(gpu):
    # strgpu_comm_init + strgpu_cluster_sharded(_device) on a communicator of ONE rank (what a single-GPU box can run; 2 and 8
    # ranks: tools/sharded_check.py and bench.py --gpus N): owner partition through the peer table, device-side record count,
    # gather, final order -- the result must be strgpu_cluster's, in call and in merge mode
    import torch

    g = sb.StrGpu(0)
    try:
        g.comm_init(0, 1, sb.StrGpu.comm_unique_id())
        treads = synth.make_treads(1500, seed=77, n_samples=4, noise_reads=30_000, unplaced=300)
        for merge_mode in (False, True):
            kw = dict(window=480, min_support=4, max_clip_dist=190, merge_mode=merge_mode)
            one_b, one_u = gpu.cluster(treads, **kw)
            many_b, many_u = g.cluster_sharded(treads, len(treads) + 5, **kw)
            assert len(one_b) == len(many_b) > 100 and one_u == many_u
            for f in FIELDS:
                assert np.array_equal(one_b[f], many_b[f]), f
            dev = torch.device("cuda", 0)
            d_t = torch.from_numpy(treads.view(np.uint8).reshape(-1).copy()).to(dev)
            cap = 2 * len(treads)
            d_out = torch.zeros(cap * 48, dtype=torch.uint8, device=dev)
            d_n = torch.zeros(1, dtype=torch.int32, device=dev)
            st = torch.cuda.current_stream().cuda_stream
            g.cluster_sharded_device(d_t.data_ptr(), len(treads), len(treads), g.cluster_params(**kw), d_out.data_ptr(), cap, d_n.data_ptr(), st)
            g.comm_status(st)
            got = d_out[: int(d_n.item()) * 48].cpu().numpy().view(sb.BOUNDS_DTYPE)
            got = got[got["tid"] >= 0]
            assert len(got) == len(one_b) and all(np.array_equal(one_b[f], got[f]) for f in FIELDS)
        # a pair capacity that is too small is reported, not silently truncated
        g.cluster_sharded_device(d_t.data_ptr(), len(treads), len(treads), g.cluster_params(**kw), d_out.data_ptr(), cap, d_n.data_ptr(), st,
                                 pair_capacity=100)
        with pytest.raises(sb.StrGpuError) as e:
            g.comm_status(st)
        assert e.value.status == -7
    finally:
        g.close()


def test_repeated_device_calls_replay_a_cuda_graph_with_the_same_result(gpu):
    # strgpu_cluster_device with unchanged arguments: first call direct, second captured into a CUDA graph, then replayed;
    # new input in the same buffers must give the new result, changed arguments must leave the graph
    import torch

    dev = torch.device("cuda", 0)
    g = sb.StrGpu(0)
    try:
        kw = dict(window=480, min_support=4, max_clip_dist=190)
        p = g.cluster_params(**kw)
        a = synth.make_treads(900, seed=5, noise_reads=20_000, unplaced=200)
        b = synth.make_treads(900, seed=6, noise_reads=20_000, unplaced=200)[: len(a)]
        assert len(b) == len(a)
        d_t = torch.from_numpy(a.view(np.uint8).reshape(-1).copy()).to(dev)
        cap = len(a)
        d_out = torch.zeros(cap * 48, dtype=torch.uint8, device=dev)
        d_n = torch.zeros(1, dtype=torch.int32, device=dev)
        st = torch.cuda.Stream(device=dev)

        def run(treads):
            d_t.copy_(torch.from_numpy(treads.view(np.uint8).reshape(-1).copy()))
            torch.cuda.synchronize()
            g.cluster_device(d_t.data_ptr(), len(treads), p, d_out.data_ptr(), cap, d_n.data_ptr(), st.cuda_stream)
            st.synchronize()
            return d_out[: int(d_n.item()) * 48].cpu().numpy().view(sb.BOUNDS_DTYPE).copy()

        exp_a, _ = gpu.cluster(a, **kw)
        exp_b, _ = gpu.cluster(b, **kw)
        for rep, (treads, exp) in enumerate([(a, exp_a), (a, exp_a), (a, exp_a), (b, exp_b), (a, exp_a), (b, exp_b)]):
            got = run(treads)
            got = got[got["tid"] >= 0]
            assert len(got) == len(exp) > 50, rep
            for f in FIELDS:
                assert np.array_equal(got[f], exp[f]), (rep, f)
        # other arguments (min_support): a different result, and back
        p2 = g.cluster_params(window=480, min_support=9, max_clip_dist=190)
        g.cluster_device(d_t.data_ptr(), len(b), p2, d_out.data_ptr(), cap, d_n.data_ptr(), st.cuda_stream)
        st.synchronize()
        exp2, _ = gpu.cluster(b, window=480, min_support=9, max_clip_dist=190)
        got2 = d_out[: int(d_n.item()) * 48].cpu().numpy().view(sb.BOUNDS_DTYPE)
        got2 = got2[got2["tid"] >= 0]
        assert len(got2) == len(exp2) and all(np.array_equal(got2[f], exp2[f]) for f in FIELDS)
    finally:
        g.close()
