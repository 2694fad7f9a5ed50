"""Parity of the CUDA repeat-unit scan (K1, through the C ABI) against the CPU oracle: bit-exact unit and
repeat_count for every segment.  Needs a B200: run with -m gpu."""
import numpy as np
import pytest

import strling_b200 as sb
from oracle import oracle as orc
from strling_b200 import synth

pytestmark = pytest.mark.gpu

P = [0.8, 0.8 - 0.07, 0.6]


@pytest.fixture(scope="module")
def gpu():
    g = sb.StrGpu(0)
    g.set_proportions(P)
    yield g
    g.close()


def oracle_for(reads_ascii_flat, off, lens, pclass):
    p = np.asarray(P)[pclass]
    return orc.get_repeat_batch(reads_ascii_flat, off, lens, p)


def check(gpu, reads, segs, stride, seq2, nmask):
    res = gpu.scan(seq2, reads.shape[0] * stride, nmask, segs)
    flat, off, lens = synth.segment_ascii(reads, segs, stride)
    units, counts = oracle_for(flat, off, lens, segs["pclass"])
    bad = np.nonzero((res["unit"] != units) | (res["repeat_count"] != counts))[0]
    if len(bad):
        i = int(bad[0])
        s = bytes(flat[int(off[i]): int(off[i]) + int(lens[i])])
        raise AssertionError(f"{len(bad)} mismatches; first seg {i} len {lens[i]} pclass {segs['pclass'][i]} "
                             f"gpu=({res['unit'][i]},{res['repeat_count'][i]}) oracle=({units[i]},{counts[i]}) read={s}")
    return res


def test_reference_known_answers(gpu):
    # tests/test_strling.nim:46-89 through the GPU path
    assert gpu.get_repeat(["A" * 150], 0.6) == [(b"A", 150)]
    assert gpu.get_repeat(["TGC" * 50 + "T"], 0.8) == [(b"CTG", 49)]
    assert gpu.get_repeat(["CAG" * 50, "ATTCT" * 30, "AC" * 75, "ACGATC" * 16 + "ACGA"], 0.8) == [
        (b"CAG", 50), (b"CTATT", 29), (b"CA", 74), (b"CACGAT", 15)]
    gpu.set_proportions(P)


def test_config2_mix_150bp(gpu):
    reads, cls, lclip, rclip = synth.make_reads(200_000, seed=2)
    seq2, nmask, stride = synth.pack_matrix(reads)
    segs, _ = synth.segments_for(reads, lclip, rclip, stride)
    res = check(gpu, reads, segs, stride, seq2, nmask)
    assert (res["repeat_count"] > 0).sum() > 1000  # the STR classes are actually found


def test_reads_with_N(gpu):
    reads, cls, lclip, rclip = synth.make_reads(50_000, seed=11, n_frac=0.3, mix=(0.5, 0.1, 0.2, 0.2))
    seq2, nmask, stride = synth.pack_matrix(reads)
    segs, _ = synth.segments_for(reads, lclip, rclip, stride)
    assert nmask is not None
    check(gpu, reads, segs, stride, seq2, nmask)


def test_iupac_codes_stay_out_of_the_N_gate(gpu):
    # utils.nim:238 counts the literal 'N' only: a read with > 20 non-ACGT bases of which <= 20 are 'N' is still scanned by
    # the reference (IUPAC codes scan as 'A' and never match in read.count), one with > 20 'N' is not
    reads, cls, lclip, rclip = synth.make_reads(40_000, seed=12, n_frac=0.3, iupac_frac=0.4, mix=(0.3, 0.1, 0.2, 0.4))
    seq2, masks, stride = synth.pack_matrix(reads)
    assert isinstance(masks, sb.Masks)
    segs, _ = synth.segments_for(reads, lclip, rclip, stride)
    res = check(gpu, reads, segs, stride, seq2, masks)
    # the same reads through the uniform-read entry point
    out = np.zeros(len(segs), dtype=sb.REPEAT_DTYPE)
    seq2_4, masks_4, stride_4 = synth.pack_matrix(reads, align_bases=4)
    segs_4, _ = synth.segments_for(reads, lclip, rclip, stride_4)
    n = reads.shape[0]
    extra = np.ascontiguousarray(segs_4[n:])
    t = gpu.scan_reads_submit(seq2_4, n, 150, stride_4, 0, masks_4, extra, int(extra["len"].max()), out)
    gpu.scan_wait(t)
    assert np.array_equal(out, res)
    # crafted: a block of non-ACGT bases (some N, some IUPAC) in front of a clean repeat -- the gate decides the result
    rng = np.random.default_rng(3)
    crafted, expect_unit = [], []
    for unit in ("CAG", "AC", "AAAG", "T", "ACGATC", "TTTCA"):
        for n_n in (0, 5, 19, 20, 21, 25):
            for n_x in (0, 1, 2, 10, 16):
                if n_n + n_x > 28:
                    continue
                block = list("N" * n_n + "".join(rng.choice(list("RYMKSWHBVD"), size=n_x)))
                rng.shuffle(block)
                body = (unit * 160)[: 150 - len(block)]
                crafted.append("".join(block) + body)
                expect_unit.append(n_n <= 20)
    assert sum(("N" in r and r.count("N") <= 20 and sum(c not in "ACGT" for c in r) > 20) for r in crafted) >= 12
    seq2c, masks_c, segs_c, n_bases = sb.pack_reads(crafted, 0)
    got = gpu.scan(seq2c, n_bases, masks_c, segs_c)
    for r, g, want in zip(crafted, got, expect_unit):
        exp = orc.get_repeat(r, P[0])
        assert (bytes(g["unit"]).rstrip(b"\0"), int(g["repeat_count"])) == exp, (r, g, exp)
        assert want or exp[1] == 0, (r, exp)       # more than 20 literal N: gated away
    assert sum(int(g["repeat_count"]) > 0 for g, want in zip(got, expect_unit) if want) > 0.9 * sum(expect_unit)
    # xmask == NULL means "every flagged base is N" (include/strgpu.h): then those reads are gated away, unlike the reference
    got1 = gpu.scan(seq2c, n_bases, masks_c[0], segs_c)
    differ = [r for r, g, g1 in zip(crafted, got, got1) if g["repeat_count"] != g1["repeat_count"]]
    assert len(differ) >= 12 and all(r.count("N") <= 20 < sum(c not in "ACGT" for c in r) for r in differ)


def test_all_p_classes_and_unaligned_segments(gpu):
    reads, cls, lclip, rclip = synth.make_reads(30_000, seed=4, mix=(0.2, 0.1, 0.3, 0.4), noise=0.03)
    seq2, nmask, stride = synth.pack_matrix(reads)
    rng = np.random.default_rng(9)
    n = reads.shape[0]
    segs = np.zeros(4 * n, dtype=sb.SEGMENT_DTYPE)
    start = rng.integers(0, 150, size=4 * n)
    ln = rng.integers(0, 151, size=4 * n)
    ln = np.minimum(ln, 150 - start)
    segs["base_off"] = np.repeat(np.arange(n), 4) * stride + start
    segs["len"] = ln
    segs["pclass"] = rng.integers(0, 3, size=4 * n)
    check(gpu, reads, segs, stride, seq2, nmask)


@pytest.mark.parametrize("length", [161, 250, 301, 510])
def test_long_segments(gpu, length):
    reads, cls, lclip, rclip = synth.make_reads(20_000, seed=length, length=length, mix=(0.4, 0.1, 0.2, 0.3),
                                                noise=0.02, n_frac=0.05)
    seq2, nmask, stride = synth.pack_matrix(reads)
    segs, _ = synth.segments_for(reads, lclip, rclip, stride)
    check(gpu, reads, segs, stride, seq2, nmask)


def test_adversarial_short_and_periodic(gpu):
    rng = np.random.default_rng(77)
    reads = []
    for L in list(range(0, 40)) + [64, 65, 66, 96, 97, 98, 99, 128, 131, 132, 133, 159, 160]:
        reads.append("".join(rng.choice(list("ACGT"), size=L)))
        for unit in ("A", "C", "G", "T", "AC", "CA", "GT", "AAC", "CAG", "GGC", "AAAG", "ACGT", "TTTTA", "AACCC",
                     "AAAAAT", "ACGATC", "CCCCCG", "ACACAT"):
            s = (unit * (L // len(unit) + 2))
            for ph in range(min(len(unit), 3)):
                reads.append(s[ph: ph + L])
    # two units competing for the leader (tie-break = first to reach the maximum)
    for a, b in (("AC", "GT"), ("CAG", "TTG"), ("AAAG", "CCCT")):
        for na in range(1, 20, 3):
            reads.append((a * na + b * na)[:160])
            reads.append((b * na + a * na)[:160])
            reads.append(((a + b) * na)[:160])
    reads = [r for r in reads if len(r) <= 160]
    for pcls, p in enumerate(P):
        gpu.set_proportions(P)
        seq2, nmask, segs, n_bases = sb.pack_reads(reads, pcls)
        res = gpu.scan(seq2, n_bases, nmask, segs)
        for r, got in zip(reads, res):
            exp = orc.get_repeat(r, p)
            assert (bytes(got["unit"]).rstrip(b"\0"), int(got["repeat_count"])) == exp, (r, p, got, exp)


def test_too_long_segment_is_reported(gpu):
    seq2, nmask, segs, n_bases = sb.pack_reads(["A" * 600])
    with pytest.raises(sb.StrGpuError) as e:  # honest max_len: rejected on the host before anything is copied
        gpu.scan(seq2, n_bases, nmask, segs)
    assert e.value.status == -4
    with pytest.raises(sb.StrGpuError) as e:  # caller lied about max_len: the kernel reports it
        gpu.scan(seq2, n_bases, nmask, segs, max_len=150)
    assert e.value.status == -4
    # a 200-base segment in a batch announced as <= 160 is still scanned correctly (handed to the long-segment path)
    seq2, nmask, segs, n_bases = sb.pack_reads(["A" * 200, "CAG" * 50])
    res = gpu.scan(seq2, n_bases, nmask, segs, max_len=160)
    assert [(bytes(r["unit"]).rstrip(b"\0"), int(r["repeat_count"])) for r in res] == [(b"A", 200), (b"CAG", 50)]


def test_prefilter_threshold_boundary(gpu):
    # The lane kernel finishes a segment early when its most frequent 2-mer occurs <= min_k int(L * p / k) times
    # (exact: no k-mer can then beat its threshold, utils.nim:254-259).  Reads built to sit on both sides of that bound:
    # c scattered copies of a k-mer with c around int(L * p / k), the rest random filler.
    rng = np.random.default_rng(123)
    cases = {0: [], 1: [], 2: []}
    for pcls, p in enumerate(P):
        for L in (150, 151, 160, 149, 120, 100, 75, 60, 48, 33, 17):
            for k in range(2, 7):
                t = int(float(L) * p / float(k))
                for c in (t - 1, t, t + 1, t + 2):
                    for rep in range(6):
                        if c < 0 or c * k > L:
                            continue
                        unit = "".join(rng.choice(list("ACGT"), size=k))
                        gaps = L - c * k
                        cuts = np.sort(rng.integers(0, gaps + 1, size=c)) if c else np.zeros(0, dtype=int)
                        other = [x for x in "ACGT" if x not in unit]
                        if rep >= 3 and len(other) >= 2:
                            # filler that adds nothing to the unit's 2-mer counts: the bound is met with equality at c = t + 1
                            filler = (other[0] + other[1]) * (gaps // 2 + 1)
                            filler = filler[:gaps]
                        else:
                            filler = "".join(rng.choice(list("ACGT"), size=gaps))
                        parts, prev = [], 0
                        for cut in cuts:
                            parts.append(filler[prev:cut])
                            parts.append(unit)
                            prev = cut
                        parts.append(filler[prev:])
                        r = "".join(parts)
                        assert len(r) == L
                        cases[pcls].append(r)
    n_found = 0
    for pcls, p in enumerate(P):
        reads = cases[pcls]
        seq2, nmask, segs, n_bases = sb.pack_reads(reads, pcls)
        res = gpu.scan(seq2, n_bases, nmask, segs)
        for r, got in zip(reads, res):
            exp = orc.get_repeat(r, p)
            assert (bytes(got["unit"]).rstrip(b"\0"), int(got["repeat_count"])) == exp, (r, p, got, exp)
            n_found += exp[1] > 0
    assert n_found > 500  # both sides of the bound are exercised


@pytest.mark.parametrize("variant", ["0", "1", "5", "7", "8"])
def test_both_kernel_variants_agree_with_oracle(variant, monkeypatch):
    # 0: pre-filter kernel + one ladder kernel per rung over dense survivor lists, 1: warp-per-segment kernel,
    # 5 / 7: 0 / 12 carry-save popcount streams in the pre-filter, 8: stage lists of 64 entries (the overflow path: segments
    # that find their list full are redone from rung 2 by the warp kernel)
    monkeypatch.setenv("STRGPU_SCAN_VARIANT", variant)
    g = sb.StrGpu(0)
    try:
        g.set_proportions(P)
        reads, cls, lclip, rclip = synth.make_reads(100_000, seed=31, mix=(0.6, 0.1, 0.15, 0.15), noise=0.02, n_frac=0.01)
        seq2, nmask, stride = synth.pack_matrix(reads)
        segs, _ = synth.segments_for(reads, lclip, rclip, stride)
        check(g, reads, segs, stride, seq2, nmask)
    finally:
        g.close()


def test_submit_wait_pipeline(gpu):
    reads, cls, lclip, rclip = synth.make_reads(60_000, seed=21)
    seq2, nmask, stride = synth.pack_matrix(reads)
    segs, _ = synth.segments_for(reads, lclip, rclip, stride)
    outs = [np.zeros(len(segs), dtype=sb.REPEAT_DTYPE) for _ in range(3)]
    tickets = [gpu.scan_submit(seq2, reads.shape[0] * stride, nmask, segs, 150, o) for o in outs]
    with pytest.raises(sb.StrGpuError) as e:  # all slots busy
        gpu.scan_submit(seq2, reads.shape[0] * stride, nmask, segs, 150, outs[0])
    assert e.value.status == -5
    for t in tickets:
        gpu.scan_wait(t)
    assert np.array_equal(outs[0], outs[1]) and np.array_equal(outs[0], outs[2])
    flat, off, lens = synth.segment_ascii(reads, segs, stride)
    units, counts = oracle_for(flat, off, lens, segs["pclass"])
    assert np.array_equal(outs[0]["unit"], units) and np.array_equal(outs[0]["repeat_count"], counts)


@pytest.mark.parametrize("length,n_frac", [(150, 0.0), (150, 0.05), (151, 0.02), (100, 0.0), (37, 0.1), (250, 0.02)])
def test_uniform_reads_without_descriptors(gpu, length, n_frac):
    # strgpu_scan_reads_submit: implicit whole-read segments (+ explicit clip segments) == the descriptor path == oracle
    reads, cls, lclip, rclip = synth.make_reads(40_000, seed=length, length=length, mix=(0.7, 0.1, 0.1, 0.1), n_frac=n_frac)
    lclip = np.minimum(lclip, length)
    rclip = np.minimum(rclip, length)
    seq2, nmask, stride = synth.pack_matrix(reads, align_bases=4)
    assert stride % 4 == 0 and stride - length < 4
    segs, _ = synth.segments_for(reads, lclip, rclip, stride)
    n = reads.shape[0]
    extra = np.ascontiguousarray(segs[n:])
    out = np.zeros(len(segs), dtype=sb.REPEAT_DTYPE)
    t = gpu.scan_reads_submit(seq2, n, length, stride, 0, nmask, extra if len(extra) else None, int(extra["len"].max()) if len(extra) else 0, out)
    gpu.scan_wait(t)
    flat, off, lens = synth.segment_ascii(reads, segs, stride)
    units, counts = oracle_for(flat, off, lens, segs["pclass"])
    assert np.array_equal(out["unit"], units) and np.array_equal(out["repeat_count"], counts)


def test_scale_permutation_invariance(gpu):
    # Size-independent property at scale: a segment's result depends only on its bases, length and class -- not on where it
    # sits in the batch, which lane / warp / queue handles it, or what its neighbours are.  8M segments = one 250k-read
    # shard referenced 32 times in shuffled order (every copy lands in a different group of 32 and queue slot).
    reads, cls, lclip, rclip = synth.make_reads(250_000, seed=77, mix=(0.8, 0.1, 0.05, 0.05), n_frac=0.002)
    seq2, nmask, stride = synth.pack_matrix(reads)
    segs, _ = synth.segments_for(reads, lclip, rclip, stride)
    flat, off, lens = synth.segment_ascii(reads, segs, stride)
    units, counts = oracle_for(flat, off, lens, segs["pclass"])
    rng = np.random.default_rng(5)
    reps = 32
    perm = np.concatenate([rng.permutation(len(segs)) for _ in range(reps)])
    big = np.ascontiguousarray(segs[perm])
    res = gpu.scan(seq2, reads.shape[0] * stride, nmask, big, max_len=150)
    assert np.array_equal(res["unit"], units[perm]) and np.array_equal(res["repeat_count"], counts[perm])
    # checksum of checksums: every copy of the shard folds to the same digest
    key = res["repeat_count"].astype(np.uint64) * np.uint64(1315423911) + np.frombuffer(res["unit"].tobytes(), dtype=np.uint8).reshape(-1, 6).astype(np.uint64).dot(np.array([1, 7, 49, 343, 2401, 16807], dtype=np.uint64))
    inv = np.argsort(perm.reshape(reps, -1), axis=1)
    digests = {int(np.bitwise_xor.reduce(key.reshape(reps, -1)[r][inv[r]] * (np.arange(len(segs), dtype=np.uint64) + np.uint64(1)))) for r in range(reps)}
    assert len(digests) == 1


@pytest.mark.parametrize("length,align,n_frac", [(150, 16, 0.0), (150, 4, 0.0), (150, 4, 0.02), (101, 4, 0.0), (160, 16, 0.01), (33, 4, 0.0)])
def test_device_resident_uniform_reads(gpu, length, align, n_frac):
    # strgpu_scan_reads_device (uniform reads staged through shared memory by TMA bulk copies + clip descriptors) ==
    # strgpu_scan_device (descriptors, per-lane loads) == oracle, on caller-owned device buffers and stream
    import torch

    reads, cls, lclip, rclip = synth.make_reads(70_001, seed=1000 + length, length=length, mix=(0.75, 0.1, 0.08, 0.07), n_frac=n_frac)
    lclip = np.minimum(lclip, length)
    rclip = np.minimum(rclip, length)
    seq2, nmask, stride = synth.pack_matrix(reads, align_bases=align)
    segs, _ = synth.segments_for(reads, lclip, rclip, stride)
    n = reads.shape[0]
    flat, off, lens = synth.segment_ascii(reads, segs, stride)
    units, counts = oracle_for(flat, off, lens, segs["pclass"])
    dev = torch.device("cuda", 0)
    pad = (-len(seq2)) % 16 + 16
    d_seq = torch.from_numpy(np.concatenate([seq2, np.zeros(pad, dtype=np.uint8)])).to(dev)
    d_nm = torch.from_numpy(nmask.view(np.uint8).copy()).to(dev) if nmask is not None else None
    d_segs = torch.from_numpy(segs.view(np.uint8).reshape(-1).copy()).to(dev)
    stream = torch.cuda.Stream(device=dev)
    results = []
    for which in ("reads", "segments"):
        d_out = torch.full((len(segs) * 8,), 0xEE, dtype=torch.uint8, device=dev)
        stream.wait_stream(torch.cuda.current_stream(dev))
        if which == "reads":
            n_extra = len(segs) - n
            gpu.scan_reads_device(d_seq.data_ptr(), n, length, stride, 0, d_nm.data_ptr() if d_nm is not None else None,
                                  d_segs.data_ptr() + n * 8 if n_extra else None, n_extra,
                                  int(segs["len"][n:].max()) if n_extra else 0, d_out.data_ptr(), stream.cuda_stream)
        else:
            gpu.scan_device(d_seq.data_ptr(), d_nm.data_ptr() if d_nm is not None else None, d_segs.data_ptr(), len(segs), length,
                            d_out.data_ptr(), stream.cuda_stream)
        gpu.device_status(stream.cuda_stream)
        results.append(d_out.cpu().numpy().view(sb.REPEAT_DTYPE))
    for res in results:
        assert np.array_equal(res["unit"], units) and np.array_equal(res["repeat_count"], counts)
