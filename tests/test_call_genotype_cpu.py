"""`strling call` downstream of the cluster kernels (collect.nim, spanning.nim, genotyper.nim,
call.nim:223-281).  CPU only: (1) the oracle restatement against the reference's own known-answer tests, (2) the C++ host
implementation (one streaming pass over the BAM for all loci) against the oracle on a synthetic BAM, through
`strling debug genotype`, which takes the cluster records from a file instead of the GPU."""
import os
import subprocess

import numpy as np
import pytest

from oracle import call_oracle as co
from oracle import extract_oracle as eo
from oracle import oracle as orc
from strling_b200 import bamio
from strling_b200 import build as sb_build
from strling_b200.bamio import Aln


@pytest.fixture(scope="module")
def cli():
    return sb_build.build_cli()


def _span(rc, ins, dele):
    return dict(type=co.SPANNING_READ, frag_len=0, frag_pct=0.0, rc=rc, ins=ins, dele=dele)


def test_reference_genotyper_vector():
    # tests/test_genotyper.nim:8-23
    a1_bp, a2_bp, a1_ru, a2_ru, n = co.spanning_read_est([_span(10, 0, 0), _span(10, 0, 0), _span(10, 0, 0), _span(9, 0, 2)])
    assert (a1_bp, a2_bp, a1_ru, a2_ru, n) == (0.0, -2.0, 10.0, 9.0, 4)


def test_reference_collect_vectors():
    # tests/test_collect.nim:8-48 (SAM positions are 1-based: POS 1 == start 0)
    a = Aln("read1", 0, 0, 0, 40, [("M", 25), ("S", 5)], -1, -1, 0, "A" * 30)
    assert co.overlapping_read(a, 0, 50, 100, "A") is None
    assert co.overlapping_read(a, 0, 5, 15, "AAAAAA")["type"] == co.OVERLAPPING_READ
    assert co.overlapping_read(a, 0, 6, 15, "AAAAAA")["type"] == co.SPANNING_READ
    assert co.overlapping_read(a, 0, 9, 10, "AAAAAA") is not None
    assert co.overlapping_read(a, 0, 10, 11, "AAAAAA")["type"] == co.SPANNING_READ
    # tests/test_collect.nim:50-76
    L = Aln("read1", 99, 0, 0, 40, [("M", 15), ("S", 5)], 0, 499, 0, "A" * 20)
    R = Aln("read1", 147, 0, 499, 40, [("M", 15), ("S", 5)], 0, 0, 0, "A" * 20)
    fs = np.zeros(4096, dtype=np.uint32)
    assert co.spanning_fragment(L, R, 100, 150, "A", fs) is not None
    assert co.spanning_fragment(L, R, 450, 513, "A", fs) is not None
    assert co.spanning_fragment(L, R, 512, 513, "A", fs) is None


def test_reference_median_depth_vectors():
    # tests/test_utils.nim:10-12
    assert co.median_depth([1, 2, 2]) == 2 and co.median_depth([2000]) == 1047


def test_formatting():
    assert co.fmt2(float("nan")) == "nan" and co.fmt2(1.005) == "%.2f" % 1.005 and co.nim_float(35.0) == "35.0"
    c = dict(chrom="chr4", start=3074876, stop=3074933, repeat="CAG", allele1=float("nan"), allele2=41.237, anchored_reads=7, spanning_reads=0,
             spanning_pairs=2, expected_spanning_fragments=np.float32(3.456), oe_pct=np.float32(0.5), left_clips=3, right_clips=4,
             unplaced_reads=0, depth=31.0, sum_str_counts=512)
    assert co.call_line(c) == "chr4\t3074876\t3074933\tCAG\tnan\t41.24\t7\t0\t2\t3.46\t0.50\t3\t4\t0\t31.0\t512"


def _spanning_extras(targets, loci, seed):
    """Hand-made proper pairs whose first mate spans a locus with insertions / deletions of tied frequencies (the CountTable
    tie-break of genotyper.nim:86-92) and whose fragments span it (collect.nim:35-48)."""
    rng = np.random.default_rng(seed)
    out = []
    for li, (tid, start, stop, unit) in enumerate(loci):
        plan = [[(3, 0)] * 12, [(2, 0)] * 9 + [(0, 3)] * 9, [(0, 1)] * 11 + [(4, 0)] * 2, [(0, 0)] * 5][li % 4]
        for j, (ins, dele) in enumerate(plan):
            width = stop - start
            pos = start - 40 - (j % 20)
            m1 = 40 + (j % 20)
            cig = [("M", m1)]
            if ins:
                cig += [("I", ins)]
            if dele:
                cig += [("D", dele)]
            rest = 150 - m1 - ins
            cig += [("M", rest)]
            seq = bamio._rand_seq(rng, m1) + bamio._unit_seq(unit, min(width, rest), 0) + bamio._rand_seq(rng, 150)
            seq = seq[:150]
            mpos = pos + 380 + 3 * j
            isz = mpos + 150 - pos
            out.append(Aln(f"sp{li}_{j}", 99, tid, pos, 60, cig, tid, mpos, isz, seq))
            out.append(Aln(f"sp{li}_{j}", 147, tid, mpos, 60, [("M", 150)], tid, pos, -isz, bamio._rand_seq(rng, 150)))
    return out


@pytest.mark.parametrize("seed,min_support,min_mapq", [(11, 3, 40), (12, 5, 20), (13, 2, 0)])
def test_host_genotyping_matches_oracle(cli, tmp_path, seed, min_support, min_mapq):
    targets = [("chr1", 300_000), ("chr2", 200_000)]
    loci = [(0, 50_000, 50_060, "CAG"), (0, 120_000, 120_040, "AAAG"), (1, 80_000, 80_090, "AC"), (1, 150_000, 150_030, "CCG")]
    recs = bamio.simulate_alignments(seed, 9000, targets, loci, str_pair_frac=0.25, unmapped_pairs=60)
    recs = [a for a in recs if a.tid >= 0] + _spanning_extras(targets, loci, seed) + [a for a in recs if a.tid < 0]
    placed = sorted([a for a in recs if a.tid >= 0], key=lambda a: (a.tid, a.pos))
    recs = placed + [a for a in recs if a.tid < 0]
    hdr = bamio.sam_header(targets)
    bam, binp, cl = str(tmp_path / "a.bam"), str(tmp_path / "a.bin"), str(tmp_path / "cl.tsv")
    bamio.write_bam(bam, hdr, targets, recs)
    data, _, _ = eo.extract(recs, targets, hdr)
    open(binp, "wb").write(data)
    u = eo.unpack_bin(data)
    frag = eo.fragment_length_distribution(recs)
    window, med = orc.median(frag, 0.99), orc.median(frag, 0.5)
    b, unplaced = orc.cluster_all(u["treads"], window, min_support, 0, 0, int(0.5 * float(med)) & 0xFFFF, merge_mode=False)
    with open(cl, "w") as fh:
        for unit, cnt in sorted(unplaced.items()):
            fh.write(f"-1 0 0 {unit.decode()} 0 0 0 0 0 0 0 {cnt}\n")
        for x in b:
            rep = bytes(x["repeat"]).rstrip(b"\0").decode()
            fh.write(f"{x['tid']} {x['left']} {x['right']} {rep} {x['left_most']} {x['right_most']} {x['center_mass']} {x['n_left']} "
                     f"{x['n_right']} {x['n_total']} {x['first_read']} {x['n_reads']}\n")
    prefix = str(tmp_path / "out")
    r = subprocess.run([cli, "debug", "genotype", bam, binp, cl, prefix, str(window), str(min_support), str(min_mapq)], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    exp_gt, exp_bounds, exp_unplaced, _ = co.call(recs, data, min_support=min_support, min_mapq=min_mapq)
    got_gt = open(prefix + "-genotype.txt").read().splitlines()
    got_bounds = open(prefix + "-bounds.txt").read().splitlines()
    assert got_gt[0] == co.GT_HEADER and got_bounds[0] == eo.BOUNDS_HEADER + "\tdepth"
    assert len(exp_gt) >= 8 and sorted(got_gt[1:]) == sorted(exp_gt)
    assert got_bounds[1:] == exp_bounds
    # the evidence pass is chunked and parallel over loci (genotype.hpp collect_evidence): the same files with chunks of two BGZF
    # blocks (windows straddle many chunk boundaries), on one thread's worth of parts, and through the record-by-record logic
    for k, env in enumerate(({"STRLING_CALL_BLOCKS": "2"}, {"STRLING_CALL_SERIAL": "1"}, {"STRLING_CALL_BLOCKS": "1", "STRLING_CALL_SERIAL": "1"})):
        alt = str(tmp_path / f"alt{k}")
        r = subprocess.run([cli, "debug", "genotype", bam, binp, cl, alt, str(window), str(min_support), str(min_mapq)], capture_output=True, text=True,
                           env=dict(os.environ, **env))
        assert r.returncode == 0, r.stderr
        for suffix in ("-genotype.txt", "-bounds.txt", "-unplaced.txt"):
            assert open(alt + suffix).read() == open(prefix + suffix).read(), (env, suffix)
    # the evidence is not trivial: spanning reads, spanning pairs and non-zero indel alleles occur
    cols = [l.split("\t") for l in exp_gt]
    assert any(int(c[7]) > 0 for c in cols) and any(int(c[8]) > 0 for c in cols) and any(c[4] not in ("0.00", "nan") for c in cols)
    got_un = dict(l.split("\t") for l in open(prefix + "-unplaced.txt").read().splitlines())
    assert {k.encode(): int(v) for k, v in got_un.items()} == exp_unplaced


def test_host_genotyping_with_bounds_and_loci_files(cli, tmp_path):
    # call.nim:158-218: -b bounds and -l loci are merged (a locus overwrites the bound it overlaps), take their reads first
    # (assign_reads_locus incl. its dropped read), are genotyped and reported first; the rest is clustered as usual
    targets = [("chr1", 300_000), ("chr2", 200_000)]
    loci = [(0, 50_000, 50_060, "CAG"), (0, 120_000, 120_040, "AAAG"), (1, 80_000, 80_090, "AC"), (1, 150_000, 150_030, "CCG")]
    recs = bamio.simulate_alignments(31, 9000, targets, loci, str_pair_frac=0.25, unmapped_pairs=40)
    recs = [a for a in recs if a.tid >= 0] + _spanning_extras(targets, loci, 31) + [a for a in recs if a.tid < 0]
    recs = sorted([a for a in recs if a.tid >= 0], key=lambda a: (a.tid, a.pos)) + [a for a in recs if a.tid < 0]
    hdr = bamio.sam_header(targets)
    bam, binp, cl = str(tmp_path / "a.bam"), str(tmp_path / "a.bin"), str(tmp_path / "cl.tsv")
    bamio.write_bam(bam, hdr, targets, recs)
    data, _, _ = eo.extract(recs, targets, hdr)
    open(binp, "wb").write(data)
    frag = eo.fragment_length_distribution(recs)
    window = orc.median(frag, 0.99)
    # a bounds file: two discovered bounds of a first pass (depth column stripped), one of them overlapped by a bed locus,
    # plus a bound wider than 1000 bp (takes reads but is not genotyped)
    _, first_bounds, _, _ = co.call(recs, data, min_support=3)
    cag = [l for l in first_bounds if l.split("\t")[3] == "CAG"][:1]
    other = [l for l in first_bounds if l.split("\t")[3] == "CA"][:1]
    bounds_lines = ["\t".join(l.split("\t")[:11]) for l in cag + other]
    bounds_lines.append("chr2\t149000\t150900\tCCG\twide\t148500\t151400\t150000\t0\t0\t0")
    bed_lines = ["chr1\t50000\t50060\tCAG\tHTT_like", "chr1 120000 120040 AAAG", "chr2\t10\t20\tGGGGGC\tempty_bucket"]
    bpath, lpath = str(tmp_path / "in-bounds.txt"), str(tmp_path / "loci.bed")
    open(bpath, "w").write("#header\n" + "\n".join(bounds_lines) + "\n")
    open(lpath, "w").write("\n".join(bed_lines) + "\n")
    exp_gt, exp_bounds, exp_unplaced, (sub, b) = co.call(recs, data, min_support=3, bounds_lines=bounds_lines, bed_lines=bed_lines)
    with open(cl, "w") as fh:
        for unit, cnt in sorted(exp_unplaced.items()):
            fh.write(f"-1 0 0 {unit.decode()} 0 0 0 0 0 0 0 {cnt}\n")
        for x in b:
            rep = bytes(x["repeat"]).rstrip(b"\0").decode()
            fh.write(f"{x['tid']} {x['left']} {x['right']} {rep} {x['left_most']} {x['right_most']} {x['center_mass']} {x['n_left']} "
                     f"{x['n_right']} {x['n_total']} {x['first_read']} {x['n_reads']}\n")
    prefix = str(tmp_path / "out")
    r = subprocess.run([cli, "debug", "genotype", bam, binp, cl, prefix, str(window), "3", "40", bpath, lpath], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    got_gt = open(prefix + "-genotype.txt").read().splitlines()
    got_bounds = open(prefix + "-bounds.txt").read().splitlines()
    assert sorted(got_gt[1:]) == sorted(exp_gt) and got_bounds[1:] == exp_bounds
    # the merged locus carries the bed name and interval, the wide bound is not reported, the listed loci come first
    assert got_bounds[1].split("\t")[:5] == ["chr1", "50000", "50060", "CAG", "HTT_like"]
    assert not any("wide" in l for l in got_bounds) and "large bounds" in r.stderr
    assert int(got_bounds[1].split("\t")[10]) > 20 and any("empty_bucket" in l for l in got_bounds)
    # the oracle's own device-style assignment (orc.cluster_all_loci) agrees on the per-locus read counts
    u = eo.unpack_bin(data)
    lb = co.merge_loci_into_bounds(co.parse_bounds_lines(bounds_lines, targets), co.parse_bed_lines(bed_lines, targets, window))
    arr = np.zeros(len(lb), dtype=orc.LOCUS_DTYPE)
    for i, x in enumerate(lb):
        arr[i]["tid"], arr[i]["left_most"], arr[i]["right_most"], arr[i]["repeat"] = x["tid"], x["left_most"], x["right_most"], x["repeat"].encode()
    med = orc.median(frag, 0.5)
    arr2, b2, _ = orc.cluster_all_loci(u["treads"], arr, window, 3, 0, 0, int(0.5 * float(med)) & 0xFFFF, merge_mode=False)
    buckets = {}
    for i in co.sorted_tread_order(u["treads"]):
        buckets.setdefault((int(u["treads"]["tid"][i]), bytes(u["treads"]["repeat"][i]).rstrip(b"\0")), []).append(int(i))
    for i, x in enumerate(lb):
        co.assign_reads_locus(x, buckets, u["treads"])
        assert (x["n_left"], x["n_right"], x["n_total"]) == (int(arr2[i]["n_left"]), int(arr2[i]["n_right"]), int(arr2[i]["n_total"]))
    assert len(b2) == len(b)


def test_high_depth_loci_are_skipped_like_the_reference(cli, tmp_path):
    # call.nim:236-241 / collect.nim:166-169: a locus with more than 5000 supporting records, or more than 20000 read names in
    # its window, is dropped from -bounds.txt and -genotype.txt (its reads stay consumed); other loci are unaffected
    rng = np.random.default_rng(5)
    targets = [("chr1", 400_000)]
    loci = [(0, 50_000, 50_060, "CAG"), (0, 150_000, 150_040, "AAAG"), (0, 300_000, 300_050, "AC")]
    recs = bamio.simulate_alignments(41, 4000, targets, loci, str_pair_frac=0.3, unmapped_pairs=10)
    extra = []
    seq = bamio._rand_seq(rng, 150)
    for i in range(5600):      # > 5000 reads overlapping the bounds of locus 1, few read names beyond that
        pos = 149_930 + (i % 40)
        extra.append(Aln(f"deep{i}", 0, 0, pos, 60, [("M", 150)], -1, -1, 0, seq))
    for i in range(23_000):    # > 20000 read names inside the window of locus 2, none of them overlapping its bounds
        pos = 299_650 + (i % 200)
        extra.append(Aln(f"wide{i}", 99, 0, pos, 60, [("M", 150)], 0, pos + 200, 350, seq))
    placed = sorted([a for a in recs if a.tid >= 0] + extra, key=lambda a: (a.tid, a.pos))
    recs = placed + [a for a in recs if a.tid < 0]
    hdr = bamio.sam_header(targets)
    bam, binp, cl = str(tmp_path / "a.bam"), str(tmp_path / "a.bin"), str(tmp_path / "cl.tsv")
    bamio.write_bam(bam, hdr, targets, recs)
    data, _, _ = eo.extract(recs, targets, hdr)
    open(binp, "wb").write(data)
    u = eo.unpack_bin(data)
    frag = eo.fragment_length_distribution(recs)
    window, med = orc.median(frag, 0.99), orc.median(frag, 0.5)
    b, unplaced = orc.cluster_all(u["treads"], window, 3, 0, 0, int(0.5 * float(med)) & 0xFFFF, merge_mode=False)
    with open(cl, "w") as fh:
        for unit, cnt in sorted(unplaced.items()):
            fh.write(f"-1 0 0 {unit.decode()} 0 0 0 0 0 0 0 {cnt}\n")
        for x in b:
            rep = bytes(x["repeat"]).rstrip(b"\0").decode()
            fh.write(f"{x['tid']} {x['left']} {x['right']} {rep} {x['left_most']} {x['right_most']} {x['center_mass']} {x['n_left']} "
                     f"{x['n_right']} {x['n_total']} {x['first_read']} {x['n_reads']}\n")
    prefix = str(tmp_path / "out")
    r = subprocess.run([cli, "debug", "genotype", bam, binp, cl, prefix, str(window), "3", "40"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    exp_gt, exp_bounds, _, _ = co.call(recs, data, min_support=3)
    got_gt = open(prefix + "-genotype.txt").read().splitlines()[1:]
    got_bounds = open(prefix + "-bounds.txt").read().splitlines()[1:]
    assert sorted(got_gt) == sorted(exp_gt) and got_bounds == exp_bounds
    pos = [int(l.split("\t")[1]) for l in got_bounds]
    assert any(abs(p - 50_000) < 500 for p in pos)                                   # the ordinary locus is reported
    assert not any(abs(p - 150_000) < 500 for p in pos) and not any(abs(p - 300_000) < 500 for p in pos)   # the deep ones are not
    assert len(got_bounds) < len(b)


def test_unsorted_bam_is_refused(cli, tmp_path):
    # `call` gathers the evidence in one streaming pass that relies on coordinate order (the reference's indexed queries need a
    # sorted BAM as well): an unsorted file must fail loudly, not produce evidence with records missing
    targets = [("chr1", 300_000)]
    loci = [(0, 50_000, 50_060, "CAG")]
    recs = bamio.simulate_alignments(5, 800, targets, loci, str_pair_frac=0.3)
    placed = [a for a in recs if a.tid >= 0]
    placed[10], placed[400] = placed[400], placed[10]
    hdr = bamio.sam_header(targets)
    bam, binp, cl = str(tmp_path / "u.bam"), str(tmp_path / "u.bin"), str(tmp_path / "cl.tsv")
    bamio.write_bam(bam, hdr, targets, placed)
    data, _, _ = eo.extract(sorted(placed, key=lambda a: (a.tid, a.pos)), targets, hdr)
    open(binp, "wb").write(data)
    open(cl, "w").write("0 50000 50060 CAG 49500 50500 50030 3 3 12 0 12\n")
    r = subprocess.run([cli, "debug", "genotype", bam, binp, cl, str(tmp_path / "o"), "500", "3", "40"], capture_output=True, text=True)
    assert r.returncode != 0 and "not coordinate-sorted" in r.stderr
