"""End-to-end drop-in checks on a B200: the `strling extract | call | merge` command line (C++ host + CUDA kernels)
against the oracle pipeline on the same synthetic BAMs -- `.bin` byte-identical, bounds / unplaced lines identical.
Configs follow SURVEY.md 8d: config 1 (1k reads, one unit) and a config-4 stand-in built on the reference's
simulation loci (tests/golden/disease_loci.json).  Run with -m gpu."""
import json
import os
import subprocess

import numpy as np
import pytest

from oracle import call_oracle as co
from oracle import extract_oracle as eo
from strling_b200 import bamio
from strling_b200 import build as sb_build

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def cli():
    return sb_build.build_cli()


def run(cli, *args):
    r = subprocess.run([cli, *args], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return r


def disease_setup():
    loci_json = json.load(open(os.path.join(HERE, "golden", "disease_loci.json")))
    chroms = sorted({l["chrom"] for l in loci_json}, key=lambda c: (len(c), c))
    targets = [(f"chr{c}", 2_000_000) for c in chroms]
    loci = []
    for l in loci_json:
        start = 50_000 + l["start"] % 1_800_000
        width = max(20, min(200, l["stop"] - l["start"]))
        loci.append((chroms.index(l["chrom"]), start, start + width, l["unit"]))
    return targets, loci


def test_config1_single_unit_bin_identical(cli, tmp_path):
    targets = [("chr1", 1_000_000)]
    loci = [(0, 400_000, 400_150, "CAG")]
    recs = bamio.simulate_alignments(1, 500, targets, loci, str_pair_frac=0.5)
    assert len(recs) == 1000
    bam, out = str(tmp_path / "c1.bam"), str(tmp_path / "c1.bin")
    hdr = bamio.sam_header(targets)
    bamio.write_bam(bam, hdr, targets, recs)
    run(cli, "extract", "-v", bam, out)
    exp, cache, _ = eo.extract(recs, targets, hdr)
    assert len(cache) > 100
    assert open(out, "rb").read() == exp


@pytest.mark.parametrize("p,q,use_bed,batch", [(0.8, 40, False, 1048576), (0.8, 40, True, 4096), (0.7, 20, True, 1000), (0.9, 0, False, 777)])
def test_config4_extract_bin_identical(cli, tmp_path, p, q, use_bed, batch):
    targets, loci = disease_setup()
    recs = bamio.simulate_alignments(7, 12_000, targets, loci, unmapped_pairs=150, n_frac=0.02)
    hdr = bamio.sam_header(targets)
    bam, out = str(tmp_path / "c4.bam"), str(tmp_path / "c4.bin")
    bamio.write_bam(bam, hdr, targets, recs)
    args = ["extract", "-p", repr(p), "-q", str(q), "--batch-reads", str(batch)]
    genome_str = None
    if use_bed:
        bed = str(tmp_path / "ref.str")
        with open(bed, "w") as fh:
            for tid, s, e, u in loci:
                fh.write(f"{targets[tid][0]}\t{s}\t{e}\t{u}\n")
        args += ["-g", bed]
        genome_str = eo.read_bed(bed)
    run(cli, *args, bam, out)
    exp, cache, _ = eo.extract(recs, targets, hdr, p, q, genome_str)
    got = open(out, "rb").read()
    if got != exp:
        ug, ue = eo.unpack_bin(got), eo.unpack_bin(exp)
        assert len(ug["treads"]) == len(ue["treads"]), (len(ug["treads"]), len(ue["treads"]))
        bad = [i for i in range(len(ue["treads"])) if ug["treads"][i] != ue["treads"][i] or ug["qnames"][i] != ue["qnames"][i]]
        raise AssertionError(f"{len(bad)} records differ; first {bad[:1]}: {ug['treads'][bad[0]]} {ug['qnames'][bad[0]]} vs {ue['treads'][bad[0]]} {ue['qnames'][bad[0]]}")
    assert len(cache) > 1000


def test_call_and_merge_outputs_identical(cli, tmp_path):
    targets, loci = disease_setup()
    hdr = bamio.sam_header(targets)
    bins, datas = [], []
    for s in range(3):
        recs = bamio.simulate_alignments(20 + s, 10_000, targets, loci, unmapped_pairs=100)
        bam, out = str(tmp_path / f"s{s}.bam"), str(tmp_path / f"s{s}.bin")
        bamio.write_bam(bam, hdr, targets, recs)
        run(cli, "extract", bam, out)
        exp, _, _ = eo.extract(recs, targets, hdr)
        assert open(out, "rb").read() == exp
        bins.append(out)
        datas.append(exp)
        if s == 0:
            first_bam, first_recs = bam, recs
    # call: cluster loop of one sample on the GPU, then spanning evidence + genotypes on the host (call.nim:223-281)
    # -> -bounds.txt (with the depth column), -unplaced.txt, -genotype.txt
    prefix = str(tmp_path / "call")
    run(cli, "call", "-m", "3", "-o", prefix, first_bam, bins[0])
    exp_gt, exp_lines, exp_unplaced, _ = co.call(first_recs, datas[0], min_support=3)
    got = open(prefix + "-bounds.txt").read().splitlines()
    assert got[0] == eo.BOUNDS_HEADER + "\tdepth"
    assert len(exp_lines) > 10 and sorted(got[1:]) == sorted(exp_lines) and got[1:] == exp_lines
    got_un = dict(l.split("\t") for l in open(prefix + "-unplaced.txt").read().splitlines())
    assert {k.encode(): int(v) for k, v in got_un.items()} == exp_unplaced and len(exp_unplaced) > 0
    got_gt = open(prefix + "-genotype.txt").read().splitlines()
    assert got_gt[0] == co.GT_HEADER and len(exp_gt) == len(exp_lines) and sorted(got_gt[1:]) == sorted(exp_gt)
    # call -b -l: listed bounds / loci are merged, take their reads first, are genotyped and reported first (call.nim:158-218)
    bounds_in = ["\t".join(l.split("\t")[:11]) for l in exp_lines[:3]]
    t0, s0, e0, u0 = loci[0]
    bed_in = [f"{targets[t0][0]}\t{s0}\t{e0}\t{u0}\tfirst_locus", f"{targets[loci[1][0]][0]} {loci[1][1]} {loci[1][2]} {loci[1][3]}"]
    bpath, lpath = str(tmp_path / "in-bounds.txt"), str(tmp_path / "in-loci.bed")
    open(bpath, "w").write(eo.BOUNDS_HEADER + "\n" + "\n".join(bounds_in) + "\n")
    open(lpath, "w").write("\n".join(bed_in) + "\n")
    prefix = str(tmp_path / "call_loci")
    run(cli, "call", "-m", "3", "-b", bpath, "-l", lpath, "-o", prefix, first_bam, bins[0])
    exp_gt2, exp_lines2, exp_unplaced2, _ = co.call(first_recs, datas[0], min_support=3, bounds_lines=bounds_in, bed_lines=bed_in)
    got = open(prefix + "-bounds.txt").read().splitlines()
    assert got[0] == eo.BOUNDS_HEADER + "\tdepth" and got[1:] == exp_lines2 and len(exp_lines2) > 10
    got_gt = open(prefix + "-genotype.txt").read().splitlines()
    assert sorted(got_gt[1:]) == sorted(exp_gt2) and len(exp_gt2) == len(exp_lines2)
    # merge: joint clustering with per-sample support (merge.nim:172-187)
    for ms, extra in ((5, []), (2, ["-c", "0", "-t", "3"]), (2, ["-c", "1", "-t", "1"]), (4, ["-w", "300"])):
        prefix = str(tmp_path / f"merge{ms}{''.join(extra)}")
        run(cli, "merge", "-m", str(ms), *extra, "-o", prefix, *bins)
        kw = dict(min_support=ms)
        if extra[:1] == ["-c"]:
            kw.update(min_clip=int(extra[1]), min_clip_total=int(extra[3]))
        if extra[:1] == ["-w"]:
            kw.update(window=300)
        exp_lines, _ = eo.merge(datas, **kw)
        got = open(prefix + "-bounds.txt").read().splitlines()
        assert got[0] == eo.BOUNDS_HEADER and (len(exp_lines) > 5 or extra[:2] == ["-c", "1"])
        assert sorted(got[1:]) == sorted(exp_lines) and got[1:] == exp_lines
    # the multi-GPU joint merge (strling_b200/joint.py; here one process / one GPU) writes the same file as `strling merge`
    import sys
    prefix = str(tmp_path / "joint")
    r = subprocess.run([sys.executable, "-m", "strling_b200.joint", "-m", "5", "-o", prefix, *bins], capture_output=True, text=True,
                       cwd=os.path.dirname(HERE))
    assert r.returncode == 0, r.stderr
    assert open(prefix + "-bounds.txt").read() == open(str(tmp_path / "merge5") + "-bounds.txt").read()
    # -l bed: listed loci take their reads before clustering and are reported first (merge.nim:154-168, callclusters.nim:14-50)
    bed = str(tmp_path / "loci.bed")
    bed_lines = [f"{targets[tid][0]}\t{s}\t{e}\t{u}\tL{i}" if i % 2 else f"{targets[tid][0]} {s} {e} {u}" for i, (tid, s, e, u) in enumerate(loci[:12])]
    bed_lines.append(f"{targets[0][0]}\t5\t9\tGGGGGC")  # a bucket that does not exist
    open(bed, "w").write("\n".join(bed_lines) + "\n")
    prefix = str(tmp_path / "merge_bed")
    run(cli, "merge", "-m", "3", "-l", bed, "-o", prefix, *bins)
    exp_lines, _ = eo.merge(datas, min_support=3, bed_lines=bed_lines)
    got = open(prefix + "-bounds.txt").read().splitlines()
    assert got[1:] == exp_lines and sum(int(l.split("\t")[10]) for l in got[1:14]) > 50
    # --chromosome restricts parsing to one contig (merge.nim:52,89)
    prefix = str(tmp_path / "merge_chr")
    run(cli, "merge", "-m", "3", "--chromosome", targets[2][0], "-o", prefix, *bins)
    got = open(prefix + "-bounds.txt").read().splitlines()[1:]
    all_lines, _ = eo.merge(datas, min_support=3)
    assert got == [l for l in all_lines if l.split("\t")[0] == targets[2][0]] and len(got) > 0


def _simulate_chunk(args):
    seed, n_pairs, targets, loci, frac, unmapped = args
    return bamio.simulate_alignments(seed, n_pairs, targets, loci, str_pair_frac=frac, unmapped_pairs=unmapped, name_prefix=f"c{seed}_")


def test_config3_30x_depth_extract_then_call(cli, tmp_path):
    # BASELINE.json configs[3] at real depth: the sim/ disease loci (tests/golden/disease_loci.json) on contigs sized so that
    # 10^6 read pairs of 150 bp are 30x coverage (sim/sim_shared.groovy:6,65), through `strling extract` -> `strling call`;
    # .bin byte-identical, bounds / unplaced / genotype lines equal to the oracle's.  STRLING_CONFIG3_PAIRS scales it.
    import multiprocessing as mp

    n_pairs = int(os.environ.get("STRLING_CONFIG3_PAIRS", "1000000"))
    loci_json = json.load(open(os.path.join(HERE, "golden", "disease_loci.json")))
    chroms = sorted({l["chrom"] for l in loci_json}, key=lambda c: (len(c), c))
    contig = max(100_000, n_pairs * 300 // 30 // len(chroms))
    targets = [(f"chr{c}", contig) for c in chroms]
    loci = []
    for l in loci_json:
        start = 20_000 + l["start"] % (contig - 40_000)
        loci.append((chroms.index(l["chrom"]), start, start + max(20, min(200, l["stop"] - l["start"])), l["unit"]))
    n_chunks = 16
    jobs = [(100 + c, n_pairs // n_chunks, targets, loci, 0.03, 40) for c in range(n_chunks)]
    with mp.get_context("fork").Pool(min(n_chunks, os.cpu_count() or 1)) as pool:
        parts = pool.map(_simulate_chunk, jobs)
    placed = [a for part in parts for a in part if a.tid >= 0]
    unplaced = [a for part in parts for a in part if a.tid < 0]
    placed.sort(key=lambda a: (a.tid, a.pos))
    recs = placed + unplaced
    assert len(recs) >= 2 * (n_pairs // n_chunks) * n_chunks
    hdr = bamio.sam_header(targets)
    bam, out = str(tmp_path / "c3.bam"), str(tmp_path / "c3.bin")
    bamio.write_bam(bam, hdr, targets, recs)
    run(cli, "extract", bam, out)
    exp, cache, _ = eo.extract(recs, targets, hdr)
    assert open(out, "rb").read() == exp and len(cache) > n_pairs // 100
    prefix = str(tmp_path / "c3")
    run(cli, "call", "-o", prefix, bam, out)
    exp_gt, exp_lines, exp_unplaced, _ = co.call(recs, exp)
    got = open(prefix + "-bounds.txt").read().splitlines()
    assert got[0] == eo.BOUNDS_HEADER + "\tdepth" and len(exp_lines) >= len(loci) // 2
    assert sorted(got[1:]) == sorted(exp_lines)
    got_un = dict(l.split("\t") for l in open(prefix + "-unplaced.txt").read().splitlines())
    assert {k.encode(): int(v) for k, v in got_un.items()} == exp_unplaced
    got_gt = open(prefix + "-genotype.txt").read().splitlines()
    assert got_gt[0] == co.GT_HEADER and sorted(got_gt[1:]) == sorted(exp_gt)


def _random_genome(seed, targets):
    rng = np.random.default_rng(seed)

    def rnd(n):
        return "".join("ACGT"[i] for i in rng.integers(0, 4, size=n))

    chroms = []
    for name, ln in targets:
        parts, used = [], 0
        while used < ln:
            n = int(rng.integers(200, 3000))
            parts.append(rnd(n))
            unit = ["CAG", "A", "AC", "AAAG", "ATTCT", "CACGAT", "GGC", "T"][int(rng.integers(0, 8))]
            rep = unit * int(rng.integers(5, 90))
            if rng.random() < 0.3:
                rep = rep.lower()            # soft-masked reference sequence is upper-cased first (genome_strs.nim:72)
            if rng.random() < 0.2:
                rep = rep[: len(rep) // 2] + "N" * int(rng.integers(1, 40)) + rep[len(rep) // 2:]
            parts.append(rep)
            used += n + len(rep)
        chroms.append((name, "".join(parts)[:ln]))
    return chroms


def test_index_matches_oracle_and_feeds_extract(cli, tmp_path):
    # `strling index` (genome_strs.nim:61-131): 100-bp windows, step 60, merge + trim -> bed of STR-like regions
    targets = [("chr1", 60_000), ("chr2", 45_000), ("chrS", 130)]
    chroms = _random_genome(3, targets)
    fasta = str(tmp_path / "ref.fa")
    with open(fasta, "w") as fh:
        for name, seq in chroms:
            fh.write(f">{name} some description\n")
            for i in range(0, len(seq), 60):
                fh.write(seq[i:i + 60] + "\n")
    bed = str(tmp_path / "ref.fa.str")
    run(cli, "index", "-g", bed, fasta)
    exp = eo.genome_repeat_lines(chroms, 0.8)
    got = open(bed).read().splitlines()
    assert len(exp) > 20 and got == exp
    # extract without an existing -g file builds the same index first (genome_strs.nim:117-128) and then filters with it
    loci = [(0, 10_000, 10_150, "CAG"), (1, 20_000, 20_090, "AAAG")]
    recs = bamio.simulate_alignments(11, 4000, targets[:2], loci, unmapped_pairs=30)
    hdr = bamio.sam_header(targets[:2])
    bam, out, bed2 = str(tmp_path / "x.bam"), str(tmp_path / "x.bin"), str(tmp_path / "built.str")
    bamio.write_bam(bam, hdr, targets[:2], recs)
    run(cli, "extract", "-f", fasta, "-g", bed2, bam, out)
    assert open(bed2).read().splitlines() == exp
    expected_bin, cache, _ = eo.extract(recs, targets[:2], hdr, 0.8, 40, eo.read_bed(bed2))
    assert open(out, "rb").read() == expected_bin and len(cache) > 100


def test_cli_errors(cli, tmp_path):
    r = subprocess.run([cli, "extract", str(tmp_path / "missing.bam"), str(tmp_path / "o.bin")], capture_output=True, text=True)
    assert r.returncode != 0 and "couldn't open bam" in r.stderr
    r = subprocess.run([cli, "merge", str(tmp_path / "missing.bin")], capture_output=True, text=True)
    assert r.returncode != 0 and "unable to open" in r.stderr
