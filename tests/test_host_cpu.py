"""CPU-side checks of the C++ host code behind the `strling` command line, against the oracle: BAM decode, fragment
distribution, `.bin` codec (also cross-checked with python msgpack), and the per-pair arithmetic
(adjust_by / unplaced_pair / canonical_repeat, extract.nim:134-190, utils.nim:61-83,304-310)."""
import os
import subprocess

import numpy as np
import pytest

from oracle import extract_oracle as eo
from oracle import oracle as orc
from strling_b200 import bamio
from strling_b200 import build as sb_build

TARGETS = [("chr1", 1_000_000), ("chr2", 800_000), ("chrUn_x", 50_000)]
LOCI = [(0, 100000, 100150, "CAG"), (0, 500000, 500090, "AAAG"), (1, 300000, 300060, "ATTCT"), (1, 600000, 600040, "A")]


@pytest.fixture(scope="module")
def cli():
    return sb_build.build_cli()


@pytest.fixture(scope="module")
def sim(tmp_path_factory):
    d = tmp_path_factory.mktemp("sim")
    recs = bamio.simulate_alignments(3, 1500, TARGETS, LOCI, unmapped_pairs=30, n_frac=0.03)
    bam = str(d / "sim.bam")
    bamio.write_bam(bam, bamio.sam_header(TARGETS), TARGETS, recs)
    return recs, bam, d


def run(cli, *args, stdin=None):
    r = subprocess.run([cli, *args], input=stdin, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return r.stdout


def test_bam_decode_matches_writer(cli, sim):
    recs, bam, _ = sim
    lines = [l for l in run(cli, "debug", "bam", bam).splitlines() if not l.startswith("@")]
    assert len(lines) == len(recs)
    for a, l in zip(recs, lines):
        f = l.split("\t")
        cig = "".join(f"{n}{op}" for op, n in a.cigar) or "*"
        assert f[:10] == [a.qname, str(a.flag), str(a.tid), str(a.pos), str(a.mapq), cig, str(a.mate_tid), str(a.mate_pos),
                          str(a.isize), a.seq]
        assert int(f[10]) == eo.aln_stop(a)


def test_fragment_distribution_and_medians(cli, sim):
    recs, bam, _ = sim
    exp = eo.fragment_length_distribution(recs)
    out = run(cli, "debug", "fragdist", bam).splitlines()
    got = np.zeros(4096, dtype=np.uint32)
    for l in out[:-1]:
        i, c = l.split("\t")
        got[int(i)] = int(c)
    assert np.array_equal(got, exp)
    assert out[-1].split("\t")[1:] == [str(orc.median(exp, 0.5)), str(orc.median(exp, 0.98)), str(orc.median(exp, 0.99))]


def test_bin_codec_roundtrip(cli, sim):
    recs, bam, d = sim
    data, cache, fd = eo.extract(recs, TARGETS, bamio.sam_header(TARGETS), 0.8, 40)
    assert len(cache) > 50
    src, dst = str(d / "oracle.bin"), str(d / "reencoded.bin")
    open(src, "wb").write(data)
    out = run(cli, "debug", "bin", src, dst).splitlines()
    assert open(dst, "rb").read() == data                      # C++ writer == python msgpack writer, byte for byte
    u = eo.unpack_bin(data)
    assert out[3] == f"n\t{len(cache)}"
    for l, t, q in zip(out[4:], u["treads"], u["qnames"]):
        unit = bytes(t["repeat"]).decode() or "."
        assert l == f"{t['tid']}\t{t['position']}\t{unit}\t{t['flag']}\t{t['split']}\t{t['mapq']}\t{t['repeat_count']}\t{t['align_length']}\t{q}"


def test_bin_reader_rejects_bad_magic(cli, tmp_path):
    p = tmp_path / "bad.bin"
    p.write_bytes(b"XYZ" + b"\0" * 20000)
    r = subprocess.run([cli, "debug", "bin", str(p)], capture_output=True, text=True)
    assert r.returncode != 0 and "expected bin file to start" in r.stderr


def _tread_fields(t):
    unit = bytes(t["repeat"][0]).decode() or "."
    return f"{t['tid'][0]} {t['position'][0]} {unit} {t['flag'][0]} {t['split'][0]} {t['mapq'][0]} {t['repeat_count'][0]} {t['align_length'][0]}"


def test_pair_arithmetic_matches_oracle(cli):
    rng = np.random.default_rng(8)
    units = [b"", b"A", b"AC", b"CAG", b"CCG", b"AAAG", b"ATTCT", b"CACGAT", b"TTTTTG", b"GGC", b"T"]
    cmds, expect = [], []
    # reference vector: tests/test_extract.nim:7-19
    A = orc.make_tread(tid=2, position=86914345, repeat=b"CCG", mapq=10, repeat_count=40, align_length=80)
    B = orc.make_tread(tid=16, position=17470852, split=orc.NONE_RIGHT, mapq=60, repeat_count=0, align_length=71)
    cases = [(A, B, 0.4, 20, 0, 17470852)]
    for _ in range(3000):
        def rt():
            u = units[int(rng.integers(0, len(units)))]
            return orc.make_tread(tid=int(rng.integers(-1, 5)), position=int(rng.choice([0, 5, 100, 4294967290, int(rng.integers(0, 2 ** 31))])),
                                  repeat=u, flag=int(rng.integers(0, 4096)), split=int(rng.integers(0, 6)), mapq=int(rng.integers(0, 61)),
                                  repeat_count=int(rng.choice([0, 1, 10, 40, 75, 150, 200])) if u else 0, align_length=int(rng.choice([0, 1, 71, 150, 151, 250])))
        cases.append((rt(), rt(), float(rng.choice([0.8, 0.73, 0.6, 0.4])), int(rng.choice([0, 20, 40])), int(rng.integers(0, 900)),
                      int(rng.choice([0, 3, 2 ** 31, int(rng.integers(0, 2 ** 32))]))))
    for A, B, p, mq, mf, bpos in cases:
        cmds.append(f"prepeat {_tread_fields(A)}")
        expect.append(f"{orc.p_repeat(A):.17g}")
        cmds.append(f"unplaced {_tread_fields(A)} {_tread_fields(B)} {p!r} {mq}")
        expect.append(str(int(orc.unplaced_pair(A, B, p, mq))))
        cmds.append(f"adjust {_tread_fields(A)} {_tread_fields(B)} {p!r} {mq} {mf} {bpos}")
        A2 = A.copy()
        r = orc.adjust_by(A2, B, p, mq, mf, bpos)
        unit = bytes(A2["repeat"][0]).decode() or "."
        expect.append(f"{int(r)}\t{A2['tid'][0]}\t{A2['position'][0]}\t{unit}\t{A2['split'][0]}\t{A2['mapq'][0]}")
    import itertools
    for k in range(1, 5):
        for tup in itertools.product("ACGT", repeat=k):
            u = "".join(tup)
            cmds.append(f"canonical {u}")
            expect.append(orc.canonical_repeat(u.encode()).decode())
            cmds.append(f"minrc {u}")
            expect.append(orc.min_rev_complement(u.encode()).decode())
    for u in ("CCCTT", "AAAAAT", "CACGAT", "TTTTTG", "ACGTAC", "GGGGGC"):
        cmds.append(f"canonical {u}")
        expect.append(orc.canonical_repeat(u.encode()).decode())
    got = run(cli, "debug", "logic", stdin="\n".join(cmds) + "\n").splitlines()
    assert len(got) == len(expect)
    for c, g, e in zip(cmds, got, expect):
        assert g == e, (c, g, e)
    assert orc.canonical_repeat(b"CCCTT") == b"AAGGG"   # tests/test_utils.nim:66-74


def test_cli_fails_loudly_without_gpu(cli, sim, tmp_path):
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    recs, bam, _ = sim
    r = subprocess.run([cli, "extract", bam, str(tmp_path / "x.bin")], capture_output=True, text=True)
    assert r.returncode != 0 and "no CPU fallback" in r.stderr
    assert not os.path.exists(tmp_path / "x.bin")
