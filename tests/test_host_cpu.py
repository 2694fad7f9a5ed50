"""CPU-side checks of the C++ host code behind the `strling` command line, against the oracle: BAM decode, fragment
distribution, `.bin` codec (also cross-checked with python msgpack), and the per-pair arithmetic
(adjust_by / unplaced_pair / canonical_repeat, extract.nim:134-190, utils.nim:61-83,304-310)."""
import os
import struct
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


def run(cli, *args, stdin=None, env=None):
    r = subprocess.run([cli, *args], input=stdin, capture_output=True, text=True, env=dict(os.environ, **env) if env else None)
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


def test_bam_decode_against_an_independent_encoder(cli, tmp_path):
    """host/bam.hpp against a BAM laid out HERE from the SAM/BAM specification (SAMv1 section 4.2) with nothing shared with
    strling_b200/bamio.py: records span BGZF block boundaries, blocks have uneven sizes and different deflate settings
    (stored, fixed and dynamic Huffman), the gzip extra field carries a second subfield, tags and qualities follow the SEQ,
    one CIGAR has every operator and one read has no CIGAR / SEQ at all."""
    import struct
    import zlib

    nib = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
    ops = {c: i for i, c in enumerate("MIDNSHP=X")}

    def record(qname, flag, tid, pos, mapq, cigar, mtid, mpos, tlen, seq, tags=b""):
        name = qname.encode() + b"\0"
        cig = b"".join(struct.pack("<I", (n << 4) | ops[o]) for n, o in cigar)
        sq = bytearray((len(seq) + 1) // 2)
        for i, ch in enumerate(seq):
            sq[i // 2] |= nib[ch] << (4 if i % 2 == 0 else 0)
        qual = bytes([0xFF] * len(seq))
        ref_span = sum(n for n, o in cigar if o in "MDN=X")
        end = pos + (ref_span if ref_span else 1)
        # reg2bin (SAMv1 section 5.3)
        b, e = max(pos, 0), max(end, 1) - 1
        bin_ = next((((1 << s) - 1) // 7 + (b >> sh) for s, sh in ((15, 14), (12, 17), (9, 20), (6, 23), (3, 26)) if b >> sh == e >> sh), 0)
        core = struct.pack("<iiBBHHHIiii", tid, pos, len(name), mapq, bin_, len(cigar), flag, len(seq), mtid, mpos, tlen)
        body = core + name + cig + bytes(sq) + qual + tags
        return struct.pack("<i", len(body)) + body

    targets = [("chrA", 100_000), ("chrB_random", 5_000)]
    text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in targets) + "@PG\tID:hand\n"
    hdr = b"BAM\1" + struct.pack("<i", len(text)) + text.encode() + struct.pack("<i", len(targets))
    for n, l in targets:
        hdr += struct.pack("<i", len(n) + 1) + n.encode() + b"\0" + struct.pack("<i", l)
    rng = np.random.default_rng(99)
    expect, body = [], b""
    every_op = [(5, "S"), (10, "M"), (2, "I"), (3, "D"), (7, "N"), (4, "="), (1, "X"), (6, "M"), (2, "P"), (3, "H")]
    for i in range(400):
        kind = i % 5
        seq = "".join(rng.choice(list("ACGTN" if i % 7 else "ACGTRYKMSWBDHVN="), size=int(rng.integers(30, 260))))
        cigar = [(len(seq), "M")]
        if kind == 1:
            cigar = [(20, "S"), (len(seq) - 20, "M")]
        elif kind == 2:
            seq = seq[:38]
            cigar = every_op   # query-consuming ops: 5+10+2+4+1+6 = 28 ... make SEQ match below
            seq = (seq * 2)[: sum(n for n, o in every_op if o in "MIS=X")]
        elif kind == 3:
            cigar = [(len(seq) - 17, "M"), (17, "S")]
        tid, pos = (0, 100 + 37 * i) if i < 390 else (-1, -1)
        if tid < 0:
            cigar = []
        flag = [99, 147, 83, 163, 77, 141, 1, 2113, 355][i % 9] if tid >= 0 else 77
        tags = struct.pack("<2sci", b"NM", b"i", i) + b"RGZgrp" + bytes([0])
        if i == 123:
            seq, cigar = "", []     # no SEQ: l_seq = 0
        body += record(f"q{i}:{'x' * (i % 40)}", flag, tid, pos, i % 61, cigar, 1 if i % 11 == 0 else tid, pos + 300 if tid >= 0 else -1,
                       (350 if i % 2 == 0 else -350) if tid >= 0 else 0, seq, tags)
        cig_s = "".join(f"{n}{o}" for n, o in cigar) or "*"
        span = sum(n for n, o in cigar if o in "MDN=X")
        expect.append((f"q{i}:{'x' * (i % 40)}", flag, tid, pos, i % 61, cig_s, 1 if i % 11 == 0 else tid, pos + 300 if tid >= 0 else -1,
                       ((350 if i % 2 == 0 else -350) if tid >= 0 else 0), seq, span))
    payload = hdr + body

    def bgzf(data, level, strategy, extra_sub):
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 9, strategy)
        comp = co.compress(data) + co.flush()
        xlen = 6 + len(extra_sub)
        bsize = 12 + xlen + len(comp) + 8 - 1
        assert bsize < 65536
        return (struct.pack("<BBBBIBBH", 31, 139, 8, 4, 0, 0, 255, xlen) + extra_sub + struct.pack("<BBHH", 66, 67, 2, bsize) + comp
                + struct.pack("<II", zlib.crc32(data), len(data)))

    out, off, k = b"", 0, 0
    sizes = [7, 60_000, 113, 1, 4096, 30_000, 65_280]
    while off < len(payload):
        n = min(sizes[k % len(sizes)], len(payload) - off)
        level, strategy = [(0, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (9, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_HUFFMAN_ONLY)][k % 4]
        extra = struct.pack("<BBH3s", 88, 89, 3, b"abc") if k % 3 == 1 else b""   # an unrelated subfield BEFORE the BC one
        out += bgzf(payload[off: off + n], level, strategy, extra)
        off += n
        k += 1
    out += bgzf(b"", 6, zlib.Z_DEFAULT_STRATEGY, b"")    # EOF marker block
    bam = str(tmp_path / "hand.bam")
    open(bam, "wb").write(out)
    lines = run(cli, "debug", "bam", bam).splitlines()
    assert lines[0] == "@targets 2" and lines[1] == "@target\tchrA\t100000" and lines[2] == "@target\tchrB_random\t5000"
    assert lines[3] == f"@text_bytes {len(text)}"
    recs = [l.split("\t") for l in lines[4:]]
    assert len(recs) == len(expect)
    for f, e in zip(recs, expect):
        qname, flag, tid, pos, mapq, cig, mtid, mpos, tlen, seq, span = e
        assert f[:10] == [qname, str(flag), str(tid), str(pos), str(mapq), cig, str(mtid), str(mpos), str(tlen), seq], (f, e)
        # hts-nim `stop` = htslib bam_endpos: pos + reference span of the CIGAR; pos + 1 for a read flagged unmapped or without
        # a reference-consuming CIGAR
        assert int(f[10]) == (pos + span if span and not flag & 4 else pos + 1), (f, e)


def _awkward_pairs(targets, rng):
    """Records that exercise the order-dependent corners of Cache.add (extract.nim:60-61,192-248): mates at the SAME position
    (after_mate then depends on the table), a qname that occurs three and four times (hasKeyOrPut's 'bad read' branch drops the
    table entry), and second-seen reads whose first-seen mate is missing."""
    out = []
    unit = "CAG" * 50
    for k in range(40):
        tid, pos = k % 2, 5_000 + 7_000 * k
        name = f"same{k}"
        seqs = [unit if k % 3 == 0 else "".join(rng.choice("ACGT") for _ in range(150)), unit if k % 3 == 1 else "".join(rng.choice("ACGT") for _ in range(150))]
        mapq = [60, 0 if k % 4 == 0 else 60]
        out.append(bamio.Aln(name, 0x1 | 0x40 | 0x20, tid, pos, mapq[0], [("M", 150)], tid, pos, 0, seqs[0]))
        out.append(bamio.Aln(name, 0x1 | 0x80 | 0x10, tid, pos, mapq[1], [("S", 40), ("M", 110)] if k % 5 == 0 else [("M", 150)], tid, pos, 0, seqs[1]))
    for k in range(12):
        tid, pos = 0, 290_000 + 500 * k
        name = f"dup{k}"
        for j in range(3 + k % 2):  # 3 or 4 records of one name, all "first-seen" by position
            out.append(bamio.Aln(name, 0x1 | 0x40, tid, pos + j, 60, [("M", 150)], tid, pos + 300, 300, unit if j == 1 else "".join(rng.choice("ACGT") for _ in range(150))))
        out.append(bamio.Aln(name, 0x1 | 0x80 | 0x10, tid, pos + 300, 60, [("M", 100), ("S", 50)], tid, pos, -300, "".join(rng.choice("ACGT") for _ in range(100)) + unit[:50]))
    for k in range(10):  # the mate never shows up / shows up as a secondary record only
        out.append(bamio.Aln(f"orphan{k}", 0x1 | 0x80 | 0x10, 1, 150_000 + 10 * k, 60, [("M", 150)], 0, 1_000, 0, unit))
    return out


@pytest.mark.parametrize("p,q,use_bed,batch,env", [
    (0.8, 40, False, 262144, None), (0.8, 40, True, 1500, None), (0.7, 20, True, 700, None), (0.9, 0, False, 333, None),
    # the same through the thread pool: parallel inflate / walk / staging and the replay split over qname-hash shards
    (0.8, 40, True, 1500, {"STRLING_DEBUG_THREADS": "8", "STRLING_DEBUG_SHARDS": "5"}),
    (0.7, 20, False, 333, {"STRLING_DEBUG_THREADS": "3", "STRLING_DEBUG_SHARDS": "16"}),
    (0.8, 40, False, 262144, {"STRLING_DEBUG_THREADS": "8", "STRLING_DEBUG_SHARDS": "2"}),
    # the --gpu-inflate plumbing (compressed bytes staged contiguously, rebased descriptors, inflate hook of the chunk reader) with
    # the blocks decoded on the host from the staging buffer
    (0.8, 40, True, 1500, {"STRLING_DEBUG_THREADS": "4", "STRLING_DEBUG_STAGED_INFLATE": "1"}),
])
def test_extract_staging_and_replay_without_the_scan(cli, tmp_path, p, q, use_bed, batch, env):
    """The host half of `strling extract` on the CPU: BAM decode -> which segments a record contributes (genome-STR filter,
    soft clips under both proportion classes, extract.nim:20-40,93-114) -> replay of Cache.add in file order (extract.nim:192-248)
    -> .bin.  `strling debug extract dump` writes the staged segments, the ORACLE scans them here in the test, and
    `strling debug extract replay` consumes those results: the .bin must be the oracle pipeline's, byte for byte, also with
    batches so small that mates land in different batches."""
    targets = [("chr1", 400_000), ("chr2", 300_000)]
    loci = [(0, 100_000, 100_150, "CAG"), (0, 250_000, 250_090, "AAAG"), (1, 120_000, 120_060, "ATTCT"), (1, 200_000, 200_040, "A")]
    recs = bamio.simulate_alignments(31, 4000, targets, loci, unmapped_pairs=60, n_frac=0.03)
    import random
    extra_recs = _awkward_pairs(targets, random.Random(5))
    placed = sorted([a for a in recs if a.tid >= 0] + extra_recs, key=lambda a: (a.tid, a.pos))  # stable: equal positions keep their order
    recs = placed + [a for a in recs if a.tid < 0]
    hdr = bamio.sam_header(targets)
    bam, segs_path, res_path, out = (str(tmp_path / n) for n in ("x.bam", "segs.tsv", "res.bin", "x.bin"))
    bamio.write_bam(bam, hdr, targets, recs)
    extra, genome_str = [], None
    if use_bed:
        bed = str(tmp_path / "ref.str")
        with open(bed, "w") as fh:
            for tid, s, e, u in loci:
                fh.write(f"{targets[tid][0]}\t{s}\t{e}\t{u}\n")
        extra = [bed]
        genome_str = eo.read_bed(bed)
    run(cli, "debug", "extract", "dump", segs_path, bam, out, repr(p), str(q), str(batch), *extra, env=env)
    classes = [p, p - 0.07, min(p, 0.6)]
    lines = open(segs_path).read().splitlines()
    assert len(lines) > 1000
    res = np.zeros(len(lines), dtype=[("unit", "S6"), ("repeat_count", "<u2")])
    for i, l in enumerate(lines):
        cls, _, seq = l.partition("\t")
        unit, count = orc.get_repeat(seq, classes[int(cls)])
        res["unit"][i], res["repeat_count"][i] = unit, count
    res.tofile(res_path)
    run(cli, "debug", "extract", "replay", res_path, bam, out, repr(p), str(q), str(batch), *extra, env=env)
    exp, cache, _ = eo.extract(recs, targets, hdr, p, q, genome_str)
    assert len(cache) > 500
    assert open(out, "rb").read() == exp


def test_inflate_decoder_against_zlib(cli):
    """The repo's whole-block DEFLATE decoder (host/inflate_fast.hpp) against zlib's encoder inside the binary: every level and
    strategy (stored / fixed / dynamic blocks, multi-block streams, sub-table codes), six kinds of data, all sizes below 64
    and random sizes up to a BGZF block; output must match byte for byte, nothing outside the output buffer may be touched,
    and a stream with a flipped bit must never write outside it either."""
    for seed in (1, 2):
        out = run(cli, "debug", "inflate-selftest", str(seed), "160")
        assert out.startswith("ok\t"), out


def _decoy_chain(n_fake: int) -> bytes:
    """n_fake byte strings that each look like a complete, minimal BAM record (block_size 34, refID 0, pos 0, a 2-byte name,
    no CIGAR, no SEQ) and chain into each other -- planted inside QUAL fields to mislead a record-start guesser."""
    one = struct.pack("<iiiBBHHHiiii", 34, 0, 0, 2, 0, 4680, 0, 0, 0, -1, -1, 0) + b"A\0"
    assert len(one) == 38
    return one * n_fake


def test_parallel_record_walk_is_exact_despite_decoys(cli, tmp_path):
    """BamChunkReader cuts the record walk of a chunk into parts that GUESS their first record start and are then verified
    against the chain from the chunk's first byte (host/bam.hpp).  Here most QUAL fields hold chains of well-formed fake
    records, so the guesses of many parts are wrong: the reader must notice (rewalked > 0) and still deliver exactly the
    records the single-threaded walk delivers (same count, same digest over tid / pos / flag / l_seq / qname)."""
    import random
    rng = random.Random(11)
    targets = [("chr1", 5_000_000)]
    recs = []
    for i in range(40_000):
        l_seq = 380 if i % 3 else 150
        qual = _decoy_chain(10) if l_seq == 380 else bytes(rng.randrange(0, 94) for _ in range(150))
        a = bamio.Aln(f"r{i:07d}", 0x1 | 0x40, 0, 100 + 50 * i, 60, [("M", l_seq)], 0, 400 + 50 * i, 450, "".join(rng.choice("ACGT") for _ in range(l_seq)))
        a.extra["qual"] = qual
        recs.append(a)
    bam = str(tmp_path / "decoy.bam")
    bamio.write_bam(bam, bamio.sam_header(targets), targets, recs)
    serial = run(cli, "debug", "chunks", bam, "1", "300", "digest").split("\t")
    par = run(cli, "debug", "chunks", bam, "8", "300", "digest").split("\t")
    f = lambda row, key: row[row.index(key) + 1].strip()
    assert f(serial, "records") == f(par, "records") == "40000"
    assert f(serial, "digest") == f(par, "digest")
    assert f(serial, "rewalked") == "0" and int(f(par, "rewalked")) > 0
    # and the record-at-a-time reader (the one `debug bam` uses) agrees on the count
    assert sum(1 for l in run(cli, "debug", "bam", bam).splitlines() if not l.startswith("@")) == 40000


def test_synth_bam_is_a_valid_sorted_bam(cli, tmp_path):
    """`strling debug synth-bam` (the measurement input of bench.py's cli leg): coordinate-sorted, mates consistent, the
    no-coordinate pairs at the end, and readable by both readers."""
    bam = str(tmp_path / "s.bam")
    run(cli, "debug", "synth-bam", bam, "5000", "9")
    rows = [l.split("\t") for l in run(cli, "debug", "bam", bam).splitlines() if not l.startswith("@")]
    assert len(rows) == 10000
    keys = [(int(r[2]) if int(r[2]) >= 0 else 1 << 30, int(r[3])) for r in rows]
    assert keys == sorted(keys)
    by_name = {}
    for r in rows:
        by_name.setdefault(r[0], []).append(r)
    assert all(len(v) == 2 for v in by_name.values())
    for a, b in by_name.values():
        assert (a[2], a[3]) == (b[6], b[7]) and (b[2], b[3]) == (a[6], a[7])
    assert sum(1 for r in rows if int(r[2]) < 0) == 100
    par = run(cli, "debug", "chunks", bam, "4", "64").split("\t")
    assert par[par.index("records") + 1] == "10000"


def test_mate_table_against_unordered_map(cli):
    """The replay's mate table (host/mate_table.hpp: open addressing over a pooled entry array, backward-shift deletion, names inline
    up to 54 bytes) under random insert / find / take traffic against std::unordered_map -- growth to >10^5 live entries, and
    degraded hashes (4096 and 64 classes) that force long probe runs."""
    out = run(cli, "debug", "matetable-selftest", "3", "600000")
    assert [l.split("\t")[0] for l in out.splitlines()] == ["mode 0 ok", "mode 1 ok", "mode 2 ok"], out


def _fuzz_records(rng, n, targets):
    """Random BAM records far from what an aligner writes: every flag combination, empty / clip-only / hard-clipped / spliced
    CIGARs, names of 1..250 bytes that repeat up to four times, mates that point anywhere (same position, other contig, nowhere),
    mapq around the threshold, reads of 0..300 bases made of repeats, noise, N and IUPAC codes."""
    pool = []
    for k in range(max(4, n // 2)):
        ln = rng.choice([1, 2, 7, 15, 16, 31, 40, 54, 55, 56, 90, 250]) if rng.random() < 0.3 else rng.randrange(8, 40)
        pool.append(("%x" % rng.getrandbits(4 * ln)).rjust(ln, "q")[:ln - len(str(k))] + str(k) if ln > len(str(k)) else str(k))
    units = ["A", "C", "AC", "AG", "CAG", "AAG", "AAAG", "ATTCT", "AAGGG", "CACGAT", "CCCCGG"]

    def seq_of(L, kind):
        if kind == 0:
            return "".join(rng.choice("ACGT") for _ in range(L))
        u = rng.choice(units)
        if len(u) == 1 and L > 250:   # a homopolymer of 256+ bases trips the reference's doAssert repeat_count < 256 (extract.nim:72)
            u = "AC"
        s = (u * (L // len(u) + 2))[rng.randrange(len(u)):][:L]
        if kind == 2:   # half repeat, half noise (either order): what a clipped STR read looks like
            h = L // 2
            s = s[:h] + "".join(rng.choice("ACGT") for _ in range(L - h)) if rng.random() < 0.5 else "".join(rng.choice("ACGT") for _ in range(h)) + s[h:]
        s = list(s)
        for i in range(L):
            r = rng.random()
            if r < 0.01:
                s[i] = rng.choice("ACGT")
            elif kind == 3 and r < 0.08:
                s[i] = rng.choice("NNNRYKM")
        return "".join(s)

    recs = []
    for i in range(n):
        placed = rng.random() < 0.92
        tid = rng.randrange(len(targets)) if placed else -1
        pos = rng.randrange(0, targets[tid][1] - 2000) if placed else -1
        flag = 0
        for bit, p in ((0x1, 0.95), (0x2, 0.7), (0x4, 0.03), (0x8, 0.05), (0x10, 0.5), (0x20, 0.5), (0x100, 0.03), (0x800, 0.03), (0x400, 0.02)):
            if rng.random() < p:
                flag |= bit
        flag |= rng.choice([0x40, 0x80])
        if not placed:
            flag |= 0x4
        r = rng.random()
        if not placed or r < 0.1:
            mtid, mpos = -1, -1
        elif r < 0.25:
            mtid, mpos = tid, pos                      # same position: after_mate asks the table
        elif r < 0.4:
            mtid, mpos = rng.randrange(len(targets)), rng.randrange(0, 100000)
        else:
            mtid, mpos = tid, max(0, pos + rng.randrange(-600, 600))
        L = rng.choice([0, 1, 30, 100, 150, 150, 150, 151, 250, 300])
        t = rng.randrange(9)
        if not placed or t == 0 or L == 0:
            cigar = []
        elif t == 1:
            cigar = [("M", L)]
        elif t == 2:
            a = min(L - 1, rng.choice([5, 16, 17, 40, 100])) if L > 1 else 0
            cigar = [("S", a), ("M", L - a)] if a else [("M", L)]
        elif t == 3:
            a = min(L - 1, rng.choice([5, 16, 17, 40, 100])) if L > 1 else 0
            cigar = [("M", L - a), ("S", a)] if a else [("M", L)]
        elif t == 4 and L >= 60:
            a, c = rng.choice([10, 17, 25]), rng.choice([16, 17, 30])
            cigar = [("S", a), ("M", L - a - c), ("S", c)]
        elif t == 5:
            cigar = [("S", L)]                         # one op: add_soft takes it as "left" twice
        elif t == 6 and L >= 120:
            cigar = [("H", 5), ("S", 20), ("M", 30), ("I", 3), ("M", 20), ("D", 4), ("N", 50), ("M", L - 20 - 30 - 3 - 20 - 18), ("S", 18), ("H", 3)]
        elif t == 7 and L >= 30:
            cigar = [("=", 10), ("X", 1), ("=", L - 11)]
        else:
            cigar = [("M", L)]
        seq = seq_of(L, rng.choice([0, 0, 1, 1, 2, 2, 3]))
        isize = rng.choice([0, 300, 350, 400, 450, 4095, 4096, -400, 10000])
        recs.append(bamio.Aln(rng.choice(pool), flag, tid, pos, rng.choice([0, 10, 19, 20, 39, 40, 41, 60]), cigar, mtid, mpos, isize, seq))
    placed = sorted((a for a in recs if a.tid >= 0), key=lambda a: (a.tid, a.pos))
    return placed + [a for a in recs if a.tid < 0]


@pytest.mark.parametrize("seed", range(12))
def test_extract_host_logic_on_random_records(cli, tmp_path, seed):
    """Fuzz of the host half of `strling extract` against the oracle's restatement of extract.nim:20-248: records no aligner would
    write (see _fuzz_records), a genome-STR bed on odd seeds, p / min_mapq varied, small batches and several replay shards; the
    .bin must equal the oracle's byte for byte."""
    import random
    rng = random.Random(1000 + seed)
    targets = [("chr1", 200_000), ("chr2", 150_000), ("chrM", 20_000)]
    recs = _fuzz_records(rng, 2500, targets)
    hdr = bamio.sam_header(targets)
    bam, segs_path, res_path, out = (str(tmp_path / n) for n in ("f.bam", "segs.tsv", "res.bin", "f.bin"))
    bamio.write_bam(bam, hdr, targets, recs)
    p, q = [(0.8, 40), (0.7, 20), (0.9, 0), (0.85, 41)][seed % 4]
    extra, genome_str = [], None
    if seed % 2:
        bed = str(tmp_path / "ref.str")
        with open(bed, "w") as fh:
            for k in range(60):
                t = rng.randrange(len(targets))
                s = rng.randrange(0, targets[t][1] - 3000)
                fh.write(f"{targets[t][0]}\t{s}\t{s + rng.randrange(1, 2500)}\tCAG\n")
        extra = [bed]
        genome_str = eo.read_bed(bed)
    env = {"STRLING_DEBUG_THREADS": str(1 + seed % 5), "STRLING_DEBUG_SHARDS": str(1 + seed % 4)}
    batch = str([100000, 64, 257][seed % 3])
    run(cli, "debug", "extract", "dump", segs_path, bam, out, repr(p), str(q), batch, *extra, env=env)
    classes = [p, p - 0.07, min(p, 0.6)]
    lines = open(segs_path).read().splitlines()
    res = np.zeros(len(lines), dtype=[("unit", "S6"), ("repeat_count", "<u2")])
    for i, l in enumerate(lines):
        cls, _, seq = l.partition("\t")
        unit, count = orc.get_repeat(seq, classes[int(cls)])
        res["unit"][i], res["repeat_count"][i] = unit, count
    res.tofile(res_path)
    run(cli, "debug", "extract", "replay", res_path, bam, out, repr(p), str(q), batch, *extra, env=env)
    exp, cache, _ = eo.extract(recs, targets, hdr, p, q, genome_str)
    assert len(cache) > 20
    assert open(out, "rb").read() == exp


@pytest.mark.parametrize("threads", ["1", "3", "8"])
def test_thread_pool_contract(cli, threads):
    """host/pool.hpp: concurrent submitters with different priorities, every task exactly once, exceptions surface in their own job."""
    assert run(cli, "debug", "pool-selftest", threads).strip() == "ok"
