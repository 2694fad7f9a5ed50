"""Minimal BAM / BGZF writer and the synthetic-alignment generator used by the tests and benchmarks (there is no
samtools / htslib / bwa in this image).  Files follow the SAM spec (BGZF blocks <= 64 KiB, EOF marker block) so the
C++ host reader -- and htslib -- can read them.  Contains no reference algorithm."""
from __future__ import annotations

import struct
import zlib
from dataclasses import dataclass, field

import numpy as np

CIGAR_OPS = "MIDNSHP=X"
SEQ_NIBBLE = {c: i for i, c in enumerate("=ACMGRSVTWYHKDBN")}
_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")


@dataclass
class Aln:
    qname: str
    flag: int
    tid: int
    pos: int            # 0-based; -1 when unplaced
    mapq: int
    cigar: list         # [(op char, length)]
    mate_tid: int
    mate_pos: int
    isize: int
    seq: str
    extra: dict = field(default_factory=dict)


def _reg2bin(beg: int, end: int) -> int:
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def ref_len(cigar) -> int:
    return sum(n for op, n in cigar if op in "MDN=X")


def encode_record(a: Aln) -> bytes:
    name = a.qname.encode() + b"\0"
    l_seq = len(a.seq)
    rl = ref_len(a.cigar) if not (a.flag & 4) else 0
    end = a.pos + (rl if rl > 0 else 1)
    binv = _reg2bin(max(a.pos, 0), max(end, 1)) if a.pos >= 0 else 4680
    cig = b"".join(struct.pack("<I", (n << 4) | CIGAR_OPS.index(op)) for op, n in a.cigar)
    nib = [SEQ_NIBBLE.get(c, 15) for c in a.seq]
    if l_seq & 1:
        nib.append(0)
    seq = bytes((nib[i] << 4) | nib[i + 1] for i in range(0, len(nib), 2))
    qual = a.extra.get("qual", b"\xff" * l_seq)   # tests plant byte patterns here
    assert len(qual) == l_seq
    body = struct.pack("<iiBBHHHiiii", a.tid, a.pos, len(name), a.mapq, binv, len(a.cigar), a.flag, l_seq, a.mate_tid,
                       a.mate_pos, a.isize) + name + cig + seq + qual
    return struct.pack("<i", len(body)) + body


def _bgzf_block(data: bytes, level: int = 1) -> bytes:
    co = zlib.compressobj(level, zlib.DEFLATED, -15)
    comp = co.compress(data) + co.flush()
    bsize = len(comp) + 25
    hdr = struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize)
    return hdr + comp + struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data) & 0xFFFFFFFF)


def write_bam(path: str, header_text: str, targets, records, level: int = 1) -> None:
    """targets: [(name, length)]; records: iterable of Aln (written in the given order)."""
    buf = bytearray()
    ht = header_text.encode()
    buf += b"BAM\x01" + struct.pack("<i", len(ht)) + ht + struct.pack("<i", len(targets))
    for name, ln in targets:
        nm = name.encode() + b"\0"
        buf += struct.pack("<i", len(nm)) + nm + struct.pack("<i", ln)
    with open(path, "wb") as fh:
        def flush(final=False):
            nonlocal buf
            while len(buf) >= 0xFF00 or (final and len(buf)):
                fh.write(_bgzf_block(bytes(buf[:0xFF00]), level))
                buf = buf[0xFF00:]
        flush()
        for a in records:
            buf += encode_record(a)
            if len(buf) >= 0xFF00:
                flush()
        flush(final=True)
        fh.write(_EOF)


def sam_header(targets, sort_order: str = "coordinate") -> str:
    return f"@HD\tVN:1.6\tSO:{sort_order}\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in targets)


# ------------------------------------------------------------------------------------------------ synthetic alignments
def _rand_seq(rng, n):
    return "".join("ACGT"[i] for i in rng.integers(0, 4, size=n))


def _unit_seq(unit: str, n: int, phase: int = 0) -> str:
    s = unit * (n // len(unit) + 3)
    return s[phase: phase + n]


def _revcomp(s: str) -> str:
    return s[::-1].translate(str.maketrans("ACGTN", "TGCAN"))


def simulate_alignments(seed: int, n_pairs: int, targets, loci, read_len: int = 150, insert_mean: float = 400, insert_sd: float = 60,
                        str_pair_frac: float = 0.3, unmapped_pairs: int = 0, n_frac: float = 0.0, name_prefix: str = ""):
    """Synthetic stand-in for `bwa mem` output around STR loci.
    loci: [(tid, start, stop, unit)].  Returns a coordinate-sorted list of Aln (unplaced pairs last).
    Pair kinds: background (both mates random sequence, 150M, proper pair), spanning-clip (one mate soft-clipped
    into the repeat at a locus edge), STR mate (one mate entirely repeat, mapq 0 / mismapped elsewhere, the other
    anchored near the locus), and fully unmapped STR pairs."""
    rng = np.random.default_rng(seed)
    out = []
    L = read_len

    def add_pair(name, r1, r2):
        # r = dict(tid,pos,mapq,cigar,seq,rev,unmapped)
        for me, mate, first in ((r1, r2, True), (r2, r1, False)):
            flag = 1 | (0x40 if first else 0x80)
            if me.get("unmapped"):
                flag |= 0x4
            if mate.get("unmapped"):
                flag |= 0x8
            if me.get("rev"):
                flag |= 0x10
            if mate.get("rev"):
                flag |= 0x20
            proper = (not me.get("unmapped") and not mate.get("unmapped") and me["tid"] == mate["tid"]
                      and abs(me["pos"] - mate["pos"]) < 1000 and me["mapq"] > 0 and mate["mapq"] > 0)
            if proper:
                flag |= 0x2
            isize = 0
            if proper:
                lo = min(me["pos"], mate["pos"])
                hi = max(me["pos"] + ref_len(me["cigar"]), mate["pos"] + ref_len(mate["cigar"]))
                isize = (hi - lo) if me["pos"] <= mate["pos"] else -(hi - lo)
                if me["pos"] == mate["pos"]:
                    isize = (hi - lo) if first else -(hi - lo)
            out.append(Aln(name_prefix + name, flag, me["tid"], me["pos"], me["mapq"], me["cigar"], mate["tid"], mate["pos"], isize, me["seq"]))

    def frag():
        return int(np.clip(rng.normal(insert_mean, insert_sd), L + 10, 4000))

    n_str = int(n_pairs * str_pair_frac) if loci else 0
    for i in range(n_pairs - n_str):
        tid = int(rng.integers(0, len(targets)))
        f = frag()
        pos = int(rng.integers(0, targets[tid][1] - f - 1))
        kind = rng.random()
        cig1 = [("M", L)]
        if kind < 0.07:  # messy: an indel or a short clip
            cig1 = [("M", 70), ("I", 2), ("M", L - 72)] if kind < 0.035 else [("S", int(rng.integers(1, 17))), ("M", 0)]
            if cig1[0][0] == "S":
                cig1[1] = ("M", L - cig1[0][1])
        add_pair(f"bg{i}", dict(tid=tid, pos=pos, mapq=60, cigar=cig1, seq=_rand_seq(rng, L), rev=False),
                 dict(tid=tid, pos=pos + f - L, mapq=60, cigar=[("M", L)], seq=_rand_seq(rng, L), rev=True))
    for i in range(n_str):
        tid, start, stop, unit = loci[int(rng.integers(0, len(loci)))]
        kind = rng.random()
        f = frag()
        phase = int(rng.integers(0, len(unit)))
        if kind < 0.35:      # left-anchored read running into the repeat: right soft clip
            clip = int(rng.integers(17, 90))
            pos = start - (L - clip)
            seq = _rand_seq(rng, L - clip) + _unit_seq(unit, clip, 0)
            add_pair(f"rc{i}", dict(tid=tid, pos=pos, mapq=int(rng.choice([60, 60, 30, 10])), cigar=[("M", L - clip), ("S", clip)], seq=seq, rev=False),
                     dict(tid=tid, pos=max(0, pos - f + L), mapq=60, cigar=[("M", L)], seq=_rand_seq(rng, L), rev=True))
        elif kind < 0.7:     # right-anchored read: left soft clip
            clip = int(rng.integers(17, 90))
            pos = stop
            seq = _unit_seq(unit, clip, phase) + _rand_seq(rng, L - clip)
            add_pair(f"lc{i}", dict(tid=tid, pos=pos, mapq=int(rng.choice([60, 60, 30, 10])), cigar=[("S", clip), ("M", L - clip)], seq=seq, rev=True),
                     dict(tid=tid, pos=pos + f - L, mapq=60, cigar=[("M", L)], seq=_rand_seq(rng, L), rev=False))
        elif kind < 0.9:     # STR mate mismapped to another place with low mapq, anchor near the locus
            anchor_left = rng.random() < 0.5
            apos = start - f + int(rng.integers(0, 60)) if anchor_left else stop + f - L - int(rng.integers(0, 60))
            apos = max(0, apos)
            otid = int(rng.integers(0, len(targets)))
            opos = int(rng.integers(0, targets[otid][1] - L - 1))
            sseq = _unit_seq(unit, L, phase)
            if rng.random() < 0.5:
                sseq = _revcomp(sseq)
            add_pair(f"sm{i}", dict(tid=tid, pos=apos, mapq=60, cigar=[("M", L)], seq=_rand_seq(rng, L), rev=not anchor_left),
                     dict(tid=otid, pos=opos, mapq=int(rng.choice([0, 0, 3, 25])), cigar=[("M", L)] if rng.random() < 0.6 else [("S", 30), ("M", L - 30)],
                          seq=sseq, rev=anchor_left))
        else:                # both mates STR, placed but mapq 0
            otid = int(rng.integers(0, len(targets)))
            opos = int(rng.integers(0, targets[otid][1] - 2 * L - 1))
            add_pair(f"ss{i}", dict(tid=otid, pos=opos, mapq=0, cigar=[("M", L)], seq=_unit_seq(unit, L, phase), rev=False),
                     dict(tid=otid, pos=opos + 40, mapq=0, cigar=[("M", L)], seq=_revcomp(_unit_seq(unit, L, 0)), rev=True))
    for i in range(unmapped_pairs):
        unit = loci[int(rng.integers(0, len(loci)))][3] if loci else "CAG"
        s1 = _unit_seq(unit, L, int(rng.integers(0, len(unit)))) if rng.random() < 0.7 else _rand_seq(rng, L)
        s2 = _revcomp(_unit_seq(unit, L, 0)) if rng.random() < 0.7 else _rand_seq(rng, L)
        add_pair(f"un{i}", dict(tid=-1, pos=-1, mapq=0, cigar=[], seq=s1, rev=False, unmapped=True),
                 dict(tid=-1, pos=-1, mapq=0, cigar=[], seq=s2, rev=False, unmapped=True))
    if n_frac > 0:
        for a in out:
            if rng.random() < n_frac:
                s = list(a.seq)
                for j in rng.choice(len(s), size=int(rng.integers(1, 26)), replace=False):
                    s[j] = "N"
                a.seq = "".join(s)
    placed = [a for a in out if a.tid >= 0]
    unplaced = [a for a in out if a.tid < 0]
    placed.sort(key=lambda a: (a.tid, a.pos))
    return placed + unplaced
