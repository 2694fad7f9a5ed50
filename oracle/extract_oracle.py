"""CPU ORACLE (test infrastructure, not product code): the order-dependent host logic of `strling extract`
(extract.nim:20-248,250-350), the `.bin` codec (cluster.nim:38-50, unpack.nim:36-133) and the merge / call
cluster drivers (merge.nim:91-187, call.nim:96-130,223-235,280-281), restated in pure Python over in-memory
alignment records for SMALL cases.  The per-read arithmetic goes through oracle/liboracle.so.

Un-vendored boundaries restated from their published behaviour (parity unpinned):
  hts-nim `aln.stop` = htslib bam_endpos (pos + reference length, or pos + 1 when unmapped / zero length);
  `aln.chrom` = "" for tid -1; lapper `find(start, stop)` = any interval with iv.start < stop and iv.stop > start;
  msgpack4nim: smallest-form ints, array[6,char] -> fixarray of six uint8, string -> fixstr/str8/str16.
"""
from __future__ import annotations

import struct

import msgpack
import numpy as np

from . import oracle as orc

STRLING_VERSION = "0.6.0"          # version.nim:1
FMT_VERSION = 0                    # version.nim:4


def aln_stop(a) -> int:
    rl = 0 if (a.flag & 4) else sum(n for op, n in a.cigar if op in "MDN=X")
    return a.pos + (rl if rl else 1)


class Lapper:
    def __init__(self, ivs):
        self.ivs = sorted(ivs)

    def find(self, start, stop) -> bool:
        return any(s < stop and e > start for s, e in self.ivs)


def read_bed(path):  # read_bed.nim:30-50
    by = {}
    for line in open(path):
        if line.startswith("track ") or line.startswith("#"):
            continue
        f = line.strip().split("\t", 5)
        if len(f) < 3:
            continue
        by.setdefault(f[0], []).append((int(f[1]), int(f[2])))
    return {c: Lapper(v) for c, v in by.items()}


class Opts:
    def __init__(self, median_fragment_length=0, proportion_repeat=0.8, min_mapq=40):
        self.median_fragment_length = median_fragment_length
        self.proportion_repeat = proportion_repeat
        self.min_mapq = min_mapq


def get_repeat_aln(a, targets, genome_str, opts):  # extract.nim:20-40 -> (unit, repeat_count, align_length)
    chrom = targets[a.tid][0] if a.tid >= 0 else ""
    if len(a.cigar) == 1 and a.cigar[0][0] == "M" and genome_str is not None and chrom in genome_str:
        if not genome_str[chrom].find(a.pos, aln_stop(a)):
            return b"", 0, a.cigar[0][1]
    unit, rc = orc.get_repeat(a.seq, opts.proportion_repeat)
    return unit, rc, len(a.seq)


def to_tread(a, targets, genome_str, opts):  # extract.nim:63-87
    unit, rc, al = get_repeat_aln(a, targets, genome_str, opts)
    assert rc < 256
    t = orc.make_tread(tid=a.tid, position=max(0, a.pos), repeat=unit, flag=a.flag, repeat_count=rc, align_length=al,
                       split=orc.NONE, mapq=a.mapq)
    L = len(a.cigar)
    if L > 1 and a.cigar[0][0] == "S" and a.cigar[0][1] > 16:
        t["split"] = orc.NONE_LEFT
    if L > 1 and a.cigar[L - 1][0] == "S" and a.cigar[L - 1][1] > 16:
        t["split"] = orc.NONE_RIGHT
    return t


class Cache:
    def __init__(self):
        self.tbl = {}        # qname -> tread (insertion-ordered dict; only membership / take are used)
        self.cache = []      # [(tread, qname)] in append order == .bin record order


def add_soft(cache, a, opts, read_repeat: bytes):  # extract.nim:93-132
    if a.mapq < opts.min_mapq:
        return
    if len(a.cigar) == 0 or (a.cigar[0][0] != "S" and a.cigar[-1][0] != "S"):
        return
    for cig_index in (0, len(a.cigar) - 1):
        op, ln = a.cigar[cig_index]
        if op != "S":
            continue
        if read_repeat == b"" and ln <= 16:
            continue
        soft = a.seq[0:ln] if cig_index == 0 else a.seq[len(a.seq) - ln:]
        unit, rc = orc.get_repeat(soft, opts.proportion_repeat)
        if rc == 0:
            continue
        position = max(0, a.pos) if cig_index == 0 else max(0, aln_stop(a))
        tr = orc.make_tread(tid=a.tid, position=position, flag=a.flag, repeat=unit, repeat_count=rc, align_length=len(soft),
                            split=orc.LEFT if cig_index == 0 else orc.RIGHT, mapq=a.mapq)
        if orc.p_repeat(tr) < 0.9:
            continue
        cache.cache.append((tr, a.qname))


def cache_add(cache, a, targets, genome_str, opts):  # extract.nim:192-248
    assert not (a.flag & 0x100 or a.flag & 0x800)
    after_mate = a.tid > a.mate_tid or (a.tid == a.mate_tid and (a.pos > a.mate_pos or (a.pos == a.mate_pos and a.qname in cache.tbl)))
    if after_mate:
        if a.qname not in cache.tbl:
            return
        mate, mate_q = cache.tbl.pop(a.qname)
        me = to_tread(a, targets, genome_str, opts)
        b = opts.proportion_repeat
        opts.proportion_repeat = min(b, 0.6)
        add_soft(cache, a, opts, bytes(me["repeat"][0]).rstrip(b"\0"))
        opts.proportion_repeat = b
        if mate["repeat_count"][0] == 0 and me["repeat_count"][0] == 0:
            return
        if orc.unplaced_pair(me, mate, opts.proportion_repeat, opts.min_mapq):
            if bytes(me["repeat"][0]).rstrip(b"\0") == b"" or bytes(mate["repeat"][0]).rstrip(b"\0") == b"":
                return
            for t in (me, mate):
                t["repeat"] = orc.canonical_repeat(bytes(t["repeat"][0]))
                t["position"] = 0
                t["tid"] = -1
            cache.cache.append((me, a.qname))
            cache.cache.append((mate, mate_q))
            return
        mp = int(mate["position"][0])
        if orc.adjust_by(mate, me, opts.proportion_repeat, opts.min_mapq, opts.median_fragment_length, int(me["position"][0])):
            cache.cache.append((mate, mate_q))
        if orc.adjust_by(me, mate, opts.proportion_repeat, opts.min_mapq, opts.median_fragment_length, mp):
            cache.cache.append((me, a.qname))
    else:
        tr = to_tread(a, targets, genome_str, opts)
        b = opts.proportion_repeat
        opts.proportion_repeat -= 0.07
        add_soft(cache, a, opts, bytes(tr["repeat"][0]).rstrip(b"\0"))
        opts.proportion_repeat = b
        if a.qname in cache.tbl:     # hasKeyOrPut: key existed -> warn and drop it (extract.nim:245-248)
            cache.tbl.pop(a.qname)
        else:
            cache.tbl[a.qname] = (tr, a.qname)


def fragment_length_distribution(records, n_reads=2_000_000, skip_reads=100_000):  # utils.nim:86-111
    res = np.zeros(4096, dtype=np.uint32)
    i = -1
    counted = 0
    skipped = []
    for a in records:
        i += 1
        if not (a.flag & 2):
            continue
        if a.flag & 0x800 or a.flag & 0x100:
            continue
        if a.isize < 0 or a.isize > 4095:
            continue
        if i < skip_reads:
            skipped.append(a)
            continue
        else:
            skipped = []
        res[a.isize] += 1
        counted += 1
        if counted > n_reads:
            break
    if res.sum() == 0:
        for a in skipped:
            if not (a.flag & 2):
                continue
            if a.isize < 0 or a.isize > 4095:
                continue
            res[a.isize] += 1
    return res


def pack_tread(t, qname: str) -> bytes:  # cluster.nim:38-50 with msgpack4nim's encodings
    q = qname.encode()
    rep = bytes(t["repeat"][0]).ljust(6, b"\0")
    return b"".join([
        msgpack.packb(int(t["tid"][0])), msgpack.packb(int(t["position"][0])), b"\x96" + b"".join(msgpack.packb(c) for c in rep),
        msgpack.packb(int(t["flag"][0])), msgpack.packb(int(t["split"][0])), msgpack.packb(int(t["mapq"][0])),
        msgpack.packb(int(t["repeat_count"][0])), msgpack.packb(int(t["align_length"][0])), msgpack.packb(len(q)),
        msgpack.packb(qname),
    ])


def extract(records, targets, header_text: str, proportion_repeat=0.8, min_mapq=40, genome_str=None, skip_reads=100_000):
    """extract_main (extract.nim:250-350) -> (.bin bytes, [(tread, qname)], frag_dist)."""
    frag_dist = fragment_length_distribution(records, skip_reads=skip_reads)
    opts = Opts(orc.median(frag_dist), proportion_repeat, min_mapq)
    cache = Cache()
    for a in records:                                   # pass 2: every record, file order
        if a.flag & 0x100 or a.flag & 0x800:
            continue
        cache_add(cache, a, targets, genome_str, opts)
    for a in records:                                   # ibam.query("*"): the no-coordinate tail AGAIN
        if a.tid >= 0:
            continue
        if a.flag & 0x100 or a.flag & 0x800:
            continue
        cache_add(cache, a, targets, genome_str, opts)
    out = bytearray()
    out += b"STR" + struct.pack("<h", FMT_VERSION) + STRLING_VERSION.encode().ljust(9, b"\0")
    out += struct.pack("<f", proportion_repeat) + struct.pack("<B", min_mapq) + frag_dist.astype("<u4").tobytes()
    h = header_text.encode()
    out += struct.pack("<i", len(h)) + h + struct.pack("<i", len(cache.cache))
    for t, q in cache.cache:
        out += pack_tread(t, q)
    return bytes(out), cache.cache, frag_dist


def unpack_bin(data: bytes):  # unpack.nim:58-133 -> dict(p, min_mapq, frag_dist, header, treads (TREAD_DTYPE), qnames)
    assert data[:3] == b"STR"
    (fmt,) = struct.unpack_from("<h", data, 3)
    assert fmt == FMT_VERSION
    (p,) = struct.unpack_from("<f", data, 14)
    min_mapq = data[18]
    frag = np.frombuffer(data, dtype="<u4", count=4096, offset=19).copy()
    off = 19 + 16384
    (hl,) = struct.unpack_from("<i", data, off)
    header = data[off + 4: off + 4 + hl].decode()
    off += 4 + hl
    (n,) = struct.unpack_from("<i", data, off)
    off += 4
    up = msgpack.Unpacker(raw=True)
    up.feed(data[off:])
    vals = list(up)
    assert len(vals) == 10 * n, (len(vals), n)
    treads = np.zeros(n, dtype=orc.TREAD_DTYPE)
    qnames = []
    for i in range(n):
        v = vals[10 * i: 10 * i + 10]
        treads[i]["tid"], treads[i]["position"] = v[0], v[1]
        treads[i]["repeat"] = bytes(v[2]).rstrip(b"\0")
        treads[i]["flag"], treads[i]["split"], treads[i]["mapq"], treads[i]["repeat_count"], treads[i]["align_length"] = v[3:8]
        assert v[8] == len(v[9])
        qnames.append(v[9].decode())
    return dict(p=p, min_mapq=min_mapq, frag_dist=frag, header=header, treads=treads, qnames=qnames)


def targets_from_header(header: str):
    t = []
    for line in header.splitlines():
        if line.startswith("@SQ"):
            f = dict(x.split(":", 1) for x in line.split("\t")[1:] if ":" in x)
            t.append((f["SN"], int(f["LN"])))
    return t


def bounds_line(b, targets) -> str:  # cluster.nim:262-266 (name is empty for discovered clusters)
    rep = bytes(b["repeat"]).rstrip(b"\0").decode()
    return (f"{targets[int(b['tid'])][0]}\t{b['left']}\t{b['right']}\t{rep}\t\t{b['left_most']}\t{b['right_most']}\t"
            f"{b['center_mass']}\t{b['n_left']}\t{b['n_right']}\t{b['n_total']}")


BOUNDS_HEADER = "#chrom\tleft\tright\trepeat\tname\tleft_most\tright_most\tcenter_mass\tn_left\tn_right\tn_total"


def parse_bed_loci(bed_lines, targets, window):
    """parse_bed (cluster.nim:111-141) -> (LOCUS_DTYPE array, [(left, right, unit, name)])."""
    names = [t[0] for t in targets]
    loci = np.zeros(len(bed_lines), dtype=orc.LOCUS_DTYPE)
    meta = []
    for i, l in enumerate(bed_lines):
        f = l.split()
        assert len(f) in (4, 5)
        tid = names.index(f[0])
        left, right = int(f[1]), int(f[2])
        loci[i]["tid"] = tid
        loci[i]["repeat"] = f[3].encode()
        loci[i]["left_most"] = max(left - window, 0)
        loci[i]["right_most"] = min(right + window, targets[tid][1])
        meta.append((left, right, f[3], f[4] if len(f) == 5 else ""))
    return loci, meta


def locus_line(locus, meta, targets) -> str:
    left, right, unit, name = meta
    return (f"{targets[int(locus['tid'])][0]}\t{left}\t{right}\t{unit}\t{name}\t{locus['left_most']}\t{locus['right_most']}\t0\t"
            f"{locus['n_left']}\t{locus['n_right']}\t{locus['n_total']}")


def merge(bins, window=-1, min_support=5, min_clip=0, min_clip_total=0, bed_lines=None):
    """merge_main without --chromosome (merge.nim:91-187) -> list of bounds lines (loci first, as merge.nim:166-168)."""
    frag = np.zeros(4096, dtype=np.uint64)
    parts = []
    targets = None
    for si, data in enumerate(bins):
        u = unpack_bin(data)
        if targets is None:
            targets = targets_from_header(u["header"])
        frag += u["frag_dist"]
        t = u["treads"][u["treads"]["tid"] >= 0].copy()   # drop_unplaced=true
        t["sample"] = si
        parts.append(t)
    assert frag.max() < 2 ** 32
    frag = frag.astype(np.uint32)
    treads = np.concatenate(parts)
    if window < 0:
        window = orc.median(frag, 0.98)
    mcd = int(0.5 * float(orc.median(frag, 0.5))) & 0xFFFF
    if bed_lines:
        loci, meta = parse_bed_loci(bed_lines, targets, window)
        loci, b, _ = orc.cluster_all_loci(treads, loci, window, min_support, min_clip, min_clip_total, mcd, merge_mode=True)
        return [locus_line(l, m, targets) for l, m in zip(loci, meta)] + [bounds_line(x, targets) for x in b], targets
    b, _ = orc.cluster_all(treads, window, min_support, min_clip, min_clip_total, mcd, merge_mode=True)
    return [bounds_line(x, targets) for x in b], targets


def call_clusters(data: bytes, frag_dist=None, min_support=5, min_clip=0, min_clip_total=0):
    """The cluster loop of call_main (call.nim:114-130,223-235,280-281) -> (bounds lines, unplaced {unit: n})."""
    u = unpack_bin(data)
    targets = targets_from_header(u["header"])
    frag = u["frag_dist"] if frag_dist is None else frag_dist
    window = orc.median(frag, 0.99)
    mcd = int(0.5 * float(orc.median(frag, 0.5))) & 0xFFFF
    b, unplaced = orc.cluster_all(u["treads"], window, min_support, min_clip, min_clip_total, mcd, merge_mode=False)
    return [bounds_line(x, targets) for x in b], unplaced, targets


# ------------------------------------------------------------------------------------------------ strling index
def _trim(start, stop, repeat: str, dna: str):  # genome_strs.nim:22-59
    assert len(dna) == stop - start
    k = len(repeat)
    expected = int(orc.slide_by(repeat, k)[0])
    for enc in orc.slide_by(dna, k):
        if int(enc) != expected:
            start += k
        else:
            break
    assert start < stop, "repeat not found in expected region"
    expected = int(orc.slide_by(repeat[::-1], k)[0])
    for enc in orc.slide_by(dna[::-1], k):
        if int(enc) != expected:
            stop -= k
        else:
            break
    assert start < stop, "repeat not found in expected region"
    return start, stop


def genome_repeat_lines(chroms, proportion_repeat=0.8, window_size=100, step=60):
    """repeat_windows + the bed lines genome_repeats writes (genome_strs.nim:61-92,121-125).  chroms: [(name, seq)]."""
    lines = []
    for name, seq in chroms:
        seq = seq.upper()
        L = len(seq)
        last = None  # [start, stop, repeat]

        def flush(w):
            if w is not None and w[1] - w[0] >= (window_size - step):
                a = max(0, w[0] - window_size)
                b = min(w[1] + window_size, L)
                s2, e2 = _trim(a, b, w[2], seq[a:b])
                lines.append(f"{name}\t{s2}\t{e2}\t{w[2]}")

        start = 0
        while start < L:
            dna = seq[start:min(L, start + window_size)]
            unit, rc = orc.get_repeat(dna, proportion_repeat)
            if rc > 0:
                w = [start, start + len(dna), unit.decode()]
                if last is None or last[2] != w[2] or w[0] > last[1] + (window_size - step):
                    flush(last)
                    last = w
                else:
                    last[1] = w[1]
            start += step
        flush(last)
    return lines
