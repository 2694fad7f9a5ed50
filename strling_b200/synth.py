"""Seeded synthetic inputs for the hot path: 150-bp reads with the config-2 class mix, packed
into the library's 2-bit layout with numpy (no per-read Python loops), plus the tread generator used by the
cluster tests.  Shared by tests/ and bench.py; contains no reference algorithm."""
from __future__ import annotations

import numpy as np

from .binding import SEG_HAS_N, SEGMENT_DTYPE, Masks

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)
# ASCII -> 2-bit code of the library (C=0 A=1 T=2 G=3, everything else 1) and the non-ACGT flag
_CODE = np.ones(256, dtype=np.uint8)
_OTHER = np.ones(256, dtype=np.uint8)
for _j, _c in enumerate(b"CATG"):
    _CODE[_c] = _j
    _OTHER[_c] = 0
    _CODE[_c + 32] = _j
    _OTHER[_c + 32] = 0

CLASS_PLAIN, CLASS_MESSY, CLASS_CLIPPED_STR, CLASS_STR = 0, 1, 2, 3


def make_reads(n: int, seed: int, length: int = 150, mix=(0.90, 0.07, 0.02, 0.01), noise: float = 0.01,
               n_frac: float = 0.0, iupac_frac: float = 0.0):
    """Returns (ascii uint8 [n, length], cls uint8 [n], lclip uint8 [n], rclip uint8 [n]).
    plain/messy: uniform ACGT.  clipped-STR: a soft clip of 17..60 bases on one end filled with a repeat.
    STR: >= 80 % of the read is a repeat of a random 1-6 bp unit, random phase, `noise` substitutions.
    n_frac: fraction of reads that get 1..30 'N' bases (0 for the benchmark; used by parity tests).
    iupac_frac: fraction of reads that get 1..40 IUPAC ambiguity codes other than N (they scan as 'A', never match in the
    recount and -- unlike N -- do not count towards the N > 20 gate of utils.nim:238)."""
    rng = np.random.default_rng(seed)
    reads = _ACGT[rng.integers(0, 4, size=(n, length), dtype=np.uint8)]
    cls = rng.choice(4, size=n, p=np.asarray(mix) / np.sum(mix)).astype(np.uint8)
    lclip = np.zeros(n, dtype=np.uint8)
    rclip = np.zeros(n, dtype=np.uint8)
    pos = np.arange(length)

    def repeat_fill(rows, start, stop):
        # rows: indices; start/stop: per-row [start, stop) filled with unit repeats
        m = len(rows)
        if m == 0:
            return
        k = rng.integers(1, 7, size=m)
        units = _ACGT[rng.integers(0, 4, size=(m, 6), dtype=np.uint8)]
        phase = rng.integers(0, 6, size=m)
        idx = (pos[None, :] + phase[:, None]) % k[:, None]
        rep = np.take_along_axis(units, idx, axis=1)
        sub = rng.random((m, length)) < noise
        rep = np.where(sub, _ACGT[rng.integers(0, 4, size=(m, length), dtype=np.uint8)], rep)
        inside = (pos[None, :] >= start[:, None]) & (pos[None, :] < stop[:, None])
        reads[rows] = np.where(inside, rep, reads[rows])

    rows = np.nonzero(cls == CLASS_STR)[0]
    span = rng.integers(int(0.8 * length), length + 1, size=len(rows))
    st = rng.integers(0, length - span + 1)
    repeat_fill(rows, st, st + span)

    rows = np.nonzero(cls == CLASS_CLIPPED_STR)[0]
    clip = rng.integers(17, 61, size=len(rows))
    left = rng.random(len(rows)) < 0.5
    lclip[rows] = np.where(left, clip, 0)
    rclip[rows] = np.where(left, 0, clip)
    repeat_fill(rows, np.where(left, 0, length - clip), np.where(left, clip, length))

    # messy reads: a short (<= 16) soft clip, never scanned on its own unless the read has a unit
    rows = np.nonzero(cls == CLASS_MESSY)[0]
    short = rng.integers(0, 17, size=len(rows)).astype(np.uint8)
    lclip[rows] = np.where(rng.random(len(rows)) < 0.5, short, 0)

    if n_frac > 0:
        rows = np.nonzero(rng.random(n) < n_frac)[0]
        for r in rows:  # small counts only (tests)
            k = int(rng.integers(1, 31))
            reads[r, rng.choice(length, size=k, replace=False)] = ord("N")
    if iupac_frac > 0:
        codes = np.frombuffer(b"RYMKSWHBVD", dtype=np.uint8)
        rows = np.nonzero(rng.random(n) < iupac_frac)[0]
        for r in rows:
            k = int(rng.integers(1, min(41, length + 1)))
            where = rng.choice(length, size=k, replace=False)
            reads[r, where] = codes[rng.integers(0, len(codes), size=k)]
    return reads, cls, lclip, rclip


def pack_matrix(reads: np.ndarray, align_bases: int = 16):
    """Packs an [n, L] ASCII matrix: every read starts at a multiple of align_bases.  Returns
    (seq2 uint8 with 8 bytes slack, nmask uint32 (+slack) or None, stride_bases); when the reads hold non-ACGT bases other than
    the literal 'N', nmask is a binding.Masks pair (nmask, xmask)."""
    n, length = reads.shape
    stride = (length + align_bases - 1) // align_bases * align_bases
    codes = np.zeros((n, stride), dtype=np.uint8)
    codes[:, :length] = _CODE[reads]
    c4 = codes.reshape(n, stride // 4, 4)
    packed = (c4[:, :, 0] << 6) | (c4[:, :, 1] << 4) | (c4[:, :, 2] << 2) | c4[:, :, 3]
    seq2 = np.zeros(n * stride // 4 + 8, dtype=np.uint8)
    seq2[: n * stride // 4] = packed.reshape(-1)
    other = np.zeros((n, stride), dtype=np.uint8)
    other[:, :length] = _OTHER[reads]
    nmask = None
    if other.any():
        bits = np.packbits(other.reshape(-1), bitorder="little")
        nmask = np.zeros((n * stride + 31) // 32 + 2, dtype=np.uint32)
        nb = np.zeros(nmask.size * 4, dtype=np.uint8)
        nb[: bits.size] = bits
        nmask[:] = nb.view("<u4")
        xother = other.copy()
        xother[:, :length] &= (reads != ord("N")).astype(np.uint8)
        if xother.any():
            bits = np.packbits(xother.reshape(-1), bitorder="little")
            xb = np.zeros(nmask.size * 4, dtype=np.uint8)
            xb[: bits.size] = bits
            nmask = Masks(nmask, xb.view("<u4").copy())
    return seq2, nmask, stride


def segments_for(reads: np.ndarray, lclip: np.ndarray, rclip: np.ndarray, stride: int, clip_min: int = 17,
                 pclass_read: int = 0, pclass_clip: int = 1):
    """Whole-read segment for every read + one segment per soft clip longer than 16 (the unconditional part of
    extract.nim:93-114).  Returns (segments, owner read index per segment)."""
    n, length = reads.shape
    has_n = (_OTHER[reads].any(axis=1)).astype(np.uint8) * SEG_HAS_N
    base = np.arange(n, dtype=np.int64) * stride
    parts = []
    full = np.zeros(n, dtype=SEGMENT_DTYPE)
    full["base_off"], full["len"], full["pclass"], full["flags"] = base, length, pclass_read, has_n
    parts.append((full, np.arange(n)))
    for clip, is_left in ((lclip, True), (rclip, False)):
        rows = np.nonzero(clip >= clip_min)[0]
        seg = np.zeros(len(rows), dtype=SEGMENT_DTYPE)
        seg["base_off"] = base[rows] + (0 if is_left else (length - clip[rows].astype(np.int64)))
        seg["len"], seg["pclass"], seg["flags"] = clip[rows], pclass_clip, has_n[rows]
        parts.append((seg, rows))
    segs = np.concatenate([p[0] for p in parts])
    owner = np.concatenate([p[1] for p in parts])
    return segs, owner


def segment_ascii(reads: np.ndarray, segs: np.ndarray, stride: int):
    """ASCII view of each segment (for the oracle): returns (flat uint8 buffer, offsets uint64, lens uint32)."""
    n, length = reads.shape
    flat = reads.reshape(-1)
    row = segs["base_off"].astype(np.int64) // stride
    col = segs["base_off"].astype(np.int64) % stride
    off = (row * length + col).astype(np.uint64)
    return flat, off, segs["len"].astype(np.uint32)


# ------------------------------------------------------------------------------------------------ cluster inputs
TREAD_DTYPE = np.dtype(
    [("tid", "<i4"), ("position", "<u4"), ("repeat", "S6"), ("flag", "<u2"), ("split", "u1"),
     ("mapq", "u1"), ("repeat_count", "u1"), ("align_length", "u1"), ("sample", "<i4")]
)
SOFT_LEFT, SOFT_RIGHT, SOFT_BOTH, SOFT_NONE, SOFT_NONE_RIGHT, SOFT_NONE_LEFT = range(6)


def make_treads(n_loci: int, seed: int, n_tids: int = 24, n_samples: int = 1, noise_reads: int = 0, unplaced: int = 0,
                contig_len: int = 100_000_000, dense: bool = False):
    """STR-read records (`.bin` order, i.e. unsorted) shaped like extract output: per locus a cloud of anchored
    reads (split none) within a fragment length of the locus, left-clipped reads piling up at the left edge and
    right-clipped reads at the right edge (with stutter so that clip-position ties occur), plus background reads
    and unplaced (tid -1) reads.  dense=True packs loci so that clusters chain, trim and split."""
    rng = np.random.default_rng(seed)
    units = [b"A", b"AC", b"AG", b"CAG", b"CCG", b"AAG", b"AAAG", b"AAGG", b"ATTCT", b"AAAAG", b"CACGAT", b"AAAAAT", b"T", b"CTG"]
    parts = []
    for _ in range(n_loci):
        tid = int(rng.integers(0, n_tids))
        unit = units[int(rng.integers(0, len(units)))]
        left = int(rng.integers(1000, 3000 if dense else contig_len))
        width = int(rng.integers(1, 120))
        right = left + width
        n_anchor = int(rng.integers(0, 30))
        n_left = int(rng.integers(0, 12))
        n_right = int(rng.integers(0, 12))
        m = n_anchor + n_left + n_right
        if m == 0:
            continue
        t = np.zeros(m, dtype=TREAD_DTYPE)
        t["tid"] = tid
        t["repeat"] = unit
        pos = np.concatenate([
            left + rng.integers(-450, 450, size=n_anchor),
            left + rng.choice([0, 0, 0, 1, -1, 2, 7], size=n_left),
            right + rng.choice([0, 0, 0, 1, -1, -2, 5], size=n_right),
        ])
        t["position"] = np.maximum(pos, 0).astype(np.uint32)
        t["split"] = np.concatenate([
            rng.choice([SOFT_NONE, SOFT_NONE, SOFT_NONE, SOFT_NONE_LEFT, SOFT_NONE_RIGHT], size=n_anchor),
            np.full(n_left, SOFT_LEFT), np.full(n_right, SOFT_RIGHT)]).astype(np.uint8)
        t["sample"] = rng.integers(0, n_samples, size=m)
        parts.append(t)
    if noise_reads:
        t = np.zeros(noise_reads, dtype=TREAD_DTYPE)
        t["tid"] = rng.integers(0, n_tids, size=noise_reads)
        t["repeat"] = rng.choice(np.array(units, dtype="S6"), size=noise_reads)
        t["position"] = rng.integers(0, 5000 if dense else contig_len, size=noise_reads)
        t["split"] = rng.choice([SOFT_NONE, SOFT_LEFT, SOFT_RIGHT, SOFT_NONE_LEFT], size=noise_reads, p=[0.6, 0.15, 0.15, 0.1])
        t["sample"] = rng.integers(0, n_samples, size=noise_reads)
        parts.append(t)
    if unplaced:
        t = np.zeros(unplaced, dtype=TREAD_DTYPE)
        t["tid"] = -1
        t["repeat"] = rng.choice(np.array([b"AAGGG", b"AC", b"AGC", b"A"], dtype="S6"), size=unplaced)
        t["split"] = SOFT_NONE
        t["sample"] = rng.integers(0, n_samples, size=unplaced)
        parts.append(t)
    out = np.concatenate(parts) if parts else np.zeros(0, dtype=TREAD_DTYPE)
    out["flag"] = rng.integers(0, 4096, size=len(out))
    out["mapq"] = rng.integers(0, 61, size=len(out))
    out["repeat_count"] = rng.integers(1, 75, size=len(out))
    out["align_length"] = 150
    return out[rng.permutation(len(out))]
