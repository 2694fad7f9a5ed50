"""Seeded synthetic inputs for the hot path: 150-bp reads with the config-2 class mix, packed
into the library's 2-bit layout with numpy (no per-read Python loops), plus the tread generator used by the
cluster tests.  Shared by tests/ and bench.py; contains no reference algorithm."""
from __future__ import annotations

import numpy as np

from .binding import SEG_HAS_N, SEGMENT_DTYPE

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
               n_frac: float = 0.0):
    """Returns (ascii uint8 [n, length], cls uint8 [n], lclip uint8 [n], rclip uint8 [n]).
    plain/messy: uniform ACGT.  clipped-STR: a soft clip of 17..60 bases on one end filled with a repeat.
    STR: >= 80 % of the read is a repeat of a random 1-6 bp unit, random phase, `noise` substitutions.
    n_frac: fraction of reads that get 1..30 'N' bases (0 for the benchmark; used by parity tests)."""
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
    return reads, cls, lclip, rclip


def pack_matrix(reads: np.ndarray, align_bases: int = 16):
    """Packs an [n, L] ASCII matrix: every read starts at a multiple of align_bases.  Returns
    (seq2 uint8 with 8 bytes slack, nmask uint32 (+slack) or None, stride_bases)."""
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
