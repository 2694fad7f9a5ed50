"""ctypes front-end of the CPU ORACLE (test infrastructure, not product code).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  It wraps oracle/liboracle.so (plain-C restatement of the reference hot path,
see strling_oracle.h for parity status) and adds the order-dependent host logic of
extract.nim:63-248 (to_tread / add_soft / Cache.add) in pure Python for small cases.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TREAD_DTYPE = np.dtype(
    [("tid", "<i4"), ("position", "<u4"), ("repeat", "S6"), ("flag", "<u2"), ("split", "u1"),
     ("mapq", "u1"), ("repeat_count", "u1"), ("align_length", "u1"), ("sample", "<i4")]
)
assert TREAD_DTYPE.itemsize == 24
BOUNDS_DTYPE = np.dtype(
    [("tid", "<i4"), ("left", "<u4"), ("left_most", "<u4"), ("right", "<u4"), ("right_most", "<u4"),
     ("center_mass", "<u4"), ("n_left", "<u2"), ("n_right", "<u2"), ("n_total", "<u2"), ("repeat", "S6"),
     ("first_read", "<u4"), ("n_reads", "<u4")]
)
assert BOUNDS_DTYPE.itemsize == 44

LOCUS_DTYPE = np.dtype([("tid", "<i4"), ("left_most", "<u4"), ("right_most", "<u4"), ("repeat", "S6"),
                        ("n_left", "<u2"), ("n_right", "<u2"), ("n_total", "<u2")])
assert LOCUS_DTYPE.itemsize == 24

LEFT, RIGHT, BOTH, NONE, NONE_RIGHT, NONE_LEFT = range(6)
SOFT_NAMES = ["left", "right", "both", "none", "none_right", "none_left"]


def build(force: bool = False) -> str:
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "strling_oracle.c")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        L.orc_get_repeat.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_char_p, C.POINTER(C.c_int)]
        L.orc_get_repeat.restype = None
        L.orc_get_repeat_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p]
        L.orc_get_repeat_batch.restype = None
        L.orc_count.argtypes = [C.c_char_p, C.c_int, C.c_int, C.POINTER(C.c_uint64)]
        L.orc_count.restype = C.c_int
        L.orc_slide_by.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.orc_slide_by.restype = C.c_int
        L.orc_reduce_repeat.argtypes = [C.c_char_p]
        L.orc_reduce_repeat.restype = C.c_int
        L.orc_min_rev_complement.argtypes = [C.c_char_p]
        L.orc_canonical_repeat.argtypes = [C.c_char_p, C.c_char_p]
        L.orc_p_repeat.argtypes = [C.c_void_p]
        L.orc_p_repeat.restype = C.c_double
        L.orc_adjust_by.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_uint8, C.c_int, C.c_uint32]
        L.orc_adjust_by.restype = C.c_int
        L.orc_unplaced_pair.argtypes = [C.c_void_p, C.c_void_p, C.c_double, C.c_uint8]
        L.orc_unplaced_pair.restype = C.c_int
        L.orc_median.argtypes = [C.c_void_p, C.c_double]
        L.orc_median.restype = C.c_int
        L.orc_hash_wangyi1.argtypes = [C.c_uint64]
        L.orc_hash_wangyi1.restype = C.c_uint64
        L.orc_counttable_largest.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.orc_counttable_largest.restype = C.c_uint32
        L.orc_bounds_of.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint16, C.c_void_p]
        L.orc_bounds_filtered.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_uint32, C.c_uint16, C.c_uint16, C.c_uint16, C.c_void_p]
        L.orc_bounds_filtered.restype = C.c_int
        L.orc_cluster_bucket.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_cluster_bucket.restype = C.c_int
        L.orc_cluster_all.argtypes = [C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_uint16, C.c_uint16, C.c_uint16, C.c_int,
                                      C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.orc_cluster_all.restype = C.c_int
        L.orc_cluster_all_loci.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_uint32, C.c_int, C.c_uint16, C.c_uint16, C.c_uint16,
                                           C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.POINTER(C.c_int)]
        L.orc_cluster_all_loci.restype = C.c_int
        _LIB = L
    return _LIB


def _unit_to_bytes(unit) -> bytes:
    b = unit.encode() if isinstance(unit, str) else bytes(unit)
    return b.rstrip(b"\0")


# ---------------------------------------------------------------- scan half
def get_repeat(read, p: float):
    """utils.nim:236 -> (unit bytes without padding, repeat_count)."""
    r = read.encode() if isinstance(read, str) else bytes(read)
    unit = C.create_string_buffer(6)
    rc = C.c_int(0)
    lib().orc_get_repeat(r, len(r), p, unit, C.byref(rc))
    return unit.raw.rstrip(b"\0"), rc.value


def get_repeat_batch(seqs: np.ndarray, off: np.ndarray, length: np.ndarray, p: np.ndarray):
    """seqs: uint8 ASCII buffer; returns (units S6 array, counts int32 array)."""
    seqs = np.ascontiguousarray(seqs, dtype=np.uint8)
    off = np.ascontiguousarray(off, dtype=np.uint64)
    length = np.ascontiguousarray(length, dtype=np.uint32)
    p = np.ascontiguousarray(p, dtype=np.float64)
    n = len(off)
    units = np.zeros(n, dtype="S6")
    counts = np.zeros(n, dtype=np.int32)
    lib().orc_get_repeat_batch(seqs.ctypes.data, off.ctypes.data, length.ctypes.data, p.ctypes.data, n,
                               units.ctypes.data, counts.ctypes.data)
    return units, counts


def slide_by(read, k: int) -> np.ndarray:
    r = read.encode() if isinstance(read, str) else bytes(read)
    out = np.zeros(max(1, len(r) // max(k, 1) + 1), dtype=np.uint64)
    n = lib().orc_slide_by(r, len(r), k, out.ctypes.data, len(out))
    return out[:n]


def count(read, k: int):
    r = read.encode() if isinstance(read, str) else bytes(read)
    leader = C.c_uint64(0)
    c = lib().orc_count(r, len(r), k, C.byref(leader))
    return c, leader.value


def reduce_repeat(unit):
    buf = C.create_string_buffer(_unit_to_bytes(unit).ljust(6, b"\0"), 6)
    m = lib().orc_reduce_repeat(buf)
    return buf.raw, m


def min_rev_complement(unit) -> bytes:
    buf = C.create_string_buffer(_unit_to_bytes(unit).ljust(6, b"\0"), 6)
    lib().orc_min_rev_complement(buf)
    return buf.raw.rstrip(b"\0")


def canonical_repeat(unit) -> bytes:
    src = C.create_string_buffer(_unit_to_bytes(unit).ljust(6, b"\0"), 6)
    dst = C.create_string_buffer(6)
    lib().orc_canonical_repeat(src, dst)
    return dst.raw.rstrip(b"\0")


def make_tread(tid=0, position=0, repeat=b"", flag=0, split=NONE, mapq=0, repeat_count=0, align_length=0, sample=0):
    t = np.zeros(1, dtype=TREAD_DTYPE)
    t["tid"], t["position"], t["repeat"], t["flag"], t["split"] = tid, position, _unit_to_bytes(repeat), flag, split
    t["mapq"], t["repeat_count"], t["align_length"], t["sample"] = mapq, repeat_count & 0xFF, align_length & 0xFF, sample
    return t


def p_repeat(t) -> float:
    return lib().orc_p_repeat(t.ctypes.data)


def adjust_by(A, B, p: float, min_mapq: int, median_frag: int, B_position: int) -> bool:
    """Mutates A (1-element TREAD_DTYPE array) in place, like extract.nim:141."""
    return bool(lib().orc_adjust_by(A.ctypes.data, B.ctypes.data, p, min_mapq, median_frag, B_position & 0xFFFFFFFF))


def unplaced_pair(A, B, p: float, min_mapq: int) -> bool:
    return bool(lib().orc_unplaced_pair(A.ctypes.data, B.ctypes.data, p, min_mapq))


def median(frag: np.ndarray, pct: float = 0.5) -> int:
    frag = np.ascontiguousarray(frag, dtype=np.uint32)
    assert frag.shape == (4096,)
    return lib().orc_median(frag.ctypes.data, pct)


# ---------------------------------------------------------------- cluster half
def counttable_largest(keys):
    k = np.ascontiguousarray(keys, dtype=np.uint32)
    v, d = C.c_int(0), C.c_int(0)
    key = lib().orc_counttable_largest(k.ctypes.data, len(k), C.byref(v), C.byref(d))
    return key, v.value, d.value


def bounds_of(reads: np.ndarray, left_most=0, right_most=0, max_clip_dist=200):
    reads = np.ascontiguousarray(reads, dtype=TREAD_DTYPE)
    out = np.zeros(1, dtype=BOUNDS_DTYPE)
    lib().orc_bounds_of(reads.ctypes.data, len(reads), left_most, right_most, max_clip_dist, out.ctypes.data)
    return out[0]


def cluster_bucket(reps: np.ndarray, max_dist: int, min_supporting_reads: int):
    """cluster.nim:364 on one position-sorted bucket -> list of (first, count, left_most, right_most)."""
    reps = np.ascontiguousarray(reps, dtype=TREAD_DTYPE)
    cap = len(reps) + 1
    a = [np.zeros(cap, dtype=np.uint32) for _ in range(4)]
    n = lib().orc_cluster_bucket(reps.ctypes.data, len(reps), max_dist, min_supporting_reads,
                                 a[0].ctypes.data, a[1].ctypes.data, a[2].ctypes.data, a[3].ctypes.data, cap)
    return [(int(a[0][i]), int(a[1][i]), int(a[2][i]), int(a[3][i])) for i in range(n)]


def cluster_all(treads: np.ndarray, window: int, min_support: int, min_clip: int = 0, min_clip_total: int = 0,
                max_clip_dist: int = 200, merge_mode: bool = False):
    """Whole cluster loop (call.nim:118-130,223-235 / merge.nim:125-187).  Returns (bounds, unplaced dict)."""
    treads = np.ascontiguousarray(treads, dtype=TREAD_DTYPE)
    n = len(treads)
    out = np.zeros(max(n, 1), dtype=BOUNDS_DTYPE)
    uu = np.zeros(max(n, 1), dtype="S6")
    uc = np.zeros(max(n, 1), dtype=np.int32)
    nu = C.c_int(0)
    nb = lib().orc_cluster_all(treads.ctypes.data, n, window, min_support, min_clip, min_clip_total, max_clip_dist,
                               int(merge_mode), out.ctypes.data, len(out), uu.ctypes.data, uc.ctypes.data, len(uu), C.byref(nu))
    assert nb >= 0
    unplaced = {bytes(uu[i]).rstrip(b"\0"): int(uc[i]) for i in range(nu.value)}
    return out[:nb].copy(), unplaced


def cluster_all_loci(treads: np.ndarray, loci: np.ndarray, window: int, min_support: int, min_clip: int = 0, min_clip_total: int = 0,
                     max_clip_dist: int = 200, merge_mode: bool = False):
    """assign_reads_locus for every locus (file order) then the cluster loop.  Returns (loci with counts, bounds, unplaced)."""
    treads = np.ascontiguousarray(treads, dtype=TREAD_DTYPE)
    loci = np.ascontiguousarray(loci, dtype=LOCUS_DTYPE).copy()
    n = len(treads)
    out = np.zeros(max(n, 1), dtype=BOUNDS_DTYPE)
    uu = np.zeros(max(n, 1), dtype="S6")
    uc = np.zeros(max(n, 1), dtype=np.int32)
    nu = C.c_int(0)
    nb = lib().orc_cluster_all_loci(treads.ctypes.data, n, loci.ctypes.data, len(loci), window, min_support, min_clip, min_clip_total,
                                    max_clip_dist, int(merge_mode), out.ctypes.data, len(out), uu.ctypes.data, uc.ctypes.data, len(uu),
                                    C.byref(nu))
    assert nb >= 0
    unplaced = {bytes(uu[i]).rstrip(b"\0"): int(uc[i]) for i in range(nu.value)}
    return loci, out[:nb].copy(), unplaced
