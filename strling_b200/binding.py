"""ctypes binding of libstrgpu.so (include/strgpu.h).  The names mirror the reference procs they replace:
`StrGpu.get_repeat` <- get_repeat (utils.nim:236), `StrGpu.cluster` <- cluster/bounds (cluster.nim:364,
callclusters.nim:52).  Nothing here computes on the CPU."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstrgpu.so")

SEGMENT_DTYPE = np.dtype([("base_off", "<u4"), ("len", "<u2"), ("pclass", "u1"), ("flags", "u1")])
REPEAT_DTYPE = np.dtype([("unit", "S6"), ("repeat_count", "<u2")])
BGZF_BLOCK_DTYPE = np.dtype([("in_off", "<u8"), ("csize", "<u4"), ("isize", "<u4"), ("out_off", "<u8")])
TREAD_DTYPE = np.dtype(
    [("tid", "<i4"), ("position", "<u4"), ("repeat", "S6"), ("flag", "<u2"), ("split", "u1"),
     ("mapq", "u1"), ("repeat_count", "u1"), ("align_length", "u1"), ("sample", "<i4")]
)
BOUNDS_DTYPE = np.dtype(
    [("tid", "<i4"), ("left", "<u4"), ("left_most", "<u4"), ("right", "<u4"), ("right_most", "<u4"),
     ("center_mass", "<u4"), ("n_left", "<u2"), ("n_right", "<u2"), ("n_total", "<u2"), ("repeat", "S6"),
     ("first_read", "<u4"), ("n_reads", "<u4"), ("reserved", "<u4")]
)
LOCUS_DTYPE = np.dtype([("tid", "<i4"), ("left_most", "<u4"), ("right_most", "<u4"), ("repeat", "S6"),
                        ("n_left", "<u2"), ("n_right", "<u2"), ("n_total", "<u2")])
CLUSTER_PARAMS_DTYPE = np.dtype(
    [("window", "<u4"), ("min_support", "<i4"), ("min_clip", "<u2"), ("min_clip_total", "<u2"),
     ("max_clip_dist", "<u2"), ("merge_mode", "<u2")]
)
assert SEGMENT_DTYPE.itemsize == 8 and REPEAT_DTYPE.itemsize == 8
assert TREAD_DTYPE.itemsize == 24 and BOUNDS_DTYPE.itemsize == 48 and CLUSTER_PARAMS_DTYPE.itemsize == 16 and LOCUS_DTYPE.itemsize == 24
SEG_HAS_N = 1
MAX_SEGMENT_LEN = 510
COMM_ID_BYTES = 128

_lib = None


class StrGpuError(RuntimeError):
    def __init__(self, status: int, msg: str):
        super().__init__(f"strgpu status {status}: {msg}")
        self.status = status


def load_library():
    """Loads the in-tree CUDA library; raises (never falls back) when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: run `python -m strling_b200.build` (or __graft_entry__.build()); "
                          "there is no CPU fallback for the scan / cluster kernels")
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    L.strgpu_version.restype = C.c_char_p
    L.strgpu_error_string.argtypes = [i32]
    L.strgpu_error_string.restype = C.c_char_p
    L.strgpu_create.argtypes = [C.POINTER(vp), i32]
    L.strgpu_destroy.argtypes = [vp]
    L.strgpu_destroy.restype = None
    L.strgpu_last_error.argtypes = [vp]
    L.strgpu_last_error.restype = C.c_char_p
    L.strgpu_launch_count.argtypes = [vp]
    L.strgpu_launch_count.restype = u64
    L.strgpu_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    L.strgpu_host_free.argtypes = [vp]
    L.strgpu_host_free.restype = None
    L.strgpu_set_proportions.argtypes = [vp, C.POINTER(C.c_double), i32]
    L.strgpu_seq2_bytes.argtypes = [u64]
    L.strgpu_seq2_bytes.restype = C.c_size_t
    L.strgpu_nmask_bytes.argtypes = [u64]
    L.strgpu_nmask_bytes.restype = C.c_size_t
    L.strgpu_pack_ascii.argtypes = [vp, u32, vp, vp, vp, u64]
    L.strgpu_pack_bam4.argtypes = [vp, u32, vp, vp, vp, u64]
    L.strgpu_scan_submit.argtypes = [vp, vp, u64, vp, vp, vp, u32, u32, vp, C.POINTER(i32)]
    L.strgpu_scan_reads_submit.argtypes = [vp, vp, u32, u32, u32, u32, vp, vp, vp, u32, u32, vp, C.POINTER(i32)]
    L.strgpu_scan_wait.argtypes = [vp, i32]
    L.strgpu_scan.argtypes = [vp, vp, u64, vp, vp, vp, u32, u32, vp]
    L.strgpu_scan_device.argtypes = [vp, vp, vp, vp, vp, u32, u32, vp, vp]
    L.strgpu_scan_reads_device.argtypes = [vp, vp, u32, u32, u32, u32, vp, vp, vp, u32, u32, vp, vp]
    L.strgpu_device_status.argtypes = [vp, vp]
    L.strgpu_cluster.argtypes = [vp, vp, u32, vp, vp, u32, C.POINTER(u32)]
    L.strgpu_cluster_loci.argtypes = [vp, vp, u32, vp, vp, u32, vp, u32, C.POINTER(u32)]
    L.strgpu_cluster_device.argtypes = [vp, vp, u32, vp, vp, u32, vp, vp]
    L.strgpu_comm_unique_id.argtypes = [vp]
    L.strgpu_comm_init.argtypes = [vp, i32, i32, vp]
    L.strgpu_comm_info.argtypes = [vp, C.POINTER(i32), C.POINTER(i32)]
    L.strgpu_comm_destroy.argtypes = [vp]
    L.strgpu_comm_destroy.restype = None
    L.strgpu_cluster_sharded_device.argtypes = [vp, vp, u32, u32, u32, vp, vp, u32, vp, vp]
    L.strgpu_comm_status.argtypes = [vp, vp]
    L.strgpu_cluster_sharded.argtypes = [vp, vp, u32, u32, vp, vp, u32, C.POINTER(u32)]
    L.strgpu_inflate_bgzf.argtypes = [vp, vp, C.c_size_t, vp, u32, vp, C.c_size_t]
    L.strgpu_inflate_bgzf.restype = i32
    _lib = L
    return L


class Masks(tuple):
    """(nmask, xmask): the non-ACGT plane and the plane of non-ACGT bases that are not the literal 'N' (include/strgpu.h).
    Accepted wherever an `nmask` argument is: a bare array means "every flagged base is N"."""

    def __new__(cls, nmask, xmask):
        return super().__new__(cls, (nmask, xmask))


def _mask_ptrs(nmask):
    if nmask is None:
        return None, None
    if isinstance(nmask, Masks):
        n, x = nmask
        return (None if n is None else n.ctypes.data), (None if x is None or n is None else x.ctypes.data)
    return nmask.ctypes.data, None


def pack_reads(reads, pclass=0, align_bases: int = 16):
    """Packs ASCII reads into one seq2 buffer (+ N masks) and one whole-read segment per read.
    Each read starts at a multiple of `align_bases` (>= 4).  Returns (seq2 u8, Masks or None, segments, n_bases)."""
    L = load_library()
    reads = [r.encode() if isinstance(r, str) else bytes(r) for r in reads]
    n = len(reads)
    lens = np.fromiter((len(r) for r in reads), dtype=np.int64, count=n)
    padded = (lens + align_bases - 1) // align_bases * align_bases
    offs = np.zeros(n, dtype=np.int64)
    if n:
        offs[1:] = np.cumsum(padded)[:-1]
    n_bases = int(padded.sum())
    seq2 = np.zeros(L.strgpu_seq2_bytes(n_bases), dtype=np.uint8)
    nmask = np.zeros(L.strgpu_nmask_bytes(n_bases) // 4, dtype=np.uint32)
    xmask = np.zeros(L.strgpu_nmask_bytes(n_bases) // 4, dtype=np.uint32)
    segs = np.zeros(n, dtype=SEGMENT_DTYPE)
    segs["base_off"] = offs
    segs["len"] = lens
    segs["pclass"] = pclass
    any_n = False
    for i, r in enumerate(reads):
        k = L.strgpu_pack_ascii(r, len(r), seq2.ctypes.data, nmask.ctypes.data, xmask.ctypes.data, int(offs[i]))
        if k < 0:
            raise StrGpuError(k, "pack_ascii")
        if k:
            segs["flags"][i] |= SEG_HAS_N
            any_n = True
    return seq2, (Masks(nmask, xmask if xmask.any() else None) if any_n else None), segs, n_bases


class StrGpu:
    """One context per process and GPU (include/strgpu.h)."""

    def __init__(self, device: int = 0):
        self.L = load_library()
        h = C.c_void_p()
        rc = self.L.strgpu_create(C.byref(h), device)
        if rc != 0:
            msg = self.L.strgpu_last_error(h).decode() if h else self.L.strgpu_error_string(rc).decode()
            if h:
                self.L.strgpu_destroy(h)
            raise StrGpuError(rc, msg)
        self.h = h
        self.device = device

    def close(self):
        if getattr(self, "h", None):
            self.L.strgpu_destroy(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc: int):
        if rc != 0:
            raise StrGpuError(rc, self.L.strgpu_last_error(self.h).decode())

    @property
    def launch_count(self) -> int:
        return int(self.L.strgpu_launch_count(self.h))

    def set_proportions(self, ps):
        arr = (C.c_double * len(ps))(*ps)
        self._check(self.L.strgpu_set_proportions(self.h, arr, len(ps)))

    # ---- BGZF inflate -------------------------------------------------------------------------
    def inflate_bgzf(self, comp: np.ndarray, blocks: np.ndarray, out_bytes: int) -> np.ndarray:
        """Inflates BGZF block payloads on the device (strgpu_inflate_bgzf).  comp: uint8 array holding the raw DEFLATE payloads,
        blocks: BGZF_BLOCK_DTYPE array (in_off, csize, isize, out_off); returns the uint8 output buffer of out_bytes bytes."""
        comp = np.ascontiguousarray(comp, dtype=np.uint8)
        blocks = np.ascontiguousarray(blocks, dtype=BGZF_BLOCK_DTYPE)
        out = np.zeros(out_bytes, dtype=np.uint8)
        self._check(self.L.strgpu_inflate_bgzf(self.h, comp.ctypes.data, comp.nbytes, blocks.ctypes.data, len(blocks), out.ctypes.data, out.nbytes))
        return out

    # ---- scan ---------------------------------------------------------------------------------
    def scan(self, seq2: np.ndarray, n_bases: int, nmask, segs: np.ndarray, max_len: int | None = None) -> np.ndarray:
        """get_repeat for every segment (host buffers in, host results out)."""
        segs = np.ascontiguousarray(segs, dtype=SEGMENT_DTYPE)
        out = np.zeros(len(segs), dtype=REPEAT_DTYPE)
        if max_len is None:
            max_len = int(segs["len"].max()) if len(segs) else 0
        nm, xm = _mask_ptrs(nmask)
        self._check(self.L.strgpu_scan(self.h, seq2.ctypes.data, n_bases, nm, xm, segs.ctypes.data, len(segs), max_len, out.ctypes.data))
        return out

    def scan_submit(self, seq2, n_bases, nmask, segs, max_len, out) -> int:
        t = C.c_int(-1)
        nm, xm = _mask_ptrs(nmask)
        self._check(self.L.strgpu_scan_submit(self.h, seq2.ctypes.data, n_bases, nm, xm, segs.ctypes.data, len(segs), max_len,
                                              out.ctypes.data, C.byref(t)))
        return t.value

    def scan_reads_submit(self, seq2, n_reads: int, read_len: int, stride_bases: int, pclass: int, nmask, extra, extra_max_len: int, out) -> int:
        """Uniform whole reads without descriptors (+ optional explicit extra segments); results: reads first, then extras."""
        t = C.c_int(-1)
        n_extra = 0 if extra is None else len(extra)
        nm, xm = _mask_ptrs(nmask)
        self._check(self.L.strgpu_scan_reads_submit(self.h, seq2.ctypes.data, n_reads, read_len, stride_bases, pclass, nm, xm,
                                                    None if extra is None else extra.ctypes.data, n_extra, extra_max_len,
                                                    out.ctypes.data, C.byref(t)))
        return t.value

    def scan_wait(self, ticket: int):
        self._check(self.L.strgpu_scan_wait(self.h, ticket))

    def scan_reads_device(self, d_seq2: int, n_reads: int, read_len: int, stride_bases: int, pclass: int, d_nmask: int | None,
                          d_extra: int | None, n_extra: int, extra_max_len: int, d_out: int, stream: int = 0, d_xmask: int | None = None):
        """Device-resident uniform-read batch (strgpu_scan_reads_device): all pointers are device addresses."""
        self._check(self.L.strgpu_scan_reads_device(self.h, d_seq2, n_reads, read_len, stride_bases, pclass, d_nmask or None,
                                                    d_xmask or None, d_extra or None, n_extra, extra_max_len, d_out, stream or None))

    def scan_device(self, d_seq2: int, d_nmask: int | None, d_segs: int, n_seg: int, max_len: int, d_out: int, stream: int = 0,
                    d_xmask: int | None = None):
        self._check(self.L.strgpu_scan_device(self.h, d_seq2, d_nmask, d_xmask or None, d_segs, n_seg, max_len, d_out, stream or None))

    def device_status(self, stream: int = 0):
        self._check(self.L.strgpu_device_status(self.h, stream or None))

    def get_repeat(self, reads, p: float = 0.8):
        """Convenience mirror of get_repeat(read, ..) (utils.nim:236) for a list of ASCII reads.
        Returns a list of (unit bytes, repeat_count)."""
        self.set_proportions([p])
        seq2, nmask, segs, n_bases = pack_reads(reads, 0)
        res = self.scan(seq2, n_bases, nmask, segs)
        return [(bytes(r["unit"]).rstrip(b"\0"), int(r["repeat_count"])) for r in res]

    # ---- cluster ------------------------------------------------------------------------------
    @staticmethod
    def cluster_params(window: int, min_support: int, min_clip: int = 0, min_clip_total: int = 0, max_clip_dist: int = 200,
                       merge_mode: bool = False) -> np.ndarray:
        p = np.zeros(1, dtype=CLUSTER_PARAMS_DTYPE)
        p["window"], p["min_support"], p["min_clip"], p["min_clip_total"] = window, min_support, min_clip, min_clip_total
        p["max_clip_dist"], p["merge_mode"] = max_clip_dist, int(merge_mode)
        return p

    def cluster(self, treads: np.ndarray, window: int, min_support: int, min_clip: int = 0, min_clip_total: int = 0,
                max_clip_dist: int = 200, merge_mode: bool = False):
        """The cluster loop of call.nim:223-235 / merge.nim:172-187 over treads in `.bin` order.
        Returns (bounds records with tid >= 0, {unit: count} of unplaced buckets)."""
        treads = np.ascontiguousarray(treads, dtype=TREAD_DTYPE)
        p = self.cluster_params(window, min_support, min_clip, min_clip_total, max_clip_dist, merge_mode)
        cap = max(16, len(treads))
        out = np.zeros(cap, dtype=BOUNDS_DTYPE)
        n_out = C.c_uint32(0)
        self._check(self.L.strgpu_cluster(self.h, treads.ctypes.data, len(treads), p.ctypes.data, out.ctypes.data, cap,
                                          C.byref(n_out)))
        out = out[: n_out.value]
        unplaced = {bytes(r["repeat"]).rstrip(b"\0"): int(r["n_reads"]) for r in out[out["tid"] < 0]}
        return out[out["tid"] >= 0].copy(), unplaced

    def cluster_loci(self, treads: np.ndarray, loci: np.ndarray, window: int, min_support: int, min_clip: int = 0,
                     min_clip_total: int = 0, max_clip_dist: int = 200, merge_mode: bool = False):
        """assign_reads_locus (callclusters.nim:14) for every locus in order, then the cluster loop.
        Returns (loci with n_left/n_right/n_total filled in, bounds, unplaced)."""
        treads = np.ascontiguousarray(treads, dtype=TREAD_DTYPE)
        loci = np.ascontiguousarray(loci, dtype=LOCUS_DTYPE).copy()
        p = self.cluster_params(window, min_support, min_clip, min_clip_total, max_clip_dist, merge_mode)
        cap = max(16, len(treads))
        out = np.zeros(cap, dtype=BOUNDS_DTYPE)
        n_out = C.c_uint32(0)
        self._check(self.L.strgpu_cluster_loci(self.h, treads.ctypes.data, len(treads), p.ctypes.data, loci.ctypes.data, len(loci),
                                               out.ctypes.data, cap, C.byref(n_out)))
        out = out[: n_out.value]
        unplaced = {bytes(r["repeat"]).rstrip(b"\0"): int(r["n_reads"]) for r in out[out["tid"] < 0]}
        return loci, out[out["tid"] >= 0].copy(), unplaced

    # ---- sharded clustering (one process per GPU, NCCL inside the library) ------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        buf = C.create_string_buffer(COMM_ID_BYTES)
        rc = load_library().strgpu_comm_unique_id(buf)
        if rc != 0:
            raise StrGpuError(rc, "strgpu_comm_unique_id (libnccl.so.2 not loadable?)")
        return buf.raw

    def comm_init(self, rank: int, world: int, uid: bytes):
        assert len(uid) == COMM_ID_BYTES
        self._check(self.L.strgpu_comm_init(self.h, rank, world, uid))

    def comm_init_torch(self):
        """Communicator over the ranks of an initialised torch.distributed job (the id travels through its store / broadcast)."""
        import torch
        import torch.distributed as dist

        rank, world = dist.get_rank(), dist.get_world_size()
        box = [self.comm_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        torch.cuda.synchronize()
        self.comm_init(rank, world, box[0])

    def cluster_sharded_device(self, d_treads: int, n: int, max_n: int, params: np.ndarray, d_out: int, cap: int, d_n_out: int,
                               stream: int = 0, pair_capacity: int = 0):
        self._check(self.L.strgpu_cluster_sharded_device(self.h, d_treads, n, max_n, pair_capacity, params.ctypes.data, d_out, cap,
                                                         d_n_out, stream or None))

    def comm_status(self, stream: int = 0):
        self._check(self.L.strgpu_comm_status(self.h, stream or None))

    def cluster_sharded(self, treads: np.ndarray, max_n: int, window: int, min_support: int, min_clip: int = 0, min_clip_total: int = 0,
                        max_clip_dist: int = 200, merge_mode: bool = False, cap: int | None = None):
        """Collective: this rank's shard in, the cluster records of the whole job out (same result and order on every rank as
        `cluster` over the concatenated shards).  Returns (bounds with tid >= 0, {unit: count} of unplaced buckets)."""
        treads = np.ascontiguousarray(treads, dtype=TREAD_DTYPE)
        p = self.cluster_params(window, min_support, min_clip, min_clip_total, max_clip_dist, merge_mode)
        cap = cap or max(1024, 2 * max_n)
        out = np.zeros(cap, dtype=BOUNDS_DTYPE)
        n_out = C.c_uint32(0)
        self._check(self.L.strgpu_cluster_sharded(self.h, treads.ctypes.data, len(treads), max_n, p.ctypes.data, out.ctypes.data, cap,
                                                  C.byref(n_out)))
        out = out[: n_out.value]
        unplaced = {bytes(r["repeat"]).rstrip(b"\0"): int(r["n_reads"]) for r in out[out["tid"] < 0]}
        return out[out["tid"] >= 0].copy(), unplaced

    def cluster_device(self, d_treads: int, n: int, params: np.ndarray, d_out: int, cap: int, d_n_out: int, stream: int = 0):
        self._check(self.L.strgpu_cluster_device(self.h, d_treads, n, params.ctypes.data, d_out, cap, d_n_out, stream or None))
