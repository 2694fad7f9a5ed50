## strgpu.nim -- Nim binding of libstrgpu.so (include/strgpu.h) for STRling's hot path.
##
## SOURCE ONLY: this image has no `nim` compiler, so this file has never been compiled; the same ABI is
## exercised from C++ (strling_b200/host) and Python ctypes (strling_b200/binding.py).  It shows the shim a
## STRling maintainer would add next to src/strpkg/utils.nim: `get_repeat_gpu` is a drop-in for
## `get_repeat*(read: var string, counts: var Seqs[uint8], repeat_count: var int, opts: Options)` (utils.nim:236)
## and `cluster_gpu` for the `cluster` iterator + `bounds` (cluster.nim:364, callclusters.nim:52).

{.passL: "-lstrgpu".}

type
  StrGpuCtx* = distinct pointer

  StrGpuSegment* {.bycopy.} = object   ## strgpu_segment
    base_off*: uint32
    len*: uint16
    pclass*: uint8
    flags*: uint8

  StrGpuRepeat* {.bycopy.} = object    ## strgpu_repeat
    unit*: array[6, char]
    repeat_count*: uint16

  StrGpuTread* {.bycopy.} = object     ## strgpu_tread == cluster.tread with qname -> sample
    tid*: int32
    position*: uint32
    repeat*: array[6, char]
    flag*: uint16
    split*: uint8
    mapping_quality*: uint8
    repeat_count*: uint8
    align_length*: uint8
    sample*: int32

  StrGpuBounds* {.bycopy.} = object    ## strgpu_bounds == cluster.Bounds as a POD
    tid*: int32
    left*, left_most*, right*, right_most*, center_mass*: uint32
    n_left*, n_right*, n_total*: uint16
    repeat*: array[6, char]
    first_read*, n_reads*, reserved*: uint32

  StrGpuClusterParams* {.bycopy.} = object
    window*: uint32
    min_support*: int32
    min_clip*, min_clip_total*, max_clip_dist*, merge_mode*: uint16

  StrGpuLocus* {.bycopy.} = object     ## strgpu_locus: a -l / -b locus after parse_bedline / parse_boundsline
    tid*: int32
    left_most*, right_most*: uint32
    repeat*: array[6, char]
    n_left*, n_right*, n_total*: uint16

const STRGPU_SEG_HAS_N* = 1'u8

proc strgpu_create*(ctx: ptr StrGpuCtx, device: cint): cint {.importc, cdecl, dynlib: "libstrgpu.so".}
proc strgpu_destroy*(ctx: StrGpuCtx) {.importc, cdecl, dynlib: "libstrgpu.so".}
proc strgpu_last_error*(ctx: StrGpuCtx): cstring {.importc, cdecl, dynlib: "libstrgpu.so".}
proc strgpu_set_proportions*(ctx: StrGpuCtx, p: ptr cdouble, n: cint): cint {.importc, cdecl, dynlib: "libstrgpu.so".}
proc strgpu_seq2_bytes*(n_bases: uint64): csize_t {.importc, cdecl, dynlib: "libstrgpu.so".}
proc strgpu_nmask_bytes*(n_bases: uint64): csize_t {.importc, cdecl, dynlib: "libstrgpu.so".}
proc strgpu_pack_ascii*(s: cstring, len: uint32, seq2: ptr uint8, nmask: ptr uint32, base_off: uint64): cint {.importc, cdecl, dynlib: "libstrgpu.so".}
proc strgpu_pack_bam4*(s: ptr uint8, len: uint32, seq2: ptr uint8, nmask: ptr uint32, base_off: uint64): cint {.importc, cdecl, dynlib: "libstrgpu.so".}
proc strgpu_scan_submit*(ctx: StrGpuCtx, seq2: ptr uint8, n_bases: uint64, nmask: ptr uint32, segs: ptr StrGpuSegment,
                         n_seg: uint32, max_len: uint32, res: ptr StrGpuRepeat, ticket: ptr cint): cint {.importc, cdecl, dynlib: "libstrgpu.so".}
proc strgpu_scan_wait*(ctx: StrGpuCtx, ticket: cint): cint {.importc, cdecl, dynlib: "libstrgpu.so".}
proc strgpu_scan*(ctx: StrGpuCtx, seq2: ptr uint8, n_bases: uint64, nmask: ptr uint32, segs: ptr StrGpuSegment,
                  n_seg: uint32, max_len: uint32, res: ptr StrGpuRepeat): cint {.importc, cdecl, dynlib: "libstrgpu.so".}
proc strgpu_scan_reads_submit*(ctx: StrGpuCtx, seq2: ptr uint8, n_reads, read_len, stride_bases, pclass: uint32, nmask: ptr uint32,
                               extra: ptr StrGpuSegment, n_extra, extra_max_len: uint32, res: ptr StrGpuRepeat, ticket: ptr cint): cint {.importc, cdecl, dynlib: "libstrgpu.so".}
proc strgpu_scan_reads_device*(ctx: StrGpuCtx, d_seq2: pointer, n_reads, read_len, stride_bases, pclass: uint32, d_nmask, d_extra: pointer,
                               n_extra, extra_max_len: uint32, d_out, cuda_stream: pointer): cint {.importc, cdecl, dynlib: "libstrgpu.so".}
proc strgpu_cluster_loci*(ctx: StrGpuCtx, treads: ptr StrGpuTread, n: uint32, params: ptr StrGpuClusterParams, loci: ptr StrGpuLocus,
                          n_loci: uint32, res: ptr StrGpuBounds, cap: uint32, n_out: ptr uint32): cint {.importc, cdecl, dynlib: "libstrgpu.so".}
proc strgpu_cluster*(ctx: StrGpuCtx, treads: ptr StrGpuTread, n: uint32, params: ptr StrGpuClusterParams,
                     res: ptr StrGpuBounds, cap: uint32, n_out: ptr uint32): cint {.importc, cdecl, dynlib: "libstrgpu.so".}

template check(ctx: StrGpuCtx, rc: cint) =
  if rc != 0: quit "[strling] gpu: " & $strgpu_last_error(ctx)

proc get_repeat_gpu*(ctx: StrGpuCtx, reads: seq[string], proportion_repeat: float): seq[tuple[unit: array[6, char], count: int]] =
  ## Batched stand-in for utils.get_repeat: one whole-read segment per string.
  var p = [proportion_repeat.cdouble]
  ctx.check strgpu_set_proportions(ctx, p[0].addr, 1)
  var total = 0'u64
  var segs = newSeq[StrGpuSegment](reads.len)
  var maxlen = 0'u32
  for i, r in reads:
    segs[i] = StrGpuSegment(base_off: total.uint32, len: r.len.uint16, pclass: 0)
    total += uint64((r.len + 15) div 16 * 16)
    maxlen = max(maxlen, r.len.uint32)
  var seq2 = newSeq[uint8](strgpu_seq2_bytes(total).int)
  var nmask = newSeq[uint32](strgpu_nmask_bytes(total).int div 4)
  for i, r in reads:
    if strgpu_pack_ascii(r.cstring, r.len.uint32, seq2[0].addr, nmask[0].addr, segs[i].base_off.uint64) > 0:
      segs[i].flags = STRGPU_SEG_HAS_N
  var res = newSeq[StrGpuRepeat](reads.len)
  ctx.check strgpu_scan(ctx, seq2[0].addr, total, nmask[0].addr, segs[0].addr, reads.len.uint32, maxlen, res[0].addr)
  for r in res: result.add((r.unit, r.repeat_count.int))
