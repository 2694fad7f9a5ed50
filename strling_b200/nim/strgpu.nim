## strgpu.nim -- Nim binding of libstrgpu.so (include/strgpu.h) for STRling's hot path.
##
## SOURCE ONLY: this image has no `nim` compiler, so this file has never been compiled; the same ABI is
## exercised from C++ (strling_b200/host) and Python ctypes (strling_b200/binding.py).  It shows the shim a
## STRling maintainer would add next to src/strpkg/utils.nim:
##   * declarations of EVERY entry point of include/strgpu.h (tests/test_abi_cpu.py checks that none is missing),
##   * `get_repeat_gpu`   -- batched drop-in for `get_repeat*(read, counts, repeat_count, opts)` (utils.nim:236),
##   * `cluster_gpu`      -- drop-in for the `cluster` iterator + `bounds` + filters (cluster.nim:364,
##                           callclusters.nim:52; loops of call.nim:223-235 and merge.nim:172-187),
##   * `ScanBatch` / `stage` / `submit` / `wait` -- the batch that replaces the per-read scan inside
##                           `extract_main`'s loop (extract.nim:308-322), with the loop itself at the bottom of the file.

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

  StrGpuBgzfBlock* {.bycopy.} = object ## strgpu_bgzf_block: one BGZF block of a batch handed to strgpu_inflate_bgzf
    in_off*: uint64
    csize*, isize*: uint32
    out_off*: uint64

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

const
  STRGPU_SEG_HAS_N* = 1'u8
  STRGPU_SLOTS* = 3
  STRGPU_MAX_SEGMENT_LEN* = 510
  STRGPU_COMM_ID_BYTES* = 128
  lib = "libstrgpu.so"

{.push importc, cdecl, dynlib: lib.}
# ---- lifetime
proc strgpu_version*(): cstring
proc strgpu_error_string*(status: cint): cstring
proc strgpu_create*(ctx: ptr StrGpuCtx, device: cint): cint
proc strgpu_destroy*(ctx: StrGpuCtx)
proc strgpu_last_error*(ctx: StrGpuCtx): cstring
proc strgpu_launch_count*(ctx: StrGpuCtx): uint64
proc strgpu_host_alloc*(p: ptr pointer, bytes: csize_t): cint
proc strgpu_host_free*(p: pointer)
# ---- scan: get_repeat (utils.nim:236)
proc strgpu_set_proportions*(ctx: StrGpuCtx, p: ptr cdouble, n: cint): cint
proc strgpu_seq2_bytes*(n_bases: uint64): csize_t
proc strgpu_nmask_bytes*(n_bases: uint64): csize_t
proc strgpu_pack_ascii*(s: cstring, len: uint32, seq2: ptr uint8, nmask, xmask: ptr uint32, base_off: uint64): cint
proc strgpu_pack_bam4*(s: ptr uint8, len: uint32, seq2: ptr uint8, nmask, xmask: ptr uint32, base_off: uint64): cint
proc strgpu_scan_submit*(ctx: StrGpuCtx, seq2: ptr uint8, n_bases: uint64, nmask, xmask: ptr uint32, segs: ptr StrGpuSegment,
                         n_seg, max_len: uint32, res: ptr StrGpuRepeat, ticket: ptr cint): cint
proc strgpu_scan_wait*(ctx: StrGpuCtx, ticket: cint): cint
proc strgpu_scan*(ctx: StrGpuCtx, seq2: ptr uint8, n_bases: uint64, nmask, xmask: ptr uint32, segs: ptr StrGpuSegment,
                  n_seg, max_len: uint32, res: ptr StrGpuRepeat): cint
proc strgpu_scan_reads_submit*(ctx: StrGpuCtx, seq2: ptr uint8, n_reads, read_len, stride_bases, pclass: uint32, nmask, xmask: ptr uint32,
                               extra: ptr StrGpuSegment, n_extra, extra_max_len: uint32, res: ptr StrGpuRepeat, ticket: ptr cint): cint
proc strgpu_scan_device*(ctx: StrGpuCtx, d_seq2, d_nmask, d_xmask, d_segs: pointer, n_seg, max_len: uint32, d_out, cuda_stream: pointer): cint
proc strgpu_scan_reads_device*(ctx: StrGpuCtx, d_seq2: pointer, n_reads, read_len, stride_bases, pclass: uint32, d_nmask, d_xmask, d_extra: pointer,
                               n_extra, extra_max_len: uint32, d_out, cuda_stream: pointer): cint
proc strgpu_device_status*(ctx: StrGpuCtx, cuda_stream: pointer): cint
# ---- cluster: cluster.nim:364, callclusters.nim:14,52
proc strgpu_cluster*(ctx: StrGpuCtx, treads: ptr StrGpuTread, n: uint32, params: ptr StrGpuClusterParams,
                     res: ptr StrGpuBounds, cap: uint32, n_out: ptr uint32): cint
proc strgpu_cluster_loci*(ctx: StrGpuCtx, treads: ptr StrGpuTread, n: uint32, params: ptr StrGpuClusterParams, loci: ptr StrGpuLocus,
                          n_loci: uint32, res: ptr StrGpuBounds, cap: uint32, n_out: ptr uint32): cint
proc strgpu_cluster_device*(ctx: StrGpuCtx, d_treads: pointer, n: uint32, params: ptr StrGpuClusterParams, d_out: pointer, cap: uint32,
                            d_n_out, cuda_stream: pointer): cint
# ---- sharded clustering (one process per GPU, NCCL inside the library)
proc strgpu_comm_unique_id*(id_out: pointer): cint
proc strgpu_comm_init*(ctx: StrGpuCtx, rank, world: cint, id: pointer): cint
proc strgpu_comm_info*(ctx: StrGpuCtx, rank, world: ptr cint): cint
proc strgpu_comm_destroy*(ctx: StrGpuCtx)
proc strgpu_cluster_sharded_device*(ctx: StrGpuCtx, d_treads: pointer, n, max_n, pair_capacity: uint32, params: ptr StrGpuClusterParams,
                                    d_out: pointer, cap: uint32, d_n_out, cuda_stream: pointer): cint
proc strgpu_comm_status*(ctx: StrGpuCtx, cuda_stream: pointer): cint
proc strgpu_cluster_sharded*(ctx: StrGpuCtx, treads: ptr StrGpuTread, n, max_n: uint32, params: ptr StrGpuClusterParams,
                             res: ptr StrGpuBounds, cap: uint32, n_out: ptr uint32): cint
proc strgpu_inflate_bgzf*(ctx: StrGpuCtx, comp: ptr uint8, comp_bytes: csize_t, blocks: ptr StrGpuBgzfBlock, n_blocks: uint32,
                          res: ptr uint8, out_bytes: csize_t): cint
{.pop.}

template check*(ctx: StrGpuCtx, rc: cint) =
  ## every export returns 0 or a negative strgpu_status; STRling reports failures with `quit` (e.g. extract.nim:276,290,334)
  if rc != 0: quit "[strling] gpu: " & $strgpu_last_error(ctx)

proc open_gpu*(device = 0, proportion_repeat = 0.8): StrGpuCtx =
  ## one context per process and GPU; the three proportion classes `add` uses (extract.nim:204-211,240-244)
  let rc = strgpu_create(result.addr, device.cint)
  if rc != 0: quit "[strling] gpu: " & $strgpu_error_string(rc)
  var p = [proportion_repeat.cdouble, (proportion_repeat - 0.07).cdouble, min(proportion_repeat, 0.6).cdouble]
  result.check strgpu_set_proportions(result, p[0].addr, 3)

# ------------------------------------------------------------------------------------------------------------------
# get_repeat (utils.nim:236), batched
proc get_repeat_gpu*(ctx: StrGpuCtx, reads: seq[string], pclass = 0'u8): seq[tuple[unit: array[6, char], count: int]] =
  ## One whole-read segment per string, scanned with proportion class `pclass` of open_gpu.
  var total = 0'u64
  var segs = newSeq[StrGpuSegment](reads.len)
  var maxlen = 0'u32
  for i, r in reads:
    segs[i] = StrGpuSegment(base_off: total.uint32, len: r.len.uint16, pclass: pclass)
    total += uint64((r.len + 31) div 32 * 32)          # 32-base alignment: no two reads share a mask word
    maxlen = max(maxlen, r.len.uint32)
  var seq2 = newSeq[uint8](strgpu_seq2_bytes(total).int)
  var nmask = newSeq[uint32](strgpu_nmask_bytes(total).int div 4)
  var xmask = newSeq[uint32](nmask.len)                # non-ACGT bases that are not the literal 'N' (utils.nim:238 counts 'N' only)
  for i, r in reads:
    if strgpu_pack_ascii(r.cstring, r.len.uint32, seq2[0].addr, nmask[0].addr, xmask[0].addr, segs[i].base_off.uint64) > 0:
      segs[i].flags = STRGPU_SEG_HAS_N
  var res = newSeq[StrGpuRepeat](reads.len)
  if reads.len > 0:
    ctx.check strgpu_scan(ctx, seq2[0].addr, total, nmask[0].addr, xmask[0].addr, segs[0].addr, reads.len.uint32, maxlen, res[0].addr)
  for r in res: result.add((r.unit, r.repeat_count.int))

# ------------------------------------------------------------------------------------------------------------------
# cluster + bounds + filters (cluster.nim:364, callclusters.nim:52), the loops of call.nim:223-235 / merge.nim:172-187
type GpuCluster* = object
  bounds*: seq[StrGpuBounds]                     ## tid >= 0, ascending (tid, repeat, position)
  unplaced*: seq[tuple[repeat: array[6, char], n: int]]   ## call.nim:226-228 (empty in merge mode)

proc cluster_gpu*(ctx: StrGpuCtx, treads: var seq[StrGpuTread], window: uint32, min_support: int, min_clip, min_clip_total,
                  max_clip_dist: uint16, merge_mode = false, loci: ptr seq[StrGpuLocus] = nil): GpuCluster =
  ## treads: `.bin` / concatenation order; `sample` = the index merge.nim:121-124 writes into qname (0 in `call`).
  ## loci (optional): the -l / -b loci in list order; their n_left / n_right / n_total are filled in
  ## (assign_reads_locus, callclusters.nim:14-50) and their reads are removed before clustering.
  var p = StrGpuClusterParams(window: window, min_support: min_support.int32, min_clip: min_clip, min_clip_total: min_clip_total,
                              max_clip_dist: max_clip_dist, merge_mode: (if merge_mode: 1'u16 else: 0'u16))
  var res = newSeq[StrGpuBounds](max(16, treads.len))
  var n_out = 0'u32
  let tp = if treads.len > 0: treads[0].addr else: nil
  if loci != nil and loci[].len > 0:
    ctx.check strgpu_cluster_loci(ctx, tp, treads.len.uint32, p.addr, loci[][0].addr, loci[].len.uint32, res[0].addr, res.len.uint32, n_out.addr)
  else:
    ctx.check strgpu_cluster(ctx, tp, treads.len.uint32, p.addr, res[0].addr, res.len.uint32, n_out.addr)
  for i in 0 ..< n_out.int:
    if res[i].tid < 0: result.unplaced.add((res[i].repeat, res[i].n_reads.int))
    else: result.bounds.add(res[i])

# ------------------------------------------------------------------------------------------------------------------
# extract: the batch that replaces the per-read scan of extract.nim:40,114
type
  StagedRead* = object        ## what the replay of Cache.add needs besides the hts Record itself
    seg_full*: int32          ## segment index of the whole read, -1 when the genome-STR filter skipped it (extract.nim:30-34)
    seg_clip*: array[2, array[2, int32]]   ## [left, right][first-seen class p-0.07, second-seen class min(p,0.6)], -1 = none
    m_len*: int32             ## align_length of a skipped read (extract.nim:33)
  ScanBatch* = object
    seq2*: seq[uint8]
    nmask*, xmask*: seq[uint32]
    segs*: seq[StrGpuSegment]
    res*: seq[StrGpuRepeat]
    reads*: seq[StagedRead]
    n_bases*: uint64
    max_len*: uint32
    any_n*: bool
    ticket*: cint

proc stage*(b: var ScanBatch, bam_seq: ptr uint8, l_seq: int, skip: bool, m_len: int, left_clip, right_clip: int) =
  ## Adds one primary record: bam_seq = bam_get_seq(aln.b) (4-bit), left_clip / right_clip = length of a leading / trailing
  ## S op when add_soft's preconditions hold (mapq >= min_mapq; extract.nim:97-104), else 0.
  var sr = StagedRead(seg_full: -1, m_len: m_len.int32)
  sr.seg_clip = [[-1'i32, -1'i32], [-1'i32, -1'i32]]
  let base = b.n_bases
  if not skip or left_clip > 0 or right_clip > 0:
    let need = strgpu_seq2_bytes(base + l_seq.uint64 + 64).int
    if b.seq2.len < need:
      b.seq2.setLen(need * 2)
      b.nmask.setLen(strgpu_nmask_bytes(uint64(need * 8)).int div 4)
      b.xmask.setLen(b.nmask.len)
    let n_other = strgpu_pack_bam4(bam_seq, l_seq.uint32, b.seq2[0].addr, b.nmask[0].addr, b.xmask[0].addr, base)
    let flags = if n_other > 0: STRGPU_SEG_HAS_N else: 0'u8
    b.any_n = b.any_n or n_other > 0
    template put(off: uint64, len: int, cls: uint8): int32 =
      b.segs.add StrGpuSegment(base_off: off.uint32, len: len.uint16, pclass: cls, flags: flags)
      b.max_len = max(b.max_len, len.uint32)
      int32(b.segs.len - 1)
    if not skip: sr.seg_full = put(base, l_seq, 0)
    if left_clip > 0:
      sr.seg_clip[0][0] = put(base, min(left_clip, l_seq), 1)
      sr.seg_clip[0][1] = put(base, min(left_clip, l_seq), 2)
    if right_clip > 0:
      let n = min(right_clip, l_seq)
      sr.seg_clip[1][0] = put(base + uint64(l_seq - n), n, 1)
      sr.seg_clip[1][1] = put(base + uint64(l_seq - n), n, 2)
    b.n_bases = base + uint64((l_seq + 31) div 32 * 32)
  b.reads.add sr

proc submit*(ctx: StrGpuCtx, b: var ScanBatch) =
  b.res.setLen(b.segs.len)
  if b.segs.len == 0: b.ticket = -1; return
  ctx.check strgpu_scan_submit(ctx, b.seq2[0].addr, b.n_bases, (if b.any_n: b.nmask[0].addr else: nil),
                               (if b.any_n: b.xmask[0].addr else: nil), b.segs[0].addr, b.segs.len.uint32, b.max_len, b.res[0].addr,
                               b.ticket.addr)

proc wait*(ctx: StrGpuCtx, b: var ScanBatch) =
  if b.ticket >= 0: ctx.check strgpu_scan_wait(ctx, b.ticket)

proc repeat_of*(b: ScanBatch, seg: int32, repeat_count: var int): array[6, char] =
  ## what `read.get_repeat(counts, repeat_count, opts)` returned for that segment (utils.nim:236-271)
  if seg < 0: repeat_count = 0; return
  repeat_count = b.res[seg].repeat_count.int
  b.res[seg].unit

# The loop of extract_main (extract.nim:308-322) then becomes -- sketch against hts-nim's Record API:
#
#   let gpu = open_gpu(0, proportion_repeat)
#   var batches: array[2, ScanBatch]; var alns: array[2, seq[Record]]; var cur = 0
#   for aln in ibam:                                   # also the `ibam.query("*")` pass, extract.nim:326-329
#     if aln.flag.secondary or aln.flag.supplementary: continue
#     let single_m = aln.cigar.len == 1 and aln.cigar[0].op == CigarOp.match
#     let skip = single_m and aln.chrom in genome_str and genome_str[aln.chrom].find(aln.start, aln.stop, res) == false   # extract.nim:30-34
#     var lc, rc = 0
#     if aln.mapping_quality >= opts.min_mapq and aln.cigar.len > 0:                                                       # extract.nim:97-104
#       if aln.cigar[0].op == CigarOp.soft_clip: lc = aln.cigar[0].len
#       if aln.cigar.len > 1 and aln.cigar[aln.cigar.len - 1].op == CigarOp.soft_clip: rc = aln.cigar[aln.cigar.len - 1].len
#     batches[cur].stage(bam_get_seq(aln.b), aln.b.core.l_qseq, skip, aln.cigar[0].len, lc, rc)
#     alns[cur].add aln.copy()
#     if alns[cur].len == batch_reads:
#       gpu.submit(batches[cur])                       # H2D + kernels + D2H run while the previous batch is replayed
#       let prev = 1 - cur
#       gpu.wait(batches[prev])
#       for i, a in alns[prev]: cache.add(a, batches[prev], i, opts)     # Cache.add with repeat_of() in place of get_repeat
#       batches[prev] = ScanBatch(); alns[prev].setLen(0); cur = prev
#
# where Cache.add / to_tread / add_soft (extract.nim:63-132,192-248) are unchanged except that
#   `aln.get_repeat(genome_str, counts, repeat_count, align_length, opts)`  ->  `batch.repeat_of(batch.reads[i].seg_full, repeat_count)`
#   (align_length = l_qseq, or reads[i].m_len when seg_full < 0), and inside add_soft
#   `soft_seq.get_repeat(counts, repeat_count, opts)`  ->  `batch.repeat_of(batch.reads[i].seg_clip[side][class], repeat_count)`
#   with class 0 for the first-seen call (p - 0.07, extract.nim:208) and 1 for the second-seen call (min(p, 0.6), extract.nim:242).
# strling_b200/host/extract.cpp is exactly this loop in C++ (plus threads for inflate / staging).
