/*
 * strgpu.h -- C ABI of libstrgpu.so, the B200 (sm_100a) implementation of STRling's data-parallel hot path.
 *
 * STRling (Nim) has no plugin/FFI surface of its own; the boundary is the set of Nim procs a shim replaces.
 * Every entry point below cites the reference interface it stands in for (paths relative to the reference
 * checkout).  INTEGRATION.md shows the Nim `importc` binding a maintainer would add.
 *
 * Conventions: plain pointers and sizes only; every function returns 0 on success or a negative
 * strgpu_status; nothing aborts or exits (the reference `quit`s / doAsserts -- the shim maps a non-zero
 * status to `quit`).  One ctx per GPU; a ctx may be shared by ONE submitting and ONE waiting host thread (slot
 * bookkeeping is locked inside the library) and, beside them, ONE thread calling strgpu_inflate_bgzf (own stream, own
 * buffers, own lock); every other use is one thread at a time.  All structs are little-endian PODs with the
 * exact layouts below (static_asserted in the implementation).
 */
#ifndef STRGPU_H
#define STRGPU_H
#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct strgpu_ctx strgpu_ctx;

typedef enum {
  STRGPU_OK = 0,
  STRGPU_ERR_INVALID = -1,       /* bad argument */
  STRGPU_ERR_CUDA = -2,          /* CUDA runtime error; see strgpu_last_error */
  STRGPU_ERR_NO_DEVICE = -3,     /* no usable sm_100 device: the library never falls back to the CPU */
  STRGPU_ERR_TOO_LONG = -4,      /* a segment is longer than STRGPU_MAX_SEGMENT_LEN (or than max_len passed) */
  STRGPU_ERR_BUSY = -5,          /* no free submit slot: wait on an earlier ticket first */
  STRGPU_ERR_TICKET = -6,        /* unknown / already consumed ticket */
  STRGPU_ERR_OVERFLOW = -7,      /* an output capacity was too small */
  STRGPU_ERR_DATA = -8           /* malformed input data (a BGZF block that does not inflate to its stated size) */
} strgpu_status;

/* The reference counts k-mers in uint8 tables with overflow checks off (utils.nim:9,113-117,192-195);
 * beyond 510 bases a count can wrap, so longer segments are rejected instead of reproducing the wrap. */
#define STRGPU_MAX_SEGMENT_LEN 510
#define STRGPU_MAX_PCLASS 4
#define STRGPU_SLOTS 3            /* submit/wait pipeline depth */

/* ---- sequence encoding -------------------------------------------------------------------------
 * seq2: 2 bits per base, 4 bases per byte, first base in bits 7..6 (byte-wise big-endian, like BAM's
 * 4-bit packing).  Codes C=0 A=1 T=2 G=3 -- the order of the `kmer` nimble package the reference scans
 * with (utils.nim:14,19,245).  A non-ACGT base is stored as 1 ('A', as that package encodes it) and is
 * additionally flagged in nmask: bit (b & 31) of 32-bit word (b >> 5), b = absolute base index.
 * nmask may be NULL when no segment has STRGPU_SEG_HAS_N set.
 * xmask (same layout, optional) flags the non-ACGT bases that are NOT the literal 'N' (IUPAC ambiguity codes, '='):
 * the reference's `read.count('N') > 20` gate (utils.nim:238) counts 'N' only, while every non-ACGT base scans as 'A'
 * and never matches in the recount (utils.nim:254 compares raw characters).  xmask == NULL: every flagged base is 'N'.
 * The seq2 buffer handed to the library must be readable for 8 bytes past the last base
 * (strgpu_seq2_bytes() includes that slack); nmask / xmask for 8 bytes past their last word likewise.
 */
#define STRGPU_SEG_HAS_N 0x01u

/* One scan unit: a read, a soft-clipped end of a read, or a reference window.
 * Replaces the `read: var string` argument of get_repeat (utils.nim:236). */
typedef struct {
  uint32_t base_off;   /* index of the segment's first base in seq2 / nmask */
  uint16_t len;        /* bases; 0..STRGPU_MAX_SEGMENT_LEN */
  uint8_t  pclass;     /* which proportion_repeat (strgpu_set_proportions) applies: opts.proportion_repeat */
  uint8_t  flags;      /* STRGPU_SEG_* */
} strgpu_segment;      /* 8 bytes */

/* Result of get_repeat (utils.nim:236-271): `result: array[6, char]` zero padded, and `repeat_count`
 * (already multiplied by reduce_repeat, utils.nim:271). */
typedef struct {
  char     unit[6];
  uint16_t repeat_count;
} strgpu_repeat;       /* 8 bytes */

/* ---- lifetime ---------------------------------------------------------------------------------- */
const char *strgpu_version(void);
const char *strgpu_error_string(int status);
/* device: CUDA ordinal.  Fails with STRGPU_ERR_NO_DEVICE unless it is compute capability 10.x.
 * On any failure *ctx is NULL and nothing is left to destroy. */
int  strgpu_create(strgpu_ctx **ctx, int device);
void strgpu_destroy(strgpu_ctx *ctx);
const char *strgpu_last_error(const strgpu_ctx *ctx);
/* kernels launched by this ctx since creation (for benchmark accounting) */
uint64_t strgpu_launch_count(const strgpu_ctx *ctx);

/* Pinned host memory for submit buffers (optional but needed for copy/compute overlap). */
int  strgpu_host_alloc(void **ptr, size_t bytes);
void strgpu_host_free(void *ptr);

/* ---- BGZF inflate (optional front end of the scan) ----------------------------------------------
 * Stands in for the htslib inflate behind hts-nim's record iterator (`for aln in ibam`, extract.nim:308,326): a batch of
 * BGZF blocks (SAM specification 4.1) is copied to the device compressed, inflated there -- one block per warp -- and the
 * inflated bytes are copied back, so the host cores only walk and stage records.  Record boundaries, field extraction and
 * mate pairing stay with the caller.  Synchronous; runs on its own stream beside scans in flight on the same ctx. */
typedef struct {
  uint64_t in_off;     /* offset in `comp` of the block's raw DEFLATE payload (after the gzip header + extra field) */
  uint32_t csize;      /* payload bytes (BSIZE + 1 - XLEN - 20) */
  uint32_t isize;      /* inflated size from the block footer; 0: nothing to do */
  uint64_t out_off;    /* where in `out` the block's isize bytes go; blocks must not overlap */
} strgpu_bgzf_block;   /* 24 bytes */
/* comp / out: host buffers (pinned memory from strgpu_host_alloc makes the copies asynchronous DMA).  Only
 * [min out_off, max out_off + isize) of `out` is written.  STRGPU_ERR_DATA: a block is not a valid DEFLATE stream of
 * exactly isize bytes (strgpu_last_error names it). */
int strgpu_inflate_bgzf(strgpu_ctx *ctx, const uint8_t *comp, size_t comp_bytes, const strgpu_bgzf_block *blocks, uint32_t n_blocks,
                        uint8_t *out, size_t out_bytes);

/* ---- scan: get_repeat(read, counts, repeat_count, opts) -- utils.nim:236, called from
 * extract.nim:40 (whole read), extract.nim:114 (soft clip), genome_strs.nim:74 (reference window) ---- */

/* The proportion_repeat values in play (extract.nim:204-211,240-244 uses p, p-0.07 and min(p,0.6)).
 * Thresholds int(len*p/k) and int(len*0.12/k) (utils.nim:251,259) are tabulated on the host in fp64
 * exactly as the reference computes them. */
int strgpu_set_proportions(strgpu_ctx *ctx, const double *p, int n);

/* bytes to allocate for n_bases of seq2 / nmask, slack included */
size_t strgpu_seq2_bytes(uint64_t n_bases);
size_t strgpu_nmask_bytes(uint64_t n_bases);
/* Host packers: write `len` bases at base index base_off.  Return the number of non-ACGT bases.
 * ascii: what hts-nim's aln.sequence() yields (extract.nim:37); bam4: the BAM record's 4-bit SEQ field.
 * base_off must be a multiple of 4 for the packers (segments themselves may start anywhere). */
int strgpu_pack_ascii(const char *seq, uint32_t len, uint8_t *seq2, uint32_t *nmask, uint32_t *xmask, uint64_t base_off);
int strgpu_pack_bam4(const uint8_t *bam_seq, uint32_t len, uint8_t *seq2, uint32_t *nmask, uint32_t *xmask, uint64_t base_off);

/* Asynchronous host API.  Copies the batch to the device on an internal stream, runs the scan, copies the
 * results back; strgpu_scan_wait blocks until `out` (n_seg records) is filled.  Buffers must stay valid
 * until the wait returns.  max_len: upper bound of segment lengths in this batch (selects the kernel
 * variant; a longer segment yields STRGPU_ERR_TOO_LONG at wait). */
int strgpu_scan_submit(strgpu_ctx *ctx, const uint8_t *seq2, uint64_t n_bases, const uint32_t *nmask, const uint32_t *xmask,
                       const strgpu_segment *segs, uint32_t n_seg, uint32_t max_len,
                       strgpu_repeat *out, int *ticket);
int strgpu_scan_wait(strgpu_ctx *ctx, int ticket);
/* submit + wait */
int strgpu_scan(strgpu_ctx *ctx, const uint8_t *seq2, uint64_t n_bases, const uint32_t *nmask, const uint32_t *xmask,
                const strgpu_segment *segs, uint32_t n_seg, uint32_t max_len, strgpu_repeat *out);
/* Uniform-read batches without descriptors (saves the 8 B/read of host-to-device traffic that whole-read descriptors
 * cost): read i occupies bases [i * stride_bases, i * stride_bases + read_len) of seq2 and is scanned with
 * proportion class `pclass`; stride_bases must be a multiple of 4 and read_len <= 160 (longer uniform reads are
 * expanded into descriptors inside the library).  With nmask != NULL every read's mask bits are inspected on the
 * device.  `extra` (may be NULL) are ordinary descriptors for the other segments of the batch (soft clips);
 * out[0 .. n_reads) receives the reads' results, out[n_reads .. n_reads + n_extra) the extra segments'. */
int strgpu_scan_reads_submit(strgpu_ctx *ctx, const uint8_t *seq2, uint32_t n_reads, uint32_t read_len, uint32_t stride_bases,
                             uint32_t pclass, const uint32_t *nmask, const uint32_t *xmask, const strgpu_segment *extra,
                             uint32_t n_extra, uint32_t extra_max_len, strgpu_repeat *out, int *ticket);

/* Device-resident variant: all pointers are device pointers on ctx's device; the kernel is enqueued on
 * `cuda_stream` (a cudaStream_t, NULL = default stream) and the call returns without synchronising.
 * A too-long segment is reported by the next strgpu_device_status(). */
int strgpu_scan_device(strgpu_ctx *ctx, const void *d_seq2, const void *d_nmask, const void *d_xmask, const void *d_segs,
                       uint32_t n_seg, uint32_t max_len, void *d_out, void *cuda_stream);
/* Device-resident variant of strgpu_scan_reads_submit (same argument meaning; d_seq2 must be 16-byte aligned for the
 * TMA-staged path, otherwise per-lane loads are used): d_out[0 .. n_reads) receives the reads' results,
 * d_out[n_reads .. n_reads + n_extra) the extra segments'. */
int strgpu_scan_reads_device(strgpu_ctx *ctx, const void *d_seq2, uint32_t n_reads, uint32_t read_len, uint32_t stride_bases,
                             uint32_t pclass, const void *d_nmask, const void *d_xmask, const void *d_extra, uint32_t n_extra,
                             uint32_t extra_max_len, void *d_out, void *cuda_stream);
/* synchronises `cuda_stream` and returns the sticky device-side status of launches since the last call */
int strgpu_device_status(strgpu_ctx *ctx, void *cuda_stream);

/* ---- cluster: the cluster loop of `strling call` (call.nim:118-130,223-235) and `strling merge`
 * (merge.nim:125-187): group by (tid, repeat), stable sort by position, cluster (cluster.nim:364),
 * bounds + filters (cluster.nim:175, callclusters.nim:52), has_per_sample_reads (merge.nim:18) ---- */

/* tread (cluster.nim:23-32) as a POD.  `sample` stands in for qname: on this path qname is only read after
 * merge.nim:121-124 has overwritten it with the sample index. */
typedef struct {
  int32_t  tid;
  uint32_t position;
  char     repeat[6];
  uint16_t flag;
  uint8_t  split;            /* Soft: left=0 right=1 both=2 none=3 none_right=4 none_left=5 (cluster.nim:14-20) */
  uint8_t  mapping_quality;
  uint8_t  repeat_count;
  uint8_t  align_length;
  int32_t  sample;
} strgpu_tread;              /* 24 bytes */

/* Bounds (cluster.nim:75-87) as a POD; `name` is always empty for discovered clusters. */
typedef struct {
  int32_t  tid;              /* -1: an unplaced bucket (call.nim:226-228): only repeat and n_reads are meaningful */
  uint32_t left;
  uint32_t left_most;
  uint32_t right;
  uint32_t right_most;
  uint32_t center_mass;
  uint16_t n_left;
  uint16_t n_right;
  uint16_t n_total;
  char     repeat[6];
  uint32_t first_read;       /* index of the cluster's first read in the (tid, repeat, position)-sorted order */
  uint32_t n_reads;          /* reads in the cluster after trim / split */
  uint32_t reserved;
} strgpu_bounds;             /* 48 bytes */

typedef struct {
  uint32_t window;           /* max_dist: opts.window (call.nim:114, merge.nim:148-152) */
  int32_t  min_support;      /* -m (call.nim:54, merge.nim:51) */
  uint16_t min_clip;         /* -c */
  uint16_t min_clip_total;   /* -t */
  uint16_t max_clip_dist;    /* uint16(0.5 * frag_dist.median(0.5)) (call.nim:232, merge.nim:181) */
  uint16_t merge_mode;       /* 1: merge.nim semantics (skip unplaced, has_per_sample_reads); 0: call.nim */
} strgpu_cluster_params;     /* 16 bytes */

/* treads: n records in `.bin` / concatenation order (host memory).  Writes up to `cap` records to `out`
 * in ascending (tid, repeat bytes, position) order -- the reference visits buckets in Nim Table hash order,
 * which only permutes output lines.  Unplaced buckets (tid == -1) come first as tid == -1 records when
 * merge_mode == 0.  STRGPU_ERR_OVERFLOW if more than `cap` records were produced (*n_out = needed). */
int strgpu_cluster(strgpu_ctx *ctx, const strgpu_tread *treads, uint32_t n, const strgpu_cluster_params *params,
                   strgpu_bounds *out, uint32_t cap, uint32_t *n_out);
/* A locus from a `-l` bed or `-b` bounds file after parse_bedline / parse_boundsline (cluster.nim:111-163):
 * the bucket key (tid, repeat) and the window [left_most, right_most] whose reads it takes. */
typedef struct {
  int32_t  tid;
  uint32_t left_most;
  uint32_t right_most;
  char     repeat[6];
  uint16_t n_left, n_right, n_total;   /* out: the counts assign_reads_locus writes into the Bounds (callclusters.nim:41-50) */
} strgpu_locus;              /* 24 bytes */

/* strgpu_cluster preceded by assign_reads_locus (callclusters.nim:14-50) for every locus in array order, as
 * merge.nim:166-168 and call.nim:189-218 do before clustering: each locus takes the not-yet-taken reads of its
 * bucket with left_most-1 <= position <= right_most, and -- like the reference (callclusters.nim:35-36) -- the
 * first remaining read after that window is dropped as well.  loci[i].n_left/n_right/n_total are filled in. */
int strgpu_cluster_loci(strgpu_ctx *ctx, const strgpu_tread *treads, uint32_t n, const strgpu_cluster_params *params,
                        strgpu_locus *loci, uint32_t n_loci, strgpu_bounds *out, uint32_t cap, uint32_t *n_out);

/* Device-resident variant: d_treads / d_out are device pointers, *d_n_out a device uint32.  The whole path (sort, chain,
 * bounds, compaction) is enqueued on `cuda_stream` and the call returns without synchronising or reading anything back;
 * every size only the device knows stays in device memory.  d_out needs room for `cap` records; records past cap are
 * dropped and counted in *d_n_out.  n must be below 2^29. */
int strgpu_cluster_device(strgpu_ctx *ctx, const void *d_treads, uint32_t n, const strgpu_cluster_params *params,
                          void *d_out, uint32_t cap, void *d_n_out, void *cuda_stream);


/* ---- sharded clustering: one process per GPU, NCCL over NVLink / NVSwitch ----------------------------------------
 * The reference's only parallelism for this stage is one `strling merge --chromosome C` process per chromosome
 * (merge.nim:52,89; pipelines/strling-joint.groovy:7-12): buckets (tid, repeat) never interact (call.nim:124-125,
 * merge.nim:125).  Here every bucket has an owner rank: each rank hands over its shard's treads, the records travel to
 * their owners, every rank runs the cluster kernels on the buckets it owns, the 48-byte cluster records are all-gathered.
 * The result on EVERY rank equals strgpu_cluster over the concatenation (rank 0's treads, then rank 1's, ...) of all shards,
 * in the same order; `first_read` indexes the owner rank's sorted records.
 * NCCL (libnccl.so.2) is loaded with dlopen by strgpu_comm_unique_id / strgpu_comm_init; nothing else in the library needs it. */
#define STRGPU_COMM_ID_BYTES 128
/* rank 0 creates the id and hands it to the other ranks by whatever means the host program has (a file, MPI, a socket) */
int  strgpu_comm_unique_id(void *id_out /* STRGPU_COMM_ID_BYTES */);
/* collective: every rank of the job calls it with the same id.  One communicator per ctx; ctx's device is the rank's GPU. */
int  strgpu_comm_init(strgpu_ctx *ctx, int rank, int world, const void *id);
int  strgpu_comm_info(const strgpu_ctx *ctx, int *rank, int *world);
void strgpu_comm_destroy(strgpu_ctx *ctx);   /* also done by strgpu_destroy */

/* Collective, device-resident, enqueued on `cuda_stream` without any host synchronisation (partition by owner, record
 * exchange, per-rank clustering with a device-side record count, all-gather, final order).
 *   n             treads of this rank's shard (d_treads, `.bin` / concatenation order)
 *   max_n         an upper bound of n that is THE SAME on every rank (it sizes the exchange slots)
 *   pair_capacity treads one rank may send to one owner; 0 = max_n / world * 1.25 + 1024 (a hash partition is that even);
 *                 max_n always fits.  Must be the same on every rank.
 *   d_out, cap    room for the cluster records of ALL ranks; one rank may contribute at most cap / world of them.
 *                 Must be the same on every rank (like max_n and pair_capacity: they size the collectives' slots).
 *   d_n_out       device uint32: records produced by all ranks together
 * A slot that was too small is reported by strgpu_comm_status (STRGPU_ERR_OVERFLOW): the output is then incomplete. */
int strgpu_cluster_sharded_device(strgpu_ctx *ctx, const void *d_treads, uint32_t n, uint32_t max_n, uint32_t pair_capacity,
                                  const strgpu_cluster_params *params, void *d_out, uint32_t cap, void *d_n_out, void *cuda_stream);
/* synchronises `cuda_stream` and reports the sticky overflow flags of the last sharded call */
int strgpu_comm_status(strgpu_ctx *ctx, void *cuda_stream);
/* Host-buffer variant (collective): copies the shard in, uses the always-sufficient pair capacity max_n, copies all ranks'
 * cluster records out. */
int strgpu_cluster_sharded(strgpu_ctx *ctx, const strgpu_tread *treads, uint32_t n, uint32_t max_n, const strgpu_cluster_params *params,
                           strgpu_bounds *out, uint32_t cap, uint32_t *n_out);

#ifdef __cplusplus
}
#endif
#endif /* STRGPU_H */
