/*
 * strling_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 *
 * A plain-C restatement of the STRling hot path (per-read repeat-unit scan +
 * STR-read clustering).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this.  The product path
 * (strling_b200/) never links or calls it.
 *
 * Parity status: PINNED by the reference's own known-answer tests
 * (tests/test_strling.nim, test_utils.nim, test_extract.nim, test_cluster.nim;
 * see tests/test_oracle_kat.py).  UNPINNED (no reference test covers them, and
 * the third-party code is not vendored): the 2-bit base order of the `kmer`
 * nimble package (C<A<T<G adopted on in-tree evidence), the code
 * given to non-ACGT bases (assumed 'A'), and Nim CountTable slot order used to
 * break ties in `largest` (emulated from memory of Nim 1.6.10 stdlib).
 *
 * All file:line citations are relative to the reference checkout.
 */
#ifndef STRLING_ORACLE_H
#define STRLING_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Soft enum, cluster.nim:14-20 */
enum { ORC_LEFT = 0, ORC_RIGHT = 1, ORC_BOTH = 2, ORC_NONE = 3, ORC_NONE_RIGHT = 4, ORC_NONE_LEFT = 5 };

/* tread, cluster.nim:23-32.  qname is replaced by an integer tag: merge.nim:121-124
 * overwrites qname with the sample index, and nothing else on the cluster path reads it. */
typedef struct {
  int32_t  tid;
  uint32_t position;
  char     repeat[6];
  uint16_t flag;
  uint8_t  split;
  uint8_t  mapping_quality;
  uint8_t  repeat_count;
  uint8_t  align_length;
  int32_t  sample;
} orc_tread;            /* 24 bytes */

/* Bounds, cluster.nim:75-87 (repeat as 6 chars, name omitted: always empty for clusters) */
typedef struct {
  int32_t  tid;
  uint32_t left;
  uint32_t left_most;
  uint32_t right;
  uint32_t right_most;
  uint32_t center_mass;
  uint16_t n_left;
  uint16_t n_right;
  uint16_t n_total;
  char     repeat[6];
  uint32_t first_read;   /* index (in the sorted order) of the cluster's first read; diagnostic */
  uint32_t n_reads;      /* reads in the (trimmed / split) cluster; diagnostic */
} orc_bounds;           /* 44 bytes */

/* ---- scan half ---- */
/* utils.nim:236-271 */
void orc_get_repeat(const char *read, int len, double proportion_repeat, char unit[6], int *repeat_count);
/* many segments: seqs is one ASCII buffer, segment i = seqs[off[i] .. off[i]+len[i]) with proportion p[i].
 * out_unit: n*6 bytes, out_count: n ints. */
void orc_get_repeat_batch(const char *seqs, const uint64_t *off, const uint32_t *len, const double *p,
                          uint64_t n, char *out_unit, int32_t *out_count);
/* utils.nim:205-211 + :197 : max multiplicity and leader code of min-rotation k-mers (leader = UINT64_MAX if no window) */
int  orc_count(const char *read, int len, int k, uint64_t *leader);
/* utils.nim:10-34 : every window's min-rotation code; returns the number of windows (writes at most cap) */
int  orc_slide_by(const char *s, int len, int k, uint64_t *out, int cap);
/* utils.nim:220-233 */
int  orc_reduce_repeat(char rep[6]);
/* utils.nim:61-80 */
void orc_min_rev_complement(char rep[6]);
/* utils.nim:304-310 */
void orc_canonical_repeat(const char in[6], char out[6]);
/* extract.nim:56-58 */
double orc_p_repeat(const orc_tread *t);
/* extract.nim:141-179 ; returns the proc's bool */
int  orc_adjust_by(orc_tread *A, const orc_tread *B, double proportion_repeat, uint8_t min_mapq,
                   int median_fragment_length, uint32_t B_position);
/* extract.nim:182-190 */
int  orc_unplaced_pair(const orc_tread *A, const orc_tread *B, double proportion_repeat, uint8_t min_mapq);
/* utils.nim:139-146 */
int  orc_median(const uint32_t frag[4096], double pct);

/* ---- cluster half ---- */
/* Nim 1.6 hashes.nim hashWangYi1 (used by CountTable[uint32]) */
uint64_t orc_hash_wangyi1(uint64_t x);
/* CountTable[uint32].largest over keys inserted in the given order (counts incremented one by one in
 * call order, as cluster.nim:194,198,292-293 do).  Returns key, *val = its count. */
uint32_t orc_counttable_largest(const uint32_t *keys_in_call_order, int n_calls, int *val, int *n_distinct);

/* cluster.nim:175-250.  reads = the cluster's reads (position-sorted).  cl_left_most/right_most = Cluster fields. */
void orc_bounds_of(const orc_tread *reads, int n, uint32_t cl_left_most, uint32_t cl_right_most,
                   uint16_t max_clip_dist, orc_bounds *out);
/* callclusters.nim:52-66 : returns 1 if the cluster passes, fills *out */
int  orc_bounds_filtered(const orc_tread *reads, int n, uint32_t cl_left_most, uint32_t cl_right_most,
                         uint16_t min_clip, uint16_t min_clip_total, uint16_t max_clip_dist, orc_bounds *out);

/* One bucket (same tid, same repeat; position-sorted).  Runs cluster.nim:364-374 (`cluster`), i.e.
 * trcluster + trim + has_anchor + split_cluster, and reports every yielded Cluster as
 * (first index into reps, n reads, left_most, right_most).  Returns the number of clusters
 * (never more than cap are written). */
int  orc_cluster_bucket(const orc_tread *reps, int n, uint32_t max_dist, int min_supporting_reads,
                        uint32_t *first, uint32_t *count, uint32_t *left_most, uint32_t *right_most, int cap);

/* The whole call.nim:118-130,223-235 / merge.nim:125-187 cluster loop over an unsorted tread array in
 * .bin / concatenation order: group by (tid, repeat), stable sort by position, cluster, bounds, filter.
 * merge_mode != 0 applies merge.nim:175-177 (skip tid<0 silently, has_per_sample_reads).
 * Buckets are visited in ascending (tid, repeat bytes) order (the reference's order is Nim Table hash
 * order, which only permutes output lines).  Unplaced buckets (tid<0) are reported through
 * unplaced_unit/unplaced_count (call.nim:226-228) when merge_mode == 0.
 * Returns number of bounds written (<= cap_bounds), or -1 if cap exceeded. */
int  orc_cluster_all(const orc_tread *treads, int n, uint32_t window, int min_support,
                     uint16_t min_clip, uint16_t min_clip_total, uint16_t max_clip_dist, int merge_mode,
                     orc_bounds *out, int cap_bounds,
                     char *unplaced_unit /* cap_unplaced*6 */, int32_t *unplaced_count, int cap_unplaced,
                     int *n_unplaced);

/* A locus from a `-l` bed / `-b` bounds file after parse_bedline (cluster.nim:111-134): key + the window to collect. */
typedef struct {
  int32_t  tid;
  uint32_t left_most;
  uint32_t right_most;
  char     repeat[6];
  uint16_t n_left, n_right, n_total;   /* outputs of assign_reads_locus (callclusters.nim:41-50) */
} orc_locus;             /* 24 bytes */

/* orc_cluster_all preceded by assign_reads_locus (callclusters.nim:14-50) for every locus in file order, as
 * merge.nim:166-168 / call.nim:189-218 do: the reads a locus takes (and the one extra read the reference drops,
 * callclusters.nim:35-36) leave their bucket before clustering.  loci[i].n_* are filled in. */
int  orc_cluster_all_loci(const orc_tread *treads, int n, orc_locus *loci, int n_loci, uint32_t window, int min_support,
                          uint16_t min_clip, uint16_t min_clip_total, uint16_t max_clip_dist, int merge_mode,
                          orc_bounds *out, int cap_bounds,
                          char *unplaced_unit, int32_t *unplaced_count, int cap_unplaced, int *n_unplaced);

#ifdef __cplusplus
}
#endif
#endif
