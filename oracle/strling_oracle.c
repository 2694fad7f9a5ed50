/*
 * strling_oracle.c -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT CODE).
 * See strling_oracle.h for the parity status.  Literal restatement of the reference algorithm:
 * same passes, same tables, same memsets, same fixed-width wraparound.
 * Build: see oracle/Makefile (-O3 -ffp-contract=off: thresholds are fp64 mul/div then trunc).
 */
#include "strling_oracle.h"
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------------------------------
 * third-party `kmer` package (brentp/nim-kmer, `kmer >= 0.2.2`, not vendored; strling.nimble:20).
 * Published algorithm: 2-bit big-endian packing; forward_add shifts a base in at the low end.
 * Base order C<A<T<G: pinned only by in-tree documentary evidence (genome_strs.nim:204 `CACGAT`,
 * `CAG` in data/hg38.STR_disease_loci.bed).  Non-ACGT -> 1 ('A'): ASSUMPTION (parity unpinned).
 * ---------------------------------------------------------------------------------------------- */
static inline uint64_t kmer_code(char b) {
  switch (b) {
    case 'C': case 'c': return 0;
    case 'T': case 't': return 2;
    case 'G': case 'g': return 3;
    default: return 1;
  }
}
static const char KMER_ALPHA[4] = {'C', 'A', 'T', 'G'};

static inline uint64_t kmer_encode(const char *s, int k) {
  uint64_t f = 0;
  for (int i = 0; i < k; i++) f = (f << 2) | kmer_code(s[i]);
  return f;
}
static inline void kmer_forward_add(uint64_t *f, char b, int k) {
  uint64_t mask = (k >= 32) ? ~0ULL : ((1ULL << (2 * k)) - 1ULL);
  *f = ((*f << 2) | kmer_code(b)) & mask;
}
static inline void kmer_decode(uint64_t e, char *out, int k) {
  for (int i = k; i > 0; i--) { out[i - 1] = KMER_ALPHA[e & 3ULL]; e >>= 2; }
}

/* utils.nim:113-117,181-203 : Seq[uint8] count table with a running leader (imax) */
typedef struct { long imax; uint8_t *A; int n; } orc_seq;
static orc_seq g_counts[7];
static int g_counts_init = 0;
static void counts_init(void) {                       /* utils.nim:181-190 */
  static const int sz[7] = {0, 0, 16, 64, 256, 1024, 4096};
  for (int i = 0; i < 7; i++) {
    g_counts[i].A = sz[i] ? (uint8_t *)calloc((size_t)sz[i], 1) : NULL;
    g_counts[i].n = sz[i];
    g_counts[i].imax = 0;                             /* Nim default-initialises imax to 0, not -1 */
  }
  g_counts_init = 1;
}
static inline void seq_clear(orc_seq *s) {            /* utils.nim:200-203 */
  if (s->imax == -1) return;
  memset(s->A, 0, (size_t)s->n);
  s->imax = -1;
}
static inline void seq_inc(orc_seq *s, uint64_t enc) { /* utils.nim:192-195 ; uint8 wraps (checks off) */
  s->A[enc] = (uint8_t)(s->A[enc] + 1);
  if (s->imax == -1 || s->A[enc] > s->A[s->imax]) s->imax = (long)enc;
}

/* utils.nim:10-34 slide_by fused with utils.nim:205-211 count */
static int count_k(const char *s, int len, int k, orc_seq *cnt) {
  seq_clear(cnt);
  if (k <= len) {
    uint64_t f = kmer_encode(s, k);
    uint64_t kmin = f;
    for (int j = 0; j < k; j++) {                     /* rotate the first k-mer */
      kmer_forward_add(&f, s[j], k);
      if (f < kmin) kmin = f;
    }
    seq_inc(cnt, kmin);
    for (int i = k; i <= (len - 1) - k + 1; i += k) { /* countup(k, s.high - k + 1, k) */
      for (int m = 0; m < k; m++) kmer_forward_add(&f, s[i + m], k);
      kmin = f;
      for (int j = 0; j < k; j++) {
        kmer_forward_add(&f, s[i + j], k);
        if (f < kmin) kmin = f;
      }
      seq_inc(cnt, kmin);
    }
  }
  if (cnt->imax == -1) return 0;
  return (int)cnt->A[cnt->imax];
}

int orc_slide_by(const char *s, int len, int k, uint64_t *out, int cap) { /* utils.nim:10-34, literal */
  int n = 0;
  if (k <= len && k > 0) {
    uint64_t f = kmer_encode(s, k);
    uint64_t kmin = f;
    for (int j = 0; j < k; j++) { kmer_forward_add(&f, s[j], k); if (f < kmin) kmin = f; }
    if (n < cap) out[n] = kmin;
    n++;
    for (int i = k; i <= (len - 1) - k + 1; i += k) {
      for (int m = 0; m < k; m++) kmer_forward_add(&f, s[i + m], k);
      kmin = f;
      for (int j = 0; j < k; j++) { kmer_forward_add(&f, s[i + j], k); if (f < kmin) kmin = f; }
      if (n < cap) out[n] = kmin;
      n++;
    }
  }
  return n;
}

int orc_count(const char *read, int len, int k, uint64_t *leader) {
  if (!g_counts_init) counts_init();
  int c = count_k(read, len, k, &g_counts[k]);
  *leader = (uint64_t)g_counts[k].imax;               /* argmax: imax.uint64 (-1 -> all ones) */
  return c;
}

/* Nim strutils.count(s, sub, overlapping=false): greedy leftmost, raw byte compare */
static int str_count(const char *s, int len, const char *sub, int k) {
  int c = 0, i = 0;
  while (i + k <= len) {
    if (memcmp(s + i, sub, (size_t)k) == 0) { c++; i += k; }
    else i++;
  }
  return c;
}

int orc_reduce_repeat(char rep[6]) {                  /* utils.nim:220-233 */
  int result = 1;
  if (rep[0] == '\0') return result;
  char seen = rep[0];
  for (int i = 1; i < 6; i++) {
    if (rep[i] == '\0') break;
    if (rep[i] != seen) return result;
  }
  for (int i = 1; i < 6; i++) {
    if (rep[i] == '\0') break;
    result++;
    rep[i] = '\0';
  }
  return result;
}

void orc_get_repeat(const char *read, int len, double p, char unit[6], int *repeat_count) { /* utils.nim:236-271 */
  if (!g_counts_init) counts_init();
  memset(unit, 0, 6);
  *repeat_count = 0;
  int nN = 0;
  for (int i = 0; i < len; i++) nN += (read[i] == 'N');
  if (nN > 20) return;
  char s[8];
  int best_score = -1;
  for (int k = 2; k <= 6; k++) {
    int count = count_k(read, len, k, &g_counts[k]);
    kmer_decode((uint64_t)g_counts[k].imax, s, k);
    int score = count * k;
    if (score <= best_score) {
      if (count < (int)((double)len * 0.12 / (double)k)) break;
      continue;
    }
    count = str_count(read, len, s, k);
    score = count * k;
    if (score < best_score) continue;
    best_score = score;
    if (count > (int)((double)len * p / (double)k)) {
      memcpy(unit, s, (size_t)k);
      *repeat_count = count;
    }
  }
  *repeat_count *= orc_reduce_repeat(unit);
}

void orc_get_repeat_batch(const char *seqs, const uint64_t *off, const uint32_t *len, const double *p,
                          uint64_t n, char *out_unit, int32_t *out_count) {
  for (uint64_t i = 0; i < n; i++) {
    int rc;
    orc_get_repeat(seqs + off[i], (int)len[i], p[i], out_unit + 6 * i, &rc);
    out_count[i] = rc;
  }
}

static char complement(char c) {                      /* utils.nim:37-47 */
  switch (c) { case 'C': return 'G'; case 'G': return 'C'; case 'A': return 'T'; case 'T': return 'A'; default: return c; }
}

void orc_min_rev_complement(char rep[6]) {            /* utils.nim:61-80 */
  char s[16];
  int l = 0;
  for (int i = 0; i < 6; i++) { if (rep[i] == 0) break; l++; }
  if (l == 0) return;                                 /* reference would index s[0..<0]; never called with empty unit */
  for (int i = 0; i < l; i++) s[l - 1 - i] = complement(rep[i]);
  for (int i = 0; i < l; i++) s[l + i] = s[i];
  /* slide_by(s & s, l): two windows, each already minimised over rotations */
  uint64_t mv = ~0ULL;
  for (int w = 0; w < 2; w++) {
    uint64_t f = kmer_encode(s + w * l, l), kmin = f;
    for (int j = 0; j < l; j++) { kmer_forward_add(&f, s[w * l + j], l); if (f < kmin) kmin = f; }
    if (kmin < mv) mv = kmin;
  }
  char ms[8];
  kmer_decode(mv, ms, l);
  for (int i = 0; i < l; i++) rep[i] = ms[i];
}

void orc_canonical_repeat(const char in[6], char out[6]) { /* utils.nim:291-310 */
  char r[6];
  memcpy(r, in, 6);
  orc_min_rev_complement(r);
  /* `<` on array[6,char]: Nim char compare is unsigned */
  int lt = 0;
  for (int i = 0; i < 6; i++) {
    if (i == 5 || r[i] != in[i]) { lt = (unsigned char)r[i] < (unsigned char)in[i]; break; }
  }
  memcpy(out, lt ? r : in, 6);
}

static inline uint8_t repeat_length(const orc_tread *t) { /* extract.nim:51-54 */
  uint8_t n = 0;
  for (int i = 0; i < 6; i++) { if (t->repeat[i] == 0) return n; n++; }
  return n;
}
double orc_p_repeat(const orc_tread *t) {             /* extract.nim:56-58 ; uint8 product wraps */
  uint8_t prod = (uint8_t)(t->repeat_count * repeat_length(t));
  uint8_t al = t->align_length > 1 ? t->align_length : 1;
  return (double)prod / (double)al;
}

#define FLAG_PROPER_PAIR 0x2
#define FLAG_REVERSE 0x10
#define FLAG_MATE_REVERSE 0x20

int orc_adjust_by(orc_tread *A, const orc_tread *B, double p, uint8_t min_mapq, int median_frag, uint32_t B_position) { /* extract.nim:141-179 */
  if (A->repeat_count == 0) return 0;
  uint32_t half = (uint32_t)((double)((float)A->align_length / 2.0f) + 0.5); /* uint32(A.align_length.float / 2'f + 0.5) */
  if (B->mapping_quality > min_mapq &&
      ((orc_p_repeat(A) > p && orc_p_repeat(B) < 0.2) ||
       (!(A->flag & FLAG_PROPER_PAIR) && A->mapping_quality < min_mapq))) {
    if (B->flag & FLAG_REVERSE) {
      A->position = B_position - (uint32_t)median_frag + (uint32_t)B->align_length + half;
      if (B->split == ORC_NONE_LEFT) A->position = B_position;
    } else {
      A->position = B_position + (uint32_t)median_frag - half;
      if (B->split == ORC_NONE_RIGHT) A->position = B_position + (uint32_t)B->align_length;
    }
    A->split = ORC_NONE;
    A->tid = B->tid;
    if (B->mapping_quality > A->mapping_quality) A->mapping_quality = B->mapping_quality;
    int should_reverse = !(A->flag & FLAG_MATE_REVERSE);   /* extract.nim:134-139 */
    if (A->flag & FLAG_REVERSE) should_reverse = !should_reverse;
    if (should_reverse) orc_min_rev_complement(A->repeat);
  } else if (A->mapping_quality >= min_mapq || (A->flag & FLAG_PROPER_PAIR)) {
    A->position += half;
    if (B->mapping_quality > A->mapping_quality) A->mapping_quality = B->mapping_quality;
  }
  return 1;
}

int orc_unplaced_pair(const orc_tread *A, const orc_tread *B, double p, uint8_t min_mapq) { /* extract.nim:182-190 */
  if (orc_p_repeat(A) > p && orc_p_repeat(B) > p) return 1;
  if (orc_p_repeat(A) > p && B->mapping_quality < min_mapq) return 1;
  if (orc_p_repeat(B) > p && A->mapping_quality < min_mapq) return 1;
  return 0;
}

int orc_median(const uint32_t frag[4096], double pct) { /* utils.nim:139-146 ; sum and count are uint32 */
  uint32_t n = 0;
  for (int i = 0; i < 4096; i++) n += frag[i];
  uint32_t count = 0;
  uint32_t target = (uint32_t)(0.5 + (double)n / (1.0 / pct));
  for (int i = 0; i < 4096; i++) {
    count += frag[i];
    if (count >= target) return i;
  }
  return 4096;
}

/* ------------------------------------------------------------------------------------------------
 * Nim 1.6.10 stdlib pieces the cluster path leans on (not in the reference tree; restated from the
 * published stdlib: lib/pure/hashes.nim hashWangYi1, lib/pure/collections/tables.nim CountTable,
 * hashcommon.nim mustRehash/nextTry/slotsNeeded).  PARITY UNPINNED: no reference test has a tie.
 * ---------------------------------------------------------------------------------------------- */
static inline uint64_t hi_xor_lo(uint64_t a, uint64_t b) {
  __uint128_t r = (__uint128_t)a * b;
  return (uint64_t)(r >> 64) ^ (uint64_t)r;
}
uint64_t orc_hash_wangyi1(uint64_t x) {
  const uint64_t P0 = 0xa0761d6478bd642fULL, P1 = 0xe7037ed1a0b428dbULL, P58 = 0xeb44accab455d165ULL ^ 8ULL;
  return hi_xor_lo(hi_xor_lo(P0, x ^ P1), P58);
}

typedef struct { uint32_t *key; int *val; int cap; int counter; } count_table;
static void ct_init(count_table *t) {                 /* initCountTable(8): slotsNeeded(8) = nextPowerOfTwo(8*3 div 2 + 4) = 16 */
  t->cap = 16; t->counter = 0;
  t->key = (uint32_t *)calloc((size_t)t->cap, sizeof(uint32_t));
  t->val = (int *)calloc((size_t)t->cap, sizeof(int));
}
static void ct_free(count_table *t) { free(t->key); free(t->val); }
static void ct_raw_insert(uint32_t *key, int *val, int cap, uint32_t k, int v) {
  uint64_t h = orc_hash_wangyi1((uint64_t)k) & (uint64_t)(cap - 1);
  while (val[h] != 0) h = (h + 1) & (uint64_t)(cap - 1);
  key[h] = k; val[h] = v;
}
static void ct_inc(count_table *t, uint32_t k) {
  uint64_t h = orc_hash_wangyi1((uint64_t)k) & (uint64_t)(t->cap - 1);
  while (t->val[h] != 0) {
    if (t->key[h] == k) { t->val[h]++; return; }
    h = (h + 1) & (uint64_t)(t->cap - 1);
  }
  /* insertImpl: mustRehash -> enlarge (x2, re-insert in old slot order) -> rawInsert -> inc counter */
  if ((t->cap * 2 < t->counter * 3) || (t->cap - t->counter < 4)) {
    int ncap = t->cap * 2;
    uint32_t *nk = (uint32_t *)calloc((size_t)ncap, sizeof(uint32_t));
    int *nv = (int *)calloc((size_t)ncap, sizeof(int));
    for (int i = 0; i < t->cap; i++) if (t->val[i] != 0) ct_raw_insert(nk, nv, ncap, t->key[i], t->val[i]);
    free(t->key); free(t->val);
    t->key = nk; t->val = nv; t->cap = ncap;
  }
  ct_raw_insert(t->key, t->val, t->cap, k, 1);
  t->counter++;
}
static uint32_t ct_largest(const count_table *t, int *val) { /* first slot holding the maximum */
  int mi = 0;
  for (int h = 1; h < t->cap; h++) if (t->val[mi] < t->val[h]) mi = h;
  *val = t->val[mi];
  return t->key[mi];
}

uint32_t orc_counttable_largest(const uint32_t *keys, int n_calls, int *val, int *n_distinct) {
  count_table t; ct_init(&t);
  for (int i = 0; i < n_calls; i++) ct_inc(&t, keys[i]);
  uint32_t k = ct_largest(&t, val);
  *n_distinct = t.counter;
  ct_free(&t);
  return k;
}

/* ------------------------------------------------------------------------------------------------
 * cluster half
 * ---------------------------------------------------------------------------------------------- */
void orc_bounds_of(const orc_tread *reads, int n, uint32_t cl_left_most, uint32_t cl_right_most,
                   uint16_t max_clip_dist, orc_bounds *b) { /* cluster.nim:175-250 */
  memset(b, 0, sizeof(*b));
  count_table lefts, rights; ct_init(&lefts); ct_init(&rights);
  memcpy(b->repeat, reads[0].repeat, 6);
  b->tid = reads[0].tid;
  b->n_reads = (uint32_t)n;
  b->center_mass = reads[n / 2].position;              /* posns[int(posns.len / 2)] */
  uint32_t pmin = reads[0].position, pmax = reads[0].position;
  for (int i = 0; i < n; i++) {
    const orc_tread *r = &reads[i];
    if (r->position < pmin) pmin = r->position;
    if (r->position > pmax) pmax = r->position;
    /* int32 casts and adds wrap (checks off) */
    int32_t pos = (int32_t)r->position;
    int32_t hi = (int32_t)((uint32_t)(int32_t)b->center_mass + (uint32_t)max_clip_dist);
    int32_t lo = (int32_t)((uint32_t)(int32_t)b->center_mass - (uint32_t)max_clip_dist);
    if (r->split == ORC_LEFT && pos < hi) { ct_inc(&lefts, r->position); b->n_left++; b->n_total++; }
    else if (r->split == ORC_RIGHT && pos > lo) { ct_inc(&rights, r->position); b->n_right++; b->n_total++; }
    else b->n_total++;
  }
  if (lefts.counter > 0) { int v; uint32_t k = ct_largest(&lefts, &v); if (v > 1) b->left = k; }
  if (rights.counter > 0) { int v; uint32_t k = ct_largest(&rights, &v); if (v > 1) b->right = k; }
  ct_free(&lefts); ct_free(&rights);
  if (b->left == 0) b->left = b->center_mass;          /* posns.len > 0 branch, cluster.nim:213-217 */
  if (b->right == 0) b->right = b->left + 1;
  if (b->left >= b->right) {
    if (b->n_left > 0 && b->n_right > 0) { uint32_t t = b->left; b->left = b->right; b->right = t; }
    else b->left = b->right - 1;
  }
  b->left_most = ((long)cl_left_most > 0) ? cl_left_most : pmin;
  b->right_most = ((long)cl_right_most > 0) ? cl_right_most : pmax;
  if (b->left_most > b->left) b->left_most = b->left;
  if (b->right_most < b->right) b->right_most = b->right;
}

int orc_bounds_filtered(const orc_tread *reads, int n, uint32_t cl_left_most, uint32_t cl_right_most,
                        uint16_t min_clip, uint16_t min_clip_total, uint16_t max_clip_dist, orc_bounds *out) { /* callclusters.nim:52-66 */
  if (n >= 65535) return 0;
  orc_bounds_of(reads, n, cl_left_most, cl_right_most, max_clip_dist, out);
  if (out->right - out->left > 1000u) return 0;
  if (out->n_left < min_clip) return 0;
  if (out->n_right < min_clip) return 0;
  if ((uint16_t)(out->n_right + out->n_left) < min_clip_total) return 0;
  return 1;
}

/* growable Cluster of indexes into reps: reads are always a contiguous run [lo, hi) of the bucket */
typedef struct { int lo, hi; uint32_t left_most, right_most; } cl_t;

static inline uint32_t posmed(const orc_tread *reps, const cl_t *c) { /* cluster.nim:59-62 */
  int n = c->hi - c->lo;
  int m = n < 9 ? n : 9;
  int mid = (int)((double)m / 2.0 - 0.5);
  return reps[c->lo + mid].position;
}
static void cl_trim(const orc_tread *reps, cl_t *c, uint32_t max_dist) { /* cluster.nim:252-257 */
  if (c->hi - c->lo == 0) return;
  long lo_l = (long)posmed(reps, c) - (long)max_dist;
  uint32_t lo = (uint32_t)(lo_l > 0 ? lo_l : 0);
  while (c->hi - c->lo > 1 && reps[c->lo].position < lo) c->lo++;
}
static int has_anchor(const orc_tread *reps, const cl_t *c) { /* cluster.nim:275-281 */
  for (int i = c->lo; i < c->hi; i++) if (reps[i].split == ORC_NONE) return 1;
  return 0;
}

typedef struct { uint32_t *first, *count, *left_most, *right_most; int cap, n; } cl_sink;
static void sink_emit(cl_sink *s, int lo, int hi, uint32_t lm, uint32_t rm) {
  if (s->n < s->cap) { s->first[s->n] = (uint32_t)lo; s->count[s->n] = (uint32_t)(hi - lo); s->left_most[s->n] = lm; s->right_most[s->n] = rm; }
  s->n++;
}

static void split_cluster(const orc_tread *reps, const cl_t *c, int min_supporting, cl_sink *sink) { /* cluster.nim:283-320 */
  count_table lefts, rights; ct_init(&lefts); ct_init(&rights);
  for (int i = c->lo; i < c->hi; i++) {
    if (reps[i].split == ORC_LEFT) ct_inc(&lefts, reps[i].position);
    else if (reps[i].split == ORC_RIGHT) ct_inc(&rights, reps[i].position);
  }
  if (rights.counter == 0 || lefts.counter == 0) {
    sink_emit(sink, c->lo, c->hi, c->left_most, c->right_most);
  } else {
    int rv, lv;
    uint32_t rk = ct_largest(&rights, &rv);
    uint32_t lk = ct_largest(&lefts, &lv);
    if (rk < lk && rv >= min_supporting && lv >= min_supporting &&
        (double)lv / (double)lefts.counter > 0.5 && (double)rv / (double)rights.counter > 0.5) {
      uint32_t mid = (uint32_t)(0.5 + ((double)rk + (double)lk) / 2.0);
      int m = c->lo;
      while (m < c->hi && reps[m].position < mid) m++;   /* sorted: `position < mid` is a prefix */
      sink_emit(sink, c->lo, m, 0, mid - 1);             /* c1: right_most = mid-1, left_most default 0 */
      sink_emit(sink, m, c->hi, mid, 0);                 /* c2: left_most = mid, right_most default 0 */
    } else {
      sink_emit(sink, c->lo, c->hi, c->left_most, c->right_most);
    }
  }
  ct_free(&lefts); ct_free(&rights);
}

static void finish(const orc_tread *reps, cl_t *c, uint32_t max_dist, int min_supporting, cl_sink *sink) { /* cluster.nim:342-349,354-362 */
  cl_trim(reps, c, max_dist + 100u);
  uint32_t pm = posmed(reps, c);
  uint32_t last = reps[c->hi - 1].position, firstp = reps[c->lo].position;
  uint32_t a = pm + max_dist;                            /* uint32 wrap */
  uint32_t b = pm - max_dist;                            /* uint32 wrap */
  c->right_most = last > a ? last : a;
  c->left_most = firstp < b ? firstp : b;
  if (c->hi - c->lo >= min_supporting && has_anchor(reps, c)) split_cluster(reps, c, min_supporting, sink);
}

int orc_cluster_bucket(const orc_tread *reps, int n, uint32_t max_dist, int min_supporting,
                       uint32_t *first, uint32_t *count, uint32_t *left_most, uint32_t *right_most, int cap) { /* cluster.nim:323-374 */
  cl_sink sink = {first, count, left_most, right_most, cap, 0};
  if (n <= 0) return 0;
  if (reps[0].tid < 0) {                                 /* cluster.nim:369-371 : Cluster(reads: reps) */
    sink_emit(&sink, 0, n, 0, 0);
    return sink.n;
  }
  int i = 0;
  cl_t c = {0, 0, 0, 0};
  while (i < n) {
    c.lo = i; c.hi = i + 1; c.left_most = 0; c.right_most = 0;
    i += 1;
    for (int j = i; j < n; j++) {
      if (reps[j].position <= posmed(reps, &c) + max_dist + 100u) {   /* uint32 wrap */
        c.hi = j + 1;
        i = j + 1;
        continue;
      }
      finish(reps, &c, max_dist, min_supporting, &sink);
      break;
    }
  }
  finish(reps, &c, max_dist, min_supporting, &sink);     /* tail cluster, cluster.nim:354-362 */
  return sink.n;
}

/* group + stable sort: (tid, repeat bytes) then position, ties keep input order */
typedef struct { orc_tread t; uint32_t idx; } sort_rec;
static int cmp_rec(const void *a, const void *b) {
  const sort_rec *x = (const sort_rec *)a, *y = (const sort_rec *)b;
  if (x->t.tid != y->t.tid) return x->t.tid < y->t.tid ? -1 : 1;
  int m = memcmp(x->t.repeat, y->t.repeat, 6);
  if (m) return m;
  if (x->t.position != y->t.position) return x->t.position < y->t.position ? -1 : 1;
  return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

static int cmp_i32(const void *a, const void *b) {
  int32_t x = *(const int32_t *)a, y = *(const int32_t *)b;
  return x < y ? -1 : (x > y);
}
static int has_per_sample_reads(const orc_tread *reads, int n, int supporting) { /* merge.nim:18-25 */
  /* largest per-sample count; value only, so no tie-break question */
  int32_t *s = (int32_t *)malloc((size_t)n * sizeof(int32_t));
  for (int i = 0; i < n; i++) s[i] = reads[i].sample;
  qsort(s, (size_t)n, sizeof(int32_t), cmp_i32);
  int best = 0, run = 0;
  for (int i = 0; i < n; i++) {
    run = (i > 0 && s[i] == s[i - 1]) ? run + 1 : 1;
    if (run > best) best = run;
  }
  free(s);
  return best >= supporting;
}

/* callclusters.nim:14-50 on one bucket held as a growable array (literal: slices and re-concatenation) */
static void assign_reads_locus(orc_locus *locus, orc_tread *trs, int *len) {
  int n = *len;
  uint32_t left_most = locus->left_most == 0 ? 0u : locus->left_most - 1u;
  int li = 0, ri = 0;
  while (li < n && trs[li].position < left_most) li++;          /* lowerBound(trs, left_most) */
  while (ri < n && trs[ri].position <= locus->right_most) ri++;  /* upperBound(trs, right_most) */
  locus->n_total = 0; locus->n_left = 0; locus->n_right = 0;
  if (n > 0) {
    for (int i = li; i < ri; i++) {                              /* result = trs[li..<ri] */
      locus->n_total++;
      if (trs[i].split == ORC_RIGHT) locus->n_right++;
      else if (trs[i].split == ORC_LEFT) locus->n_left++;
    }
    int m = li < n ? li : n;                                     /* trs[0..<li] */
    if (ri < n - 1) {                                            /* if ri < trs.high: add trs[ri+1..high] */
      for (int i = ri + 1; i < n; i++) trs[m++] = trs[i];
    }
    if (ri < li) m = n;  /* empty slice with ri < li cannot happen for left_most <= right_most; keep all if it did */
    *len = m;
  }
}

int orc_cluster_all_loci(const orc_tread *treads, int n, orc_locus *loci, int n_loci, uint32_t window, int min_support,
                         uint16_t min_clip, uint16_t min_clip_total, uint16_t max_clip_dist, int merge_mode,
                         orc_bounds *out, int cap_bounds,
                         char *unplaced_unit, int32_t *unplaced_count, int cap_unplaced, int *n_unplaced) {
  if (n_unplaced) *n_unplaced = 0;
  for (int j = 0; j < n_loci; j++) { loci[j].n_left = loci[j].n_right = loci[j].n_total = 0; }
  if (n <= 0) return 0;
  sort_rec *recs = (sort_rec *)malloc((size_t)n * sizeof(sort_rec));
  orc_tread *sorted = (orc_tread *)malloc((size_t)n * sizeof(orc_tread));
  for (int i = 0; i < n; i++) { recs[i].t = treads[i]; recs[i].idx = (uint32_t)i; }
  qsort(recs, (size_t)n, sizeof(sort_rec), cmp_rec);
  for (int i = 0; i < n; i++) sorted[i] = recs[i].t;
  free(recs);
  /* bucket table: start / length of every (tid, repeat) run; loci edit their bucket in place */
  int nbk = 0;
  int *bs = (int *)malloc((size_t)n * sizeof(int)), *bl = (int *)malloc((size_t)n * sizeof(int));
  for (int s0 = 0; s0 < n;) {
    int e = s0 + 1;
    while (e < n && sorted[e].tid == sorted[s0].tid && memcmp(sorted[e].repeat, sorted[s0].repeat, 6) == 0) e++;
    bs[nbk] = s0; bl[nbk] = e - s0; nbk++;
    s0 = e;
  }
  for (int j = 0; j < n_loci; j++) {
    for (int b = 0; b < nbk; b++) {
      if (sorted[bs[b]].tid == loci[j].tid && memcmp(sorted[bs[b]].repeat, loci[j].repeat, 6) == 0 && bl[b] > 0) {
        assign_reads_locus(&loci[j], sorted + bs[b], &bl[b]);
        break;
      }
    }
  }
  /* compact the surviving reads (bucket order is unchanged) and cluster them */
  orc_tread *kept = (orc_tread *)malloc((size_t)n * sizeof(orc_tread));
  int nk = 0;
  for (int b = 0; b < nbk; b++) for (int i = 0; i < bl[b]; i++) kept[nk++] = sorted[bs[b] + i];
  free(bs); free(bl); free(sorted);
  int r = orc_cluster_all(kept, nk, window, min_support, min_clip, min_clip_total, max_clip_dist, merge_mode, out, cap_bounds,
                          unplaced_unit, unplaced_count, cap_unplaced, n_unplaced);
  free(kept);
  return r;
}

int orc_cluster_all(const orc_tread *treads, int n, uint32_t window, int min_support,
                    uint16_t min_clip, uint16_t min_clip_total, uint16_t max_clip_dist, int merge_mode,
                    orc_bounds *out, int cap_bounds,
                    char *unplaced_unit, int32_t *unplaced_count, int cap_unplaced, int *n_unplaced) {
  if (n_unplaced) *n_unplaced = 0;
  if (n <= 0) return 0;
  sort_rec *recs = (sort_rec *)malloc((size_t)n * sizeof(sort_rec));
  orc_tread *sorted = (orc_tread *)malloc((size_t)n * sizeof(orc_tread));
  for (int i = 0; i < n; i++) { recs[i].t = treads[i]; recs[i].idx = (uint32_t)i; }
  qsort(recs, (size_t)n, sizeof(sort_rec), cmp_rec);
  for (int i = 0; i < n; i++) sorted[i] = recs[i].t;
  free(recs);
  int nb = 0, overflow = 0;
  int cap = n + 1;
  uint32_t *first = (uint32_t *)malloc((size_t)cap * 4), *count = (uint32_t *)malloc((size_t)cap * 4);
  uint32_t *lm = (uint32_t *)malloc((size_t)cap * 4), *rm = (uint32_t *)malloc((size_t)cap * 4);
  int s = 0;
  while (s < n) {
    int e = s + 1;
    while (e < n && sorted[e].tid == sorted[s].tid && memcmp(sorted[e].repeat, sorted[s].repeat, 6) == 0) e++;
    int nc = orc_cluster_bucket(sorted + s, e - s, window, min_support, first, count, lm, rm, cap);
    for (int ci = 0; ci < nc; ci++) {
      const orc_tread *cr = sorted + s + first[ci];
      int cn = (int)count[ci];
      if (cr[0].tid == -1) {                              /* call.nim:226-228 / merge.nim:175-176 */
        if (!merge_mode && n_unplaced) {
          if (*n_unplaced < cap_unplaced) { memcpy(unplaced_unit + 6 * (*n_unplaced), cr[0].repeat, 6); unplaced_count[*n_unplaced] = cn; }
          (*n_unplaced)++;
        }
        continue;
      }
      if (merge_mode && !has_per_sample_reads(cr, cn, min_support)) continue;
      orc_bounds b;
      if (!orc_bounds_filtered(cr, cn, lm[ci], rm[ci], min_clip, min_clip_total, max_clip_dist, &b)) continue;
      b.first_read = (uint32_t)s + first[ci];
      if (nb < cap_bounds) out[nb] = b; else overflow = 1;
      nb++;
    }
    s = e;
  }
  free(first); free(count); free(lm); free(rm); free(sorted);
  return overflow ? -1 : nb;
}
