/* freddy_oracle.c — CPU ORACLE. TEST INFRASTRUCTURE ONLY (see freddy_oracle.h).
 *
 * Restates, function by function, the search arithmetic of the reference
 * PostgreSQL extension over in-memory arrays.  "ref:" comments give the
 * reference file:line (relative to /root/reference/freddy_extension/) that
 * the code below follows.  Written from scratch; no reference source is
 * copied.  Build: gcc -O2 -ffp-contract=off (no -march, hence no FMA), which
 * is what PGXS gives the reference.
 */
#include "freddy_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ------------------------------------------------------------------------- */
/* kernel-level restatements                                                  */
/* ------------------------------------------------------------------------- */

/* ref: index_utils.c:500-508 squareDistance — sequential fp32 sum of
 * (a-b)*(a-b), accumulator starts at 0. */
float fo_square_distance(const float* v1, const float* v2, int n) {
  float acc = 0;
  for (int i = 0; i < n; i++) {
    float diff = v1[i] - v2[i];
    float sq = diff * diff;
    acc = acc + sq;
  }
  return acc;
}

/* ref: index_utils.c:19-33 updateTopK.  Scan from the tail for the last slot
 * whose distance is strictly smaller, insert just behind it, shift the rest
 * one slot right (the old last entry falls off).  Callers gate on
 * `distance < tk[k-1].distance`.  NOTE (kept from the reference): when the
 * gate is violated the write lands on tk[k] — callers here never do that. */
void fo_update_topk(FoTopKEntry* tk, float distance, int id, int k) {
  int slot = k;
  while (slot > 0 && !(tk[slot - 1].distance < distance)) slot--;
  if (slot >= k) return; /* gate violated: the reference would write tk[k] (out of bounds) */
  for (int j = k - 1; j > slot; j--) tk[j] = tk[j - 1];
  tk[slot].distance = distance;
  tk[slot].id = id;
}

/* ref: index_utils.c:66-72 initTopK */
void fo_init_topk(FoTopKEntry* tk, int k, float max_dist) {
  for (int i = 0; i < k; i++) {
    tk[i].distance = max_dist;
    tk[i].id = -1;
  }
}

/* ref: index_utils.c:445-455 getPrecomputedDistances.  The reference walks a
 * list of (pos, code, vector) codebook rows and writes preDists[pos*K+code];
 * with a dense [m][K][sub] codebook that is the double loop below (the result
 * does not depend on row order). */
void fo_precomputed_distances(float* pre_dists, int positions, int codes, int sub,
                              const float* query, const float* codebook) {
  for (int pos = 0; pos < positions; pos++) {
    for (int code = 0; code < codes; code++) {
      const float* cw = codebook + ((size_t)pos * codes + code) * sub;
      pre_dists[pos * codes + code] = fo_square_distance(query + pos * sub, cw, sub);
    }
  }
}

/* ref: index_utils.c:1126-1133 computePQDistanceInt16 (and the inline copies
 * at freddy.c:364-368, :958-965, :1135-1138): left-to-right fp32 sum from 0. */
float fo_pq_distance_int16(const float* pre_dists, const int16_t* codes, int positions, int ncodes) {
  float dist = 0;
  for (int l = 0; l < positions; l++) dist = dist + pre_dists[ncodes * l + codes[l]];
  return dist;
}

/* ref: freddy.c:401-408 / output_utils.c:8-28 — distances leave every SRF as
 * snprintf(buf,16,"%f") text and re-enter SQL through float4in. */
float fo_round_through_text(float distance) {
  char buf[16];
  snprintf(buf, sizeof buf, "%f", distance);
  return strtof(buf, NULL);
}

/* ------------------------------------------------------------------------- */
/* index image helpers (not in the reference: stand in for the SPI fetch)     */
/* ------------------------------------------------------------------------- */

int fo_index_prepare(FoIndex* ix) {
  ix->list_offsets = NULL;
  ix->list_rows = NULL;
  ix->list_codes = NULL;
  ix->list_ids = NULL;
  if (ix->C <= 0 || ix->coarse_ids == NULL) return 0;
  ix->list_offsets = calloc((size_t)ix->C + 1, sizeof(int32_t));
  ix->list_rows = malloc(sizeof(int32_t) * (size_t)(ix->N > 0 ? ix->N : 1));
  if (!ix->list_offsets || !ix->list_rows) return -1;
  for (int r = 0; r < ix->N; r++) {
    int c = ix->coarse_ids[r];
    if (c < 0 || c >= ix->C) return -2;
    ix->list_offsets[c + 1]++;
  }
  for (int c = 0; c < ix->C; c++) ix->list_offsets[c + 1] += ix->list_offsets[c];
  int32_t* cursor = malloc(sizeof(int32_t) * (size_t)ix->C);
  memcpy(cursor, ix->list_offsets, sizeof(int32_t) * (size_t)ix->C);
  for (int r = 0; r < ix->N; r++) ix->list_rows[cursor[ix->coarse_ids[r]]++] = r; /* stays id-ascending */
  free(cursor);
  /* list-contiguous copies so that the CPU baseline streams each list instead of chasing table rows */
  ix->list_codes = malloc(sizeof(int16_t) * (size_t)(ix->N > 0 ? ix->N : 1) * ix->m);
  ix->list_ids = malloc(sizeof(int32_t) * (size_t)(ix->N > 0 ? ix->N : 1));
  if (!ix->list_codes || !ix->list_ids) return -1;
  for (int s = 0; s < ix->N; s++) {
    const int r = ix->list_rows[s];
    memcpy(ix->list_codes + (size_t)s * ix->m, ix->codes + (size_t)r * ix->m, sizeof(int16_t) * (size_t)ix->m);
    ix->list_ids[s] = ix->ids[r];
  }
  return 0;
}

void fo_index_release(FoIndex* ix) {
  free(ix->list_offsets);
  free(ix->list_rows);
  free(ix->list_codes);
  free(ix->list_ids);
  ix->list_offsets = NULL;
  ix->list_rows = NULL;
  ix->list_codes = NULL;
  ix->list_ids = NULL;
}

/* ------------------------------------------------------------------------- */
/* ivfadc_search                                                              */
/* ------------------------------------------------------------------------- */

/* ref: freddy.c:247-378 (first-call body of ivfadc_search), parameter w from
 * get_w() (freddy.c:229).  The SQL `SELECT id, vector, coarse_id FROM fine
 * WHERE coarse_id IN (...)` (freddy.c:324-338) returns heap order = ascending
 * id; here that is a w-way merge of the probed lists by row number. */
/* scratch the reference pallocs per call (freddy.c:296-311); the multi-threaded
 * runner keeps one per thread so that 100+ threads do not serialise on mmap */
typedef struct { unsigned char* blacklisted; FoTopKEntry* sel; float* residual; float* luts; int* cursor; } FoScratch;

static int scratch_alloc(FoScratch* s, const FoIndex* ix, int w) {
  s->blacklisted = malloc((size_t)(ix->C > 0 ? ix->C : 1));
  s->sel = malloc(sizeof(FoTopKEntry) * (size_t)w);
  s->residual = malloc(sizeof(float) * (size_t)ix->d);
  s->luts = malloc(sizeof(float) * (size_t)w * ix->m * ix->K);
  s->cursor = malloc(sizeof(int) * (size_t)w);
  return (s->blacklisted && s->sel && s->residual && s->luts && s->cursor) ? 0 : -1;
}
static void scratch_free(FoScratch* s) {
  free(s->blacklisted); free(s->sel); free(s->residual); free(s->luts); free(s->cursor);
}

static int ivfadc_search_ws(const FoIndex* ix, const float* query, int k, int w,
                            FoTopKEntry* topk, int64_t* stats, FoScratch* ws) {
  const float MAX_DIST = 1000;                       /* ref: freddy.c:184 */
  const int d = ix->d, m = ix->m, K = ix->K, C = ix->C, sub = d / m;
  int rc = 0;
  int found = 0;                                     /* ref: :255 foundInstances */
  int64_t rows_scanned = 0, rounds = 0;
  unsigned char* blacklisted = ws->blacklisted;      /* ref: :256, index_utils.c:157-176 */
  FoTopKEntry* sel = ws->sel;
  float* residual = ws->residual;
  float* luts = ws->luts;
  int* cursor = ws->cursor;
  int n_blacklisted = 0;
  memset(blacklisted, 0, (size_t)C);

  fo_init_topk(topk, k, MAX_DIST);                   /* ref: :258-259 */
  float max_dist = MAX_DIST;                         /* ref: :260 */

  while (found < k) {                                /* ref: :262 */
    rounds++;
    if (C - n_blacklisted < w) { rc = -1; break; }   /* reference would index cq[-1] */
    float min_dist = 1000.0f;                        /* ref: :266 */
    for (int i = 0; i < w; i++) { sel[i].distance = 100.0f; sel[i].id = -1; } /* ref: :268-271 */
    for (int i = 0; i < C; i++) {                    /* ref: :272-283 */
      if (blacklisted[i]) continue;
      float dist = fo_square_distance(query, ix->coarse + (size_t)i * d, d);
      if (dist < min_dist) {
        if (!(dist < 100.0f)) { rc = -2; break; }    /* reference would write sel[w] */
        fo_update_topk(sel, dist, i, w);
        min_dist = sel[w - 1].distance;
      }
    }
    if (rc) break;
    for (int j = 0; j < w; j++) { blacklisted[sel[j].id] = 1; n_blacklisted++; } /* ref: :289-293 */

    for (int j = 0; j < w; j++) {                    /* ref: :296-314 */
      const float* cvec = ix->coarse + (size_t)sel[j].id * d;
      for (int t = 0; t < d; t++) residual[t] = query[t] - cvec[t];
      fo_precomputed_distances(luts + (size_t)j * m * K, m, K, sub, residual, ix->codebook);
    }

    /* ref: :324-377 — rows with coarse_id in sel, ascending id */
    int n_rows = 0;
    for (int j = 0; j < w; j++) {
      cursor[j] = ix->list_offsets[sel[j].id];
      n_rows += ix->list_offsets[sel[j].id + 1] - ix->list_offsets[sel[j].id];
    }
    for (int r = 0; r < n_rows; r++) {
      int best = -1, best_row = 0;
      for (int j = 0; j < w; j++) {
        if (cursor[j] < ix->list_offsets[sel[j].id + 1]) {
          int row = ix->list_rows[cursor[j]];
          if (best < 0 || row < best_row) { best = j; best_row = row; }
        }
      }
      const int slot = cursor[best]++;
      const int16_t* codes = ix->list_codes + (size_t)slot * m;
      const float* lut = luts + (size_t)best * m * K;
      float dist = 0;                                /* ref: :364-368 */
      for (int l = 0; l < m; l++) dist = dist + lut[l * K + codes[l]];
      if (dist < max_dist) {                         /* ref: :369-372 */
        fo_update_topk(topk, dist, ix->list_ids[slot], k);
        max_dist = topk[k - 1].distance;
      }
    }
    rows_scanned += n_rows;
    found += n_rows;                                 /* ref: :377 */
  }
  if (stats) { stats[0] = rows_scanned; stats[1] = rounds; }
  return rc;
}

int fo_ivfadc_search(const FoIndex* ix, const float* query, int k, int w,
                     FoTopKEntry* topk, int64_t* stats) {
  FoScratch ws;
  if (scratch_alloc(&ws, ix, w)) return -3;
  int rc = ivfadc_search_ws(ix, query, k, w, topk, stats, &ws);
  scratch_free(&ws);
  return rc;
}

/* ------------------------------------------------------------------------- */
/* flat PQ                                                                    */
/* ------------------------------------------------------------------------- */

/* ref: freddy.c:74-134 pq_search: LUT on the raw query, every row of
 * pq_quantization in table order, sentinel 100.0 (freddy.c:90-92). */
int fo_pq_search(const FoIndex* ix, const float* query, int k, FoTopKEntry* topk) {
  const int m = ix->m, K = ix->K, sub = ix->d / m;
  float* lut = malloc(sizeof(float) * (size_t)m * K);
  fo_precomputed_distances(lut, m, K, sub, query, ix->codebook);
  fo_init_topk(topk, k, 100.0f);
  float max_dist = 100.0f;
  for (int r = 0; r < ix->N; r++) {
    float dist = fo_pq_distance_int16(lut, ix->codes + (size_t)r * m, m, K);
    if (dist < max_dist) {
      fo_update_topk(topk, dist, ix->ids[r], k);
      max_dist = topk[k - 1].distance;
    }
  }
  free(lut);
  return 0;
}

static int cmp_i32(const void* a, const void* b) {
  int32_t x = *(const int32_t*)a, y = *(const int32_t*)b;
  return (x > y) - (x < y);
}

/* Rows selected by `WHERE id IN (targets)` in TABLE order, every matching row once (duplicates and unknown ids
 * in the IN-list select nothing extra; an id carried by several rows selects them all).  Strictly ascending id
 * columns (the bulk-loaded tables) take the fast path: one binary search per listed id. */
static int select_target_rows(const FoIndex* ix, const int32_t* targets, int n_targets, int32_t** rows_out) {
  int32_t* sorted = malloc(sizeof(int32_t) * (size_t)(n_targets > 0 ? n_targets : 1));
  memcpy(sorted, targets, sizeof(int32_t) * (size_t)n_targets);
  qsort(sorted, (size_t)n_targets, sizeof(int32_t), cmp_i32);
  int ascending = 1;
  for (int64_t r = 1; r < ix->N && ascending; r++) ascending = ix->ids[r - 1] < ix->ids[r];
  int32_t* rows;
  int n = 0;
  if (ascending) {
    rows = malloc(sizeof(int32_t) * (size_t)(n_targets > 0 ? n_targets : 1));
    for (int i = 0; i < n_targets; i++) {
      if (i > 0 && sorted[i] == sorted[i - 1]) continue;
      int lo = 0, hi = ix->N - 1;
      while (lo <= hi) {
        int mid = lo + (hi - lo) / 2;
        if (ix->ids[mid] < sorted[i]) lo = mid + 1;
        else if (ix->ids[mid] > sorted[i]) hi = mid - 1;
        else { rows[n++] = mid; break; }
      }
    }
  } else {
    rows = malloc(sizeof(int32_t) * (size_t)(ix->N > 0 ? ix->N : 1));
    for (int64_t r = 0; r < ix->N; r++) {
      int lo = 0, hi = n_targets - 1, hit = 0;
      while (lo <= hi && !hit) {
        int mid = lo + (hi - lo) / 2;
        if (sorted[mid] < ix->ids[r]) lo = mid + 1;
        else if (sorted[mid] > ix->ids[r]) hi = mid - 1;
        else hit = 1;
      }
      if (hit) rows[n++] = (int32_t)r;
    }
  }
  free(sorted);
  *rows_out = rows;
  return n;
}

/* ref: freddy.c:1070-1143 pq_search_in: sentinel 1000.0 (freddy.c:1094-1098) */
int fo_pq_search_in(const FoIndex* ix, const float* query, int k,
                    const int32_t* targets, int n_targets, FoTopKEntry* topk) {
  return fo_pq_search_in_batch(ix, query, 1, k, targets, n_targets, 0, topk);
}

/* ref: freddy.c:514-631 pq_search_in_batch.  use_target_lists only changes
 * the loop nest (row-major :600-608 vs query-major :613-631); each query sees
 * the rows in the same order either way.  Both nests are kept so the oracle
 * exercises what the reference executes. */
int fo_pq_search_in_batch(const FoIndex* ix, const float* queries, int nq, int k,
                          const int32_t* targets, int n_targets, int use_target_lists,
                          FoTopKEntry* topks) {
  const float MAX_DIST = 1000.0f;                    /* ref: freddy.c:415 */
  const int d = ix->d, m = ix->m, K = ix->K, sub = d / m;
  float* luts = malloc(sizeof(float) * (size_t)nq * m * K);
  float* max_dists = malloc(sizeof(float) * (size_t)(nq > 0 ? nq : 1));
  for (int i = 0; i < nq; i++) {                     /* ref: :517-525 */
    fo_init_topk(topks + (size_t)i * k, k, MAX_DIST);
    max_dists[i] = MAX_DIST;
    fo_precomputed_distances(luts + (size_t)i * m * K, m, K, sub, queries + (size_t)i * d, ix->codebook);
  }
  int32_t* rows;
  int n_rows = select_target_rows(ix, targets, n_targets, &rows); /* ref: :544-562 */
  if (!use_target_lists) {
    for (int r = 0; r < n_rows; r++) {               /* ref: :600-608 */
      const int16_t* codes = ix->codes + (size_t)rows[r] * m;
      for (int j = 0; j < nq; j++) {
        float dist = fo_pq_distance_int16(luts + (size_t)j * m * K, codes, m, K);
        if (dist < max_dists[j]) {
          fo_update_topk(topks + (size_t)j * k, dist, ix->ids[rows[r]], k);
          max_dists[j] = topks[(size_t)j * k + k - 1].distance;
        }
      }
    }
  } else {
    for (int i = 0; i < nq; i++) {                   /* ref: :613-631 */
      for (int r = 0; r < n_rows; r++) {
        const int16_t* codes = ix->codes + (size_t)rows[r] * m;
        float dist = 0;
        for (int l = 0; l < m; l++) dist = dist + luts[(size_t)i * m * K + K * l + codes[l]];
        if (dist < max_dists[i]) {
          fo_update_topk(topks + (size_t)i * k, dist, ix->ids[rows[r]], k);
          max_dists[i] = topks[(size_t)i * k + k - 1].distance;
        }
      }
    }
  }
  free(rows); free(luts); free(max_dists);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* multi-threaded runner for the CPU baseline (one shard per thread)          */
/* ------------------------------------------------------------------------- */

typedef struct {
  const FoIndex* ix; const float* queries; int begin, end, k, w;
  FoTopKEntry* out; int64_t rows; int rc;
} ManyArgs;

static void* many_worker(void* p) {
  ManyArgs* a = p;
  a->rows = 0; a->rc = 0;
  FoScratch ws;
  if (scratch_alloc(&ws, a->ix, a->w)) { a->rc = -3; return NULL; }
  for (int q = a->begin; q < a->end; q++) {
    int64_t st[2];
    int rc = ivfadc_search_ws(a->ix, a->queries + (size_t)q * a->ix->d, a->k, a->w,
                              a->out + (size_t)q * a->k, st, &ws);
    if (rc) a->rc = rc;
    a->rows += st[0];
  }
  scratch_free(&ws);
  return NULL;
}

int fo_ivfadc_search_many(const FoIndex* ix, const float* queries, int nq, int k, int w,
                          int n_threads, FoTopKEntry* out_topk, int64_t* rows_scanned) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > nq) n_threads = nq > 0 ? nq : 1;
  pthread_t* th = malloc(sizeof(pthread_t) * (size_t)n_threads);
  ManyArgs* args = malloc(sizeof(ManyArgs) * (size_t)n_threads);
  int rc = 0;
  int64_t rows = 0;
  for (int t = 0; t < n_threads; t++) {
    args[t] = (ManyArgs){ix, queries, (int)((int64_t)nq * t / n_threads),
                         (int)((int64_t)nq * (t + 1) / n_threads), k, w, out_topk, 0, 0};
    pthread_create(&th[t], NULL, many_worker, &args[t]);
  }
  for (int t = 0; t < n_threads; t++) {
    pthread_join(th[t], NULL);
    if (args[t].rc) rc = args[t].rc;
    rows += args[t].rows;
  }
  if (rows_scanned) *rows_scanned = rows;
  free(th); free(args);
  return rc;
}

/* ------------------------------------------------------------------------- */
/* vector UDFs                                                                */
/* ------------------------------------------------------------------------- */

/* ref: cosine_similarity.c:12-37 cosine_similarity_simple: three double accumulators,
 * each term a double product of float inputs; 0 when either squared norm is 0. */
double fo_cosine_similarity(const float* v1, const float* v2, int n) {
  double scalar = 0, sq1 = 0, sq2 = 0;
  for (int i = 0; i < n; i++) {
    scalar += ((double)v1[i]) * ((double)v2[i]);
    sq2 += ((double)v2[i]) * ((double)v2[i]);
    sq1 += ((double)v1[i]) * ((double)v1[i]);
  }
  if (sq1 > 0 && sq2 > 0) return scalar / (sqrt(sq1) * sqrt(sq2));
  return 0;
}

/* ref: cosine_similarity.c:39-45 cosine_similarity_simple_norm: plain double dot */
double fo_cosine_similarity_norm(const float* v1, const float* v2, int n) {
  double scalar = 0;
  for (int i = 0; i < n; i++) scalar += ((double)v1[i]) * ((double)v2[i]);
  return scalar;
}

/* ref: core_functions.c:67-81 cosine_similarity_bytea: fp32 dot, product and sum rounded
 * separately, no normalisation */
float fo_cosine_similarity_bytea(const float* v1, const float* v2, int n) {
  float scalar = 0;
  for (int i = 0; i < n; i++) {
    float prod = v1[i] * v2[i];
    scalar = scalar + prod;
  }
  return scalar;
}

/* ref: core_functions.c:120-139 / :179-196 */
void fo_vec_minus(const float* a, const float* b, int n, float* out) {
  for (int i = 0; i < n; i++) out[i] = a[i] - b[i];
}
void fo_vec_plus(const float* a, const float* b, int n, float* out) {
  for (int i = 0; i < n; i++) out[i] = a[i] + b[i];
}

/* ref: core_functions.c:243-269 vec_normalize_bytea: fp32 sum of squares, sqrt through
 * double (the C library sqrt), fp32 division */
void fo_vec_normalize(const float* a, int n, float* out) {
  float sq = 0;
  for (int i = 0; i < n; i++) {
    float p = a[i] * a[i];
    sq = sq + p;
  }
  float length = (float)sqrt((double)sq);
  for (int i = 0; i < n; i++) out[i] = a[i] / length;
}

/* ref: freddy--0.0.1.sql:1270-1288 analogy_3cosadd */
int fo_analogy_3cosadd(const float* vectors, int N, int d, int row_a, int row_b, int row_c, float* score) {
  float* q = malloc(sizeof(float) * (size_t)d);
  float* t = malloc(sizeof(float) * (size_t)d);
  fo_vec_minus(vectors + (size_t)row_c * d, vectors + (size_t)row_a * d, d, t);   /* v3 - v1 */
  fo_vec_plus(t, vectors + (size_t)row_b * d, d, q);                               /* + v2    */
  int best = -1;
  float best_s = 0;
  for (int r = 0; r < N; r++) {
    if (r == row_a || r == row_b || r == row_c) continue;                          /* word NOT IN (...) */
    float s = fo_cosine_similarity_bytea(q, vectors + (size_t)r * d, d);
    if (best < 0 || s > best_s) { best = r; best_s = s; }                          /* DESC, first row wins ties */
  }
  if (score) *score = best_s;
  free(q); free(t);
  return best;
}

typedef struct { const float* v; int N, d; const int32_t* abc; int begin, end; int32_t* rows; float* scores; } AnaArgs;
static void* ana_worker(void* p) {
  AnaArgs* a = p;
  for (int q = a->begin; q < a->end; q++)
    a->rows[q] = fo_analogy_3cosadd(a->v, a->N, a->d, a->abc[3 * q], a->abc[3 * q + 1], a->abc[3 * q + 2], &a->scores[q]);
  return NULL;
}
int fo_analogy_3cosadd_many(const float* vectors, int N, int d, const int32_t* rows_abc, int nq, int n_threads,
                            int32_t* out_rows, float* out_scores) {
  if (n_threads < 1) n_threads = 1;
  if (n_threads > nq) n_threads = nq > 0 ? nq : 1;
  pthread_t* th = malloc(sizeof(pthread_t) * (size_t)n_threads);
  AnaArgs* args = malloc(sizeof(AnaArgs) * (size_t)n_threads);
  for (int t = 0; t < n_threads; t++) {
    args[t] = (AnaArgs){vectors, N, d, rows_abc, (int)((int64_t)nq * t / n_threads),
                        (int)((int64_t)nq * (t + 1) / n_threads), out_rows, out_scores};
    pthread_create(&th[t], NULL, ana_worker, &args[t]);
  }
  for (int t = 0; t < n_threads; t++) pthread_join(th[t], NULL);
  free(th); free(args);
  return 0;
}

/* ------------------------------------------------------------------------- */
/* kNN-join: ivpq_search_in                                                   */
/* ------------------------------------------------------------------------- */

/* ref: index_utils.c:673-682 getConfidenceHyp — normal approximation of a
 * hypergeometric tail; float/double mix kept exactly as written there. */
float fo_confidence_hyp(int expect, int size, float p, int stat_size) {
  if (expect > size) return 0;
  float mu = size * p;
  float sig = sqrt(size * p * (1.0 - p)) * (((float)stat_size - size) / ((float)stat_size - 1.0));
  return 1.0 - 0.5 * (1.0 + erf((((float)expect) - 0.5 - mu) / (sig * sqrt(2))));
}

typedef struct { int id; float distance; int pos0, pos1; } HeapNode;   /* index_utils.h:33-37 QueueEntry */

/* ref: index_utils.c:118-131 push — sift up while the parent is strictly larger */
static void heap_push(HeapNode* nodes, int* len, float distance, int id, int p0, int p1) {
  int i = *len;
  int parent = (i - 1) / 2;
  while (i > 0 && nodes[parent].distance > distance) {
    nodes[i] = nodes[parent];
    i = parent;
    parent = (parent - 1) / 2;
  }
  nodes[i].distance = distance; nodes[i].id = id; nodes[i].pos0 = p0; nodes[i].pos1 = p1;
  (*len)++;
}

/* ref: index_utils.c:133-155 pop — the old last element (still readable at nodes[len])
 * is sifted down from the root through a hole; a child moves up only if strictly
 * smaller; note the reference also reads nodes[j+1] when j+1 == len (the stale copy). */
static HeapNode heap_pop(HeapNode* nodes, int* len) {
  HeapNode result = nodes[0];
  nodes[0] = nodes[*len - 1];
  (*len)--;
  int i = 0;
  while (i != *len) {
    int pick = *len;
    int j = 1 + 2 * i;
    if (j <= *len - 1 && nodes[j].distance < nodes[pick].distance) pick = j;
    if (j <= *len - 1 && nodes[j + 1].distance < nodes[pick].distance) pick = j + 1;
    nodes[i] = nodes[pick];
    i = pick;
  }
  return result;
}

static int cmp_topk_dist(const void* a, const void* b) {   /* index_utils.c:111-116 cmpTopKEntry */
  float x = ((const FoTopKEntry*)a)->distance, y = ((const FoTopKEntry*)b)->distance;
  return (x > y) - (x < y);
}

typedef struct { int id; float distance; const float* vector; } PvEntry;   /* index_utils.h:27-31 TopKPVEntry */
static int cmp_pv_dist(const void* a, const void* b) {      /* index_utils.c:104-109 cmpTopKPVEntry */
  float x = ((const PvEntry*)a)->distance, y = ((const PvEntry*)b)->distance;
  return (x > y) - (x < y);
}

/* ref: ivpq_search_in.c:40-44 reorderTopKPV */
static void pv_reorder(PvEntry* tk, int kk, int* fill, float* max_dist) {
  qsort(tk, (size_t)*fill, sizeof(PvEntry), cmp_pv_dist);
  *fill = kk;
  *max_dist = tk[kk - 1].distance;
}
/* ref: ivpq_search_in.c:46-57 updateTopKPVFast */
static void pv_append(PvEntry* tk, int batch, int kk, int* fill, float* max_dist, int id, float distance, const float* vec) {
  tk[*fill].id = id; tk[*fill].distance = distance; tk[*fill].vector = vec;
  (*fill)++;
  if (*fill == batch - 1) pv_reorder(tk, kk, fill, max_dist);
}
static void pv_init(PvEntry* tk, int n, float max_dist) {   /* index_utils.c:84-92 initTopKPV */
  for (int i = 0; i < n; i++) { tk[i].distance = max_dist; tk[i].id = -1; tk[i].vector = NULL; }
}

/* ref: index_utils.c:252-443 determineCoarseIdsMultiWithStatisticsMulti (USE_PROPERTY_QUEUE,
 * two positions): per active query, multi-sequence traversal of the Kc x Kc grid in ascending
 * d0 + d1 until getConfidenceHyp(minTarget, |targets|, sum freq, total) >= confidence.
 * Registers query q with every visited cell (cell_q[cell] in registration order). */
static int ivpq_select_cells(const FoIvpqIndex* ix, const float* queries, const int* active, int n_active,
                             int n_targets, int min_target, float confidence,
                             int** cell_q, int* cell_n) {
  const int Kc = ix->Kc, cells = Kc * Kc, half = ix->d / 2;
  int last_iteration = 1;
  FoTopKEntry* md0 = malloc(sizeof(FoTopKEntry) * (size_t)Kc);
  FoTopKEntry* md1 = malloc(sizeof(FoTopKEntry) * (size_t)Kc);
  float* all = malloc(sizeof(float) * (size_t)cells);
  HeapNode* heap = malloc(sizeof(HeapNode) * (size_t)(cells + 2));
  unsigned char* traversed = malloc((size_t)cells);
  unsigned char* in_queue = malloc((size_t)cells);
  for (int c = 0; c < cells; c++) cell_n[c] = 0;
  for (int x = 0; x < n_active; x++) {
    const int q = active[x];
    const float* qv = queries + (size_t)q * ix->d;
    int visited = 0;
    float prob = 0.0f;
    for (int j = 0; j < Kc; j++) {                                             /* :296-305 */
      md0[j].id = j; md0[j].distance = fo_square_distance(qv, ix->coarse_multi + (size_t)j * half, half);
      md1[j].id = j; md1[j].distance = fo_square_distance(qv + half, ix->coarse_multi + (size_t)(Kc + j) * half, half);
    }
    for (int i = 0; i < cells; i++) {                                          /* :306-313 */
      float s = 0;
      s += md0[i % Kc].distance;
      s += md1[i / Kc].distance;
      all[i] = s;
    }
    qsort(md0, (size_t)Kc, sizeof(FoTopKEntry), cmp_topk_dist);                 /* :317-319 */
    qsort(md1, (size_t)Kc, sizeof(FoTopKEntry), cmp_topk_dist);
    memset(traversed, 0, (size_t)cells);
    memset(in_queue, 0, (size_t)cells);
    int len = 0;
    {
      int first = md0[0].id + Kc * md1[0].id;                                   /* :342-348 */
      heap[0].pos0 = 0; heap[0].pos1 = 0; heap[0].id = first; heap[0].distance = all[first];
      len = 1;
    }
    while (fo_confidence_hyp(min_target, n_targets, prob, (int)ix->stats[cells]) < confidence && visited < cells) { /* :349-351 */
      HeapNode next = heap_pop(heap, &len);
      traversed[next.pos0 + Kc * next.pos1] = 1;
      if (next.pos0 < Kc - 1 &&
          (next.pos1 == 0 || traversed[next.pos0 + 1 + Kc * (next.pos1 - 1)])) {      /* :356-372 */
        int np = (next.pos0 + 1) + Kc * next.pos1;
        if (!in_queue[np]) {
          int nid = md0[next.pos0 + 1].id + Kc * md1[next.pos1].id;
          heap_push(heap, &len, all[nid], nid, next.pos0 + 1, next.pos1);
          in_queue[np] = 1;
        }
      }
      if (next.pos1 < Kc - 1 &&
          (next.pos0 == 0 || traversed[next.pos0 - 1 + Kc * (next.pos1 + 1)])) {      /* :373-392 */
        int np = next.pos0 + Kc * (next.pos1 + 1);
        if (!in_queue[np]) {
          int nid = md0[next.pos0].id + Kc * md1[next.pos1 + 1].id;
          heap_push(heap, &len, all[nid], nid, next.pos0, next.pos1 + 1);
          in_queue[np] = 1;
        }
      }
      prob += ix->stats[next.id];                                               /* :394 */
      visited++;
      cell_q[next.id][cell_n[next.id]++] = q;                                   /* :398-402 */
    }
    if (visited < cells) last_iteration = 0;                                    /* :404-406 */
  }
  free(md0); free(md1); free(all); free(heap); free(traversed); free(in_queue);
  return last_iteration;
}

static const float* ivpq_vector_of(const FoIvpqIndex* ix, int32_t id) {   /* INNER JOIN vecs ON fq.id = vecs.id */
  int lo = 0, hi = ix->Nv - 1;
  while (lo <= hi) {
    int mid = lo + (hi - lo) / 2;
    if (ix->vec_ids[mid] < id) lo = mid + 1; else if (ix->vec_ids[mid] > id) hi = mid - 1;
    else return ix->vectors + (size_t)mid * ix->d;
  }
  return NULL;
}

/* computePQDistanceInt16 over getPrecomputedDistancesDouble's table (index_utils.c:457-475, :1126-1133) */
static float pq_distance_pairs(const float* lut, const int16_t* codes, int m, int K) {
  float distance = 0;
  for (int l = 0; l < m / 2; l++) {
    float pair = lut[(size_t)(2 * l) * K + codes[2 * l]] + lut[(size_t)(2 * l + 1) * K + codes[2 * l + 1]];
    distance += pair;
  }
  return distance;
}

/* ref: ivpq_search_in.c:197-684 (first-call body).  The target-list mode only reorders the
 * loop nest (row-major collect, then query-major evaluate, :546-607) except for the
 * `targetCounts < k*alpha_original` skip (:553-557), which is kept. */
int fo_ivpq_search_in(const FoIvpqIndex* ix, const float* queries, int nq, int k, const int32_t* targets, int n_targets,
                      int alpha_original, int pvf, int method, int use_tl, float confidence, int double_threshold,
                      FoTopKEntry* topks, int64_t* stats_out) {
  const float MAX_DIST = 1000.0f;                    /* :62 */
  const int BATCH = 200;                             /* :63 TOPK_BATCH_SIZE */
  const int d = ix->d, m = ix->m, K = ix->K, sub = d / m, cells = ix->Kc * ix->Kc;
  int alpha = alpha_original;
  if (pvf < 1) pvf = 1;                              /* :206-208 */
  /* pair-LUT variant (:261-275, index_utils.c:457-475): preDists[pair][c0 + K*c1] = d(pos 2l, c0) + d(pos 2l+1, c1),
   * the row's distance = sum over the positions/2 pairs (a trailing odd position is never looked at).  The pair
   * sums are formed on the fly here instead of materialising the K*K table; same fp32 additions in the same order.
   * codes2[] is an int16 array (:417,:447-451): the pair code overflows it when K*K > 32768. */
  const int double_codes = (method == 0 || method == 2) && alpha * k > double_threshold;
  if (double_codes && (int64_t)K * K > 32768) return -10;
  const int kk = k * pvf;
  int64_t rounds = 0, pairs = 0;

  float* max_dists = malloc(sizeof(float) * (size_t)(nq ? nq : 1));
  int* target_counts = calloc((size_t)(nq ? nq : 1), sizeof(int));
  int* fill = calloc((size_t)(nq ? nq : 1), sizeof(int));
  PvEntry* pv = NULL;
  float* luts = NULL;
  for (int i = 0; i < nq; i++) { fo_init_topk(topks + (size_t)i * k, k, MAX_DIST); max_dists[i] = MAX_DIST; }   /* :238 */
  if (method == 2) {                                 /* :243-251 */
    pv = malloc(sizeof(PvEntry) * (size_t)(nq ? nq : 1) * (BATCH + kk));
    for (int i = 0; i < nq; i++) pv_init(pv + (size_t)i * (BATCH + kk), BATCH + kk, MAX_DIST);
  }
  if (method == 0 || method == 2) {                  /* :279-290 */
    luts = malloc(sizeof(float) * (size_t)(nq ? nq : 1) * m * K);
    for (int i = 0; i < nq; i++) fo_precomputed_distances(luts + (size_t)i * m * K, m, K, sub, queries + (size_t)i * d, ix->codebook);
  }
  /* rows selected by `fq.id IN (targets)`, table order */
  int32_t* tsorted = malloc(sizeof(int32_t) * (size_t)(n_targets ? n_targets : 1));
  memcpy(tsorted, targets, sizeof(int32_t) * (size_t)n_targets);
  qsort(tsorted, (size_t)n_targets, sizeof(int32_t), cmp_i32);

  int* active = malloc(sizeof(int) * (size_t)(nq ? nq : 1));
  int n_active = nq;
  for (int i = 0; i < nq; i++) active[i] = i;
  int** cell_q = malloc(sizeof(int*) * (size_t)cells);
  for (int c = 0; c < cells; c++) cell_q[c] = malloc(sizeof(int) * (size_t)(nq ? nq : 1));
  int* cell_n = malloc(sizeof(int) * (size_t)cells);
  /* per-query candidate lists for the target-list mode (query-major evaluation) */
  int** tl_rows = NULL; int* tl_n = NULL; int* tl_cap = NULL;
  if (use_tl) { tl_rows = calloc((size_t)(nq ? nq : 1), sizeof(int*)); tl_n = calloc((size_t)(nq ? nq : 1), sizeof(int)); tl_cap = calloc((size_t)(nq ? nq : 1), sizeof(int)); }

  while (n_active > 0) {                              /* :299 */
    rounds++;
    int last_iteration = ivpq_select_cells(ix, queries, active, n_active, n_targets, k * alpha, confidence, cell_q, cell_n);
    if (use_tl) for (int i = 0; i < nq; i++) tl_n[i] = 0;
    for (int r = 0; r < ix->N; r++) {                 /* rows: coarse_id IN (cells with queries) AND id IN (targets) */
      const int cell = ix->coarse_ids[r];
      if (cell_n[cell] == 0) continue;
      if (!bsearch(&ix->ids[r], tsorted, (size_t)n_targets, sizeof(int32_t), cmp_i32)) continue;
      const float* vec = NULL;
      if (method != 0) { vec = ivpq_vector_of(ix, ix->ids[r]); if (!vec) continue; }
      const int16_t* codes = ix->codes + (size_t)r * m;
      for (int j = 0; j < cell_n[cell]; j++) {        /* :459-541 */
        const int q = cell_q[cell][j];
        target_counts[q] += 1;
        pairs++;
        if (use_tl) {
          if (tl_n[q] == tl_cap[q]) { tl_cap[q] = tl_cap[q] ? 2 * tl_cap[q] : 256; tl_rows[q] = realloc(tl_rows[q], sizeof(int) * (size_t)tl_cap[q]); }
          tl_rows[q][tl_n[q]++] = r;
          continue;
        }
        float dist;
        if (method == 1) dist = fo_square_distance(queries + (size_t)q * d, vec, d);
        else dist = double_codes ? pq_distance_pairs(luts + (size_t)q * m * K, codes, m, K)
                                 : fo_pq_distance_int16(luts + (size_t)q * m * K, codes, m, K);
        if (dist < max_dists[q]) {
          if (method == 2) pv_append(pv + (size_t)q * (BATCH + kk), BATCH + kk, kk, &fill[q], &max_dists[q], ix->ids[r], dist, vec);
          else { fo_update_topk(topks + (size_t)q * k, dist, ix->ids[r], k); max_dists[q] = topks[(size_t)q * k + k - 1].distance; }
        }
      }
    }
    if (use_tl) {                                      /* :546-607 */
      for (int x = 0; x < n_active; x++) {
        const int q = active[x];
        if (target_counts[q] < k * alpha_original && !last_iteration) { target_counts[q] = 0; continue; }   /* :553-557 */
        for (int t = 0; t < tl_n[q]; t++) {
          const int r = tl_rows[q][t];
          const float* vec = (method != 0) ? ivpq_vector_of(ix, ix->ids[r]) : NULL;
          float dist;
          if (method == 1) dist = fo_square_distance(queries + (size_t)q * d, vec, d);
          else dist = double_codes ? pq_distance_pairs(luts + (size_t)q * m * K, ix->codes + (size_t)r * m, m, K)
                                   : fo_pq_distance_int16(luts + (size_t)q * m * K, ix->codes + (size_t)r * m, m, K);
          if (dist < max_dists[q]) {
            if (method == 2) pv_append(pv + (size_t)q * (BATCH + kk), BATCH + kk, kk, &fill[q], &max_dists[q], ix->ids[r], dist, vec);
            else { fo_update_topk(topks + (size_t)q * k, dist, ix->ids[r], k); max_dists[q] = topks[(size_t)q * k + k - 1].distance; }
          }
        }
      }
    }
    if (method == 2) {                                 /* :611-629 + index_utils.c:477-498 postverify */
      for (int x = 0; x < n_active; x++) {
        const int q = active[x];
        pv_reorder(pv + (size_t)q * (BATCH + kk), kk, &fill[q], &max_dists[q]);
      }
      for (int x = 0; x < n_active; x++) {
        const int q = active[x];
        PvEntry* tk = pv + (size_t)q * (BATCH + kk);
        float md = MAX_DIST;
        for (int j = 0; j < kk; j++) {
          if (tk[j].id != -1) {
            float dist = fo_square_distance(queries + (size_t)q * d, tk[j].vector, d);
            if (dist < md) { fo_update_topk(topks + (size_t)q * k, dist, tk[j].id, k); md = topks[(size_t)q * k + k - 1].distance; }
          }
        }
      }
    }
    if (!last_iteration) {                             /* :639-666 */
      int n_new = 0;
      int* next = malloc(sizeof(int) * (size_t)n_active);
      for (int x = 0; x < n_active; x++) {
        const int q = active[x];
        if (topks[(size_t)q * k + k - 1].distance == MAX_DIST) {
          next[n_new++] = q;
          fo_init_topk(topks + (size_t)q * k, k, MAX_DIST);
          max_dists[q] = MAX_DIST;
          if (method == 2) { pv_init(pv + (size_t)q * (BATCH + kk), BATCH + kk, MAX_DIST); fill[x] = 0; /* sic: :654 indexes by x */ }
        }
      }
      memcpy(active, next, sizeof(int) * (size_t)n_new);
      free(next);
      n_active = n_new;
    } else {
      n_active = 0;
    }
    alpha += alpha;                                    /* :680 */
  }
  if (stats_out) { stats_out[0] = rounds; stats_out[1] = pairs; }
  for (int c = 0; c < cells; c++) free(cell_q[c]);
  free(cell_q); free(cell_n); free(active); free(tsorted); free(max_dists); free(target_counts); free(fill); free(pv); free(luts);
  if (use_tl) { for (int i = 0; i < nq; i++) free(tl_rows[i]); free(tl_rows); free(tl_n); free(tl_cap); }
  return 0;
}


/* ref: freddy.c:1567-1582 (coarse assignment + residual of insert_batch) and index_utils.c:923-939
 * (updateCodebook's nearest-centroid loop; its codebook drift update is not part of the quantisation). */
int fo_encode(const float* vectors, int n, int d, const float* coarse, int C,
              const float* codebook, int m, int K, int32_t* out_coarse_ids, int16_t* out_codes) {
  const int sub = d / m;
  float* res = malloc(sizeof(float) * (size_t)d);
  int rc = 0;
  for (int i = 0; i < n; i++) {
    const float* raw = vectors + (size_t)i * d;
    const float* v = raw;
    if (coarse != NULL) {
      float min_dist = 100;
      int best = -1;
      for (int j = 0; j < C; j++) {
        float dist = fo_square_distance(raw, coarse + (size_t)j * d, d);
        if (dist < min_dist) { best = j; min_dist = dist; }
      }
      if (best < 0) { rc = -1; best = 0; }
      out_coarse_ids[i] = best;
      for (int j = 0; j < d; j++) res[j] = raw[j] - coarse[(size_t)best * d + j];
      v = res;
    }
    for (int pos = 0; pos < m; pos++) {
      float min_dist = 100;   /* "sufficient high value" */
      int best = -1;
      for (int code = 0; code < K; code++) {
        float dist = fo_square_distance(v + pos * sub, codebook + ((size_t)pos * K + code) * sub, sub);
        if (dist < min_dist) { best = code; min_dist = dist; }
      }
      if (best < 0) { rc = -1; best = 0; }
      out_codes[(size_t)i * m + pos] = (int16_t)best;
    }
  }
  free(res);
  return rc;
}


/* ref: freddy.c:1178-1401 grouping_pq */
int fo_grouping_pq(const FoIndex* ix, const float* vectors, const int32_t* vec_ids, int n_vec,
                   const int32_t* ids, int n_ids, const int32_t* group_ids, int n_groups,
                   int32_t* out_ids, int32_t* out_group_ids) {
  const int m = ix->m, K = ix->K, d = ix->d, sub = d / m;
  int32_t* groups = malloc(sizeof(int32_t) * (size_t)(n_groups > 0 ? n_groups : 1));
  memcpy(groups, group_ids, sizeof(int32_t) * (size_t)n_groups);
  qsort(groups, (size_t)n_groups, sizeof(int32_t), cmp_i32);                  /* :1222 */
  /* `SELECT id, vector ... WHERE id IN (groups) ORDER BY id ASC`: one row per distinct existing id (:1224-1246) */
  float* luts = malloc(sizeof(float) * (size_t)(n_groups > 0 ? n_groups : 1) * m * K);
  int found = 0;
  for (int g = 0; g < n_groups; g++) {
    if (g > 0 && groups[g] == groups[g - 1]) continue;
    int lo = 0, hi = n_vec;
    while (lo < hi) { int mid = (lo + hi) / 2; if (vec_ids[mid] < groups[g]) lo = mid + 1; else hi = mid; }
    if (lo < n_vec && vec_ids[lo] == groups[g]) {
      /* row `found` of the result pairs with groups[found] in the reference: only equal when nothing is missing */
      fo_precomputed_distances(luts + (size_t)found * m * K, m, K, sub, vectors + (size_t)lo * d, ix->codebook);   /* :1291-1299 */
      found++;
    }
  }
  if (found != n_groups) { free(groups); free(luts); return -1; }             /* "Group ids do not exist" */
  int32_t* rows = NULL;
  const int n = select_target_rows(ix, ids, n_ids, &rows);                    /* :1303-1321 */
  int rc = n;
  for (int i = 0; i < n; i++) {
    const int16_t* codes = ix->codes + (size_t)rows[i] * m;
    float min_dist = 100;                                                     /* :1326 */
    int nearest = -1;
    for (int g = 0; g < n_groups; g++) {
      float distance = fo_pq_distance_int16(luts + (size_t)g * m * K, codes, m, K);   /* :1341-1347 */
      if (distance < min_dist) { min_dist = distance; nearest = g; }
    }
    if (nearest < 0) { rc = -2; nearest = 0; }
    out_ids[i] = ix->ids[rows[i]];
    out_group_ids[i] = n_groups > 0 ? groups[nearest] : -1;
  }
  free(rows); free(groups); free(luts);
  return rc;
}
