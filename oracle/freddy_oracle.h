/* freddy_oracle.h — CPU ORACLE. TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of the search arithmetic of guenthermi/postgres-word2vec
 * (freddy_extension C files), over in-memory arrays instead of SPI rows.  It exists
 * to check the CUDA engine; it is never linked into, imported by or called from
 * the product library (postgres-word2vec_b200/).  Only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() may use it.
 *
 * Parity pin: every kernel-level function below is checked bit-for-bit against
 * the reference's own index_utils.c / cosine_similarity.c compiled unmodified
 * into oracle/_ref/libfreddy_ref.so (tests/test_oracle_vs_ref.py) and against
 * golden vectors generated from that library (tests/golden/).  The SRF driver
 * bodies (which interleave arithmetic with SPI row iteration and cannot run
 * without a Postgres server) are restated line by line; each cites the
 * reference file:line it follows.
 *
 * All arithmetic: IEEE fp32, each + - * individually rounded (compile with
 * -O2 -ffp-contract=off and no -march: the PGXS default), loops left to right.
 */
#ifndef FREDDY_ORACLE_H
#define FREDDY_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct { int id; float distance; } FoTopKEntry; /* index_utils.h:17-20 */

/* In-memory image of the index tables.  Rows are in table
 * (= ascending id) order, exactly what a heap scan of the bulk-loaded tables
 * returns. */
typedef struct {
  int d;   /* vector dimensionality                                        */
  int m;   /* sub-quantizer positions (codebook "pos" 0..m-1)              */
  int K;   /* codes per position ("code" 0..K-1)                           */
  int C;   /* coarse centroids (0 for a flat PQ index)                     */
  int N;   /* rows of the fine / pq quantization table                     */
  const float* coarse;        /* [C][d]   coarse_quantization.vector, row = id      */
  const float* codebook;      /* [m][K][d/m]  (residual_)codebook rows by (pos,code) */
  const int32_t* ids;         /* [N] ascending                                       */
  const int32_t* coarse_ids;  /* [N] fine_quantization.coarse_id (NULL for flat PQ)  */
  const int16_t* codes;       /* [N][m] int2[] payload                               */
  /* derived by fo_index_prepare(): CSR over coarse_id (rows stay id-ascending) */
  int32_t* list_offsets;      /* [C+1] */
  int32_t* list_rows;         /* [N] row numbers grouped by coarse id */
  int16_t* list_codes;        /* [N][m] codes in list_rows order (contiguous per list: what a clustered scan reads) */
  int32_t* list_ids;          /* [N] ids in list_rows order */
} FoIndex;

int fo_index_prepare(FoIndex* ix);   /* builds the CSR; returns 0 on success */
void fo_index_release(FoIndex* ix);

/* ---- kernel-level restatements (checked against oracle/_ref) ---- */
float fo_square_distance(const float* v1, const float* v2, int n);
void fo_update_topk(FoTopKEntry* tk, float distance, int id, int k);
void fo_init_topk(FoTopKEntry* tk, int k, float max_dist);
void fo_precomputed_distances(float* pre_dists, int positions, int codes, int sub,
                              const float* query, const float* codebook);
float fo_pq_distance_int16(const float* pre_dists, const int16_t* codes, int positions, int ncodes);
/* "%f" through text, as every SRF emits distances (freddy.c:401-408) */
float fo_round_through_text(float distance);

/* ---- SRF driver restatements ---- */
/* ivfadc_search (freddy.c:174-393).  Returns 0, or <0 where the reference
 * would hit undefined behaviour (fewer than w unprobed lists left; coarse
 * distance >= 100).  stats (optional, may be NULL): [0]=rows scanned,
 * [1]=rounds of the re-probe loop. */
int fo_ivfadc_search(const FoIndex* ix, const float* query, int k, int w,
                     FoTopKEntry* out_topk, int64_t* stats);
/* pq_search (freddy.c:28-134): exhaustive ADC over all rows, sentinel 100.0 */
int fo_pq_search(const FoIndex* ix, const float* query, int k, FoTopKEntry* out_topk);
/* pq_search_in (freddy.c:1028-1143): ADC over the rows whose id is in targets */
int fo_pq_search_in(const FoIndex* ix, const float* query, int k,
                    const int32_t* targets, int n_targets, FoTopKEntry* out_topk);
/* pq_search_in_batch (freddy.c:414-631): out_topk is [nq][k] */
int fo_pq_search_in_batch(const FoIndex* ix, const float* queries, int nq, int k,
                          const int32_t* targets, int n_targets, int use_target_lists,
                          FoTopKEntry* out_topk);

/* ---- quantisation of new rows (insert_batch) ---- */
/* freddy.c:1567-1582: nearest coarse centroid (strict `<` from 100, first minimum wins) and residual;
 * index_utils.c:923-939 (updateCodebook): nearest codeword per position (strict `<` from 100, table
 * order = (pos, code) ascending).  coarse == NULL: the raw vectors are quantised (pq / ivpq tables).
 * Returns 0, or -1 where the reference would read an uninitialised assignment (every distance >= 100). */
int fo_encode(const float* vectors, int n, int d, const float* coarse, int C,
              const float* codebook, int m, int K, int32_t* out_coarse_ids, int16_t* out_codes);

/* ---- grouping_pq (freddy.c:1178-1401) ---- */
/* ix: flat PQ index (codebook = pq codebook, codes = pq codes).  vectors/vec_ids: the normalized word-vector
 * table (ids ascending).  Group vectors are taken in ascending group-id order; every selected pq row (table
 * order, `WHERE id IN`) gets the first group with the smallest ADC distance below 100.
 * Returns the number of rows written, -1 "Group ids do not exist" (:1243-1245), -2 where the reference reads
 * an uninitialised assignment. */
int fo_grouping_pq(const FoIndex* ix, const float* vectors, const int32_t* vec_ids, int n_vec,
                   const int32_t* ids, int n_ids, const int32_t* group_ids, int n_groups,
                   int32_t* out_ids, int32_t* out_group_ids);

/* ---- vector UDFs (core_functions.c, cosine_similarity.c) ---- */
double fo_cosine_similarity(const float* v1, const float* v2, int n);      /* cosine_similarity.c:12-37 */
double fo_cosine_similarity_norm(const float* v1, const float* v2, int n); /* cosine_similarity.c:39-45 */
float fo_cosine_similarity_bytea(const float* v1, const float* v2, int n); /* core_functions.c:67-81 */
void fo_vec_minus(const float* a, const float* b, int n, float* out);      /* core_functions.c:120-139 */
void fo_vec_plus(const float* a, const float* b, int n, float* out);       /* core_functions.c:179-196 */
void fo_vec_normalize(const float* a, int n, float* out);                  /* core_functions.c:243-269 */
/* analogy_3cosadd (freddy--0.0.1.sql:1270-1288) over an in-memory word-vector table:
 * argmax over rows r (table order, rows a/b/c excluded) of
 * cosine_similarity_bytea(vec_plus_bytea(vec_minus_bytea(v[c], v[a]), v[b]), v[r]);
 * ORDER BY ... DESC FETCH FIRST 1: the first row reaching the maximum wins.
 * rows are table row numbers; returns the winning row (or -1), score in *score. */
int fo_analogy_3cosadd(const float* vectors, int N, int d, int row_a, int row_b, int row_c, float* score);
int fo_analogy_3cosadd_many(const float* vectors, int N, int d, const int32_t* rows_abc, int nq, int n_threads,
                            int32_t* out_rows, float* out_scores);

/* ---- kNN-join: ivpq_search_in (ivpq_search_in.c:61-721) ---- */
typedef struct {
  int d, m, K;                  /* fine PQ over the raw vectors (codebook_ivpq)              */
  int Kc;                       /* codes per half of the 2-way multi-index coarse quantizer   */
  int N;                        /* rows of fine_quantization_ivpq                             */
  const float* coarse_multi;    /* [2][Kc][d/2]  coarse_quantization_ivpq rows by (pos, code) */
  const float* codebook;        /* [m][K][d/m]                                                */
  const int32_t* ids;           /* [N] table order                                            */
  const int32_t* coarse_ids;    /* [N] c0 + Kc*c1 (ivpq.py:18)                                 */
  const int16_t* codes;         /* [N][m]                                                     */
  const float* stats;           /* [Kc*Kc + 1] cell frequencies, last entry = total count     */
  int Nv;                       /* rows of the normalized word-vector table                   */
  const int32_t* vec_ids;       /* [Nv] ascending                                             */
  const float* vectors;         /* [Nv][d]                                                    */
} FoIvpqIndex;
/* method: 0 PQ, 1 exact, 2 PQ + post verification.  out_topk is [nq][k].
 * stats_out (optional): [0] rounds of the retry loop, [1] candidate (query,row) pairs. */
int fo_ivpq_search_in(const FoIvpqIndex* ix, const float* queries, int nq, int k, const int32_t* targets, int n_targets,
                      int alpha, int pvf, int method, int use_target_lists, float confidence, int double_threshold,
                      FoTopKEntry* out_topk, int64_t* stats_out);
float fo_confidence_hyp(int expect, int size, float p, int stat_size);   /* index_utils.c:673-682 */

/* Run fo_ivfadc_search over nq queries on n_threads host threads (disjoint
 * query shards = n_threads concurrent backends).  out_topk is [nq][k]. */
int fo_ivfadc_search_many(const FoIndex* ix, const float* queries, int nq, int k, int w,
                          int n_threads, FoTopKEntry* out_topk, int64_t* rows_scanned);

#ifdef __cplusplus
}
#endif
#endif
