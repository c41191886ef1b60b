/* freddy_b200.h — C-ABI of the B200-native IVFADC / PQ k-NN engine.
 *
 * This is the drop-in boundary: plain C, opaque handle, caller-
 * allocated outputs, int status codes, no exceptions, no torch / C++ types.
 * Each entry point names the interface of the reference PostgreSQL extension
 * (guenthermi/postgres-word2vec, paths relative to freddy_extension/) that it
 * replaces.  The Postgres-side shim that would call it is shown in
 * INTEGRATION.md; tests/ and bench.py call it through ctypes.
 *
 * Model: one engine per process (= one Postgres backend; a CUDA context cannot
 * cross fork()).  The index tables are uploaded ONCE per session
 * (fb_load_*), transformed into the device layout and pinned in HBM; every
 * search call then runs entirely on the GPU.  The reference instead re-reads
 * codebook, coarse quantizer and inverted lists through SPI on every call
 * (freddy.c:239-241, :324-343).
 *
 * Result contract (identical to the reference's SRF output):
 *   - exactly k (id, distance) pairs per query, ascending distance; among
 *     equal distances the later table row comes first; admission is strict
 *     (`d < kth`), so which of several boundary ties survive follows the
 *     reference's insertion order (index_utils.c:19-33) exactly;
 *   - unfilled slots are id = -1 with the SRF's sentinel distance;
 *   - distances are the reference's fp32 values bit for bit (sequential,
 *     unfused sub/mul/add); the "%f" text rounding of the SRF is applied by
 *     the shim / fb_round_through_text, not here.
 */
#ifndef FREDDY_B200_H
#define FREDDY_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fb_engine fb_engine;

enum {
  FB_OK = 0,
  FB_ERR_INVALID = -1,      /* bad argument / index not loaded            */
  FB_ERR_CUDA = -2,         /* CUDA runtime error (see fb_last_error)     */
  FB_ERR_UNSUPPORTED = -3,  /* shape outside what the kernels support     */
  FB_ERR_REFERENCE_UB = -4  /* the reference would run into undefined
                               behaviour here (e.g. fewer than w unprobed
                               lists left, freddy.c:262-302 with id = -1) */
};

/* which codebook / code table a call refers to (index_utils.h:87-99 tableType) */
enum {
  FB_CB_RESIDUAL = 0,  /* residual_codebook  + fine_quantization  (IVFADC)   */
  FB_CB_PQ = 1,        /* pq_codebook        + pq_quantization    (flat PQ)  */
  FB_CB_IVPQ = 2,      /* codebook_ivpq      + fine_quantization_ivpq (kNN-join) */
  FB_CB_KINDS = 3
};

/* ---- lifecycle ---------------------------------------------------------- */
int fb_create(int device, fb_engine** out);
void fb_destroy(fb_engine* e);
/* message of the last failing call on this engine (or of fb_create if e==NULL) */
const char* fb_last_error(const fb_engine* e);

/* ---- index upload: "pin once per session" -------------------------------- */
/* coarse_quantization table, row = coarse id (replaces getCoarseQuantizer,
 * index_utils.c:531-575).  coarse: host [C][d] fp32. */
int fb_load_coarse(fb_engine* e, const float* coarse, int C, int d);
/* (residual_)codebook table as a dense [m][K][sub] fp32 array indexed by
 * (pos, code) (replaces getCodebook, index_utils.c:577-630). */
int fb_load_codebook(fb_engine* e, int kind, const float* codebook, int m, int K, int sub);
/* fine_quantization rows in table order: id, coarse_id, int2[m] codes
 * (replaces the per-call `SELECT id, vector, coarse_id FROM fine WHERE
 * coarse_id IN (...)`, freddy.c:324-343).  Table order is the arrival order
 * of the reference's top-k loop. */
int fb_load_fine(fb_engine* e, const int32_t* ids, const int32_t* coarse_ids,
                 const int16_t* codes, int64_t N, int m);
/* pq_quantization rows in table order (freddy.c:99-103, :544-562, :1099-1113) */
int fb_load_pq(fb_engine* e, const int32_t* ids, const int16_t* codes, int64_t N, int m);

/* ---- search -------------------------------------------------------------- */
/* ivfadc_search(bytea query, int k) with parameter w = get_w()
 * (freddy.c:174-410), batched over nq independent queries.
 * queries: [nq][d] fp32; out_ids/out_dists: [nq][k].  Host buffers. */
int fb_ivfadc_search(fb_engine* e, const float* queries, int nq, int k, int w,
                     int32_t* out_ids, float* out_dists);
/* Same, all pointers are DEVICE pointers on the engine's device; the call is
 * enqueued on the engine stream and is complete on return only after
 * fb_synchronize().  No host<->device copies. */
int fb_ivfadc_search_dev(fb_engine* e, const float* d_queries, int nq, int k, int w,
                         int32_t* d_out_ids, float* d_out_dists);

/* ivfadc_batch_search(int[] ids, int k) (freddy.c:677-1024): the query vectors are the rows of the
 * word-vector table (fb_load_vectors) whose id is in query_ids, in table order, each once
 * (freddy.c:767-804); per query one list per round, nearest unprobed first, until k rows were seen;
 * sentinel distance 100.0.  out_query_ids[n], out_ids/out_dists [n][k] with n = *n_queries_out
 * <= n_ids (buffers sized for n_ids). */
int fb_ivfadc_batch_search(fb_engine* e, const int32_t* query_ids, int n_ids, int k, int32_t* out_query_ids,
                           int32_t* out_ids, float* out_dists, int* n_queries_out);

/* pq_search(bytea, int) (freddy.c:28-170): exhaustive ADC over pq_quantization,
 * sentinel distance 100.0. */
int fb_pq_search(fb_engine* e, const float* queries, int nq, int k,
                 int32_t* out_ids, float* out_dists);
/* pq_search_in(bytea, int, int[]) / pq_search_in_batch(bytea[], int[], int,
 * int[], bool) (freddy.c:1028-1174, :414-675): ADC over the rows whose id is
 * in targets; sentinel 1000.0.  use_target_lists only changes the reference's
 * loop nest, not its results; accepted for signature parity. */
int fb_pq_search_in_batch(fb_engine* e, const float* queries, int nq, int k,
                          const int32_t* targets, int n_targets, int use_target_lists,
                          int32_t* out_ids, float* out_dists);

/* ---- kNN-join ------------------------------------------------------------------ */
/* IVPQ index: 2-way multi-index coarse quantizer coarse_multi [2][Kc][d/2]
 * (coarse_quantization_ivpq, ivpq.py:27-30), fine_quantization_ivpq rows in table order
 * (id, coarse_id = c0 + Kc*c1, int2[m] codes over the RAW vectors; ivpq.py:35, :18) and the
 * statistics table stats[Kc*Kc + 1] (cell frequencies, last = total; freddy--0.0.1.sql:150-171).
 * The fine codebook goes through fb_load_codebook(FB_CB_IVPQ) first; post verification /
 * exact mode read the word vectors of fb_load_vectors (the `vecs` side of the SQL join). */
int fb_load_ivpq(fb_engine* e, const float* coarse_multi, int Kc, int d, const int32_t* ids,
                 const int32_t* coarse_ids, const int16_t* codes, int64_t N, int m, const float* stats);
/* ivpq_search_in(bytea[] q, int[] qids, k, int[] targets, alpha, pvf, method, use_targetlist,
 * confidence, double_threshold) (ivpq_search_in.c:61-721), the C side of knn_join /
 * knn_in_ivpq_batch.  method 0 = PQ, 1 = exact, 2 = PQ + post verification.  Returns the
 * [nq][k] (target id, distance) rows; the caller pairs them with its query ids.
 * alpha*k > double_threshold (pair-LUT variant, off by default) is FB_ERR_UNSUPPORTED. */
int fb_ivpq_search_in(fb_engine* e, const float* queries, int nq, int k, const int32_t* targets, int n_targets,
                      int alpha, int pvf, int method, int use_target_lists, float confidence,
                      int double_threshold, int32_t* out_ids, float* out_dists);

/* ---- dense word-vector UDFs ------------------------------------------------ */
/* word-vector table (google_vecs / google_vecs_norm: id, vector bytea = float4[d],
 * vec2database.py:25) in table order; pinned in HBM, dimension-major in 32-row blocks */
int fb_load_vectors(fb_engine* e, const int32_t* ids, const float* vectors, int64_t N, int d);
/* n independent pairs a[i], b[i] of d floats -> out[i]
 *   variant 0: cosine_similarity(float4[], float4[]) -> float8   core_functions.c:23-42, cosine_similarity.c:12-37
 *   variant 1: cosine_similarity_norm                -> float8   core_functions.c:44-63, cosine_similarity.c:39-45
 *   variant 2: cosine_similarity_bytea               -> float4 (widened)   core_functions.c:65-81 */
int fb_cosine_similarity(fb_engine* e, int variant, const float* a, const float* b, int n, int d, double* out);
/* vec_minus_bytea / vec_plus_bytea / vec_normalize_bytea on n vectors
 * (core_functions.c:120-139, :179-196, :243-269); op 0: a-b, 1: a+b, 2: normalize(a) (b ignored) */
int fb_vec_op(fb_engine* e, int op, const float* a, const float* b, int n, int d, float* out);
/* analogy_3cosadd(w1, w2, w3) (freddy--0.0.1.sql:1270-1288) for nq triples of word ids
 * (a, b, c): arg-max over the whole table, rows of the three ids excluded, of
 * cosine_similarity_bytea(vec_plus_bytea(vec_minus_bytea(v_c, v_a), v_b), v_row); the first
 * table row reaching the maximum wins (ORDER BY ... DESC FETCH FIRST 1).
 * out_ids[q] = id of the winner (-1 if none), out_scores[q] = its score. */
int fb_analogy_3cosadd(fb_engine* e, const int32_t* ids_abc, int nq, int32_t* out_ids, float* out_scores);
/* same scan with explicit query vectors [nq][d] and up to three excluded ids per query
 * (-1 = none): the building block for vocabulary-sharded multi-GPU runs */
int fb_analogy_scan(fb_engine* e, const float* qvecs, const int32_t* exclude_ids, int nq,
                    int32_t* out_ids, float* out_scores);

/* ---- exact cosine k-NN and post-verification ----------
 * fb_knn_exact: k_nearest_neighbour(bytea, k) (freddy--0.0.1.sql:426-439) when targets == NULL, else
 * knn_in_exact(bytea, k, int[]) (:1026-1038): similarity = cosine_similarity_bytea (core_functions.c:67-81,
 * sequential float4 dot), ORDER BY similarity DESC FETCH FIRST k.  Equal similarities are ordered by table
 * row (SQL leaves that order to the executor).  Fewer than k rows: id -1, similarity 0.  k <= 32.
 * The whole-table form (and the analogy scans above) runs as a tensor-core pre-filter (bf16 tcgen05.mma over the
 * TMA-staged table, error bound proven in csrc/prefilter_kernels.cuh) followed by the fp32 chain on the
 * candidates: same ids, same similarity bits as the plain fp32 scan (FB_OPT_PREFILTER = 0).
 * fb_ivfadc_search_pv: k_nearest_neighbour_ivfadc_pv(bytea, k) (:574-591) with pvf = get_pvf(), w = get_w():
 * ivfadc_search(v, pvf*k) INNER JOIN vectors ON idx = id, re-ranked by cosine_similarity_bytea.
 * fb_pq_search_pv: k_nearest_neighbour_pq_pv(bytea, k) (:624-662): the same with pq_search(v, pvf*k) as the
 * candidate source (the reference's SQL passes the candidate's word to cosine_similarity_bytea and cannot run
 * as written; this is the query its ivfadc twin spells out).                                              */
int fb_knn_exact(fb_engine* e, const float* queries, int nq, int k, const int32_t* targets, int n_targets,
                 int32_t* out_ids, float* out_sims);
int fb_ivfadc_search_pv(fb_engine* e, const float* queries, int nq, int k, int pvf, int w,
                        int32_t* out_ids, float* out_sims);
int fb_pq_search_pv(fb_engine* e, const float* queries, int nq, int k, int pvf, int32_t* out_ids, float* out_sims);

/* ---- quantisation of new rows (the encode step of index build / insert_batch) ------
 * As insert_batch assigns them: nearest coarse centroid with strict `<` from 100, first minimum wins
 * (freddy.c:1567-1577), residual = row - centroid (:1578-1581), then per position the nearest codeword of the
 * residual codebook, same rule in (pos, code) table order (updateCodebook, index_utils.c:923-939).
 * fb_encode_pq quantises the raw rows with the pq / ivpq codebook (kind = FB_CB_PQ / FB_CB_IVPQ).
 * out_codes = [n][m] int16.  FB_ERR_REFERENCE_UB where every candidate distance is >= 100 (the reference
 * then reads an uninitialised assignment).  The codebook drift update of insert_batch is not done here. */
int fb_encode_ivfadc(fb_engine* e, const float* vectors, int64_t n, int32_t* out_coarse_ids, int16_t* out_codes);
int fb_encode_pq(fb_engine* e, int kind, const float* vectors, int64_t n, int16_t* out_codes);
/* the same with DEVICE pointers (index build: index_creation/ivfadc.py:36-96, ivpq.py:100-193 quantise the whole
 * table; the rows are in HBM already) */
int fb_encode_ivfadc_dev(fb_engine* e, const float* d_vectors, int64_t n, int32_t* d_out_coarse_ids, int16_t* d_out_codes);
int fb_encode_pq_dev(fb_engine* e, int kind, const float* d_vectors, int64_t n, int16_t* d_out_codes);

/* ---- in-place append (insert_batch must refresh the pinned copy) -----------------------
 * insert_batch (freddy.c:1403-1658) INSERTs the new rows into pq_quantization, fine_quantization and
 * fine_quantization_ivpq (index_utils.c:1003-1043).  These calls add the same rows to the pinned tables without a
 * re-upload: the table is re-packed on the device (the grown list moves the lists behind it), only the new rows
 * cross PCIe.  Appended rows arrive after every existing row (they are the last rows of the heap), in the order
 * given.  Codebook changes of the same insert_batch go through fb_load_codebook (1.2 MB).
 * fb_append_pq: kind = FB_CB_PQ (cells ignored) or FB_CB_IVPQ (cells = multi-index cell c0 + Kc*c1 per row). */
int fb_append_fine(fb_engine* e, const int32_t* ids, const int32_t* coarse_ids, const int16_t* codes, int64_t n);
int fb_append_pq(fb_engine* e, int kind, const int32_t* ids, const int32_t* cells, const int16_t* codes, int64_t n);
/* rows appended to the word-vector table (updateWordVectorsRelation, index_utils.c:1045-1074) */
int fb_append_vectors(fb_engine* e, const int32_t* ids, const float* vectors, int64_t n);

/* ---- grouping_pq ---------------------------------------------------------------
 * grouping_pq(int[] ids, int[] group_ids) (freddy.c:1178-1401): the rows of the flat pq table whose id is in
 * `ids` (table order, each once), each assigned to the nearest group vector (word-vector rows of `group_ids`,
 * taken in ascending id order) by ADC distance; strict `<` from 100, first minimum wins.
 * out_ids / out_group_ids: [n_ids] caller-allocated, *n_out rows written.
 * FB_ERR_INVALID "Group ids do not exist" as the SRF raises it (:1243-1245); FB_ERR_REFERENCE_UB where every
 * distance is >= 100.  Needs fb_load_pq + fb_load_vectors.                                                */
int fb_grouping_pq(fb_engine* e, const int32_t* ids, int n_ids, const int32_t* group_ids, int n_groups,
                   int32_t* out_ids, int32_t* out_group_ids, int* n_out);

int fb_synchronize(fb_engine* e);
/* Run all subsequent work on the caller's CUDA stream (a cudaStream_t passed as
 * void*; NULL restores the engine's own stream).  Lets a host runtime order the
 * engine's kernels with its own copies/events without extra synchronisation. */
int fb_set_stream(fb_engine* e, void* cuda_stream);

/* ---- knobs / introspection ---------------------------------------------- */
enum {
  FB_OPT_FORCE_EXACT_PATH = 1, /* 1: send every query through the general
                                  (tie/re-probe exact) kernel; testing aid    */
  FB_OPT_PROFILE = 2,          /* 1: bracket every kernel with CUDA events on
                                  the engine stream and accumulate fb_counters */
  FB_OPT_QUERY_CHUNK = 3,      /* queries per pipeline chunk (LUT scratch =
                                  chunk * w * m * K * 4 bytes)                */
  FB_OPT_PACKED_FP32 = 5,      /* 1 (default): LUT build uses the packed f32x2
                                  forms of the same rounded operations; 0: scalar */
  FB_OPT_LUT_TILE = 6,         /* codes per LUT-build CTA: 256 / 512 / 1024 (then
                                  4 / 2 / 1 CTAs per SM); anything else: generic  */
  FB_OPT_LUT_CTAS_PER_SM = 7,  /* cap on resident LUT-build CTAs per SM (0 = as many as fit)      */
  FB_OPT_OVERLAP = 8,          /* 1: LUT build of chunk c+1 overlaps the scan of chunk c (two streams) */
  FB_OPT_PIPELINE = 9,         /* 1 (default): large batches of the headline shapes run the warp-
                                  specialised pipeline kernel (LUT build of chunk c+1 and ADC scan of
                                  chunk c share every SM inside one launch); 0: separate kernels   */
  FB_OPT_PIPE_CHUNK = 10,      /* queries per pipeline beat (default 2048)                         */
  FB_OPT_PLACEMENT_WINDOW = 12, /* fb_load_fine: rows of a list are placed into 32-row blocks so that
                                  the lanes of a warp hit different shared-memory banks of the LUT as
                                  far as possible (arrival order travels with each row, results do
                                  not change); value = candidate rows examined per slot (default 256,
                                  0/1 = keep arrival order).  Applies to the next fb_load_fine.      */
  FB_OPT_PIPE_SHAPE = 13,      /* role split of the pipeline CTA (producer warps, scan warps): 0 = (12,14), 1 = (8,18), 2 = (8,14) with a 4x16 producer tile,
                                  3 = (16,10) for LUT-bound indexes (short probed lists) */
  FB_OPT_PIPE_RAMP = 14,       /* 1: the first and last pipeline chunks are shortened (quarter, half); default 0  */
  FB_OPT_CUDA_GRAPHS = 15,     /* 1 (default): host-buffer fb_ivfadc_search calls with <= 16 queries replay a
                                  captured CUDA graph of the whole call (upload, kernels, download)      */
  FB_OPT_ZERO_COPY_UPLOAD = 16, /* 1 (default): fb_ivfadc_search batches whose query buffer is pinned (page-locked) host
                                  memory are read by the coarse kernel directly (mapped pointer), which also writes
                                  the device copy: the upload overlaps the coarse step; pageable buffers are copied */
  FB_OPT_PIPE_DEBUG = 11,      /* timing aid, results are NOT valid: 1 = producers only (no scan),
                                  2 = scan only (LUT scratch left as is)                           */
  FB_OPT_PREFILTER = 17,       /* 1 (default): fb_knn_exact (whole table) and the analogy scans select candidates with a bf16
                                  tcgen05 GEMM whose error is bounded, then decide with the reference's fp32 chain
                                  (identical results); 0: the fp32 scan kernels do all the work                    */
  FB_OPT_PREFILTER_LOCKSTEP = 19, /* table tiles a pre-filter CTA may run ahead of the slowest CTA that streams the same
                                  slab for another query tile (default 0 = free running; >= 8 throttles): keeps shared tiles in L2 */
  FB_OPT_BYTE_CODES = 18,      /* 1 (default): tables with K <= 256 and m <= 16 also keep a true uint8 image of their codes
                                  (16 bytes per row, one 16-byte load per row in the scan kernels: the layout the
                                  reference's index_creation/config JSON files, k = 256, call for); 0: 16-bit units only      */
  FB_OPT_DEVICE_BUILD = 20,    /* 1 (default): fb_load_fine / fb_load_pq / fb_load_ivpq / fb_load_vectors group the rows by list,
                                  place them (FB_OPT_PLACEMENT_WINDOW <= 1024) and pack them with kernels, and sort the id
                                  columns on the device; the host sends the raw columns once.  0: the host-side builder
                                  (same layout; kept as the cross-check of tests/test_build_gpu.py)                      */
  FB_OPT_SUBSET_PLACEMENT = 21, /* 1 (default): the per-call target subset of pq_search_in_batch gets the conflict-aware row placement
                                  (groups of 512 rows, one CTA each) when queries x targets >= 2^26; 0: table order; 2: always */
  FB_OPT_QSCAN_MIN_QUERIES = 4 /* chunks with at least this many queries use the
                                  one-CTA-per-query scan (default 64); smaller
                                  ones use one CTA per (query, list)          */
};
int fb_set_option(fb_engine* e, int option, int64_t value);

typedef struct {
  int64_t queries;          /* queries answered since the last reset            */
  int64_t rows_scanned;     /* fine/pq rows whose ADC distance was computed     */
  int64_t scan_bytes;       /* algorithmic bytes of those rows: m*2 + 4 each    */
  int64_t exact_path_queries; /* queries re-done by the general kernel          */
  int64_t kernel_launches;  /* kernels launched by this library                 */
  double ms_coarse, ms_lut, ms_scan, ms_finalize, ms_exact; /* FB_OPT_PROFILE   */
  int64_t n_scan_launches;
  /* why queries took the general kernel (a query can have several reasons) */
  int64_t exact_coarse_tie, exact_coarse_far, exact_few_rows, exact_scan_tie, exact_forced;
  double ms_pipe;           /* pipeline-kernel launches (LUT build + scan in one) */
  int64_t n_pipe_launches;
  /* tensor-core pre-filter of the exact scans: queries it answered, queries handed back to the fp32 scan
   * (candidate buffer overflow), candidates emitted (counted under FB_OPT_PROFILE only) */
  int64_t prefilter_queries, prefilter_overflow_queries, prefilter_candidates;
} fb_counters;
int fb_get_counters(fb_engine* e, fb_counters* out);  /* synchronizes the stream */
int fb_reset_counters(fb_engine* e);

/* create_statistics (freddy--0.0.1.sql:150-171) over the pinned IVPQ table, on the device: stats[c] = (rows of cell c
 * whose id is listed)::float8 / (all listed rows found) cast to float4, stats[Kc*Kc] = that total; a table row counts
 * once per occurrence of its id in `ids` (the reference counts rows of the JOIN with the user's column).  ids = NULL:
 * every row of the table.  out_stats (host, [Kc*Kc + 1]) may be NULL; install != 0 makes them the statistics
 * fb_ivpq_search_in uses.  fb_load_ivpq(stats = NULL) calls this over all rows. */
int fb_ivpq_statistics(fb_engine* e, const int32_t* ids, int64_t n_ids, float* out_stats, int install);

/* Sidecar (include/freddy_sidecar.h): a thread of this process serves the single-query requests that backends post
 * into the shared-memory segment `name`, batching whatever is pending into one fb_ivfadc_search call.  The engine must
 * not be used by its owner between start and stop.  counters3 (may be NULL) = batches, queries, largest batch. */
typedef struct fb_sidecar fb_sidecar;
int fb_sidecar_start(fb_engine* e, const char* name, int max_k, int slots, int max_batch, int linger_us, fb_sidecar** out);
int fb_sidecar_running(fb_sidecar* sc);     /* 0 once a client asked the loop to stop (fbsc_client_request_stop) */
int fb_sidecar_stop(fb_sidecar* sc, int64_t* counters3);

/* Diagnostics: order-sensitive checksums of the layout of a pinned table (0 fine, 1 pq, 2 ivpq): out[0] over (slot,
 * row number), out[1] over the packed code units.  Equal layouts give equal sums (device build vs host build). */
int fb_table_checksum(fb_engine* e, int table, uint64_t* out);

/* Host-only helper (no device needed): the slot order fb_load_fine gives the rows of ONE inverted list
 * under FB_OPT_PLACEMENT_WINDOW = window.  codes = [n][m] int16 of that list in arrival order;
 * order_out[s] = arrival index of the row placed in slot s (a permutation of 0..n-1).            */
int fb_placement_order(const int16_t* codes, int n, int m, int K, int window, int32_t* order_out);

/* snprintf("%f") -> float4in, as every SRF returns distances (freddy.c:401-408) */
float fb_round_through_text(float distance);

const char* fb_version(void);

#ifdef __cplusplus
}
#endif
#endif
