/* freddy_shim.c — the PostgreSQL side of the drop-in: the six search SRFs of the FREDDY
 * extension (freddy--0.0.1.sql:370-392) with their first-call bodies replaced by calls into
 * libfreddy_b200.so (include/freddy_b200.h).  Built with PGXS next to the reference's
 * index_utils.c / output_utils.c / core_functions.c (table-name and parameter lookup, bytea
 * converters, loaders and the non-search UDFs stay the reference's own code).
 *
 * One engine per backend process, created on first use (a CUDA context cannot cross fork());
 * each index is read through SPI ONCE per session and pinned in HBM — the reference re-reads
 * codebook, coarse table and inverted lists on every call (freddy.c:239-241, :324-343).
 *
 * In this repository the file is compiled against oracle/pg_stubs + oracle/pg_emul.c (no
 * Postgres in the image) and driven through the fmgr/SRF protocol by tests/test_shim_gpu.py.
 */
#include "postgres.h"
#include "fmgr.h"
#include "funcapi.h"
#include "executor/spi.h"
#include "utils/array.h"
#include "catalog/pg_type.h"

#include "index_utils.h"
#include "output_utils.h"

#include "freddy_b200.h"

static fb_engine* engine = NULL;
static int pinned_d = 0, vecs_d = 0;
static bool pinned_ivfadc = false, pinned_pq = false, pinned_ivpq = false, pinned_vecs = false;

static void fb_check(int rc) {
  if (rc != FB_OK) elog(ERROR, "freddy_b200: %s", fb_last_error(engine));
}

static void ensure_engine(void) {
  if (engine == NULL) {
    int rc = fb_create(0, &engine);
    if (rc != FB_OK) elog(ERROR, "freddy_b200: %s", fb_last_error(NULL));
  }
}

/* codebook rows (pos, code, vector) -> dense [m][K][sub] */
static float* flatten_codebook(CodebookCompound cb, int sub) {
  float* out = palloc(sizeof(float) * cb.positions * cb.codeSize * sub);
  for (int i = 0; i < cb.positions * cb.codeSize; i++)
    memcpy(out + ((size_t)cb.codebook[i].pos * cb.codeSize + cb.codebook[i].code) * sub, cb.codebook[i].vector,
           sizeof(float) * sub);
  return out;
}

/* SELECT id, <int column or nothing>, vector FROM <table>: the whole code table in heap order */
static int fetch_code_table(const char* table, bool with_coarse, int32** ids, int32** cids, int16** codes, int* m) {
  char command[200];
  int n;
  snprintf(command, sizeof command, with_coarse ? "SELECT id, coarse_id, vector FROM %s" : "SELECT id, vector FROM %s", table);
  SPI_connect();
  if (SPI_exec(command, 0) <= 0 || SPI_tuptable == NULL) elog(ERROR, "cannot read %s", table);
  n = (int)SPI_processed;
  *ids = SPI_palloc(sizeof(int32) * (n ? n : 1));
  if (with_coarse) *cids = SPI_palloc(sizeof(int32) * (n ? n : 1));
  *codes = NULL;
  *m = 0;
  for (int i = 0; i < n; i++) {
    bool isnull;
    HeapTuple t = SPI_tuptable->vals[i];
    bytea* v = DatumGetByteaP(SPI_getbinval(t, SPI_tuptable->tupdesc, with_coarse ? 3 : 2, &isnull));
    int len = (VARSIZE(v) - VARHDRSZ) / sizeof(int16);
    if (*codes == NULL) { *m = len; *codes = SPI_palloc(sizeof(int16) * (size_t)n * len); }
    (*ids)[i] = DatumGetInt32(SPI_getbinval(t, SPI_tuptable->tupdesc, 1, &isnull));
    if (with_coarse) (*cids)[i] = DatumGetInt32(SPI_getbinval(t, SPI_tuptable->tupdesc, 2, &isnull));
    memcpy(*codes + (size_t)i * len, VARDATA(v), sizeof(int16) * len);
  }
  SPI_finish();
  return n;
}

static void pin_vectors(void) {
  char name[100], command[200];
  int n, d = 0;
  int32* ids;
  float* vecs = NULL;
  if (pinned_vecs) return;
  ensure_engine();
  getTableName(NORMALIZED, name, 100);
  snprintf(command, sizeof command, "SELECT id, vector FROM %s", name);
  SPI_connect();
  if (SPI_exec(command, 0) <= 0 || SPI_tuptable == NULL) elog(ERROR, "cannot read %s", name);
  n = (int)SPI_processed;
  ids = SPI_palloc(sizeof(int32) * (n ? n : 1));
  for (int i = 0; i < n; i++) {
    bool isnull;
    HeapTuple t = SPI_tuptable->vals[i];
    bytea* v = DatumGetByteaP(SPI_getbinval(t, SPI_tuptable->tupdesc, 2, &isnull));
    if (vecs == NULL) { d = (VARSIZE(v) - VARHDRSZ) / sizeof(float4); vecs = SPI_palloc(sizeof(float) * (size_t)n * d); }
    ids[i] = DatumGetInt32(SPI_getbinval(t, SPI_tuptable->tupdesc, 1, &isnull));
    memcpy(vecs + (size_t)i * d, VARDATA(v), sizeof(float) * d);
  }
  SPI_finish();
  fb_check(fb_load_vectors(engine, ids, vecs, n, d));
  vecs_d = d;
  pinned_vecs = true;
}

static void pin_ivfadc(int d) {
  char cbname[100], finename[100];
  CodebookCompound cb;
  CoarseQuantizer cq;
  int C, n, m;
  int32 *ids, *cids;
  int16* codes;
  float* coarse;
  if (pinned_ivfadc && pinned_d == d) return;
  ensure_engine();
  getTableName(RESIDUAL_CODEBOOK, cbname, 100);
  getTableName(RESIDUAL_QUANTIZATION, finename, 100);
  cb = getCodebook(cbname);                                   /* index_utils.c:577-630 */
  cq = getCoarseQuantizer(&C);                                /* index_utils.c:531-575 */
  coarse = palloc(sizeof(float) * (size_t)C * d);
  for (int i = 0; i < C; i++) memcpy(coarse + (size_t)i * d, cq[i].vector, sizeof(float) * d);   /* row i = coarse id i (freddy.c:280) */
  fb_check(fb_load_coarse(engine, coarse, C, d));
  fb_check(fb_load_codebook(engine, FB_CB_RESIDUAL, flatten_codebook(cb, d / cb.positions), cb.positions, cb.codeSize,
                            d / cb.positions));
  n = fetch_code_table(finename, true, &ids, &cids, &codes, &m);
  fb_check(fb_load_fine(engine, ids, cids, codes, n, m));
  pinned_ivfadc = true;
  pinned_d = d;
}

static void pin_pq(int d) {
  char cbname[100], tname[100];
  CodebookCompound cb;
  int n, m;
  int32* ids;
  int16* codes;
  if (pinned_pq) return;
  ensure_engine();
  getTableName(CODEBOOK, cbname, 100);
  getTableName(PQ_QUANTIZATION, tname, 100);
  cb = getCodebook(cbname);
  fb_check(fb_load_codebook(engine, FB_CB_PQ, flatten_codebook(cb, d / cb.positions), cb.positions, cb.codeSize, d / cb.positions));
  n = fetch_code_table(tname, false, &ids, NULL, &codes, &m);
  fb_check(fb_load_pq(engine, ids, codes, n, m));
  pinned_pq = true;
}

static void pin_ivpq(int d) {
  char cbname[100], tname[100], cqname[100];
  CodebookCompound cb, cqm;
  float* stats;
  int n, m;
  int32 *ids, *cids;
  int16* codes;
  if (pinned_ivpq) return;
  ensure_engine();
  getTableName(IVPQ_CODEBOOK, cbname, 100);
  getTableName(IVPQ_QUANTIZATION, tname, 100);
  getTableName(COARSE_QUANTIZATION_MULTI, cqname, 100);
  cb = getCodebook(cbname);                                   /* ivpq_search_in.c:216-218 */
  cqm = getCodebook(cqname);                                  /* ivpq_search_in.c:222-224 */
  stats = getStatistics();                                    /* ivpq_search_in.c:232 */
  if (cqm.positions != 2) elog(ERROR, "multi-index coarse quantizer with %d positions", cqm.positions);
  fb_check(fb_load_codebook(engine, FB_CB_IVPQ, flatten_codebook(cb, d / cb.positions), cb.positions, cb.codeSize, d / cb.positions));
  n = fetch_code_table(tname, true, &ids, &cids, &codes, &m);
  fb_check(fb_load_ivpq(engine, flatten_codebook(cqm, d / 2), cqm.codeSize, d, ids, cids, codes, n, m, stats));
  pinned_ivpq = true;
}

/* ---- SRF plumbing: the value-per-call emission every search SRF shares (freddy.c:394-409) ---- */
static void setup_result(FuncCallContext* funcctx, int natts) {
  TupleDesc desc = CreateTemplateTupleDesc(natts);
  if (natts == 2) {
    TupleDescInitEntry(desc, 1, "Id", INT4OID, -1, 0);
    TupleDescInitEntry(desc, 2, "Distance", FLOAT4OID, -1, 0);
  } else {
    TupleDescInitEntry(desc, 1, "QueryId", INT4OID, -1, 0);
    TupleDescInitEntry(desc, 2, "TargetId", INT4OID, -1, 0);
    TupleDescInitEntry(desc, 3, "Distance", FLOAT4OID, -1, 0);
  }
  funcctx->attinmeta = TupleDescGetAttInMetadata(desc);
}

static Datum emit_single(FunctionCallInfo fcinfo) {
  FuncCallContext* funcctx = SRF_PERCALL_SETUP();
  UsrFctx* u = (UsrFctx*)funcctx->user_fctx;
  if (u->iter >= u->k) SRF_RETURN_DONE(funcctx);
  snprintf(u->values[0], 16, "%d", u->tk[u->iter].id);
  snprintf(u->values[1], 16, "%f", u->tk[u->iter].distance);
  u->iter++;
  SRF_RETURN_NEXT(funcctx, HeapTupleGetDatum(BuildTupleFromCStrings(funcctx->attinmeta, u->values)));
}

static Datum emit_batch(FunctionCallInfo fcinfo) {
  FuncCallContext* funcctx = SRF_PERCALL_SETUP();
  UsrFctxBatch* u = (UsrFctxBatch*)funcctx->user_fctx;
  if (u->iter >= u->k * u->queryIdsSize) SRF_RETURN_DONE(funcctx);
  snprintf(u->values[0], 16, "%d", u->queryIds[u->iter / u->k]);
  snprintf(u->values[1], 16, "%d", u->tk[u->iter / u->k][u->iter % u->k].id);
  snprintf(u->values[2], 16, "%f", u->tk[u->iter / u->k][u->iter % u->k].distance);
  u->iter++;
  SRF_RETURN_NEXT(funcctx, HeapTupleGetDatum(BuildTupleFromCStrings(funcctx->attinmeta, u->values)));
}

static void finish_single(FuncCallContext* funcctx, const int32* ids, const float* dist, int k) {
  TopK tk = palloc(sizeof(TopKEntry) * k);
  UsrFctx* u = palloc(sizeof(UsrFctx));
  for (int i = 0; i < k; i++) { tk[i].id = ids[i]; tk[i].distance = dist[i]; }
  fillUsrFctx(u, tk, k);
  funcctx->user_fctx = u;
  setup_result(funcctx, 2);
}

static void finish_batch(FuncCallContext* funcctx, int* qids, int nq, const int32* ids, const float* dist, int k) {
  TopK* tks = palloc(sizeof(TopK) * (nq ? nq : 1));
  UsrFctxBatch* u = palloc(sizeof(UsrFctxBatch));
  for (int q = 0; q < nq; q++) {
    tks[q] = palloc(sizeof(TopKEntry) * k);
    for (int i = 0; i < k; i++) { tks[q][i].id = ids[(size_t)q * k + i]; tks[q][i].distance = dist[(size_t)q * k + i]; }
  }
  fillUsrFctxBatch(u, qids, nq, tks, k);
  funcctx->user_fctx = u;
  setup_result(funcctx, 3);
}

static float* bytea_array_to_matrix(ArrayType* arr, int* nq, int* d) {
  Datum* data;
  float* out = NULL;
  getArray(arr, &data, nq);
  *d = 0;
  for (int i = 0; i < *nq; i++) {
    bytea* b = DatumGetByteaP(data[i]);
    int n = (VARSIZE(b) - VARHDRSZ) / sizeof(float4);
    if (out == NULL) { *d = n; out = palloc(sizeof(float) * (size_t)(*nq) * n); }
    memcpy(out + (size_t)i * n, VARDATA(b), sizeof(float) * n);
  }
  return out;
}

static int* int_array(ArrayType* arr, int* n) {
  Datum* data;
  int* out;
  getArray(arr, &data, n);
  out = palloc(sizeof(int) * (*n ? *n : 1));
  for (int i = 0; i < *n; i++) out[i] = DatumGetInt32(data[i]);
  return out;
}

/* ---- the SRFs --------------------------------------------------------------------------- */
PG_FUNCTION_INFO_V1(ivfadc_search);
Datum ivfadc_search(PG_FUNCTION_ARGS) {
  if (SRF_IS_FIRSTCALL()) {
    FuncCallContext* funcctx = SRF_FIRSTCALL_INIT();
    MemoryContext old = MemoryContextSwitchTo(funcctx->multi_call_memory_ctx);
    int k = PG_GETARG_INT32(1), w, n = 0;
    float4* q;
    int32* ids = palloc(sizeof(int32) * k);
    float* dist = palloc(sizeof(float) * k);
    getParameter(PARAM_W, &w);                                       /* freddy.c:229 */
    convert_bytea_float4(PG_GETARG_BYTEA_P(0), &q, &n);               /* freddy.c:249 */
    pin_ivfadc(n);
    fb_check(fb_ivfadc_search(engine, q, 1, k, w, ids, dist));        /* replaces freddy.c:251-378 */
    finish_single(funcctx, ids, dist, k);
    MemoryContextSwitchTo(old);
  }
  return emit_single(fcinfo);
}

PG_FUNCTION_INFO_V1(pq_search);
Datum pq_search(PG_FUNCTION_ARGS) {
  if (SRF_IS_FIRSTCALL()) {
    FuncCallContext* funcctx = SRF_FIRSTCALL_INIT();
    MemoryContext old = MemoryContextSwitchTo(funcctx->multi_call_memory_ctx);
    int k = PG_GETARG_INT32(1), n = 0;
    float4* q;
    int32* ids = palloc(sizeof(int32) * k);
    float* dist = palloc(sizeof(float) * k);
    convert_bytea_float4(PG_GETARG_BYTEA_P(0), &q, &n);
    pin_pq(n);
    fb_check(fb_pq_search(engine, q, 1, k, ids, dist));               /* replaces freddy.c:74-134 */
    finish_single(funcctx, ids, dist, k);
    MemoryContextSwitchTo(old);
  }
  return emit_single(fcinfo);
}

PG_FUNCTION_INFO_V1(pq_search_in);
Datum pq_search_in(PG_FUNCTION_ARGS) {
  if (SRF_IS_FIRSTCALL()) {
    FuncCallContext* funcctx = SRF_FIRSTCALL_INIT();
    MemoryContext old = MemoryContextSwitchTo(funcctx->multi_call_memory_ctx);
    int k = PG_GETARG_INT32(1), n = 0, nt = 0;
    float4* q;
    int* targets = int_array(PG_GETARG_ARRAYTYPE_P(2), &nt);
    int32* ids = palloc(sizeof(int32) * k);
    float* dist = palloc(sizeof(float) * k);
    convert_bytea_float4(PG_GETARG_BYTEA_P(0), &q, &n);
    pin_pq(n);
    fb_check(fb_pq_search_in_batch(engine, q, 1, k, targets, nt, 0, ids, dist));   /* replaces freddy.c:1070-1143 */
    finish_single(funcctx, ids, dist, k);
    MemoryContextSwitchTo(old);
  }
  return emit_single(fcinfo);
}

PG_FUNCTION_INFO_V1(pq_search_in_batch);
Datum pq_search_in_batch(PG_FUNCTION_ARGS) {
  if (SRF_IS_FIRSTCALL()) {
    FuncCallContext* funcctx = SRF_FIRSTCALL_INIT();
    MemoryContext old = MemoryContextSwitchTo(funcctx->multi_call_memory_ctx);
    int nq, d, nqid, nt, k = PG_GETARG_INT32(2);
    float* q = bytea_array_to_matrix(PG_GETARG_ARRAYTYPE_P(0), &nq, &d);
    int* qids = int_array(PG_GETARG_ARRAYTYPE_P(1), &nqid);
    int* targets = int_array(PG_GETARG_ARRAYTYPE_P(3), &nt);
    int32* ids = palloc(sizeof(int32) * (size_t)(nq ? nq : 1) * k);
    float* dist = palloc(sizeof(float) * (size_t)(nq ? nq : 1) * k);
    if (nqid != nq) elog(ERROR, "Number of query vectors and query vector ids differs!");   /* freddy.c:495 */
    pin_pq(d);
    fb_check(fb_pq_search_in_batch(engine, q, nq, k, targets, nt, PG_GETARG_BOOL(4), ids, dist));   /* replaces freddy.c:514-631 */
    finish_batch(funcctx, qids, nq, ids, dist, k);
    MemoryContextSwitchTo(old);
  }
  return emit_batch(fcinfo);
}

PG_FUNCTION_INFO_V1(ivfadc_batch_search);
Datum ivfadc_batch_search(PG_FUNCTION_ARGS) {
  if (SRF_IS_FIRSTCALL()) {
    FuncCallContext* funcctx = SRF_FIRSTCALL_INIT();
    MemoryContext old = MemoryContextSwitchTo(funcctx->multi_call_memory_ctx);
    int n = 0, nq = 0, k = PG_GETARG_INT32(1);
    int* qids = int_array(PG_GETARG_ARRAYTYPE_P(0), &n);
    int* out_q = palloc(sizeof(int) * (n ? n : 1));
    int32* ids = palloc(sizeof(int32) * (size_t)(n ? n : 1) * k);
    float* dist = palloc(sizeof(float) * (size_t)(n ? n : 1) * k);
    pin_vectors();                                                    /* the query vectors live in the normalized table */
    {
      /* d of the index = d of the vectors table: pin with the first vector's length */
      char name[100], command[200];
      bool isnull;
      int d;
      getTableName(NORMALIZED, name, 100);
      snprintf(command, sizeof command, "SELECT id, vector FROM %s", name);
      SPI_connect();
      SPI_exec(command, 0);
      d = SPI_processed ? (VARSIZE(DatumGetByteaP(SPI_getbinval(SPI_tuptable->vals[0], SPI_tuptable->tupdesc, 2, &isnull))) - VARHDRSZ) / (int)sizeof(float4) : 0;
      SPI_finish();
      pin_ivfadc(d);
    }
    fb_check(fb_ivfadc_batch_search(engine, qids, n, k, out_q, ids, dist, &nq));   /* replaces freddy.c:757-982 */
    finish_batch(funcctx, out_q, nq, ids, dist, k);
    MemoryContextSwitchTo(old);
  }
  return emit_batch(fcinfo);
}

PG_FUNCTION_INFO_V1(ivpq_search_in);
Datum ivpq_search_in(PG_FUNCTION_ARGS) {
  if (SRF_IS_FIRSTCALL()) {
    FuncCallContext* funcctx = SRF_FIRSTCALL_INIT();
    MemoryContext old = MemoryContextSwitchTo(funcctx->multi_call_memory_ctx);
    int nq, d, nqid, nt, k = PG_GETARG_INT32(2);
    float* q = bytea_array_to_matrix(PG_GETARG_ARRAYTYPE_P(0), &nq, &d);
    int* qids = int_array(PG_GETARG_ARRAYTYPE_P(1), &nqid);
    int* targets = int_array(PG_GETARG_ARRAYTYPE_P(3), &nt);
    int method = PG_GETARG_INT32(6);
    int32* ids = palloc(sizeof(int32) * (size_t)(nq ? nq : 1) * k);
    float* dist = palloc(sizeof(float) * (size_t)(nq ? nq : 1) * k);
    if (nqid != nq) elog(ERROR, "Number of query vectors and query vector ids differs! ( %d, %d)", nqid, nq);   /* ivpq_search_in.c:180 */
    pin_ivpq(d);
    if (method != 0) pin_vectors();                                   /* the `vecs` side of the join (ivpq_search_in.c:363-373) */
    fb_check(fb_ivpq_search_in(engine, q, nq, k, targets, nt, PG_GETARG_INT32(4), PG_GETARG_INT32(5), method,
                               PG_GETARG_BOOL(7), PG_GETARG_FLOAT4(8), PG_GETARG_INT32(9), ids, dist));   /* replaces ivpq_search_in.c:197-684 */
    finish_batch(funcctx, qids, nq, ids, dist, k);
    MemoryContextSwitchTo(old);
  }
  return emit_batch(fcinfo);
}

/* test hook (emulator builds only): forget the pinned tables so another index can be registered */
void freddy_shim_reset(void) {
  if (engine) fb_destroy(engine);
  engine = NULL;
  pinned_ivfadc = pinned_pq = pinned_ivpq = pinned_vecs = false;
  pinned_d = 0;
}

/* grouping_pq(int[] ids, int[] group_ids) -> SETOF (id int4, group id int4)   replaces freddy.c:1185-1371 */
PG_FUNCTION_INFO_V1(grouping_pq);
Datum grouping_pq(PG_FUNCTION_ARGS) {
  FuncCallContext* funcctx;
  UsrFctxGrouping* u;
  if (SRF_IS_FIRSTCALL()) {
    MemoryContext old;
    TupleDesc desc;
    int n = 0, ng = 0, n_out = 0;
    int *ids, *groups;
    int32 *out_ids, *out_groups;
    funcctx = SRF_FIRSTCALL_INIT();
    old = MemoryContextSwitchTo(funcctx->multi_call_memory_ctx);
    ids = int_array(PG_GETARG_ARRAYTYPE_P(0), &n);
    groups = int_array(PG_GETARG_ARRAYTYPE_P(1), &ng);
    pin_vectors();                                   /* the group vectors come from the normalized table */
    pin_pq(vecs_d);
    out_ids = palloc(sizeof(int32) * (n ? n : 1));
    out_groups = palloc(sizeof(int32) * (n ? n : 1));
    fb_check(fb_grouping_pq(engine, ids, n, groups, ng, out_ids, out_groups, &n_out));
    u = palloc(sizeof(UsrFctxGrouping));
    u->ids = out_ids;
    u->size = n_out;
    u->nearestGroup = out_groups;                    /* already group ids, not indices */
    u->groups = NULL;
    u->iter = 0;
    u->groupsSize = ng;
    u->values = palloc(2 * sizeof(char*));
    u->values[0] = palloc(18);
    u->values[1] = palloc(18);
    funcctx->user_fctx = u;
    desc = CreateTemplateTupleDesc(2);
    TupleDescInitEntry(desc, 1, "Ids", INT4OID, -1, 0);
    TupleDescInitEntry(desc, 2, "GroupIds", INT4OID, -1, 0);
    funcctx->attinmeta = TupleDescGetAttInMetadata(desc);
    MemoryContextSwitchTo(old);
  }
  funcctx = SRF_PERCALL_SETUP();
  u = (UsrFctxGrouping*)funcctx->user_fctx;
  if (u->iter >= u->size) SRF_RETURN_DONE(funcctx);
  snprintf(u->values[0], 18, "%d", u->ids[u->iter]);
  snprintf(u->values[1], 18, "%d", u->nearestGroup[u->iter]);
  u->iter++;
  SRF_RETURN_NEXT(funcctx, HeapTupleGetDatum(BuildTupleFromCStrings(funcctx->attinmeta, u->values)));
}
