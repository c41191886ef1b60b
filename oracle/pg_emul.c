/* pg_emul.c — a tiny in-memory stand-in for the parts of PostgreSQL the reference
 * extension touches (SPI over registered tables, fmgr argument passing, the
 * value-per-call SRF protocol, int4[]/bytea[] arrays).  TEST INFRASTRUCTURE ONLY.
 *
 * Purpose: run the reference's OWN set-returning functions (freddy.c,
 * ivpq_search_in.c, core_functions.c — compiled unmodified from /root/reference,
 * see oracle/Makefile) in this process, on in-memory copies of the index tables,
 * so that the oracle's restatement of those drivers can be checked against the
 * real thing.  Rows come back in table (heap) order, which is what the bulk-loaded
 * Postgres tables return.
 *
 * SQL understood (exactly the statements the reference's C code issues):
 *   SELECT * FROM get_<something>()                       -> config value (text / int)
 *   SELECT * FROM <coarse|codebook table> [ORDER BY pos]
 *   SELECT coarse_id, coarse_freq FROM <stat table>
 *   SELECT <cols> FROM <t> [AS fq [INNER JOIN <v> AS vecs ON fq.id = vecs.id]]
 *          [WHERE ... coarse_id IN (...) ... id IN (...)]
 */
#include "postgres.h"
#include "fmgr.h"
#include "funcapi.h"
#include "executor/spi.h"
#include "utils/array.h"
#include "catalog/pg_type.h"
#include "index_utils.h"
#include "output_utils.h"

#include <ctype.h>

int fb_stub_module_magic = 0;
jmp_buf* fb_emul_error_jmp = NULL;
char fb_emul_error_msg[256];

/* ---------------------------------------------------------------- tables -- */
enum { COL_ID = 1, COL_VEC, COL_COARSE, COL_POS, COL_CODE, COL_COUNT, COL_VEC2, COL_FREQ, COL_WORD };
enum { T_COARSE = 1, T_CODEBOOK, T_FINE, T_PQ, T_VECS, T_STATS };

typedef struct EmTable {
  char name[64];
  int kind;
  int64 nrows;
  const int32* id;      /* id / coarse_id(stat)      */
  const int32* a;       /* coarse_id (fine) / pos    */
  const int32* b;       /* code                      */
  const void* vec;      /* payload of `vector`       */
  int vec_bytes;        /* bytes per row             */
  const float* freq;    /* stat: coarse_freq         */
  bytea** cache;        /* lazily built varlenas     */
  /* what the real tables have: a btree on coarse_id (index_creation/ivfadc.py:209-210, ivpq.py) and one on id.
   * Built lazily; a statement with `coarse_id IN (...)` / `id IN (...)` then touches only the matching rows and
   * returns them in heap (table) order, as a bitmap heap scan does. */
  int64* by_a_start;    /* [max_a + 2] CSR over column a     */
  int32* by_a_rows;     /* [nrows] rows grouped by a, ascending inside a group */
  int32 max_a;
  int ids_ascending;    /* -1 unknown, 0 no, 1 strictly ascending id column */
} EmTable;

typedef struct { char key[64]; char val[64]; } EmConfig;

static EmTable g_tables[32];
static int g_ntables = 0;
static EmConfig g_config[64];
static int g_nconfig = 0;

void ref_reset(void) {
  for (int t = 0; t < g_ntables; t++) {
    if (g_tables[t].cache) {
      for (int64 r = 0; r < g_tables[t].nrows; r++) free(g_tables[t].cache[r]);
      free(g_tables[t].cache);
    }
    free(g_tables[t].by_a_start);
    free(g_tables[t].by_a_rows);
  }
  g_ntables = 0;
  g_nconfig = 0;
}

/* config-as-functions: key = the SQL function call text, e.g. "get_w()" */
void ref_set_config(const char* key, const char* val) {
  for (int i = 0; i < g_nconfig; i++)
    if (!strcmp(g_config[i].key, key)) { snprintf(g_config[i].val, 64, "%s", val); return; }
  snprintf(g_config[g_nconfig].key, 64, "%s", key);
  snprintf(g_config[g_nconfig].val, 64, "%s", val);
  g_nconfig++;
}

int ref_register_table(const char* name, int kind, int64 nrows, const int32* id, const int32* a, const int32* b,
                       const void* vec, int vec_bytes, const float* freq) {
  if (g_ntables >= 32) return -1;
  EmTable* t = &g_tables[g_ntables++];
  memset(t, 0, sizeof *t);
  snprintf(t->name, 64, "%s", name);
  t->kind = kind; t->nrows = nrows; t->id = id; t->a = a; t->b = b; t->vec = vec; t->vec_bytes = vec_bytes; t->freq = freq;
  t->ids_ascending = -1;
  return 0;
}

static EmTable* find_table(const char* name) {
  for (int t = 0; t < g_ntables; t++)
    if (!strcmp(g_tables[t].name, name)) return &g_tables[t];
  return NULL;
}

static bytea* row_bytea(EmTable* t, int64 r) {
  if (!t->cache) t->cache = calloc((size_t)(t->nrows ? t->nrows : 1), sizeof(bytea*));
  if (!t->cache[r]) {
    bytea* b = malloc(VARHDRSZ + (size_t)t->vec_bytes);
    SET_VARSIZE(b, VARHDRSZ + t->vec_bytes);
    memcpy(VARDATA(b), (const char*)t->vec + (size_t)r * t->vec_bytes, (size_t)t->vec_bytes);
    t->cache[r] = b;
  }
  return t->cache[r];
}

/* ------------------------------------------------------------------- SPI -- */
uint64 SPI_processed = 0;
SPITupleTable* SPI_tuptable = NULL;
static TupleDescData g_spi_desc = {0};
static int g_proj[16];

int SPI_connect(void) { return 1; }
int SPI_finish(void) { return 1; }  /* result memory is simply kept (tests are short-lived) */
void* SPI_palloc(Size size) { return malloc(size ? size : 1); }

static int cmp_i32(const void* x, const void* y) {
  int32 a = *(const int32*)x, b = *(const int32*)y;
  return (a > b) - (a < b);
}

/* parse "( 1, 2,3 )" starting at the '(' ; returns sorted unique array */
static int32* parse_in_list(const char* p, int* n_out) {
  int cap = 1024, n = 0;
  int32* v = malloc(sizeof(int32) * cap);
  while (*p && *p != '(') p++;
  if (*p == '(') p++;
  while (*p && *p != ')') {
    while (*p == ' ' || *p == ',') p++;
    if (*p == ')' || !*p) break;
    char* end;
    long x = strtol(p, &end, 10);
    if (end == p) break;
    if (n == cap) { cap *= 2; v = realloc(v, sizeof(int32) * cap); }
    v[n++] = (int32)x;
    p = end;
  }
  qsort(v, (size_t)n, sizeof(int32), cmp_i32);
  int m = 0;
  for (int i = 0; i < n; i++)
    if (i == 0 || v[i] != v[i - 1]) v[m++] = v[i];
  *n_out = m;
  return v;
}

static bool in_sorted(const int32* v, int n, int32 x) {
  int lo = 0, hi = n - 1;
  while (lo <= hi) {
    int mid = (lo + hi) / 2;
    if (v[mid] < x) lo = mid + 1; else if (v[mid] > x) hi = mid - 1; else return true;
  }
  return false;
}

static int64 find_id_row(const EmTable* t, int32 id) {
  /* ids ascending (bulk-loaded tables): binary search, else linear */
  int64 lo = 0, hi = t->nrows - 1;
  while (lo <= hi) {
    int64 mid = (lo + hi) / 2;
    if (t->id[mid] < id) lo = mid + 1; else if (t->id[mid] > id) hi = mid - 1; else return mid;
  }
  for (int64 r = 0; r < t->nrows; r++) if (t->id[r] == id) return r;
  return -1;
}

static void build_a_index(EmTable* t) {
  if (t->by_a_start || !t->a) return;
  int32 mx = -1;
  for (int64 r = 0; r < t->nrows; r++) if (t->a[r] > mx) mx = t->a[r];
  t->max_a = mx;
  t->by_a_start = calloc((size_t)mx + 3, sizeof(int64));
  t->by_a_rows = malloc(sizeof(int32) * (size_t)(t->nrows ? t->nrows : 1));
  for (int64 r = 0; r < t->nrows; r++) if (t->a[r] >= 0) t->by_a_start[t->a[r] + 1]++;
  for (int32 c = 0; c <= mx; c++) t->by_a_start[c + 1] += t->by_a_start[c];
  int64* cur = malloc(sizeof(int64) * ((size_t)mx + 2));
  memcpy(cur, t->by_a_start, sizeof(int64) * ((size_t)mx + 2));
  for (int64 r = 0; r < t->nrows; r++) if (t->a[r] >= 0) t->by_a_rows[cur[t->a[r]]++] = (int32)r;
  free(cur);
}

static int ids_ascending(EmTable* t) {
  if (t->ids_ascending < 0) {
    int asc = 1;
    for (int64 r = 1; r < t->nrows && asc; r++) asc = t->id[r - 1] < t->id[r];
    t->ids_ascending = asc;
  }
  return t->ids_ascending;
}

/* Results of table statements are freed a few statements later (Postgres frees them at SPI_finish; the reference
 * never looks at a tuptable after issuing four more statements): a ring of the last kKeep results. */
enum { kKeep = 4 };
static struct { HeapTuple* rows; HeapTupleData* block; SPITupleTable* tt; } g_ring[kKeep];
static int g_ring_pos = 0;

static void remember_result(HeapTuple* rows, HeapTupleData* block, SPITupleTable* tt) {
  free(g_ring[g_ring_pos].rows);
  free(g_ring[g_ring_pos].block);
  free(g_ring[g_ring_pos].tt);
  g_ring[g_ring_pos].rows = rows;
  g_ring[g_ring_pos].block = block;
  g_ring[g_ring_pos].tt = tt;
  g_ring_pos = (g_ring_pos + 1) % kKeep;
}

static void set_result(HeapTuple* rows, int64 n) {
  SPITupleTable* tt = malloc(sizeof *tt);
  tt->tupdesc = &g_spi_desc;
  tt->vals = rows;
  SPI_tuptable = tt;
  SPI_processed = (uint64)n;
}

static void next_word(const char** pp, char* out, int cap) {
  const char* p = *pp;
  while (*p == ' ') p++;
  int n = 0;
  while (*p && *p != ' ' && *p != ',' && *p != '(' && n < cap - 1) out[n++] = *p++;
  out[n] = 0;
  *pp = p;
}

/* ---- what insert_batch needs: its tokenisation statement answered from registered rows, DML recorded ---- */
static struct { int n; char** tokens; bytea** norm; bytea** raw; } g_tok;
static char* g_log = NULL;
static size_t g_log_len = 0, g_log_cap = 0;

static bytea* make_bytea0(const void* data, size_t bytes) {
  bytea* b = malloc(VARHDRSZ + bytes);
  SET_VARSIZE(b, VARHDRSZ + bytes);
  memcpy(VARDATA(b), data, bytes);
  return b;
}

/* rows `SELECT replace(term...) AS token, tokenize(term), tokenize_raw(term) FROM unnest(...)` returns */
void ref_set_tokenization(int n, const char** tokens, const float* norm, const float* raw, int d) {
  g_tok.n = n;
  g_tok.tokens = malloc(sizeof(char*) * (size_t)(n ? n : 1));
  g_tok.norm = malloc(sizeof(bytea*) * (size_t)(n ? n : 1));
  g_tok.raw = malloc(sizeof(bytea*) * (size_t)(n ? n : 1));
  for (int i = 0; i < n; i++) {
    g_tok.tokens[i] = strdup(tokens[i]);
    g_tok.norm[i] = make_bytea0(norm + (size_t)i * d, sizeof(float) * (size_t)d);
    g_tok.raw[i] = make_bytea0(raw + (size_t)i * d, sizeof(float) * (size_t)d);
  }
}
const char* ref_statement_log(void) { return g_log ? g_log : ""; }
void ref_clear_statement_log(void) { g_log_len = 0; if (g_log) g_log[0] = 0; }
static void log_statement(const char* src) {
  const size_t n = strlen(src);
  if (g_log_len + n + 2 > g_log_cap) { g_log_cap = (g_log_len + n + 2) * 2; g_log = realloc(g_log, g_log_cap); }
  memcpy(g_log + g_log_len, src, n);
  g_log_len += n;
  g_log[g_log_len++] = '\n';
  g_log[g_log_len] = 0;
}

int SPI_exec(const char* src, long tcount) {
  (void)tcount;
  SPI_processed = 0;
  SPI_tuptable = NULL;
  if (!strncmp(src, "INSERT ", 7) || !strncmp(src, "UPDATE ", 7)) {   /* recorded, not applied (tables are read-only images) */
    log_statement(src);
    SPI_processed = 1;
    return 1;
  }
  if (strstr(src, "tokenize(") != NULL) {
    HeapTuple* rows = malloc(sizeof(HeapTuple) * (size_t)(g_tok.n ? g_tok.n : 1));
    for (int i = 0; i < g_tok.n; i++) {
      rows[i] = calloc(1, sizeof(HeapTupleData));
      rows[i]->natts = 3;
      rows[i]->values = malloc(sizeof(char*) * 3);
      rows[i]->values[0] = g_tok.tokens[i];
      rows[i]->values[1] = rows[i]->values[2] = NULL;
      rows[i]->bins = malloc(sizeof(Datum) * 3);
      rows[i]->bins[0] = PointerGetDatum(g_tok.tokens[i]);
      rows[i]->bins[1] = PointerGetDatum(g_tok.norm[i]);
      rows[i]->bins[2] = PointerGetDatum(g_tok.raw[i]);
    }
    set_result(rows, g_tok.n);
    return 1;
  }
  if (!strncmp(src, "SELECT max(id) FROM ", 20)) {                    /* the shim's cheap "did the table grow" probe */
    char tn[64];
    const char* pp = src + 20;
    next_word(&pp, tn, sizeof tn);
    EmTable* mt = find_table(tn);
    if (!mt) { elog(ERROR, "pg_emul: unknown table %s", tn); }
    int32 mx = 0;
    for (int64 r = 0; r < mt->nrows; r++) if (r == 0 || mt->id[r] > mx) mx = mt->id[r];
    HeapTuple* rows = malloc(sizeof(HeapTuple));
    rows[0] = calloc(1, sizeof(HeapTupleData));
    rows[0]->natts = 1;
    rows[0]->values = malloc(sizeof(char*));
    rows[0]->values[0] = malloc(16);
    snprintf(rows[0]->values[0], 16, "%d", mx);
    set_result(rows, 1);
    return 1;
  }
  const char* from = strstr(src, " FROM ");
  if (strncmp(src, "SELECT ", 7) != 0 || !from) { elog(ERROR, "pg_emul: unsupported statement: %.80s", src); }
  const char* p = from + 6;
  char tname[64];
  next_word(&p, tname, sizeof tname);

  /* --- config-as-functions --- */
  if (!strncmp(tname, "get_", 4)) {
    char key[80];
    snprintf(key, sizeof key, "%s()", tname);  /* next_word stopped at '(' */
    for (int i = 0; i < g_nconfig; i++) {
      if (!strcmp(g_config[i].key, key)) {
        HeapTuple* rows = malloc(sizeof(HeapTuple));
        rows[0] = calloc(1, sizeof(HeapTupleData));
        rows[0]->natts = 1;
        rows[0]->values = malloc(sizeof(char*));
        rows[0]->values[0] = g_config[i].val;
        set_result(rows, 1);
        return 1;
      }
    }
    elog(ERROR, "pg_emul: unknown config function %s", key);
  }

  EmTable* t = find_table(tname);
  if (!t) { elog(ERROR, "pg_emul: unknown table %s", tname); }
  EmTable* t2 = NULL;
  const char* join = strstr(src, " INNER JOIN ");
  if (join) {
    const char* q = join + 12;
    char jname[64];
    next_word(&q, jname, sizeof jname);
    t2 = find_table(jname);
    if (!t2) { elog(ERROR, "pg_emul: unknown join table %s", jname); }
  }

  /* --- projection --- */
  int nproj = 0;
  int* proj = malloc(sizeof(int) * 8);
  {
    char list[256];
    size_t len = (size_t)(from - (src + 7));
    if (len >= sizeof list) len = sizeof list - 1;
    memcpy(list, src + 7, len);
    list[len] = 0;
    if (strchr(list, '*')) {
      if (t->kind == T_COARSE) { proj[0] = COL_ID; proj[1] = COL_VEC; proj[2] = COL_COUNT; nproj = 3; }
      else if (t->kind == T_CODEBOOK) { proj[0] = COL_ID; proj[1] = COL_POS; proj[2] = COL_CODE; proj[3] = COL_VEC; proj[4] = COL_COUNT; nproj = 5; }
      else { elog(ERROR, "pg_emul: SELECT * on table kind %d", t->kind); }
    } else {
      char* save = NULL;
      for (char* tok = strtok_r(list, ",", &save); tok; tok = strtok_r(NULL, ",", &save)) {
        while (*tok == ' ') tok++;
        char* e = tok + strlen(tok);
        while (e > tok && e[-1] == ' ') *--e = 0;
        const char* col = strrchr(tok, '.');
        col = col ? col + 1 : tok;
        int c = 0;
        if (!strcmp(col, "id")) c = COL_ID;
        else if (!strcmp(col, "vector")) c = (!strncmp(tok, "vecs.", 5)) ? COL_VEC2 : COL_VEC;
        else if (!strcmp(col, "coarse_id")) c = (t->kind == T_STATS) ? COL_ID : COL_COARSE;
        else if (!strcmp(col, "coarse_freq")) c = COL_FREQ;
        else { elog(ERROR, "pg_emul: unknown column %s", tok); }
        proj[nproj++] = c;
      }
    }
  }

  /* --- predicates --- */
  int n_cids = 0, n_ids = 0;
  int32 *cids = NULL, *ids = NULL;
  const char* w = strstr(src, " WHERE ");
  if (w) {
    const char* c = strstr(w, "coarse_id IN");
    if (c) cids = parse_in_list(c + 12, &n_cids);
    const char* s = w;
    const char* hit = NULL;
    while ((s = strstr(s, "id IN")) != NULL) {       /* "id IN" not preceded by "coarse_" */
      if (!(s - src >= 7 && !strncmp(s - 7, "coarse_", 7))) { hit = s; break; }
      s += 5;
    }
    if (hit) ids = parse_in_list(hit + 5, &n_ids);
  }
  const bool order_by_pos = strstr(src, "ORDER BY pos") != NULL;

  /* candidate rows in table order: through the coarse_id / id indexes when the statement has such a predicate */
  int64 n_cand = 0;
  int32* cand = NULL;           /* NULL: every row */
  if (cids && t->a) {
    build_a_index(t);
    int64 tot = 0;
    for (int i = 0; i < n_cids; i++)
      if (cids[i] >= 0 && cids[i] <= t->max_a) tot += t->by_a_start[cids[i] + 1] - t->by_a_start[cids[i]];
    cand = malloc(sizeof(int32) * (size_t)(tot ? tot : 1));
    for (int i = 0; i < n_cids; i++)
      if (cids[i] >= 0 && cids[i] <= t->max_a)
        for (int64 x = t->by_a_start[cids[i]]; x < t->by_a_start[cids[i] + 1]; x++) cand[n_cand++] = t->by_a_rows[x];
    if (n_cids > 1) qsort(cand, (size_t)n_cand, sizeof(int32), cmp_i32);   /* heap order across the lists */
  } else if (ids && ids_ascending(t)) {
    cand = malloc(sizeof(int32) * (size_t)(n_ids ? n_ids : 1));
    for (int i = 0; i < n_ids; i++) {                                      /* ids[] sorted unique, id column ascending */
      int64 lo = 0, hi = t->nrows - 1;
      while (lo <= hi) {
        int64 mid = (lo + hi) / 2;
        if (t->id[mid] < ids[i]) lo = mid + 1; else if (t->id[mid] > ids[i]) hi = mid - 1; else { cand[n_cand++] = (int32)mid; break; }
      }
    }
  }
  const int64 n_scan = cand ? n_cand : t->nrows;
  HeapTuple* rows = malloc(sizeof(HeapTuple) * (size_t)(n_scan ? n_scan : 1));
  HeapTupleData* block = calloc((size_t)(n_scan ? n_scan : 1), sizeof(HeapTupleData));
  int64 n = 0;
  for (int64 x = 0; x < n_scan; x++) {
    const int64 r = cand ? cand[x] : x;
    if (cids && !in_sorted(cids, n_cids, t->a[r])) continue;
    if (ids && !in_sorted(ids, n_ids, t->id[r])) continue;
    int64 r2 = -1;
    if (t2) { r2 = find_id_row(t2, t->id[r]); if (r2 < 0) continue; }
    HeapTuple h = &block[n];
    h->table = t; h->row = r; h->table2 = t2; h->row2 = r2; h->proj = proj; h->nproj = nproj;
    rows[n++] = h;
  }
  free(cand);
  if (order_by_pos && t->kind == T_CODEBOOK) {       /* stable sort by pos */
    HeapTuple* sorted = malloc(sizeof(HeapTuple) * (size_t)(n ? n : 1));
    int maxpos = 0;
    for (int64 i = 0; i < n; i++) if (t->a[rows[i]->row] > maxpos) maxpos = t->a[rows[i]->row];
    int64 k = 0;
    for (int pos = 0; pos <= maxpos; pos++)
      for (int64 i = 0; i < n; i++) if (t->a[rows[i]->row] == pos) sorted[k++] = rows[i];
    free(rows);
    rows = sorted;
  }
  free(cids); free(ids);
  set_result(rows, n);
  remember_result(rows, block, SPI_tuptable);
  return 1;
}

int SPI_execute(const char* src, bool read_only, long tcount) { (void)read_only; return SPI_exec(src, tcount); }

Datum SPI_getbinval(HeapTuple h, TupleDesc desc, int fnumber, bool* isnull) {
  (void)desc;
  if (isnull) *isnull = false;
  if (h->table == NULL && h->bins != NULL) return h->bins[fnumber - 1];         /* ad-hoc row with binary columns */
  if (h->table == NULL) return Int32GetDatum(atoi(h->values[fnumber - 1]));  /* config row */
  EmTable* t = (EmTable*)h->table;
  if (fnumber < 1 || fnumber > h->nproj) { elog(ERROR, "pg_emul: column %d out of range", fnumber); }
  switch (h->proj[fnumber - 1]) {
    case COL_ID: return Int32GetDatum(t->id[h->row]);
    case COL_COARSE: case COL_POS: return Int32GetDatum(t->a[h->row]);
    case COL_CODE: return Int32GetDatum(t->b[h->row]);
    case COL_COUNT: return Int32GetDatum(100);   /* codebook `count` column (insert_batch's running means) */
    case COL_VEC: return PointerGetDatum(row_bytea(t, h->row));
    case COL_VEC2: return PointerGetDatum(row_bytea((EmTable*)h->table2, h->row2));
    case COL_FREQ: return Float4GetDatum(t->freq[h->row]);
  }
  elog(ERROR, "pg_emul: bad projection");
  return 0;
}

char* SPI_getvalue(HeapTuple h, TupleDesc desc, int fnumber) {
  (void)desc;
  if (h->table == NULL) return h->values[fnumber - 1];
  elog(ERROR, "pg_emul: SPI_getvalue on a table row");
  return NULL;
}

/* ---------------------------------------------------------------- arrays -- */
void get_typlenbyvalalign(Oid typid, int16* typlen, bool* typbyval, char* typalign) {
  (void)typid; *typlen = 4; *typbyval = true; *typalign = 'i';
}
void deconstruct_array(ArrayType* a, Oid e, int l, bool b, char al, Datum** elemsp, bool** nullsp, int* nelemsp) {
  (void)e; (void)l; (void)b; (void)al;
  *elemsp = malloc(sizeof(Datum) * (size_t)(a->nelems ? a->nelems : 1));
  memcpy(*elemsp, a->elems, sizeof(Datum) * (size_t)a->nelems);
  if (nullsp) *nullsp = calloc((size_t)(a->nelems ? a->nelems : 1), sizeof(bool));
  *nelemsp = a->nelems;
}
ArrayType* construct_md_array(Datum* elems, bool* nulls, int ndims, int* dims, int* lbs, Oid elmtype, int elmlen,
                              bool elmbyval, char elmalign) {
  (void)nulls; (void)ndims; (void)lbs; (void)elmlen; (void)elmbyval; (void)elmalign;
  ArrayType* a = calloc(1, sizeof *a);
  a->ndim = 1; a->elemtype = elmtype; a->nelems = dims[0];
  a->elems = malloc(sizeof(Datum) * (size_t)(dims[0] ? dims[0] : 1));
  memcpy(a->elems, elems, sizeof(Datum) * (size_t)dims[0]);
  return a;
}

/* ------------------------------------------------------------- tuples/SRF -- */
TupleDesc CreateTemplateTupleDesc(int natts) { TupleDesc d = calloc(1, sizeof *d); d->natts = natts; return d; }
void TupleDescInitEntry(TupleDesc d, int attnum, const char* name, Oid typid, int32 typmod, int attdim) {
  (void)d; (void)attnum; (void)name; (void)typid; (void)typmod; (void)attdim;
}
AttInMetadata* TupleDescGetAttInMetadata(TupleDesc d) { AttInMetadata* a = calloc(1, sizeof *a); a->tupdesc = d; return a; }
HeapTuple BuildTupleFromCStrings(AttInMetadata* am, char** values) {
  HeapTuple h = calloc(1, sizeof(HeapTupleData));
  h->natts = am->tupdesc->natts;
  h->values = malloc(sizeof(char*) * (size_t)h->natts);
  for (int i = 0; i < h->natts; i++) h->values[i] = strdup(values[i]);
  return h;
}

/* ---------------------------------------------------- call wrappers (API) -- */
extern Datum ivfadc_search(PG_FUNCTION_ARGS);
extern Datum pq_search(PG_FUNCTION_ARGS);
extern Datum pq_search_in(PG_FUNCTION_ARGS);
extern Datum pq_search_in_batch(PG_FUNCTION_ARGS);
extern Datum ivfadc_batch_search(PG_FUNCTION_ARGS);
extern Datum ivpq_search_in(PG_FUNCTION_ARGS);
extern Datum cosine_similarity_bytea(PG_FUNCTION_ARGS);
extern Datum vec_minus_bytea(PG_FUNCTION_ARGS);
extern Datum vec_plus_bytea(PG_FUNCTION_ARGS);
extern Datum vec_normalize_bytea(PG_FUNCTION_ARGS);

static bytea* make_bytea(const void* data, size_t bytes) {
  bytea* b = malloc(VARHDRSZ + bytes);
  SET_VARSIZE(b, VARHDRSZ + bytes);
  memcpy(VARDATA(b), data, bytes);
  return b;
}
static ArrayType* make_int_array(const int32* v, int n) {
  ArrayType* a = calloc(1, sizeof *a);
  a->ndim = 1; a->elemtype = INT4OID; a->nelems = n;
  a->elems = malloc(sizeof(Datum) * (size_t)(n ? n : 1));
  for (int i = 0; i < n; i++) a->elems[i] = Int32GetDatum(v[i]);
  return a;
}
static ArrayType* make_bytea_array(const float* q, int nq, int d) {
  ArrayType* a = calloc(1, sizeof *a);
  a->ndim = 1; a->elemtype = BYTEAOID; a->nelems = nq;
  a->elems = malloc(sizeof(Datum) * (size_t)(nq ? nq : 1));
  for (int i = 0; i < nq; i++) a->elems[i] = PointerGetDatum(make_bytea(q + (size_t)i * d, sizeof(float) * (size_t)d));
  return a;
}

/* Drives a value-per-call SRF to completion.  Rows are returned as text exactly as
 * the SRF emits them (ncols strings of <= 15 chars per row). Returns #rows or -1. */
static int run_srf(Datum (*fn)(FunctionCallInfo), FunctionCallInfo fcinfo, int ncols, int max_rows, char* text_out,
                   void** user_fctx_out) {
  jmp_buf env;
  FmgrInfo fl = {0};
  fcinfo->flinfo = &fl;
  int n = 0;
  fb_emul_error_jmp = &env;
  if (setjmp(env)) { fb_emul_error_jmp = NULL; return -1; }
  for (;;) {
    fcinfo->srf_state = 0;
    Datum r = fn(fcinfo);
    if (fcinfo->srf_state != 1) break;
    if (user_fctx_out && n == 0) *user_fctx_out = ((FuncCallContext*)fl.fn_extra)->user_fctx;
    HeapTuple h = (HeapTuple)DatumGetPointer(r);
    if (n < max_rows)
      for (int c = 0; c < ncols; c++) snprintf(text_out + ((size_t)n * ncols + c) * 16, 16, "%s", h->values[c]);
    n++;
  }
  fb_emul_error_jmp = NULL;
  return n;
}

const char* ref_last_error(void) { return fb_emul_error_msg; }

/* ivfadc_search(bytea, int): ids / raw fp32 distances (from the SRF's own TopK array) /
 * text-rounded distances as emitted.  returns k or -1 */
int ref_ivfadc_search(const float* q, int d, int k, int32* ids, float* raw, float* as_text) {
  FunctionCallInfoData fc = {0};
  fc.args[0] = PointerGetDatum(make_bytea(q, sizeof(float) * (size_t)d));
  fc.args[1] = Int32GetDatum(k);
  fc.nargs = 2;
  char* text = malloc((size_t)k * 2 * 16);
  void* uf = NULL;
  int n = run_srf(ivfadc_search, &fc, 2, k, text, &uf);
  if (n == k) {
    UsrFctx* u = (UsrFctx*)uf;
    for (int i = 0; i < k; i++) {
      ids[i] = atoi(text + (size_t)(i * 2) * 16);
      as_text[i] = strtof(text + (size_t)(i * 2 + 1) * 16, NULL);
      raw[i] = u->tk[i].distance;
      if (u->tk[i].id != ids[i]) n = -2;
    }
  }
  free(text);
  return n;
}

int ref_pq_search(const float* q, int d, int k, int32* ids, float* raw) {
  FunctionCallInfoData fc = {0};
  fc.args[0] = PointerGetDatum(make_bytea(q, sizeof(float) * (size_t)d));
  fc.args[1] = Int32GetDatum(k);
  char* text = malloc((size_t)k * 2 * 16);
  void* uf = NULL;
  int n = run_srf(pq_search, &fc, 2, k, text, &uf);
  if (n == k) for (int i = 0; i < k; i++) { ids[i] = ((UsrFctx*)uf)->tk[i].id; raw[i] = ((UsrFctx*)uf)->tk[i].distance; }
  free(text);
  return n;
}

int ref_pq_search_in(const float* q, int d, int k, const int32* targets, int nt, int32* ids, float* raw) {
  FunctionCallInfoData fc = {0};
  fc.args[0] = PointerGetDatum(make_bytea(q, sizeof(float) * (size_t)d));
  fc.args[1] = Int32GetDatum(k);
  fc.args[2] = PointerGetDatum(make_int_array(targets, nt));
  char* text = malloc((size_t)k * 2 * 16);
  void* uf = NULL;
  int n = run_srf(pq_search_in, &fc, 2, k, text, &uf);
  if (n == k) for (int i = 0; i < k; i++) { ids[i] = ((UsrFctx*)uf)->tk[i].id; raw[i] = ((UsrFctx*)uf)->tk[i].distance; }
  free(text);
  return n;
}

/* batch SRFs: rows (query_id, target_id, distance) -> ids[nq*k], raw[nq*k], qids_out[nq*k] */
static int collect_batch(void* uf, int n, int nq, int k, const char* text, int32* qids_out, int32* ids, float* raw) {
  if (n != nq * k) return n < 0 ? n : -3;
  UsrFctxBatch* u = (UsrFctxBatch*)uf;
  for (int i = 0; i < nq; i++)
    for (int j = 0; j < k; j++) {
      qids_out[i * k + j] = atoi(text + (size_t)((i * k + j) * 3) * 16);
      ids[i * k + j] = u->tk[i][j].id;
      raw[i * k + j] = u->tk[i][j].distance;
      if (atoi(text + (size_t)((i * k + j) * 3 + 1) * 16) != ids[i * k + j]) return -2;
    }
  return n;
}

int ref_pq_search_in_batch(const float* q, int nq, int d, const int32* qids, int k, const int32* targets, int nt,
                           int use_tl, int32* qids_out, int32* ids, float* raw) {
  FunctionCallInfoData fc = {0};
  fc.args[0] = PointerGetDatum(make_bytea_array(q, nq, d));
  fc.args[1] = PointerGetDatum(make_int_array(qids, nq));
  fc.args[2] = Int32GetDatum(k);
  fc.args[3] = PointerGetDatum(make_int_array(targets, nt));
  fc.args[4] = BoolGetDatum(use_tl);
  char* text = malloc((size_t)nq * k * 3 * 16 + 16);
  void* uf = NULL;
  int n = run_srf(pq_search_in_batch, &fc, 3, nq * k, text, &uf);
  n = collect_batch(uf, n, nq, k, text, qids_out, ids, raw);
  free(text);
  return n;
}

/* ivfadc_batch_search(int[] ids, int k): queries are fetched by id from the normalized table;
 * the SRF reports rows per fetched vector (table order), nq_out = number of vectors found */
int ref_ivfadc_batch_search(const int32* query_ids, int nq, int k, int max_q, int* nq_out, int32* qids_out, int32* ids,
                            float* raw) {
  FunctionCallInfoData fc = {0};
  fc.args[0] = PointerGetDatum(make_int_array(query_ids, nq));
  fc.args[1] = Int32GetDatum(k);
  char* text = malloc((size_t)max_q * k * 3 * 16 + 16);
  void* uf = NULL;
  int n = run_srf(ivfadc_batch_search, &fc, 3, max_q * k, text, &uf);
  if (n < 0) { free(text); return n; }
  UsrFctxBatch* u = (UsrFctxBatch*)uf;
  *nq_out = u->queryIdsSize;
  n = collect_batch(uf, n, u->queryIdsSize, k, text, qids_out, ids, raw);
  free(text);
  return n;
}

int ref_ivpq_search_in(const float* q, int nq, int d, const int32* qids, int k, const int32* targets, int nt, int alpha,
                       int pvf, int method, int use_tl, float confidence, int dbl_threshold, int32* qids_out,
                       int32* ids, float* raw) {
  FunctionCallInfoData fc = {0};
  fc.args[0] = PointerGetDatum(make_bytea_array(q, nq, d));
  fc.args[1] = PointerGetDatum(make_int_array(qids, nq));
  fc.args[2] = Int32GetDatum(k);
  fc.args[3] = PointerGetDatum(make_int_array(targets, nt));
  fc.args[4] = Int32GetDatum(alpha);
  fc.args[5] = Int32GetDatum(pvf);
  fc.args[6] = Int32GetDatum(method);
  fc.args[7] = BoolGetDatum(use_tl);
  fc.args[8] = Float4GetDatum(confidence);
  fc.args[9] = Int32GetDatum(dbl_threshold);
  char* text = malloc((size_t)nq * k * 3 * 16 + 16);
  void* uf = NULL;
  int n = run_srf(ivpq_search_in, &fc, 3, nq * k, text, &uf);
  n = collect_batch(uf, n, nq, k, text, qids_out, ids, raw);
  free(text);
  return n;
}

/* scalar / bytea UDFs of core_functions.c */
static Datum call_plain(Datum (*fn)(FunctionCallInfo), FunctionCallInfo fc) {
  FmgrInfo fl = {0};
  fc->flinfo = &fl;
  return fn(fc);
}
float ref_cosine_similarity_bytea(const float* a, const float* b, int d) {
  FunctionCallInfoData fc = {0};
  fc.args[0] = PointerGetDatum(make_bytea(a, sizeof(float) * (size_t)d));
  fc.args[1] = PointerGetDatum(make_bytea(b, sizeof(float) * (size_t)d));
  return DatumGetFloat4(call_plain(cosine_similarity_bytea, &fc));
}
static void bytea_binop(Datum (*fn)(FunctionCallInfo), const float* a, const float* b, int d, float* out) {
  FunctionCallInfoData fc = {0};
  fc.args[0] = PointerGetDatum(make_bytea(a, sizeof(float) * (size_t)d));
  if (b) fc.args[1] = PointerGetDatum(make_bytea(b, sizeof(float) * (size_t)d));
  bytea* r = DatumGetByteaP(call_plain(fn, &fc));
  memcpy(out, VARDATA(r), sizeof(float) * (size_t)d);
}
void ref_vec_minus_bytea(const float* a, const float* b, int d, float* out) { bytea_binop(vec_minus_bytea, a, b, d, out); }
void ref_vec_plus_bytea(const float* a, const float* b, int d, float* out) { bytea_binop(vec_plus_bytea, a, b, d, out); }
void ref_vec_normalize_bytea(const float* a, int d, float* out) { bytea_binop(vec_normalize_bytea, a, NULL, d, out); }

/* grouping_pq(int[] ids, int[] group_ids) -> rows (id, group id).  returns the number of rows or -1 (elog ERROR) */
extern Datum grouping_pq(FunctionCallInfo fcinfo);
int ref_grouping_pq(const int32* ids, int n, const int32* groups, int ng, int32* out_ids, int32* out_groups) {
  FunctionCallInfoData fc = {0};
  fc.args[0] = PointerGetDatum(make_int_array(ids, n));
  fc.args[1] = PointerGetDatum(make_int_array(groups, ng));
  char* text = malloc((size_t)(n > 0 ? n : 1) * 2 * 16 + 16);
  int rows = run_srf(grouping_pq, &fc, 2, n, text, NULL);
  for (int i = 0; i < rows && i < n; i++) {
    out_ids[i] = atoi(text + (size_t)(i * 2) * 16);
    out_groups[i] = atoi(text + (size_t)(i * 2 + 1) * 16);
  }
  free(text);
  return rows;
}

/* ---- insert_batch(varchar[]) through fmgr: the reference's (freddy.c:1403-1658) or the shim's.  DML statements
 * land in the statement log (ref_statement_log).  returns 0 or -1 (elog ERROR) ---- */
extern Datum insert_batch(PG_FUNCTION_ARGS);
int ref_insert_batch(int n, const char** terms) {
  FunctionCallInfoData fc = {0};
  ArrayType* a = calloc(1, sizeof *a);
  a->ndim = 1; a->elemtype = 1043 /* VARCHAROID */; a->nelems = n;
  a->elems = malloc(sizeof(Datum) * (size_t)(n ? n : 1));
  for (int i = 0; i < n; i++) a->elems[i] = PointerGetDatum(make_bytea(terms[i], strlen(terms[i])));
  fc.args[0] = PointerGetDatum(a);
  fc.nargs = 1;
  jmp_buf env;
  FmgrInfo fl = {0};
  fc.flinfo = &fl;
  fb_emul_error_jmp = &env;
  if (setjmp(env)) { fb_emul_error_jmp = NULL; return -1; }
  insert_batch(&fc);
  fb_emul_error_jmp = NULL;
  return 0;
}

/* ---- SRFs that exist only in the shim (GPU paths the reference reaches through plpgsql + per-row UDF calls) ---- */
extern Datum knn_exact_search(PG_FUNCTION_ARGS) __attribute__((weak));
extern Datum knn_in_exact_search(PG_FUNCTION_ARGS) __attribute__((weak));
extern Datum ivfadc_search_pv(PG_FUNCTION_ARGS) __attribute__((weak));
extern Datum pq_search_pv(PG_FUNCTION_ARGS) __attribute__((weak));
extern Datum analogy_3cosadd_batch(PG_FUNCTION_ARGS) __attribute__((weak));
extern Datum cosine_similarity_batch(PG_FUNCTION_ARGS) __attribute__((weak));
extern Datum freddy_repin(PG_FUNCTION_ARGS) __attribute__((weak));
extern Datum freddy_sidecar_serve(PG_FUNCTION_ARGS) __attribute__((weak));
extern Datum freddy_sidecar_stop(PG_FUNCTION_ARGS) __attribute__((weak));

static int collect_single_f32(int n, int k, const char* text, int32* ids, float* vals) {
  for (int i = 0; i < n && i < k; i++) {
    ids[i] = atoi(text + (size_t)(i * 2) * 16);
    vals[i] = strtof(text + (size_t)(i * 2 + 1) * 16, NULL);
  }
  return n;
}
/* targets == NULL: k_nearest_neighbour(bytea, k); else knn_in_exact(bytea, k, int[]) */
int ref_knn_exact_search(const float* q, int d, int k, const int32* targets, int nt, int32* ids, float* sims) {
  if (!knn_exact_search || !knn_in_exact_search) return -3;
  FunctionCallInfoData fc = {0};
  fc.args[0] = PointerGetDatum(make_bytea(q, sizeof(float) * (size_t)d));
  fc.args[1] = Int32GetDatum(k);
  if (targets) fc.args[2] = PointerGetDatum(make_int_array(targets, nt));
  char* text = malloc((size_t)k * 2 * 16 + 16);
  int n = run_srf(targets ? knn_in_exact_search : knn_exact_search, &fc, 2, k, text, NULL);
  if (n > 0) collect_single_f32(n, k, text, ids, sims);
  free(text);
  return n;
}
int ref_ivfadc_search_pv(const float* q, int d, int k, int32* ids, float* sims) {
  if (!ivfadc_search_pv) return -3;
  FunctionCallInfoData fc = {0};
  fc.args[0] = PointerGetDatum(make_bytea(q, sizeof(float) * (size_t)d));
  fc.args[1] = Int32GetDatum(k);
  char* text = malloc((size_t)k * 2 * 16 + 16);
  int n = run_srf(ivfadc_search_pv, &fc, 2, k, text, NULL);
  if (n > 0) collect_single_f32(n, k, text, ids, sims);
  free(text);
  return n;
}
int ref_pq_search_pv(const float* q, int d, int k, int32* ids, float* sims) {
  if (!pq_search_pv) return -3;
  FunctionCallInfoData fc = {0};
  fc.args[0] = PointerGetDatum(make_bytea(q, sizeof(float) * (size_t)d));
  fc.args[1] = Int32GetDatum(k);
  char* text = malloc((size_t)k * 2 * 16 + 16);
  int n = run_srf(pq_search_pv, &fc, 2, k, text, NULL);
  if (n > 0) collect_single_f32(n, k, text, ids, sims);
  free(text);
  return n;
}
/* ids_abc: [n][3] word ids; rows (query index, winner id, score) */
int ref_analogy_3cosadd_batch(const int32* ids_abc, int n, int32* out_ids, float* out_scores) {
  if (!analogy_3cosadd_batch) return -3;
  FunctionCallInfoData fc = {0};
  fc.args[0] = PointerGetDatum(make_int_array(ids_abc, 3 * n));
  char* text = malloc((size_t)(n ? n : 1) * 3 * 16 + 16);
  int rows = run_srf(analogy_3cosadd_batch, &fc, 3, n, text, NULL);
  for (int i = 0; i < rows && i < n; i++) {
    if (atoi(text + (size_t)(i * 3) * 16) != i) { rows = -2; break; }
    out_ids[i] = atoi(text + (size_t)(i * 3 + 1) * 16);
    out_scores[i] = strtof(text + (size_t)(i * 3 + 2) * 16, NULL);
  }
  free(text);
  return rows;
}
/* variant 0: cosine_similarity, 1: cosine_similarity_norm, 2: cosine_similarity_bytea; rows (index, similarity as %.17g) */
int ref_cosine_similarity_batch(const float* a, const float* b, int n, int d, int variant, double* out) {
  if (!cosine_similarity_batch) return -3;
  FunctionCallInfoData fc = {0};
  fc.args[0] = PointerGetDatum(make_bytea_array(a, n, d));
  fc.args[1] = PointerGetDatum(make_bytea_array(b, n, d));
  fc.args[2] = Int32GetDatum(variant);
  jmp_buf env;
  FmgrInfo fl = {0};
  fc.flinfo = &fl;
  int rows = 0;
  fb_emul_error_jmp = &env;
  if (setjmp(env)) { fb_emul_error_jmp = NULL; return -1; }
  for (;;) {
    fc.srf_state = 0;
    Datum r = cosine_similarity_batch(&fc);
    if (fc.srf_state != 1) break;
    HeapTuple h = (HeapTuple)DatumGetPointer(r);
    if (rows < n) out[rows] = strtod(h->values[1], NULL);
    rows++;
  }
  fb_emul_error_jmp = NULL;
  return rows;
}
/* the shim's sidecar entry points (blocking: call the first from a thread of its own) */
int ref_freddy_sidecar_serve(int d, int max_k, int seconds) {
  if (!freddy_sidecar_serve) return -3;
  FunctionCallInfoData fc = {0};
  fc.args[0] = Int32GetDatum(d);
  fc.args[1] = Int32GetDatum(max_k);
  fc.args[2] = Int32GetDatum(seconds);
  return DatumGetInt32(call_plain(freddy_sidecar_serve, &fc));
}
int ref_freddy_sidecar_stop(void) {
  if (!freddy_sidecar_stop) return -3;
  FunctionCallInfoData fc = {0};
  return DatumGetInt32(call_plain(freddy_sidecar_stop, &fc));
}
int ref_freddy_repin(void) {
  if (!freddy_repin) return -3;
  FunctionCallInfoData fc = {0};
  call_plain(freddy_repin, &fc);
  return 0;
}
