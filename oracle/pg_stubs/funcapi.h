/* stub of funcapi.h: value-per-call SRF protocol + C-string tuple building */
#ifndef FB_STUB_FUNCAPI_H
#define FB_STUB_FUNCAPI_H
#include "fmgr.h"
typedef void* MemoryContext;
typedef struct TupleDescData { int natts; } TupleDescData;
typedef TupleDescData* TupleDesc;
typedef struct AttInMetadata { TupleDesc tupdesc; } AttInMetadata;
/* one struct serves both uses of HeapTuple: a C-string output row (natts/values) and
 * an SPI result row (a row of an emulated table + optional joined row + projection) */
typedef struct HeapTupleData {
  int natts; char** values;
  const void* table; int64 row; const void* table2; int64 row2; const int* proj; int nproj;
  Datum* bins;            /* ad-hoc result rows (no table): binary values per column, or NULL */
} HeapTupleData;
typedef HeapTupleData* HeapTuple;
typedef struct FuncCallContext {
  uint64 call_cntr;
  uint64 max_calls;
  void* user_fctx;
  AttInMetadata* attinmeta;
  MemoryContext multi_call_memory_ctx;
  TupleDesc tuple_desc;
} FuncCallContext;
#define SRF_IS_FIRSTCALL() (fcinfo->flinfo->fn_extra == NULL)
#define SRF_FIRSTCALL_INIT() ((FuncCallContext*)(fcinfo->flinfo->fn_extra = calloc(1, sizeof(FuncCallContext))))
#define SRF_PERCALL_SETUP() ((FuncCallContext*)fcinfo->flinfo->fn_extra)
#define SRF_RETURN_NEXT(funcctx, result) do { (funcctx)->call_cntr++; fcinfo->srf_state = 1; return (result); } while (0)
#define SRF_RETURN_DONE(funcctx) do { fcinfo->srf_state = 2; return (Datum)0; } while (0)
static inline MemoryContext MemoryContextSwitchTo(MemoryContext c) { return c; }
TupleDesc CreateTemplateTupleDesc(int natts);
void TupleDescInitEntry(TupleDesc desc, int attnum, const char* name, Oid typid, int32 typmod, int attdim);
AttInMetadata* TupleDescGetAttInMetadata(TupleDesc desc);
HeapTuple BuildTupleFromCStrings(AttInMetadata* attinmeta, char** values);
#define HeapTupleGetDatum(t) PointerGetDatum(t)
#endif
