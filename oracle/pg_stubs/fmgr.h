/* stub of fmgr.h (see postgres.h in this directory): version-1 calling convention
 * reduced to an argument vector */
#ifndef FB_STUB_FMGR_H
#define FB_STUB_FMGR_H
#include "postgres.h"
typedef struct FmgrInfo { void* fn_extra; } FmgrInfo;
typedef struct FunctionCallInfoData {
  FmgrInfo* flinfo;
  Datum args[16];
  int nargs;
  bool isnull;
  int srf_state; /* set by SRF_RETURN_NEXT (1) / SRF_RETURN_DONE (2) */
} FunctionCallInfoData;
typedef FunctionCallInfoData* FunctionCallInfo;
#define PG_FUNCTION_ARGS FunctionCallInfo fcinfo
#define PG_FUNCTION_INFO_V1(f) extern Datum f(PG_FUNCTION_ARGS)
#define PG_MODULE_MAGIC extern int fb_stub_module_magic
#define PG_GETARG_DATUM(n) (fcinfo->args[n])
#define PG_GETARG_INT32(n) DatumGetInt32(PG_GETARG_DATUM(n))
#define PG_GETARG_BOOL(n) DatumGetBool(PG_GETARG_DATUM(n))
#define PG_GETARG_FLOAT4(n) DatumGetFloat4(PG_GETARG_DATUM(n))
#define PG_GETARG_BYTEA_P(n) DatumGetByteaP(PG_GETARG_DATUM(n))
#define PG_GETARG_ARRAYTYPE_P(n) ((ArrayType*)DatumGetPointer(PG_GETARG_DATUM(n)))
#define PG_RETURN_DATUM(x) return (x)
#define PG_RETURN_INT32(x) return Int32GetDatum(x)
#define PG_RETURN_FLOAT4(x) return Float4GetDatum(x)
#define PG_RETURN_FLOAT8(x) return Float8GetDatum(x)
#define PG_RETURN_BYTEA_P(x) return PointerGetDatum(x)
#define PG_RETURN_ARRAYTYPE_P(x) return PointerGetDatum(x)
#define PG_RETURN_NULL() do { fcinfo->isnull = true; return (Datum)0; } while (0)
#endif
