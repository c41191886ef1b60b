/* Minimal stand-in for PostgreSQL's postgres.h — TEST INFRASTRUCTURE ONLY.
 *
 * Written from scratch for this repo: just enough typedefs/macros that the
 * reference's C translation units (freddy_extension/*.c) compile UNMODIFIED,
 * from where they lie under /root/reference, into oracle/_ref/libfreddy_ref.so
 * (see oracle/Makefile), on top of the in-memory SPI/fmgr emulator in
 * oracle/pg_emul.c.  Nothing here is shipped or linked into the product. */
#ifndef FB_STUB_POSTGRES_H
#define FB_STUB_POSTGRES_H
#include <math.h>
#include <setjmp.h>
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef uintptr_t Datum;
typedef float float4;
typedef double float8;
typedef int16_t int16;
typedef int32_t int32;
typedef int64_t int64;
typedef uint32_t uint32;
typedef uint64_t uint64;
typedef unsigned int Oid;
typedef size_t Size;

/* varlena with a plain 4-byte length header (no TOAST, no short headers) */
typedef struct varlena { int32 vl_len_; char vl_dat[]; } varlena;
typedef varlena bytea;
typedef varlena text;
#define VARHDRSZ ((int32)sizeof(int32))
#define VARDATA(p) (((varlena*)(p))->vl_dat)
#define VARSIZE(p) (((varlena*)(p))->vl_len_)
#define SET_VARSIZE(p, n) (((varlena*)(p))->vl_len_ = (int32)(n))

#define palloc(n) malloc((n) ? (n) : 1)
#define palloc0(n) calloc(1, (n) ? (n) : 1)
#define repalloc(p, n) realloc((p), (n))
#define pfree(p) free(p)

#define DEBUG1 14
#define LOG 15
#define INFO 17
#define NOTICE 18
#define WARNING 19
#define ERROR 20
/* elog(ERROR) = longjmp out of the UDF, as in Postgres; the emulator's call
 * wrappers catch it (pg_emul.c: fb_emul_error) */
extern jmp_buf* fb_emul_error_jmp;
extern char fb_emul_error_msg[256];
#define elog(level, ...)                                                 \
  do {                                                                   \
    if ((level) >= ERROR) {                                              \
      snprintf(fb_emul_error_msg, sizeof fb_emul_error_msg, __VA_ARGS__);\
      if (fb_emul_error_jmp) longjmp(*fb_emul_error_jmp, 1);             \
      fprintf(stderr, "[pg-emul ERROR] %s\n", fb_emul_error_msg);        \
      abort();                                                           \
    }                                                                    \
  } while (0)

static inline float4 DatumGetFloat4(Datum d) {
  union { int32 i; float4 f; } u; u.i = (int32)d; return u.f;
}
static inline Datum Float4GetDatum(float4 f) {
  union { int32 i; float4 f; } u; u.f = f; return (Datum)(uint32)u.i;
}
static inline float8 DatumGetFloat8(Datum d) {
  union { uint64 i; float8 f; } u; u.i = (uint64)d; return u.f;
}
static inline Datum Float8GetDatum(float8 f) {
  union { uint64 i; float8 f; } u; u.f = f; return (Datum)u.i;
}
#define DatumGetInt32(d) ((int32)(d))
#define DatumGetInt16(d) ((int16)(d))
#define DatumGetBool(d) ((bool)((d) != 0))
#define Int32GetDatum(i) ((Datum)(uint32)(i))
#define Int16GetDatum(i) ((Datum)(uint32)(uint16_t)(i))
#define BoolGetDatum(b) ((Datum)((b) ? 1 : 0))
#define DatumGetPointer(d) ((void*)(d))
#define PointerGetDatum(p) ((Datum)(p))
#define DatumGetByteaP(d) ((bytea*)DatumGetPointer(d))
#define DatumGetCString(d) ((char*)DatumGetPointer(d))
#endif
