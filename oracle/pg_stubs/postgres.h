/* Minimal stand-in for PostgreSQL's postgres.h — TEST INFRASTRUCTURE ONLY.
 *
 * Written from scratch for this repo: just enough typedefs/macros that the
 * reference's arithmetic translation units (freddy_extension/index_utils.c,
 * cosine_similarity.c) compile unmodified, from where they lie under
 * /root/reference, into oracle/_ref/libfreddy_ref.so (see oracle/Makefile).
 * Nothing here is shipped or linked into the product library. */
#ifndef FB_STUB_POSTGRES_H
#define FB_STUB_POSTGRES_H
#include <stdbool.h>
#include <stddef.h>
#include <stdint.h>
#include <math.h>
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
typedef unsigned int Oid;

/* varlena with a plain 4-byte length header (no TOAST, no short headers) */
typedef struct varlena { int32 vl_len_; char vl_dat[]; } varlena;
typedef varlena bytea;
typedef varlena text;
#define VARHDRSZ ((int32)sizeof(int32))
#define VARDATA(p) (((varlena*)(p))->vl_dat)
#define VARSIZE(p) (((varlena*)(p))->vl_len_)
#define SET_VARSIZE(p, n) (((varlena*)(p))->vl_len_ = (int32)(n))

#define palloc(n) malloc(n)
#define palloc0(n) calloc(1, (n))
#define repalloc(p, n) realloc((p), (n))
#define pfree(p) free(p)

#define INFO 17
#define NOTICE 18
#define WARNING 19
#define ERROR 20
#define elog(level, ...)                                   \
  do {                                                     \
    if ((level) >= WARNING) {                              \
      fprintf(stderr, "[pg-stub elog %d] ", (level));      \
      fprintf(stderr, __VA_ARGS__);                        \
      fputc('\n', stderr);                                 \
    }                                                      \
    if ((level) >= ERROR) abort();                         \
  } while (0)

static inline float4 DatumGetFloat4(Datum d) {
  union { int32 i; float4 f; } u; u.i = (int32)d; return u.f;
}
static inline Datum Float4GetDatum(float4 f) {
  union { int32 i; float4 f; } u; u.f = f; return (Datum)(uint32)u.i;
}
#define DatumGetInt32(d) ((int32)(d))
#define Int32GetDatum(i) ((Datum)(uint32)(i))
#define DatumGetPointer(d) ((void*)(d))
#define PointerGetDatum(p) ((Datum)(p))
#define DatumGetByteaP(d) ((bytea*)DatumGetPointer(d))
#endif
