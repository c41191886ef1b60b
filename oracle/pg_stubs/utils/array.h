/* stub of utils/array.h: a one-dimensional array is a counted Datum vector */
#ifndef FB_STUB_ARRAY_H
#define FB_STUB_ARRAY_H
#include "postgres.h"
typedef struct ArrayType { int32 vl_len_; int ndim; int32 dataoffset; Oid elemtype; int nelems; Datum* elems; } ArrayType;
#define ARR_ELEMTYPE(a) ((a)->elemtype)
void get_typlenbyvalalign(Oid typid, int16* typlen, bool* typbyval, char* typalign);
void deconstruct_array(ArrayType* array, Oid elmtype, int elmlen, bool elmbyval, char elmalign,
                       Datum** elemsp, bool** nullsp, int* nelemsp);
ArrayType* construct_md_array(Datum* elems, bool* nulls, int ndims, int* dims, int* lbs, Oid elmtype,
                              int elmlen, bool elmbyval, char elmalign);
#endif
