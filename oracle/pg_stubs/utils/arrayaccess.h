/* stub of utils/arrayaccess.h: only so that core_functions.c:centroid() (out of scope,
 * never called through the emulator) compiles */
#ifndef FB_STUB_ARRAYACCESS_H
#define FB_STUB_ARRAYACCESS_H
#include "utils/array.h"
typedef ArrayType AnyArrayType;
typedef struct array_iter { ArrayType* a; } array_iter;
static int fb_stub_dims[2] = {0, 0};
#define AARR_NDIM(a) ((a)->ndim)
#define AARR_DIMS(a) (fb_stub_dims)
#define AARR_ELEMTYPE(a) ((a)->elemtype)
static inline int ArrayGetNItems(int ndim, const int* dims) { (void)ndim; (void)dims; return 0; }
static inline void array_iter_setup(array_iter* it, AnyArrayType* a) { it->a = a; }
static inline Datum array_iter_next(array_iter* it, bool* isnull, int i, int elmlen, bool elmbyval, char elmalign) {
  (void)isnull; (void)elmlen; (void)elmbyval; (void)elmalign; return it->a->elems[i];
}
#endif
