/* stub (see ../postgres.h) */
#ifndef FB_STUB_UTILS_LSYSCACHE_H
#define FB_STUB_UTILS_LSYSCACHE_H
#include "postgres.h"
#endif
