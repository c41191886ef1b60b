/* stub (see ../postgres.h) */
#ifndef FB_STUB_UTILS_BUILTINS_H
#define FB_STUB_UTILS_BUILTINS_H
#include "postgres.h"
#endif
