/* stub (see ../postgres.h) */
#ifndef FB_STUB_BUILTINS_H
#define FB_STUB_BUILTINS_H
#endif
