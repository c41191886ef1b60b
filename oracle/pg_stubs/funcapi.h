/* stub (see postgres.h in this directory) */
#ifndef FB_STUB_FUNCAPI_H
#define FB_STUB_FUNCAPI_H
#endif
