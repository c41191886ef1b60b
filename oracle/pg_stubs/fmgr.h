/* stub (see postgres.h in this directory) */
#ifndef FB_STUB_FMGR_H
#define FB_STUB_FMGR_H
#endif
