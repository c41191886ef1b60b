/* stub: the three type OIDs the reference names */
#ifndef FB_STUB_PG_TYPE_H
#define FB_STUB_PG_TYPE_H
#define INT2OID 21
#define INT4OID 23
#define FLOAT4OID 700
#define FLOAT8OID 701
#define BYTEAOID 17
#endif
