/* stub (see ../postgres.h): SPI is never reachable from the arithmetic the
 * oracle exercises; the symbols exist so the reference file links. */
#ifndef FB_STUB_SPI_H
#define FB_STUB_SPI_H
#include "postgres.h"
typedef struct TupleDescData* TupleDesc;
typedef struct HeapTupleData* HeapTuple;
typedef struct SPITupleTable { TupleDesc tupdesc; HeapTuple* vals; } SPITupleTable;
extern uint64_t SPI_processed;
extern SPITupleTable* SPI_tuptable;
int SPI_connect(void);
int SPI_finish(void);
int SPI_exec(const char* src, long tcount);
Datum SPI_getbinval(HeapTuple tuple, TupleDesc tupdesc, int fnumber, bool* isnull);
char* SPI_getvalue(HeapTuple tuple, TupleDesc tupdesc, int fnumber);
void* SPI_palloc(size_t size);
#endif
