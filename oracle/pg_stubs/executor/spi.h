/* stub of executor/spi.h, backed by the in-memory tables of oracle/pg_emul.c */
#ifndef FB_STUB_SPI_H
#define FB_STUB_SPI_H
#include "funcapi.h"
typedef struct SPITupleTable { TupleDesc tupdesc; HeapTuple* vals; } SPITupleTable;
extern uint64 SPI_processed;
extern SPITupleTable* SPI_tuptable;
int SPI_connect(void);
int SPI_finish(void);
int SPI_exec(const char* src, long tcount);
int SPI_execute(const char* src, bool read_only, long tcount);
Datum SPI_getbinval(HeapTuple tuple, TupleDesc tupdesc, int fnumber, bool* isnull);
char* SPI_getvalue(HeapTuple tuple, TupleDesc tupdesc, int fnumber);
void* SPI_palloc(Size size);
#endif
