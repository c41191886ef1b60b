/* Link-time stand-ins for the Postgres symbols the reference's index_utils.c
 * references but the oracle never reaches (SPI, array deconstruction).
 * TEST INFRASTRUCTURE ONLY — reaching one of these aborts. */
#include "postgres.h"
#include "utils/array.h"
#include "executor/spi.h"
#define UNREACHABLE(name) do { fprintf(stderr, "pg stub reached: %s\n", name); abort(); } while (0)
uint64_t SPI_processed = 0;
SPITupleTable* SPI_tuptable = NULL;
int SPI_connect(void) { UNREACHABLE("SPI_connect"); }
int SPI_finish(void) { UNREACHABLE("SPI_finish"); }
int SPI_exec(const char* src, long tcount) { (void)src; (void)tcount; UNREACHABLE("SPI_exec"); }
Datum SPI_getbinval(HeapTuple t, TupleDesc d, int f, bool* n) { (void)t; (void)d; (void)f; (void)n; UNREACHABLE("SPI_getbinval"); }
char* SPI_getvalue(HeapTuple t, TupleDesc d, int f) { (void)t; (void)d; (void)f; UNREACHABLE("SPI_getvalue"); }
void* SPI_palloc(size_t size) { return malloc(size); }
void get_typlenbyvalalign(Oid typid, int16* typlen, bool* typbyval, char* typalign) {
  (void)typid; (void)typlen; (void)typbyval; (void)typalign; UNREACHABLE("get_typlenbyvalalign");
}
void deconstruct_array(ArrayType* a, Oid e, int l, bool b, char al, Datum** ep, bool** np, int* n) {
  (void)a; (void)e; (void)l; (void)b; (void)al; (void)ep; (void)np; (void)n; UNREACHABLE("deconstruct_array");
}
