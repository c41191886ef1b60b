/* freddy_sidecar.h — cross-backend batching for single-query SQL calls.
 *
 * The reference answers `SELECT * FROM ivfadc_search(bytea, k)` inside the calling Postgres backend
 * (freddy--0.0.1.sql:370-372 -> freddy.c:414): one process, one query, one call.  A GPU engine per backend gives every
 * backend its own CUDA context; contexts time-slice the device, so P backends are SLOWER than one
 * (profiles/r2_concurrent_callers.json).  The sidecar keeps ONE engine in ONE process and lets the backends hand it
 * their queries through a shared-memory segment: whatever arrived while the previous launch was running forms the
 * next batch (group commit), so P concurrent single-query callers get batch throughput.
 *
 *   backend (shim, no CUDA):   fbsc_client_open(name) once, fbsc_client_search(...) per SQL call
 *   sidecar process:           fbsc_server_create(name, ...), fbsc_server_run(server, batch_fn, ctx, ...)
 *                              (fb_sidecar_serve in freddy_b200.h runs it over fb_ivfadc_search)
 *
 * Plain C, no CUDA and no Postgres headers: the client half links into the extension, the server half into the
 * sidecar.  Synchronisation is futex words inside the segment (Linux). */
#ifndef FREDDY_SIDECAR_H
#define FREDDY_SIDECAR_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct fbsc_server fbsc_server;
typedef struct fbsc_client fbsc_client;

enum { FBSC_OK = 0, FBSC_ERR_ARG = -1, FBSC_ERR_SYS = -2, FBSC_ERR_GONE = -3, FBSC_ERR_BUSY = -4 };

/* One batch for the engine: nq queries of dimension d, all with the same k and w; results row-major [nq][k].
 * Returns 0 or an engine error code, which every caller of the batch receives. */
typedef int (*fbsc_batch_fn)(void* ctx, const float* queries, int nq, int k, int w, int32_t* out_ids, float* out_dists);

/* ---- sidecar side ---- */
/* Creates the POSIX shared-memory segment `name` ("/freddy" ...) with `slots` request slots for vectors of dimension d
 * and up to max_k results each.  An existing segment of that name is replaced. */
int fbsc_server_create(const char* name, int d, int max_k, int slots, fbsc_server** out);
/* Batch buffers of the server (so the caller can page-lock them): queries [max_batch][d], ids / dists [max_batch][max_k]. */
int fbsc_server_buffers(fbsc_server* s, int max_batch, float** queries, int32_t** ids, float** dists);
/* Serves until fbsc_server_stop.  A batch is launched as soon as one request is pending; with linger_us > 0 the
 * server waits up to that long for more requests before it launches a batch smaller than max_batch. */
int fbsc_server_run(fbsc_server* s, fbsc_batch_fn fn, void* ctx, int max_batch, int linger_us);
void fbsc_server_stop(fbsc_server* s);                     /* from another thread or a signal handler */
void fbsc_server_counters(fbsc_server* s, int64_t* batches, int64_t* queries, int64_t* largest_batch);
void fbsc_server_destroy(fbsc_server* s);                  /* unlinks the segment */

/* ---- backend side ---- */
int fbsc_client_open(const char* name, fbsc_client** out);
int fbsc_client_dim(const fbsc_client* c);
/* Blocks until the sidecar answered; FBSC_ERR_GONE if the sidecar died, FBSC_ERR_BUSY if no slot came free within
 * timeout_ms (<= 0: wait for ever), else the engine's return code for the batch. */
int fbsc_client_search(fbsc_client* c, const float* query, int k, int w, int32_t* out_ids, float* out_dists, int timeout_ms);
int fbsc_client_request_stop(fbsc_client* c);               /* asks the sidecar to leave fbsc_server_run */
void fbsc_client_close(fbsc_client* c);

#ifdef __cplusplus
}
#endif
#endif
