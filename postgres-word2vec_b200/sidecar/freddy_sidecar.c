/* freddy_sidecar.c — shared-memory request slots between Postgres backends and the one process that owns the GPU
 * engine (include/freddy_sidecar.h).  No CUDA, no Postgres headers.
 *
 * Segment: header | slots[n].  A slot walks FREE -> FILLING (claimed by a backend) -> READY (query written) ->
 * RUNNING (picked into a batch) -> DONE (results written) -> FREE.  Callers whose request is in flight sleep on the
 * header's `done_gen` futex word, which the server bumps and wakes ONCE per batch (a wake per slot cost the server
 * ~1.5 us of system call per answered query); the server sleeps on the header's doorbell while nothing is pending. */
#define _GNU_SOURCE
#include "../../include/freddy_sidecar.h"

#include <errno.h>
#include <fcntl.h>
#include <limits.h>
#include <linux/futex.h>
#include <sched.h>
#include <signal.h>
#include <stdatomic.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <sys/syscall.h>
#include <time.h>
#include <unistd.h>

enum { ST_FREE = 0, ST_FILLING = 1, ST_READY = 2, ST_RUNNING = 3, ST_DONE = 4 };
#define FBSC_MAGIC 0x46425343u /* "FBSC" */

typedef struct {
  _Atomic uint32_t state;
  _Atomic uint32_t reserved;
  int32_t k, w, rc;
  int32_t owner_pid;
  char pad[40];
  /* float query[d]; int32 ids[max_k]; float dists[max_k] follow */
} slot_t;

typedef struct {
  uint32_t magic;
  int32_t d, max_k, n_slots;
  uint64_t slot_stride, total_bytes;
  _Atomic uint32_t doorbell;        /* bumped by every posted request: futex word of the sleeping server */
  _Atomic uint32_t server_sleeping;
  _Atomic int32_t server_pid;       /* 0 once the server is gone */
  _Atomic uint32_t stop;
  _Atomic uint32_t done_gen;        /* bumped after every batch: futex word of the sleeping callers (one wake per batch) */
  _Atomic uint32_t sleepers;        /* callers that are (about to be) asleep on done_gen */
  char pad[8];
} header_t;

struct fbsc_server {
  header_t* h;
  char name[128];
  float* bq; int32_t* bi; float* bd; int32_t* bslot;
  int batch_cap;
  int64_t batches, queries, largest;
};
struct fbsc_client { header_t* h; uint32_t hint; };

static long futex(_Atomic uint32_t* addr, int op, uint32_t val, const struct timespec* to) {
  return syscall(SYS_futex, (uint32_t*)addr, op, val, to, NULL, 0);
}
static slot_t* slot_of(header_t* h, int i) { return (slot_t*)((char*)h + sizeof(header_t) + (size_t)i * h->slot_stride); }
static float* slot_query(slot_t* s) { return (float*)((char*)s + sizeof(slot_t)); }
static int32_t* slot_ids(header_t* h, slot_t* s) { return (int32_t*)(slot_query(s) + h->d); }
static float* slot_dists(header_t* h, slot_t* s) { return (float*)(slot_ids(h, s) + h->max_k); }
static int64_t now_ns(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return (int64_t)t.tv_sec * 1000000000ll + t.tv_nsec; }
static void cpu_relax(void) {
#if defined(__x86_64__) || defined(__i386__)
  __builtin_ia32_pause();
#endif
}

/* ------------------------------------------------------------------ server */
int fbsc_server_create(const char* name, int d, int max_k, int slots, fbsc_server** out) {
  if (!name || name[0] != '/' || strlen(name) >= sizeof(((fbsc_server*)0)->name) || d < 1 || max_k < 1 || slots < 1 || !out) return FBSC_ERR_ARG;
  size_t stride = sizeof(slot_t) + (size_t)d * sizeof(float) + (size_t)max_k * (sizeof(int32_t) + sizeof(float));
  stride = (stride + 63) & ~(size_t)63;
  const size_t total = sizeof(header_t) + stride * (size_t)slots;
  shm_unlink(name);
  int fd = shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
  if (fd < 0) return FBSC_ERR_SYS;
  if (ftruncate(fd, (off_t)total) != 0) { close(fd); shm_unlink(name); return FBSC_ERR_SYS; }
  void* p = mmap(NULL, total, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) { shm_unlink(name); return FBSC_ERR_SYS; }
  memset(p, 0, total);
  fbsc_server* s = (fbsc_server*)calloc(1, sizeof *s);
  if (!s) { munmap(p, total); shm_unlink(name); return FBSC_ERR_SYS; }
  s->h = (header_t*)p;
  strcpy(s->name, name);
  s->h->d = d; s->h->max_k = max_k; s->h->n_slots = slots; s->h->slot_stride = stride; s->h->total_bytes = total;
  atomic_store(&s->h->server_pid, (int32_t)getpid());
  __atomic_store_n(&s->h->magic, FBSC_MAGIC, __ATOMIC_RELEASE);   /* last: clients check it */
  *out = s;
  return FBSC_OK;
}

int fbsc_server_buffers(fbsc_server* s, int max_batch, float** queries, int32_t** ids, float** dists) {
  if (!s || max_batch < 1) return FBSC_ERR_ARG;
  if (max_batch > s->h->n_slots) max_batch = s->h->n_slots;
  if (max_batch > s->batch_cap) {
    free(s->bq); free(s->bi); free(s->bd); free(s->bslot);
    const size_t al = 4096;
    s->bq = NULL; s->bi = NULL; s->bd = NULL;
    if (posix_memalign((void**)&s->bq, al, (((size_t)max_batch * s->h->d * sizeof(float)) + al - 1) / al * al) ||
        posix_memalign((void**)&s->bi, al, (((size_t)max_batch * s->h->max_k * sizeof(int32_t)) + al - 1) / al * al) ||
        posix_memalign((void**)&s->bd, al, (((size_t)max_batch * s->h->max_k * sizeof(float)) + al - 1) / al * al))
      return FBSC_ERR_SYS;
    s->bslot = (int32_t*)malloc((size_t)max_batch * sizeof(int32_t));
    if (!s->bslot) return FBSC_ERR_SYS;
    s->batch_cap = max_batch;
  }
  if (queries) *queries = s->bq;
  if (ids) *ids = s->bi;
  if (dists) *dists = s->bd;
  return FBSC_OK;
}

/* READY slots with the (k, w) of the first one found join the batch; the scan starts where the last one stopped so no
 * slot is starved */
static int collect(fbsc_server* s, int have, int max_batch, int* k, int* w, int* cursor) {
  header_t* h = s->h;
  for (int n = 0; n < h->n_slots && have < max_batch; n++) {
    const int i = (*cursor + n) % h->n_slots;
    slot_t* sl = slot_of(h, i);
    if (atomic_load_explicit(&sl->state, memory_order_acquire) != ST_READY) continue;
    if (have == 0) { *k = sl->k; *w = sl->w; }
    else if (sl->k != *k || sl->w != *w) continue;
    atomic_store_explicit(&sl->state, ST_RUNNING, memory_order_relaxed);
    memcpy(s->bq + (size_t)have * h->d, slot_query(sl), (size_t)h->d * sizeof(float));
    s->bslot[have++] = i;
  }
  if (have > 0) *cursor = (s->bslot[have - 1] + 1) % h->n_slots;
  return have;
}

static void reclaim_dead_owners(header_t* h) {
  for (int i = 0; i < h->n_slots; i++) {
    slot_t* sl = slot_of(h, i);
    const uint32_t st = atomic_load(&sl->state);
    if ((st == ST_FILLING || st == ST_DONE) && sl->owner_pid > 0 && kill(sl->owner_pid, 0) != 0 && errno == ESRCH)
      atomic_store(&sl->state, ST_FREE);
  }
}

int fbsc_server_run(fbsc_server* s, fbsc_batch_fn fn, void* ctx, int max_batch, int linger_us) {
  if (!s || !fn) return FBSC_ERR_ARG;
  int rc = fbsc_server_buffers(s, max_batch, NULL, NULL, NULL);
  if (rc) return rc;
  header_t* h = s->h;
  if (max_batch > s->batch_cap) max_batch = s->batch_cap;
  int cursor = 0, idle_rounds = 0;
  while (!atomic_load(&h->stop)) {
    int k = 0, w = 0;
    const uint32_t bell = atomic_load(&h->doorbell);
    int n = collect(s, 0, max_batch, &k, &w, &cursor);
    if (n == 0) {
      if (++idle_rounds < 2000) { cpu_relax(); continue; }            /* ~tens of microseconds of polling, then sleep */
      atomic_store(&h->server_sleeping, 1);
      if (atomic_load(&h->doorbell) == bell && !atomic_load(&h->stop)) {
        n = collect(s, 0, max_batch, &k, &w, &cursor);                 /* a request posted before the flag went up */
        if (n == 0) {
          const struct timespec to = {0, 2000000};                    /* 2 ms: also the cadence of the dead-owner sweep */
          futex(&h->doorbell, FUTEX_WAIT, bell, &to);
        }
      }
      atomic_store(&h->server_sleeping, 0);
      if (n == 0) { reclaim_dead_owners(h); idle_rounds = 0; continue; }
    }
    idle_rounds = 0;
    if (linger_us > 0 && n < max_batch) {
      const int64_t until = now_ns() + (int64_t)linger_us * 1000;
      while (n < max_batch && now_ns() < until) { n = collect(s, n, max_batch, &k, &w, &cursor); cpu_relax(); }
    }
    int brc = (k < 1 || k > h->max_k) ? FBSC_ERR_ARG : fn(ctx, s->bq, n, k, w, s->bi, s->bd);
    for (int j = 0; j < n; j++) {
      slot_t* sl = slot_of(h, s->bslot[j]);
      if (brc == 0) {
        memcpy(slot_ids(h, sl), s->bi + (size_t)j * k, (size_t)k * sizeof(int32_t));
        memcpy(slot_dists(h, sl), s->bd + (size_t)j * k, (size_t)k * sizeof(float));
      }
      sl->rc = brc;
      atomic_store(&sl->state, ST_DONE);
    }
    atomic_fetch_add(&h->done_gen, 1);
    if (atomic_load(&h->sleepers)) futex(&h->done_gen, FUTEX_WAKE, INT_MAX, NULL);
    s->batches++; s->queries += n;
    if (n > s->largest) s->largest = n;
  }
  return FBSC_OK;
}

void fbsc_server_stop(fbsc_server* s) {
  if (!s) return;
  atomic_store(&s->h->stop, 1);
  atomic_fetch_add(&s->h->doorbell, 1);
  futex(&s->h->doorbell, FUTEX_WAKE, 1, NULL);
}

void fbsc_server_counters(fbsc_server* s, int64_t* batches, int64_t* queries, int64_t* largest_batch) {
  if (batches) *batches = s->batches;
  if (queries) *queries = s->queries;
  if (largest_batch) *largest_batch = s->largest;
}

void fbsc_server_destroy(fbsc_server* s) {
  if (!s) return;
  header_t* h = s->h;
  atomic_store(&h->server_pid, 0);
  atomic_fetch_add(&h->done_gen, 1);                             /* release whoever still waits */
  futex(&h->done_gen, FUTEX_WAKE, INT_MAX, NULL);
  shm_unlink(s->name);
  munmap(h, h->total_bytes);
  free(s->bq); free(s->bi); free(s->bd); free(s->bslot);
  free(s);
}

/* ------------------------------------------------------------------ client */
int fbsc_client_open(const char* name, fbsc_client** out) {
  if (!name || !out) return FBSC_ERR_ARG;
  int fd = shm_open(name, O_RDWR, 0600);
  if (fd < 0) return FBSC_ERR_GONE;
  struct stat st;
  if (fstat(fd, &st) != 0 || (size_t)st.st_size < sizeof(header_t)) { close(fd); return FBSC_ERR_GONE; }
  void* p = mmap(NULL, (size_t)st.st_size, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
  close(fd);
  if (p == MAP_FAILED) return FBSC_ERR_SYS;
  header_t* h = (header_t*)p;
  if (__atomic_load_n(&h->magic, __ATOMIC_ACQUIRE) != FBSC_MAGIC || h->total_bytes != (uint64_t)st.st_size) { munmap(p, (size_t)st.st_size); return FBSC_ERR_GONE; }
  fbsc_client* c = (fbsc_client*)calloc(1, sizeof *c);
  if (!c) { munmap(p, (size_t)st.st_size); return FBSC_ERR_SYS; }
  c->h = h;
  c->hint = (uint32_t)getpid() * 2654435761u;
  *out = c;
  return FBSC_OK;
}

int fbsc_client_dim(const fbsc_client* c) { return c ? c->h->d : 0; }

static int server_alive(header_t* h) {
  const int32_t pid = atomic_load(&h->server_pid);
  return pid > 0 && (kill(pid, 0) == 0 || errno != ESRCH);
}

int fbsc_client_search(fbsc_client* c, const float* query, int k, int w, int32_t* out_ids, float* out_dists, int timeout_ms) {
  if (!c || !query || !out_ids || !out_dists || k < 1 || k > c->h->max_k) return FBSC_ERR_ARG;
  header_t* h = c->h;
  const int64_t deadline = timeout_ms > 0 ? now_ns() + (int64_t)timeout_ms * 1000000 : INT64_MAX;
  slot_t* sl = NULL;
  for (uint32_t tries = 0; sl == NULL; tries++) {
    const int i = (int)((c->hint + tries) % (uint32_t)h->n_slots);
    slot_t* cand = slot_of(h, i);
    uint32_t expect = ST_FREE;
    if (atomic_compare_exchange_strong(&cand->state, &expect, ST_FILLING)) { sl = cand; c->hint = (uint32_t)i; break; }
    if ((tries + 1) % (uint32_t)h->n_slots == 0) {                 /* every slot taken: more callers than slots */
      if (!server_alive(h)) return FBSC_ERR_GONE;
      if (now_ns() > deadline) return FBSC_ERR_BUSY;
      sched_yield();
    }
  }
  sl->owner_pid = (int32_t)getpid();
  sl->k = k; sl->w = w; sl->rc = 0;
  memcpy(slot_query(sl), query, (size_t)h->d * sizeof(float));
  atomic_store(&sl->state, ST_READY);
  atomic_fetch_add(&h->doorbell, 1);
  if (atomic_load(&h->server_sleeping)) futex(&h->doorbell, FUTEX_WAKE, 1, NULL);
  /* in flight: poll briefly, then sleep until the server announces a finished batch */
  int spins = 0;
  for (;;) {
    if (atomic_load_explicit(&sl->state, memory_order_acquire) == ST_DONE) break;
    if (++spins < 400) { cpu_relax(); continue; }
    const uint32_t gen = atomic_load(&h->done_gen);
    atomic_fetch_add(&h->sleepers, 1);
    if (atomic_load(&sl->state) != ST_DONE) {
      const struct timespec to = {0, 50000000};                   /* 50 ms, then check that the sidecar still lives */
      futex(&h->done_gen, FUTEX_WAIT, gen, &to);
    }
    atomic_fetch_sub(&h->sleepers, 1);
    if (atomic_load(&sl->state) != ST_DONE && atomic_load(&h->done_gen) == gen && !server_alive(h)) {
      atomic_store(&sl->state, ST_FREE);
      return FBSC_ERR_GONE;
    }
  }
  const int rc = sl->rc;
  if (rc == 0) {
    memcpy(out_ids, slot_ids(h, sl), (size_t)k * sizeof(int32_t));
    memcpy(out_dists, slot_dists(h, sl), (size_t)k * sizeof(float));
  }
  atomic_store(&sl->state, ST_FREE);
  return rc;
}

int fbsc_client_request_stop(fbsc_client* c) {
  if (!c) return FBSC_ERR_ARG;
  atomic_store(&c->h->stop, 1);
  atomic_fetch_add(&c->h->doorbell, 1);
  futex(&c->h->doorbell, FUTEX_WAKE, 1, NULL);
  return FBSC_OK;
}

void fbsc_client_close(fbsc_client* c) {
  if (!c) return;
  munmap(c->h, c->h->total_bytes);
  free(c);
}
