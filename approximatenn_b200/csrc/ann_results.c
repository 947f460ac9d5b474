/* ann_results.c — egress of result rows: device (u32 ids, ftype distances) -> the
 * malloc()ed size_t / ftype arrays the API returns (ann.h).
 *
 * A fresh malloc of this size is an mmap of untouched pages; copying straight into it
 * page-faults every 4 KB and costs more than the whole GPU computation, and page-locking
 * it (cudaHostRegister) is no cheaper.  So:
 *   - the arrays are allocated before the GPU starts and a few host threads touch their
 *     pages while the GPU works;
 *   - result rows leave the device in chunks as soon as the last stage has produced them:
 *     a copy stream moves each chunk (ids still 32-bit) into a cached pinned staging
 *     buffer, and the same host threads widen the ids to size_t and copy the distances
 *     into the caller-visible arrays while later chunks are still being computed.
 */
#include <pthread.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <cuda_runtime_api.h>

#include "ann_host.h"
#include "annb200.h"

#define CK(call)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      fprintf(stderr, "approximatenn_b200: %s failed at %s:%d: %s\n", #call, __FILE__,  \
              __LINE__, cudaGetErrorString(e_));                                        \
      exit(1);                                                                          \
    }                                                                                   \
  } while (0)

#define EGRESS_MAX_CHUNKS 64
#define EGRESS_MAX_THREADS 16

typedef struct { size_t r0, r1; cudaEvent_t done; } egress_chunk;
typedef struct { struct annh_egress *e; int idx; } egress_arg;

struct annh_egress {
  size_t rows, k;
  size_t *ids;            /* [rows][k] result, caller-owned after end()        */
  size_t *ids2;           /* optional second copy (save->graph), same contents  */
  ftype *dist;            /* [rows][k] or NULL                                 */
  annb_u32 *stage_ids;    /* pinned staging (cached in S)                      */
  ftype *stage_dist;
  int device, nthreads;
  pthread_t th[EGRESS_MAX_THREADS];
  egress_arg arg[EGRESS_MAX_THREADS];
  pthread_mutex_t mu;
  pthread_cond_t cv;
  pthread_barrier_t touched;   /* all pages faulted in before any result cell is written */
  int submitted, closed;
  egress_chunk chunk[EGRESS_MAX_CHUNKS];
};

static __thread struct {
  void *stage;
  size_t stage_bytes;
  cudaStream_t copy;
  cudaEvent_t produced;
  int ready;
} S;

static void tables_release(void);
void annh_egress_release(void) {
  tables_release();
  if (S.stage) CK(cudaFreeHost(S.stage));
  S.stage = NULL;
  S.stage_bytes = 0;
  if (S.ready) {
    CK(cudaEventDestroy(S.produced));
    CK(cudaStreamDestroy(S.copy));
    S.ready = 0;
  }
}

/* Makes the pages of [p, p+bytes) of a fresh malloc() resident by touching them.  (Asking the
 * kernel to populate the range instead — madvise(MADV_POPULATE_WRITE), with or without huge
 * pages — was measured 15-25 ms SLOWER per call at cfg3 on the GPU boxes: the call holds the
 * address-space lock the CUDA driver and the other threads need.)                             */
static void make_resident(void *p, size_t bytes) {
  if (!bytes) return;
  volatile char *c = (volatile char *)p;
  for (size_t o = 0; o < bytes; o += 4096) c[o] = 0;
  c[bytes - 1] = 0;
}

static void *egress_worker(void *p) {
  struct annh_egress *e = ((egress_arg *)p)->e;
  const int me = ((egress_arg *)p)->idx;
  cudaSetDevice(e->device);
  /* 1. fault in this thread's share of the result pages while the GPU computes.  Shares are
   *    whole rows; the barrier below keeps every zero written here ahead of every result
   *    cell written in step 2 (a chunk may land while a slow thread is still touching).     */
  {
    size_t lo = (e->rows * (size_t)me / e->nthreads) * e->k, hi = (e->rows * (size_t)(me + 1) / e->nthreads) * e->k;
    make_resident(e->ids + lo, (hi - lo) * sizeof(size_t));
    if (e->ids2) make_resident(e->ids2 + lo, (hi - lo) * sizeof(size_t));
    if (e->dist) make_resident(e->dist + lo, (hi - lo) * sizeof(ftype));
  }
  pthread_barrier_wait(&e->touched);
  /* 2. every chunk, as it lands: this thread widens/copies its 1/nthreads share of the rows,
   *    so the tail after the last chunk is short                                            */
  for (int c = 0;; c++) {
    pthread_mutex_lock(&e->mu);
    while (c >= e->submitted && !e->closed) pthread_cond_wait(&e->cv, &e->mu);
    int have = c < e->submitted;
    pthread_mutex_unlock(&e->mu);
    if (!have) break;
    egress_chunk *ch = &e->chunk[c];
    if (cudaEventSynchronize(ch->done) != cudaSuccess) {
      fprintf(stderr, "approximatenn_b200: result copy failed: %s\n", cudaGetErrorString(cudaGetLastError()));
      exit(1);
    }
    size_t rows = ch->r1 - ch->r0;
    size_t lo = (ch->r0 + rows * (size_t)me / e->nthreads) * e->k;
    size_t hi = (ch->r0 + rows * (size_t)(me + 1) / e->nthreads) * e->k;
    const annb_u32 *src = e->stage_ids;
    size_t *dst = e->ids;
    if (e->ids2) {
      size_t *dst2 = e->ids2;
      for (size_t i = lo; i < hi; i++) dst[i] = dst2[i] = src[i];
    } else {
      for (size_t i = lo; i < hi; i++) dst[i] = src[i];
    }
    if (e->dist) memcpy(e->dist + lo, e->stage_dist + lo, (hi - lo) * sizeof(ftype));
  }
  return NULL;
}

annh_egress *annh_egress_begin(size_t rows, size_t k, int want_dist, int want_second_ids, int device,
                               const annh_egress_target *into) {
  annh_egress *e = calloc(1, sizeof *e);
  e->rows = rows; e->k = k; e->device = device;
  if (into && into->ids) {
    e->ids = into->ids;
    e->dist = want_dist ? into->dist : NULL;
    e->ids2 = want_second_ids ? into->ids2 : NULL;
  } else {
    const size_t cells = rows * k > 0 ? rows * k : 1;
    e->ids = malloc(sizeof(size_t) * cells);
    e->dist = want_dist ? malloc(sizeof(ftype) * cells) : NULL;
    e->ids2 = want_second_ids ? malloc(sizeof(size_t) * cells) : NULL;
  }
  if (!e->ids || (want_dist && !e->dist) || (want_second_ids && !e->ids2)) annh_fatal("%s", "out of host memory for the result arrays");
  if (!S.ready) {
    CK(cudaStreamCreateWithFlags(&S.copy, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&S.produced, cudaEventDisableTiming));
    S.ready = 1;
  }
  size_t need = rows * k * (sizeof(annb_u32) + sizeof(ftype)) + 256;
  if (need > S.stage_bytes) {
    if (S.stage) CK(cudaFreeHost(S.stage));
    CK(cudaMallocHost(&S.stage, need));
    S.stage_bytes = need;
  }
  e->stage_dist = (ftype *)S.stage;                                   /* 8-byte aligned first */
  e->stage_ids = (annb_u32 *)((char *)S.stage + ((rows * k * sizeof(ftype) + 255) & ~(size_t)255));
  pthread_mutex_init(&e->mu, NULL);
  pthread_cond_init(&e->cv, NULL);
  long cores = sysconf(_SC_NPROCESSORS_ONLN);
  int nt = cores >= 12 ? 6 : 4;
  const char *env = getenv("ANN_B200_HOST_THREADS");
  if (env && *env) nt = atoi(env);
  else if (annh_dist_world() > 1) {               /* one process per GPU shares the box's cores */
    int share = (int)(cores / annh_dist_world()) - 1;
    if (nt > share) nt = share < 1 ? 1 : share;
  }
  if (cores > 1 && nt > cores - 1) nt = (int)cores - 1;
  if (rows * k < ((size_t)1 << 18)) nt = 1;
  if (nt < 1) nt = 1;
  if (nt > EGRESS_MAX_THREADS) nt = EGRESS_MAX_THREADS;
  e->nthreads = nt;
  pthread_barrier_init(&e->touched, NULL, (unsigned)nt);
  for (int i = 0; i < nt; i++) {
    e->arg[i].e = e;
    e->arg[i].idx = i;
    if (pthread_create(&e->th[i], NULL, egress_worker, &e->arg[i]) != 0)
      annh_fatal("%s", "pthread_create failed");
  }
  return e;
}

int annh_egress_chunks(const annh_egress *e) {
  /* chunks of at least 32k rows, so that the copy and the widening of a chunk hide behind the
   * computation of the next ones also when a rank owns a small slice (every chunk costs three
   * launches; 16k-row chunks cost a 125k-row slice 0.6 ms of device time at 8 ranks)          */
  size_t c = e->rows >> 15;
  return c >= 8 ? 8 : c >= 2 ? (int)c : 1;
}

void annh_egress_chunk(annh_egress *e, size_t r0, size_t r1, const void *dev_ids,
                       const void *dev_dist, void *producer_stream) {
  if (e->submitted >= EGRESS_MAX_CHUNKS) annh_fatal("%s", "internal: too many egress chunks");
  egress_chunk *ch = &e->chunk[e->submitted];
  ch->r0 = r0; ch->r1 = r1;
  CK(cudaEventCreateWithFlags(&ch->done, cudaEventDisableTiming | cudaEventBlockingSync));
  CK(cudaEventRecord(S.produced, (cudaStream_t)producer_stream));
  CK(cudaStreamWaitEvent(S.copy, S.produced, 0));
  size_t cells = (r1 - r0) * e->k;
  CK(cudaMemcpyAsync(e->stage_ids + r0 * e->k, dev_ids, cells * sizeof(annb_u32), cudaMemcpyDeviceToHost, S.copy));
  if (e->dist)
    CK(cudaMemcpyAsync(e->stage_dist + r0 * e->k, dev_dist, cells * sizeof(ftype), cudaMemcpyDeviceToHost, S.copy));
  CK(cudaEventRecord(ch->done, S.copy));
  pthread_mutex_lock(&e->mu);
  e->submitted++;
  pthread_cond_broadcast(&e->cv);
  pthread_mutex_unlock(&e->mu);
}

size_t *annh_egress_end(annh_egress *e, ftype **dists_o, size_t **second_ids_o) {
  pthread_mutex_lock(&e->mu);
  e->closed = 1;
  pthread_cond_broadcast(&e->cv);
  pthread_mutex_unlock(&e->mu);
  for (int i = 0; i < e->nthreads; i++) pthread_join(e->th[i], NULL);
  for (int c = 0; c < e->submitted; c++) CK(cudaEventDestroy(e->chunk[c].done));
  pthread_mutex_destroy(&e->mu);
  pthread_cond_destroy(&e->cv);
  pthread_barrier_destroy(&e->touched);
  size_t *ids = e->ids;
  if (dists_o) *dists_o = e->dist;
  if (second_ids_o) *second_ids_o = e->ids2;
  free(e);
  return ids;
}

/* Touches the pages of a freshly malloc()ed block with a few threads, so that a following
 * device->host copy into it does not page-fault 4 KB at a time on one thread.              */
typedef struct { volatile char *p; size_t bytes; } touch_job;
static void *touch_main(void *a) {
  touch_job *j = a;
  for (size_t o = 0; o < j->bytes; o += 4096) j->p[o] = 0;
  if (j->bytes) j->p[j->bytes - 1] = 0;
  return NULL;
}
void annh_prefault(void *ptr, size_t bytes) {
  enum { NT = 4 };
  if (bytes < ((size_t)4 << 20)) return;
  pthread_t th[NT];
  touch_job job[NT];
  size_t per = ((bytes / NT) + 4095) & ~(size_t)4095;
  int started = 0;
  for (int i = 0; i < NT; i++) {
    size_t lo = per * (size_t)i;
    if (lo >= bytes) break;
    job[i].p = (volatile char *)ptr + lo;
    job[i].bytes = bytes - lo < per ? bytes - lo : per;
    if (pthread_create(&th[i], NULL, touch_main, &job[i]) != 0) break;
    started++;
  }
  for (int i = 0; i < started; i++) pthread_join(th[i], NULL);
}

/* ---- bucket tables of save_t (which_par[t], alg.c:270-271) ------------------------------------
 * The device exports a table with 32-bit cells; a copy stream moves it into pinned staging as
 * soon as it exists and a few host threads widen it into the malloc()ed size_t array the API
 * hands out, while the GPU works on the following tries.  Nothing here blocks the caller
 * before annh_tables_end().                                                                  */
#define TABLE_THREADS 3
struct annh_tables {
  int tries, device, nthreads;
  size_t *cells;               /* [tries] */
  size_t *at;                  /* [tries] offset (cells) into the staging buffer */
  size_t **host;               /* [tries] destination arrays (owned by the caller's save_t) */
  const annb_u32 *stage;       /* the owning thread's pinned staging (TS is per thread)     */
  cudaEvent_t *done;           /* [tries] */
  pthread_t th[TABLE_THREADS];
  struct { struct annh_tables *tb; int idx; } arg[TABLE_THREADS];
  pthread_mutex_t mu;
  pthread_cond_t cv;
  int submitted, closed;
};
static __thread struct { annb_u32 *stage; size_t bytes; } TS;

static void *tables_worker(void *p) {
  struct annh_tables *tb = *(struct annh_tables **)p;        /* first member of the argument record */
  const int me = (int)((char *)p - (char *)tb->arg) / (int)sizeof tb->arg[0];
  cudaSetDevice(tb->device);
  for (int t = 0;; t++) {
    pthread_mutex_lock(&tb->mu);
    while (t >= tb->submitted && !tb->closed) pthread_cond_wait(&tb->cv, &tb->mu);
    int have = t < tb->submitted;
    pthread_mutex_unlock(&tb->mu);
    if (!have) break;
    if (cudaEventSynchronize(tb->done[t]) != cudaSuccess) {
      fprintf(stderr, "approximatenn_b200: table copy failed: %s\n", cudaGetErrorString(cudaGetLastError()));
      exit(1);
    }
    size_t lo = tb->cells[t] * (size_t)me / tb->nthreads, hi = tb->cells[t] * (size_t)(me + 1) / tb->nthreads;
    const annb_u32 *src = tb->stage + tb->at[t];
    size_t *dst = tb->host[t];
    for (size_t i = lo; i < hi; i++) dst[i] = src[i];
  }
  return NULL;
}

annh_tables *annh_tables_begin(int tries, const size_t *cells, size_t **host_tables, int device) {
  annh_tables *tb = calloc(1, sizeof *tb);
  tb->tries = tries; tb->device = device;
  tb->cells = malloc(sizeof(size_t) * tries);
  tb->at = malloc(sizeof(size_t) * tries);
  tb->host = malloc(sizeof(size_t *) * tries);
  tb->done = malloc(sizeof(cudaEvent_t) * tries);
  size_t total = 0;
  for (int t = 0; t < tries; t++) {
    tb->cells[t] = cells[t];
    tb->at[t] = total;
    total += (cells[t] + 63) & ~(size_t)63;
    tb->host[t] = host_tables[t];
    CK(cudaEventCreateWithFlags(&tb->done[t], cudaEventDisableTiming | cudaEventBlockingSync));
  }
  if (total * 4 + 256 > TS.bytes) {
    if (TS.stage) CK(cudaFreeHost(TS.stage));
    CK(cudaMallocHost((void **)&TS.stage, total * 4 + 256));
    TS.bytes = total * 4 + 256;
  }
  tb->stage = TS.stage;
  if (!S.ready) {
    CK(cudaStreamCreateWithFlags(&S.copy, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&S.produced, cudaEventDisableTiming));
    S.ready = 1;
  }
  pthread_mutex_init(&tb->mu, NULL);
  pthread_cond_init(&tb->cv, NULL);
  long cores = sysconf(_SC_NPROCESSORS_ONLN);
  tb->nthreads = cores >= 8 ? TABLE_THREADS : 1;
  for (int i = 0; i < tb->nthreads; i++) {
    tb->arg[i].tb = tb;
    tb->arg[i].idx = i;
    if (pthread_create(&tb->th[i], NULL, tables_worker, &tb->arg[i]) != 0) annh_fatal("%s", "pthread_create failed");
  }
  return tb;
}

/* tables must be submitted in try order 0, 1, ... */
void annh_tables_submit(annh_tables *tb, int t, const void *dev_table32, void *producer_stream) {
  if (t != tb->submitted) annh_fatal("%s", "internal: bucket tables submitted out of order");
  cudaEvent_t produced;
  CK(cudaEventCreateWithFlags(&produced, cudaEventDisableTiming));
  CK(cudaEventRecord(produced, (cudaStream_t)producer_stream));
  CK(cudaStreamWaitEvent(S.copy, produced, 0));
  CK(cudaEventDestroy(produced));                  /* released once the wait has consumed it */
  if (tb->cells[t])
    CK(cudaMemcpyAsync(TS.stage + tb->at[t], dev_table32, tb->cells[t] * 4, cudaMemcpyDeviceToHost, S.copy));
  CK(cudaEventRecord(tb->done[t], S.copy));
  pthread_mutex_lock(&tb->mu);
  tb->submitted++;
  pthread_cond_broadcast(&tb->cv);
  pthread_mutex_unlock(&tb->mu);
}

void annh_tables_end(annh_tables *tb) {
  pthread_mutex_lock(&tb->mu);
  tb->closed = 1;
  pthread_cond_broadcast(&tb->cv);
  pthread_mutex_unlock(&tb->mu);
  for (int i = 0; i < tb->nthreads; i++) pthread_join(tb->th[i], NULL);
  for (int t = 0; t < tb->tries; t++) CK(cudaEventDestroy(tb->done[t]));
  pthread_mutex_destroy(&tb->mu);
  pthread_cond_destroy(&tb->cv);
  free(tb->cells); free(tb->at); free(tb->host); free(tb->done);
  free(tb);
}

static void tables_release(void) {
  if (TS.stage) CK(cudaFreeHost(TS.stage));
  TS.stage = NULL;
  TS.bytes = 0;
}
