/* ann_multi.c — the whole node behind the reference's own call: ANN_B200_GPUS=N.
 *
 * The reference's programs call precomp()/query() from one process (ann.c:6-23); the sharded
 * path of ann_dist.c wants one rank per GPU.  With ANN_B200_GPUS=N (N <= devices present) this
 * file keeps N worker threads, one per device.  Every piece of library state is per thread
 * (__thread / thread_local: streams, arenas, staging buffers, the NCCL communicator, the query
 * cache), so a worker is exactly what a rank process is under torchrun and runs the same code:
 * tries split over the ranks, NCCL all-gather / all-to-all between the devices, row-sliced merge
 * and supercharge.  What differs from the multi-process mode:
 *   - the transforms are drawn ONCE, by the calling thread, from libc random() — the stream the
 *     reference consumes (alg.c:388-392) — and handed to the workers;
 *   - the result is ONE malloc()ed array as the API promises: every worker writes the rows it
 *     owns straight into its slice of it (no gather over NVLink, no second copy);
 *   - precomp with a save_t, query and free_save run on worker 0 alone (its device holds the
 *     index cache).
 * gpu_init() starts the workers, gpu_cleanup() stops them.                                     */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>

#include "ann.h"
#include "ann_host.h"
#include "annb200_dist.h"

typedef enum { J_NONE, J_PRECOMP, J_QUERY, J_FORGET, J_CLEANUP } job_kind;

typedef struct {
  job_kind kind;
  size_t n, k, d, rb, lb, ra, la, ycnt;
  const ftype *points, *y;
  int tries;
  save_t *save;
  const save_t *csave;
  ftype **dists_o;
  annh_call_ctx ctx;
  size_t *ret;
} job;

typedef struct {
  pthread_t th;
  int rank, pending, done, ready;
  pthread_mutex_t mu;
  pthread_cond_t cv;
  job j;
} worker;

#define MAX_WORKERS 16
static worker W[MAX_WORKERS];
static int M_n = -1, M_started;
static char M_id[128];
static __thread int in_worker;

int annh_multi_in_worker(void) { return in_worker; }

int annh_multi_gpus(void) {
  if (M_n < 0) {
    const char *e = getenv("ANN_B200_GPUS");
    int want = e && *e ? atoi(e) : 1, have = 0;
    if (want > 1) {
      if (cudaGetDeviceCount(&have) != cudaSuccess) have = 0;
      if (want > have) want = have;
      if (want > MAX_WORKERS) want = MAX_WORKERS;
    }
    M_n = want > 1 ? want : 1;
  }
  return M_n;
}

static void run(worker *w) {
  job *j = &w->j;
  switch (j->kind) {
    case J_PRECOMP:
      j->ret = annh_precomp_impl(j->n, j->k, j->d, j->points, j->tries, j->rb, j->lb, j->ra, j->la, j->save,
                                 j->dists_o, &j->ctx);
      break;
    case J_QUERY: j->ret = annh_query_impl(j->csave, j->points, j->ycnt, j->y, j->dists_o); break;
    case J_FORGET: annh_forget_save_impl(j->csave); break;
    default: break;
  }
}

static void *worker_main(void *p) {
  worker *w = p;
  in_worker = 1;
  annh_gpu_init_on(w->rank);
  annb200_dist_init(w->rank, M_n, M_id);             /* collective over the workers */
  pthread_mutex_lock(&w->mu);
  w->ready = 1;
  pthread_cond_broadcast(&w->cv);
  for (;;) {
    while (!w->pending) pthread_cond_wait(&w->cv, &w->mu);
    w->pending = 0;
    pthread_mutex_unlock(&w->mu);
    const int last = w->j.kind == J_CLEANUP;
    if (last) annh_gpu_cleanup_impl();               /* hooks: query cache, NCCL communicator */
    else run(w);
    pthread_mutex_lock(&w->mu);
    w->done = 1;
    pthread_cond_broadcast(&w->cv);
    if (last) break;
  }
  pthread_mutex_unlock(&w->mu);
  return NULL;
}

static void post(worker *w) {
  pthread_mutex_lock(&w->mu);
  w->done = 0;
  w->pending = 1;
  pthread_cond_broadcast(&w->cv);
  pthread_mutex_unlock(&w->mu);
}
static void wait_done(worker *w) {
  pthread_mutex_lock(&w->mu);
  while (!w->done) pthread_cond_wait(&w->cv, &w->mu);
  pthread_mutex_unlock(&w->mu);
}

void annh_multi_start(void) {
  if (M_started) return;
  annb200_dist_unique_id(M_id);
  for (int r = 0; r < M_n; r++) {
    worker *w = &W[r];
    memset(w, 0, sizeof *w);
    w->rank = r;
    pthread_mutex_init(&w->mu, NULL);
    pthread_cond_init(&w->cv, NULL);
    if (pthread_create(&w->th, NULL, worker_main, w) != 0) annh_fatal("%s", "pthread_create failed");
  }
  for (int r = 0; r < M_n; r++) {
    worker *w = &W[r];
    pthread_mutex_lock(&w->mu);
    while (!w->ready) pthread_cond_wait(&w->cv, &w->mu);
    pthread_mutex_unlock(&w->mu);
  }
  M_started = 1;
}

void annh_multi_stop(void) {
  if (!M_started) return;
  for (int r = 0; r < M_n; r++) { W[r].j.kind = J_CLEANUP; post(&W[r]); }
  for (int r = 0; r < M_n; r++) {
    pthread_join(W[r].th, NULL);
    pthread_mutex_destroy(&W[r].mu);
    pthread_cond_destroy(&W[r].cv);
  }
  M_started = 0;
}

size_t *annh_multi_precomp(size_t n, size_t k, size_t d, const ftype *points, int tries, size_t rots_before,
                           size_t rot_len_before, size_t rots_after, size_t rot_len_after, save_t *save,
                           ftype **dists_o) {
  annh_multi_start();
  job base;
  memset(&base, 0, sizeof base);
  base.kind = J_PRECOMP;
  base.n = n; base.k = k; base.d = d; base.points = points; base.tries = tries;
  base.rb = rots_before; base.lb = rot_len_before; base.ra = rots_after; base.la = rot_len_after;
  if (save) {                                        /* the index lives on device 0: one device does it all */
    W[0].j = base;
    W[0].j.save = save;
    W[0].j.dists_o = dists_o;
    W[0].j.ctx.force_single = 1;
    post(&W[0]);
    wait_done(&W[0]);
    return W[0].j.ret;
  }
  if (n < 2 || k < 1 || k >= n || tries < 1) annh_fatal("%s", "n >= 2, 1 <= k < n, tries >= 1 required");
  void *tf = annh_draw_transforms(n, k, d, tries, rots_before, rot_len_before, rots_after, rot_len_after);
  size_t *ids = malloc(sizeof(size_t) * n * k);
  ftype *dist = dists_o ? malloc(sizeof(ftype) * n * k) : NULL;
  if (!ids || (dists_o && !dist)) annh_fatal("%s", "out of host memory for the result arrays");
  ftype *slice_dist[MAX_WORKERS];
  for (int r = 0; r < M_n; r++) {
    size_t lo, hi;
    annb200_dist_slice(n, r, M_n, &lo, &hi);
    W[r].j = base;
    W[r].j.dists_o = dists_o ? &slice_dist[r] : NULL;
    W[r].j.ctx.transforms = tf;
    W[r].j.ctx.out.ids = ids + lo * k;
    W[r].j.ctx.out.dist = dist ? dist + lo * k : NULL;
    post(&W[r]);
  }
  for (int r = 0; r < M_n; r++) wait_done(&W[r]);
  annh_free_transforms(tf, tries);
  if (dists_o) *dists_o = dist;
  return ids;
}

size_t *annh_multi_query(const save_t *save, const ftype *points, size_t ycnt, const ftype *y, ftype **dists_o) {
  annh_multi_start();
  memset(&W[0].j, 0, sizeof W[0].j);
  W[0].j.kind = J_QUERY;
  W[0].j.csave = save; W[0].j.points = points; W[0].j.ycnt = ycnt; W[0].j.y = y; W[0].j.dists_o = dists_o;
  post(&W[0]);
  wait_done(&W[0]);
  return W[0].j.ret;
}

void annh_multi_forget(const save_t *save) {
  if (!M_started) return;
  memset(&W[0].j, 0, sizeof W[0].j);
  W[0].j.kind = J_FORGET;
  W[0].j.csave = save;
  post(&W[0]);
  wait_done(&W[0]);
}
