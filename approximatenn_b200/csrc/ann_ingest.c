/* ann_ingest.c — host -> device upload of the caller's point array.
 *
 * The API takes plain host pointers (ann.h); the reference's callers malloc() them, i.e. the
 * memory is pageable and a single cudaMemcpy moves it at the speed of ONE thread copying into
 * the driver's bounce buffer (~11 GB/s measured: 23 ms for the 256 MB of cfg3, more than half of
 * the whole GPU computation).  Here several host threads copy chunks into a ring of pinned
 * staging slots and each chunk goes out by DMA as soon as it is staged, so the copy runs at the
 * memory bandwidth of several cores and overlaps with the DMA.  Pinned (or registered) input
 * is recognised and sent with one cudaMemcpyAsync as before.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <cuda_runtime_api.h>

#include "ann_host.h"

#define CK(call)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      fprintf(stderr, "approximatenn_b200: %s failed at %s:%d: %s\n", #call, __FILE__,  \
              __LINE__, cudaGetErrorString(e_));                                        \
      exit(1);                                                                          \
    }                                                                                   \
  } while (0)

#define INGEST_MAX_THREADS 16
#define INGEST_PER_THREAD 2                  /* staging slots per thread                   */
#define INGEST_SLOTS (INGEST_MAX_THREADS * INGEST_PER_THREAD)
#define INGEST_CHUNK ((size_t)4 << 20)       /* bytes per slot                            */

struct ingest_state {
  char *ring;                                /* INGEST_SLOTS * INGEST_CHUNK pinned bytes   */
  cudaStream_t stream[INGEST_MAX_THREADS];
  cudaEvent_t slot_free[INGEST_SLOTS];
  cudaEvent_t done[INGEST_MAX_THREADS];
  int ready;
};
static __thread struct ingest_state I;       /* per calling thread = per device (ann_multi.c) */

void annh_ingest_release(void) {
  if (!I.ready) return;
  for (int i = 0; i < INGEST_MAX_THREADS; i++) { cudaStreamDestroy(I.stream[i]); cudaEventDestroy(I.done[i]); }
  for (int i = 0; i < INGEST_SLOTS; i++) cudaEventDestroy(I.slot_free[i]);
  cudaFreeHost(I.ring);
  memset(&I, 0, sizeof I);
}

typedef struct { const char *src; char *dst; size_t bytes; int me, device, threads; struct ingest_state *st; } ingest_job;

/* One memcpy thread moves ~6-8 GB/s into the pinned ring and PCIe 5 takes ~54 GB/s, so the
 * copy needs about eight of them; the count follows the cores the box has (the egress
 * threads are only touching pages at this point of the call).                               */
static int ingest_threads(void) {
  int nt = 8;
  const char *env = getenv("ANN_B200_INGEST_THREADS");
  if (env && *env) nt = atoi(env);
  long cores = sysconf(_SC_NPROCESSORS_ONLN);
  int world = annh_dist_world();                  /* one process per GPU shares the box's cores */
  if (world > 1 && !(env && *env)) { int share = (int)(cores / world) - 2; if (nt > share) nt = share < 2 ? 2 : share; }
  if (cores > 2 && nt > cores - 2) nt = (int)cores - 2;
  if (nt < 1) nt = 1;
  if (nt > INGEST_MAX_THREADS) nt = INGEST_MAX_THREADS;
  return nt;
}

static void *ingest_worker(void *p) {
  ingest_job *j = p;
  cudaSetDevice(j->device);
  size_t chunks = (j->bytes + INGEST_CHUNK - 1) / INGEST_CHUNK;
  int use = 0;
  for (size_t c = j->me; c < chunks; c += j->threads, use++) {
    int slot = j->me * INGEST_PER_THREAD + (use % INGEST_PER_THREAD);
    size_t off = c * INGEST_CHUNK, len = j->bytes - off < INGEST_CHUNK ? j->bytes - off : INGEST_CHUNK;
    char *stage = j->st->ring + (size_t)slot * INGEST_CHUNK;
    if (cudaEventSynchronize(j->st->slot_free[slot]) != cudaSuccess) exit(1);   /* DMA of its last use done */
    memcpy(stage, j->src + off, len);
    if (cudaMemcpyAsync(j->dst + off, stage, len, cudaMemcpyHostToDevice, j->st->stream[j->me]) != cudaSuccess ||
        cudaEventRecord(j->st->slot_free[slot], j->st->stream[j->me]) != cudaSuccess) {
      fprintf(stderr, "approximatenn_b200: staged upload failed: %s\n", cudaGetErrorString(cudaGetLastError()));
      exit(1);
    }
  }
  if (cudaEventRecord(j->st->done[j->me], j->st->stream[j->me]) != cudaSuccess) exit(1);
  return NULL;
}

/* Enqueues the upload of `bytes` from host `src` to device `dst`; work later enqueued on
 * `consumer` waits for it.  Returns after the host side has been handed over.              */
void annh_ingest(void *dst, const void *src, size_t bytes, void *consumer, int device) {
  cudaStream_t st = (cudaStream_t)consumer;
  if (bytes == 0) return;
  struct cudaPointerAttributes attr;
  int pageable = 1;
  if (cudaPointerGetAttributes(&attr, src) == cudaSuccess) pageable = attr.type == cudaMemoryTypeUnregistered;
  else (void)cudaGetLastError();
  const char *env = getenv("ANN_B200_STAGED_UPLOAD");
  if (env && *env == '0') pageable = 0;
  long cores = sysconf(_SC_NPROCESSORS_ONLN);
  if (!pageable || bytes < ((size_t)8 << 20) || cores < 4) {
    CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, st));
    return;
  }
  if (!I.ready) {
    CK(cudaMallocHost((void **)&I.ring, (size_t)INGEST_SLOTS * INGEST_CHUNK));
    for (int i = 0; i < INGEST_MAX_THREADS; i++) {
      CK(cudaStreamCreateWithFlags(&I.stream[i], cudaStreamNonBlocking));
      CK(cudaEventCreateWithFlags(&I.done[i], cudaEventDisableTiming));
    }
    for (int i = 0; i < INGEST_SLOTS; i++)
      CK(cudaEventCreateWithFlags(&I.slot_free[i], cudaEventDisableTiming | cudaEventBlockingSync));
    I.ready = 1;
  }
  int nt = ingest_threads();
  {                                                  /* a thread per 4 MB chunk at most */
    size_t chunks = (bytes + INGEST_CHUNK - 1) / INGEST_CHUNK;
    if ((size_t)nt > chunks) nt = (int)chunks;
  }
  pthread_t th[INGEST_MAX_THREADS];
  ingest_job job[INGEST_MAX_THREADS];
  for (int i = 0; i < nt; i++) {
    job[i].src = src; job[i].dst = dst; job[i].bytes = bytes; job[i].me = i; job[i].device = device;
    job[i].threads = nt;
    job[i].st = &I;
    if (pthread_create(&th[i], NULL, ingest_worker, &job[i]) != 0) annh_fatal("%s", "pthread_create failed");
  }
  for (int i = 0; i < nt; i++) {
    pthread_join(th[i], NULL);
    CK(cudaStreamWaitEvent(st, I.done[i], 0));
  }
}
