/* ann_query.c — query_gpu (alg.c:458-519 on the device).
 *
 * The reference re-wraps points, graph, bucket tables and bases as device buffers on every
 * call (alg.c:464-508).  Here the index is uploaded once and kept on the device, keyed on
 * the save_t it came from; free_save() drops it.  Because save_t has
 * no spare field, the key is (graph pointer, points pointer, shape) plus a fingerprint of
 * sampled graph/table/point entries, so a recycled allocation is not mistaken for the old
 * index.  ANN_B200_QUERY_CACHE=0 rebuilds the device copy on every call.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>

#include "ann.h"
#include "algg.h"
#include "gpu_comp.h"
#include "annb200.h"
#include "ann_host.h"

#include <time.h>
static double q_now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
/* ANN_B200_HOSTPROF=1: phase times of query_gpu on stderr (synchronises after each phase) */
#define QP(label)                                                                       \
  do {                                                                                  \
    if (qprof) {                                                                        \
      cudaStreamSynchronize(st);                                                        \
      double t_ = q_now_ms();                                                           \
      fprintf(stderr, "[queryprof] %-22s +%8.3f ms\n", label, t_ - qlast);              \
      qlast = t_;                                                                       \
    }                                                                                   \
  } while (0)

#define CK(call)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      fprintf(stderr, "approximatenn_b200: %s failed at %s:%d: %s\n", #call, __FILE__,  \
              __LINE__, cudaGetErrorString(e_));                                        \
      exit(1);                                                                          \
    }                                                                                   \
  } while (0)

typedef struct {
  int live;
  const size_t *graph_key;
  const ftype *points_key;
  size_t n, k, d_short, d, tries;
  uint64_t fingerprint;
  ftype *d_points, *d_mean, *d_bases;
  annb_u32 *d_graph;
  void *d_points16, *d_nrm;    /* fp16 copy of the points + norms for the screened candidate rows */
  unsigned *d_scale;
  size_t cap_points16, cap_nrm, cap_scale;
  int screened;
  annb_u32 **d_tab;            /* host array of device pointers to the 32-bit tables */
  /* capacities in bytes: the buffers are grow-only and survive a dropped index, because
   * cudaMalloc/cudaFree of a few hundred MB cost more than a whole query                 */
  size_t cap_points, cap_mean, cap_bases, cap_graph, *cap_tab, cap_tries;
} device_index;

/* per host thread = per device (ann_multi.c runs one worker thread per GPU) */
static __thread device_index IDX;
static __thread int cleanup_registered;

static uint64_t mix(uint64_t h, uint64_t v) {
  h ^= v + 0x9e3779b97f4a7c15ull + (h << 6) + (h >> 2);
  return h;
}

static uint64_t fingerprint(const save_t *s, const ftype *points) {
  uint64_t h = 1469598103934665603ull;
  size_t cells = s->n * s->k, step = cells / 61 + 1;
  for (size_t i = 0; i < cells; i += step) h = mix(h, s->graph[i]);
  for (int t = 0; t < s->tries; t++) {
    h = mix(h, s->par_maxes[t]);
    size_t tc = s->par_maxes[t] << s->d_short, ts = tc / 17 + 1;
    for (size_t i = 0; i < tc; i += ts) h = mix(h, s->which_par[t][i]);
  }
  size_t pc = s->n * s->d_long, ps = pc / 61 + 1;
  for (size_t i = 0; i < pc; i += ps) {
    uint64_t b = 0;
    memcpy(&b, points + i, sizeof(ftype));
    h = mix(h, b);
  }
  return h;
}

static void reserve(void **ptr, size_t *cap, size_t bytes) {
  if (bytes <= *cap && *ptr) return;
  if (*ptr) {
    CK(cudaStreamSynchronize((cudaStream_t)annh_stream()));
    CK(cudaFree(*ptr));
  }
  CK(cudaMalloc(ptr, bytes ? bytes : 1));
  *cap = bytes ? bytes : 1;
}

/* forget the current index; its device buffers stay allocated for the next one */
static void drop_index(void) {
  IDX.live = 0;
  IDX.graph_key = NULL;
  IDX.points_key = NULL;
}

/* gpu_cleanup(): really give the memory back */
static void release_index(void) {
  CK(cudaStreamSynchronize((cudaStream_t)annh_stream()));
  if (IDX.d_points) CK(cudaFree(IDX.d_points));
  if (IDX.d_mean) CK(cudaFree(IDX.d_mean));
  if (IDX.d_bases) CK(cudaFree(IDX.d_bases));
  if (IDX.d_graph) CK(cudaFree(IDX.d_graph));
  if (IDX.d_points16) CK(cudaFree(IDX.d_points16));
  if (IDX.d_nrm) CK(cudaFree(IDX.d_nrm));
  if (IDX.d_scale) CK(cudaFree(IDX.d_scale));
  for (size_t t = 0; t < IDX.cap_tries; t++)
    if (IDX.d_tab[t]) CK(cudaFree(IDX.d_tab[t]));
  free(IDX.d_tab);
  free(IDX.cap_tab);
  memset(&IDX, 0, sizeof IDX);
  cleanup_registered = 0;                          /* gpu_cleanup() ran the hook list: register again next time */
}

static void reserve_index(size_t n, size_t k, size_t d_short, size_t d, size_t tries) {
  const size_t w = sizeof(ftype);
  drop_index();
  IDX.n = n; IDX.k = k; IDX.d_short = d_short; IDX.d = d; IDX.tries = tries;
  reserve((void **)&IDX.d_points, &IDX.cap_points, n * d * w);
  reserve((void **)&IDX.d_mean, &IDX.cap_mean, d * w);
  reserve((void **)&IDX.d_bases, &IDX.cap_bases, (tries * d_short * d + 1) * w);
  reserve((void **)&IDX.d_graph, &IDX.cap_graph, n * k * 4);
  IDX.screened = annb_query_screen_applies(d, k);
  if (IDX.screened) {
    reserve(&IDX.d_points16, &IDX.cap_points16, n * d * 2);
    reserve(&IDX.d_nrm, &IDX.cap_nrm, n * 8);
    reserve((void **)&IDX.d_scale, &IDX.cap_scale, 256);
  }
  if (tries > IDX.cap_tries) {
    IDX.d_tab = realloc(IDX.d_tab, tries * sizeof(annb_u32 *));
    IDX.cap_tab = realloc(IDX.cap_tab, tries * sizeof(size_t));
    for (size_t t = IDX.cap_tries; t < tries; t++) { IDX.d_tab[t] = NULL; IDX.cap_tab[t] = 0; }
    IDX.cap_tries = tries;
  }
  if (!cleanup_registered) {
    register_cleanup(release_index);
    cleanup_registered = 1;
  }
}

void annh_forget_save_impl(const save_t *save) {
  if (IDX.live && IDX.graph_key == save->graph) drop_index();
}
void annh_forget_save(const save_t *save) {
  if (annh_multi_gpus() > 1 && !annh_multi_in_worker()) annh_multi_forget(save);
  else annh_forget_save_impl(save);
}

/* include/annb200_io.h: callers that rewrite `points` or a save_t IN PLACE (same addresses, same
 * shape) must say so; the cache key cannot see that (it samples, it does not hash everything).  */
void annb200_query_cache_invalidate(void) { drop_index(); }

static void upload_narrow(const size_t *host, size_t count, annb_u32 *dst, size_t *tmp, cudaStream_t st) {
  CK(cudaMemcpyAsync(tmp, host, count * sizeof(size_t), cudaMemcpyHostToDevice, st));
  annb_narrow_ids(tmp, count, dst, st);
}

/* fp16 copy + norms of the indexed points (device-resident points and means must be enqueued) */
static void prepare_screen(cudaStream_t st) {
  if (!IDX.screened) return;
  annb_screen_scale(IDX.d_points, IDX.d_mean, IDX.n, IDX.d, IDX.d_scale, st);
  annb_screen_prep_points(IDX.d_points, IDX.d_mean, IDX.n, IDX.d, IDX.d_scale, IDX.d_points16, IDX.d_nrm, st);
}

static void build_index(const save_t *s, const ftype *points, uint64_t fp) {
  cudaStream_t st = (cudaStream_t)annh_stream();
  const size_t w = sizeof(ftype), T = (size_t)s->tries, B = (size_t)1 << s->d_short;
  reserve_index(s->n, s->k, s->d_short, s->d_long, T);
  size_t tmp_cells = s->n * s->k;
  for (size_t t = 0; t < T; t++) {
    size_t cells = B * s->par_maxes[t];
    reserve((void **)&IDX.d_tab[t], &IDX.cap_tab[t], cells * 4);
    if (cells > tmp_cells) tmp_cells = cells;
  }
  size_t *tmp = NULL;
  CK(cudaMalloc((void **)&tmp, tmp_cells * sizeof(size_t)));
  annh_ingest(IDX.d_points, points, s->n * s->d_long * w, st, annh_device());
  CK(cudaMemcpyAsync(IDX.d_mean, s->row_means, s->d_long * w, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(IDX.d_bases, s->bases, T * s->d_short * s->d_long * w, cudaMemcpyHostToDevice, st));
  prepare_screen(st);
  upload_narrow(s->graph, s->n * s->k, IDX.d_graph, tmp, st);
  for (size_t t = 0; t < T; t++) upload_narrow(s->which_par[t], B * s->par_maxes[t], IDX.d_tab[t], tmp, st);
  CK(cudaStreamSynchronize(st));
  CK(cudaFree(tmp));
  IDX.graph_key = s->graph; IDX.points_key = points;
  IDX.fingerprint = fp;
  IDX.live = 1;
}

/* ---- adoption: precomp_gpu(save != NULL) leaves the index on the device -------------------
 * Everything query_gpu needs already sits in device memory at the end of precomp (points,
 * merged graph, bucket tables), so it is copied device-to-device into the cache instead of
 * being uploaded again by the first query.                                                 */
static int cache_enabled(void) {
  const char *env = getenv("ANN_B200_QUERY_CACHE");
  return !(env && *env == '0');
}

int annh_index_adopt_begin(size_t n, size_t k, size_t d_short, size_t d, size_t tries) {
  if (!cache_enabled()) return 0;
  reserve_index(n, k, d_short, d, tries);
  return 1;
}

/* the cache's own buffer for try t's 32-bit table: precomp_gpu exports straight into it */
annb_u32 *annh_index_table_buffer(int t, size_t cells) {
  reserve((void **)&IDX.d_tab[t], &IDX.cap_tab[t], cells * 4);
  return IDX.d_tab[t];
}

void annh_index_adopt_finish(const save_t *s, const ftype *host_points, const ftype *dev_points,
                             const ftype *dev_mean, const annb_u32 *dev_graph) {
  cudaStream_t st = (cudaStream_t)annh_stream();
  const size_t w = sizeof(ftype);
  CK(cudaMemcpyAsync(IDX.d_points, dev_points, s->n * s->d_long * w, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(IDX.d_mean, dev_mean, s->d_long * w, cudaMemcpyDeviceToDevice, st));
  CK(cudaMemcpyAsync(IDX.d_bases, s->bases, (size_t)s->tries * s->d_short * s->d_long * w, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(IDX.d_graph, dev_graph, s->n * s->k * 4, cudaMemcpyDeviceToDevice, st));
  prepare_screen(st);
  CK(cudaStreamSynchronize(st));
  IDX.graph_key = s->graph;
  IDX.points_key = host_points;
  IDX.fingerprint = fingerprint(s, host_points);
  IDX.live = 1;
}

size_t *query_gpu(const save_t *save, const ftype *points, size_t ycnt, const ftype *y,
                  ftype **dists_o) {
  if (annh_multi_gpus() > 1 && !annh_multi_in_worker()) return annh_multi_query(save, points, ycnt, y, dists_o);
  return annh_query_impl(save, points, ycnt, y, dists_o);
}

size_t *annh_query_impl(const save_t *save, const ftype *points, size_t ycnt, const ftype *y,
                        ftype **dists_o) {
  gpu_init();
  cudaStream_t st = (cudaStream_t)annh_stream();
  const char *qenv = getenv("ANN_B200_HOSTPROF");
  const int qprof = qenv && *qenv && *qenv != '0';
  double qlast = q_now_ms();
  const size_t n = save->n, k = save->k, d = save->d_long, ds = save->d_short, w = sizeof(ftype);
  const size_t T = (size_t)save->tries;
  if (n >= 0xFFFFFFFFull) annh_fatal("%s", "n must be below 2^32 - 1");
  if (k > 256) annh_fatal("%s", "k > 256 is not supported");
  if (ycnt == 0) {
    if (dists_o) *dists_o = malloc(1);
    return malloc(1);
  }
  int use_cache = cache_enabled();
  uint64_t fp = fingerprint(save, points);
  if (!(use_cache && IDX.live && IDX.graph_key == save->graph && IDX.points_key == points &&
        IDX.n == n && IDX.k == k && IDX.d_short == ds && IDX.d == d && IDX.tries == T &&
        IDX.fingerprint == fp))
    build_index(save, points, fp);
  QP("fingerprint+index");

  annh_egress *eg = annh_egress_begin(ycnt, k, dists_o != NULL, 0, annh_device(), NULL);
  const size_t scratch_bytes = ycnt * 4 + 1024 + ((size_t)64 << 20);
  size_t need = (ycnt * d * w + 256) + (T * ycnt * 4 + 256) + 2 * (ycnt * k * 4 + 256) +
                2 * (ycnt * k * w + 256) + scratch_bytes + 4096;
  annh_arena_reserve(need);
  ftype *dy = annh_arena_take(ycnt * d * w);
  annb_u32 *dsign = annh_arena_take(T * ycnt * 4);
  annb_u32 *down_ids = annh_arena_take(ycnt * k * 4), *dout_ids = annh_arena_take(ycnt * k * 4);
  ftype *down_dist = annh_arena_take(ycnt * k * w), *dout_dist = annh_arena_take(ycnt * k * w);
  void *dscratch = annh_arena_take(scratch_bytes);
  int *dstatus = annh_arena_take(sizeof(int));
  CK(cudaMemsetAsync(dstatus, 0, sizeof(int), st));

  const int same_set = (y == points);                        /* compute.cl:145 pointer test */
  const ftype *dq = dy;
  if (same_set && ycnt <= n) dq = IDX.d_points;
  else CK(cudaMemcpyAsync(dy, y, ycnt * d * w, cudaMemcpyHostToDevice, st));

  QP("alloc+upload y");
  annb_query_hash(dq, IDX.d_mean, IDX.d_bases, ycnt, d, ds, (int)T, dsign, st);
  QP("query_hash");
  annb_query_screen qs;
  qs.points16 = IDX.d_points16; qs.nrm = IDX.d_nrm; qs.mean = IDX.d_mean; qs.scale_bits = IDX.d_scale;
  annb_query_rows(dq, IDX.d_points, (const annb_u32 *const *)IDX.d_tab, save->par_maxes, (int)T,
                  dsign, n, ycnt, d, ds, k, same_set, down_ids, down_dist, dscratch, scratch_bytes,
                  dstatus, IDX.screened ? &qs : NULL, st);
  QP("query_rows");
  int nch = annh_egress_chunks(eg);
  for (int c = 0; c < nch; c++) {
    size_t r0 = (ycnt * (size_t)c / nch) & ~(size_t)31;
    size_t r1 = c + 1 == nch ? ycnt : (ycnt * (size_t)(c + 1) / nch) & ~(size_t)31;
    annb_supercharge(dq, IDX.d_points, down_ids, down_dist, IDX.d_graph, n, d, k, r0, r1, same_set,
                     dout_ids + r0 * k, dists_o ? dout_dist + r0 * k : NULL, dscratch, scratch_bytes,
                     dstatus, NULL, st);
    annh_egress_chunk(eg, r0, r1, dout_ids + r0 * k, dout_dist + r0 * k, st);
  }
  int h_status = 0;
  CK(cudaMemcpyAsync(&h_status, dstatus, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  QP("supercharge");
  size_t *result = annh_egress_end(eg, dists_o, NULL);
  QP("egress_end");
  if (h_status) annh_fatal("%s", "scratch too small for a literal candidate row");
  if (!use_cache) drop_index();
  return result;
}
