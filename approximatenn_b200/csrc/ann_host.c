/* ann_host.c — C host driver of the B200 backend: precomp_gpu / query_gpu and the device
 * lifecycle hooks, i.e. the symbols the reference's ann.c and test programs bind
 * (/root/reference/algg.h:5-11, gpu_comp.h:11-13).
 *
 * The host does what the reference's host does (alg.c:342-434) and nothing else on the
 * CPU: derive d_short/d_max, draw the random transforms from libc random() in the
 * reference's order, sequence the device stages, assemble save_t.  All per-point work
 * runs in the sm_100a kernels behind include/annb200.h.  There is no CPU fallback: every
 * CUDA failure prints a message and exits, like the reference's OpenCL glue
 * (gpu_comp.c:15-19, alggp.c:35-41).
 */
#include <math.h>
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
#include "annb200_dist.h"

/* ------------------------------------------------------------------------------------ */
/* errors, lifecycle                                                                     */

void annh_fatal(const char *fmt, const char *detail) {
  fprintf(stderr, "approximatenn_b200: ");
  fprintf(stderr, fmt, detail);
  fprintf(stderr, "\n");
  exit(1);
}

#define CK(call)                                                                        \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      fprintf(stderr, "approximatenn_b200: %s failed at %s:%d: %s\n", #call, __FILE__,  \
              __LINE__, cudaGetErrorString(e_));                                        \
      exit(1);                                                                          \
    }                                                                                   \
  } while (0)

typedef struct hook { void (*f)(void); struct hook *next; } hook;
#define ANNH_MAX_SPANS 1024

static __thread struct {
  int ready;
  int device;
  cudaStream_t stream;
  hook *hooks;
  /* one device arena reused across calls; grown on demand, released by gpu_cleanup()  */
  char *arena;
  size_t arena_bytes, arena_used;
  annb_u32 *table_buf;         /* exported 32-bit bucket tables when no query cache takes them */
  size_t table_cap;            /* cells */
  annh_stage_times last;
  int timing;
  cudaEvent_t ev[2 * ANNH_MAX_SPANS];
  int span_stage[ANNH_MAX_SPANS];
  int spans;
  /* second stream: the bucket tables and sorted copies of try j+1 are built while try j's lists
   * are computed (precomp step 6); events order the two, buffers alternate                    */
  cudaStream_t stream2;
  cudaEvent_t ev_ready, ev_s2[2], ev_leaf[2];
} G;

static void gpu_init_device(int forced_device);

void gpu_init(void) {
  if (G.ready) return;
  if (annh_multi_gpus() > 1 && !annh_multi_in_worker()) { annh_multi_start(); return; }
  gpu_init_device(-1);
}

void annh_gpu_init_on(int device) { if (!G.ready) gpu_init_device(device); }

static void gpu_init_device(int forced_device) {
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    annh_fatal("no CUDA device: %s (this library has no CPU fallback)",
               e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
  const char *env = getenv("ANN_B200_DEVICE");
  int dev = 0;
  if (forced_device >= 0) dev = forced_device;
  else if (env && *env) dev = atoi(env);
  else CK(cudaGetDevice(&dev));
  if (dev < 0 || dev >= count) annh_fatal("ANN_B200_DEVICE=%s is out of range", env ? env : "?");
  CK(cudaSetDevice(dev));
  G.device = dev;
  CK(cudaStreamCreateWithFlags(&G.stream, cudaStreamNonBlocking));
  CK(cudaStreamCreateWithFlags(&G.stream2, cudaStreamNonBlocking));
  for (int i = 0; i < 2 * ANNH_MAX_SPANS; i++) CK(cudaEventCreate(&G.ev[i]));
  CK(cudaEventCreateWithFlags(&G.ev_ready, cudaEventDisableTiming));
  for (int i = 0; i < 2; i++) {
    CK(cudaEventCreateWithFlags(&G.ev_s2[i], cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&G.ev_leaf[i], cudaEventDisableTiming));
  }
  const char *tm = getenv("ANN_B200_TIMING");
  G.timing = tm && *tm && *tm != '0';
  G.ready = 1;
}

void register_cleanup(void (*f)(void)) {
  hook *h = malloc(sizeof *h);
  h->f = f;
  h->next = G.hooks;
  G.hooks = h;
}

void gpu_cleanup(void) {
  if (annh_multi_gpus() > 1 && !annh_multi_in_worker()) annh_multi_stop();
  annh_gpu_cleanup_impl();
}

void annh_gpu_cleanup_impl(void) {
  if (!G.ready) return;
  while (G.hooks) {
    hook *h = G.hooks;
    G.hooks = h->next;
    h->f();
    free(h);
  }
  CK(cudaStreamSynchronize(G.stream));
  CK(cudaStreamSynchronize(G.stream2));
  annh_egress_release();
  annh_ingest_release();
  if (G.table_buf) CK(cudaFree(G.table_buf));
  G.table_buf = NULL;
  G.table_cap = 0;
  if (G.arena) CK(cudaFree(G.arena));
  G.arena = NULL;
  G.arena_bytes = G.arena_used = 0;
  for (int i = 0; i < 2 * ANNH_MAX_SPANS; i++) CK(cudaEventDestroy(G.ev[i]));
  CK(cudaEventDestroy(G.ev_ready));
  for (int i = 0; i < 2; i++) { CK(cudaEventDestroy(G.ev_s2[i])); CK(cudaEventDestroy(G.ev_leaf[i])); }
  CK(cudaStreamDestroy(G.stream2));
  CK(cudaStreamDestroy(G.stream));
  G.ready = 0;
}

void *annh_stream(void) { gpu_init(); return G.stream; }
int annh_device(void) { gpu_init(); return G.device; }
const annh_stage_times *annh_last_times(void) { return &G.last; }
void annh_set_timing(int on) { gpu_init(); G.timing = on; }

/* ------------------------------------------------------------------------------------ */
/* device arena                                                                          */

void annh_arena_reserve(size_t bytes) {
  if (bytes <= G.arena_bytes) { G.arena_used = 0; return; }
  if (G.arena) {
    CK(cudaStreamSynchronize(G.stream));
    CK(cudaFree(G.arena));
  }
  size_t free_b = 0, total_b = 0;
  CK(cudaMemGetInfo(&free_b, &total_b));
  if (bytes > free_b) {
    char msg[128];
    snprintf(msg, sizeof msg, "%.2f GB needed, %.2f GB free", bytes / 1e9, free_b / 1e9);
    annh_fatal("device memory: %s", msg);
  }
  CK(cudaMalloc((void **)&G.arena, bytes));
  G.arena_bytes = bytes;
  G.arena_used = 0;
}

static size_t pad256(size_t b) { return (b + 255) & ~(size_t)255; }

void *annh_arena_take(size_t bytes) {
  size_t at = G.arena_used;
  G.arena_used += pad256(bytes);
  if (G.arena_used > G.arena_bytes) annh_fatal("internal: %s", "device arena overrun");
  return G.arena + at;
}

/* ------------------------------------------------------------------------------------ */
/* sizes and transforms                                                                  */

void annh_params(size_t n, size_t k, size_t d, size_t *d_short_o, size_t *d_max_o) {
  /* alg.c:347-357: the division is in ftype, the logarithm in double                   */
  size_t d_short = ceil(log2((ftype)n / k));
  size_t d_max = 1;
  while (d_max < d) d_max <<= 1;
  if (d_short > d_max) d_short = d_max;
  *d_short_o = d_short;
  *d_max_o = d_max;
}

static int floor_log2_sz(size_t v) {
  int r = 0;
  while (v >>= 1) r++;
  return r;
}

static double draw_unit(void) {                        /* rand_pr.c:6-8 */
  return (double)(unsigned long)random() / ((double)RAND_MAX + 1);
}

/* rand_pr.c:17-30: `picks` swaps of a Fisher-Yates shuffle over [0,range)                */
static void draw_subperm(size_t picks, size_t range, size_t *p) {
  for (size_t i = 0; i < range; i++) p[i] = i;
  for (size_t i = 0; i < picks; i++) {
    size_t j = (unsigned long)random() % (range - i) + i;
    size_t t = p[i]; p[i] = p[j]; p[j] = t;
  }
}

/* One transform: sweeps are stored "before" first, then "after"; (ci,cj,angle) per plane */
typedef struct {
  size_t *ci, *cj;
  ftype *ang;
  size_t *perm_b, *perm_ai;
} host_transform;

static void draw_sweeps(size_t sweeps, size_t planes, size_t range, size_t *ci, size_t *cj,
                        ftype *ang, size_t *tmp) {
  for (size_t s = 0; s < sweeps; s++) {               /* rand_pr.c:10-16 */
    draw_subperm(2 * planes, range, tmp);
    for (size_t q = 0; q < planes; q++) {
      ci[s * planes + q] = tmp[2 * q];
      cj[s * planes + q] = tmp[2 * q + 1];
      ang[s * planes + q] = draw_unit() * M_PI;
    }
  }
}

static void draw_transform(host_transform *t, size_t rots_b, size_t len_b, size_t rots_a,
                           size_t len_a, size_t d_short, size_t d, size_t d_max) {
  size_t pb = rots_b * len_b, pa = rots_a * len_a;
  t->ci = malloc(sizeof(size_t) * (pb + pa + 1));
  t->cj = malloc(sizeof(size_t) * (pb + pa + 1));
  t->ang = malloc(sizeof(ftype) * (pb + pa + 1));
  t->perm_b = malloc(sizeof(size_t) * d_max);
  t->perm_ai = malloc(sizeof(size_t) * d_max);
  size_t *tmp = malloc(sizeof(size_t) * d_max);
  draw_sweeps(rots_b, len_b, d, t->ci, t->cj, t->ang, tmp);                    /* alg.c:65 */
  draw_sweeps(rots_a, len_a, d_short, t->ci + pb, t->cj + pb, t->ang + pb, tmp); /* alg.c:66 */
  draw_subperm(d, d_max, t->perm_b);                                           /* alg.c:67 */
  draw_subperm(d_short, d_max, t->perm_ai);                                    /* alg.c:70 */
  free(tmp);
}

static void free_transform(host_transform *t) {
  free(t->ci); free(t->cj); free(t->ang); free(t->perm_b); free(t->perm_ai);
}

void *annh_draw_transforms(size_t n, size_t k, size_t d, int tries, size_t rots_before, size_t rot_len_before,
                           size_t rots_after, size_t rot_len_after) {
  size_t d_short, d_max;
  annh_params(n, k, d, &d_short, &d_max);
  host_transform *tf = malloc(sizeof(host_transform) * (size_t)tries);
  for (int t = 0; t < tries; t++)
    draw_transform(tf + t, rots_before, rot_len_before, rots_after, rot_len_after, d_short, d, d_max);
  return tf;
}
void annh_free_transforms(void *p, int tries) {
  host_transform *tf = p;
  for (int t = 0; t < tries; t++) free_transform(tf + t);
  free(tf);
}

/* ---- the d_short x d projection matrix of one transform (save->bases), alg.c:189-217 ----
 * Tiny (d_short rows of d_max values), so it is evaluated on the host with the reference's
 * exact operation order: identity -> perm_ai embed -> reversed "after" sweeps with the
 * plane coordinates swapped -> Walsh-Hadamard -> perm_b projection -> reversed "before"
 * sweeps.  Compiled with -ffp-contract=off.                                              */
static void host_sweep(ftype *v, size_t planes, const size_t *ci, const size_t *cj,
                       const ftype *ang) {
  for (size_t q = 0; q < planes; q++) {
    ftype c = cos(ang[q]), s = sin(ang[q]);
    ftype a = v[ci[q]], b = v[cj[q]];
    ftype na = a * c - b * s, nb = a * s + b * c;
    v[ci[q]] = na;
    v[cj[q]] = nb;
  }
}

static void host_walsh(ftype *z, size_t len) {
  int levels = floor_log2_sz(len);
  for (int lev = 0; lev < levels; lev++) {
    size_t stride = (size_t)1 << lev;
    ftype div = lev % 2 + 1;
    for (size_t base = 0; base < len; base += 2 * stride)
      for (size_t o = 0; o < stride; o++) {
        ftype a = z[base + o], b = z[base + o + stride];
        z[base + o] = (a + b) / div;
        z[base + o + stride] = (a - b) / div;
      }
    if (lev == 0 && levels % 2) {
      ftype r = 1 / sqrt(2.0);
      for (size_t i = 0; i < len; i++) z[i] *= r;
    }
  }
}

static void projection_rows(const host_transform *t, size_t rots_b, size_t len_b, size_t rots_a,
                            size_t len_a, size_t d_short, size_t d, size_t d_max, ftype *out) {
  size_t pb = rots_b * len_b;
  ftype *z = malloc(sizeof(ftype) * d_max);
  for (size_t r = 0; r < d_short; r++) {
    for (size_t y = 0; y < d_max; y++) z[y] = (t->perm_ai[y] == r) ? 1 : 0;
    for (size_t s = rots_a; s-- > 0;)
      host_sweep(z, len_a, t->cj + pb + s * len_a, t->ci + pb + s * len_a, t->ang + pb + s * len_a);
    host_walsh(z, d_max);
    ftype *o = out + r * d;
    for (size_t y = 0; y < d_max; y++)
      if (t->perm_b[y] < d) o[t->perm_b[y]] = z[y];
    for (size_t s = rots_b; s-- > 0;)
      host_sweep(o, len_b, t->cj + s * len_b, t->ci + s * len_b, t->ang + s * len_b);
  }
  free(z);
}

/* ------------------------------------------------------------------------------------ */
/* stage timing (CUDA events on the library stream; off unless asked for)                */

/* a span = [begin, end) event pair on the library stream, attributed to one stage          */
static int span_begin_on(int stage, cudaStream_t s) {
  if (!G.timing || G.spans >= ANNH_MAX_SPANS) return -1;
  int id = G.spans++;
  G.span_stage[id] = stage;
  CK(cudaEventRecord(G.ev[2 * id], s));
  return id;
}
static void span_end_on(int id, cudaStream_t s) {
  if (id >= 0) CK(cudaEventRecord(G.ev[2 * id + 1], s));
}
static int span_begin(int stage) { return span_begin_on(stage, G.stream); }
static void span_end(int id) { span_end_on(id, G.stream); }
static void collect_times(void) {
  memset(&G.last, 0, sizeof G.last);
  if (!G.timing || G.spans == 0) { G.spans = 0; return; }
  for (int i = 0; i < G.spans; i++) {
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, G.ev[2 * i], G.ev[2 * i + 1]));
    G.last.ms[G.span_stage[i]] += ms;
  }
  float total = 0;
  CK(cudaEventElapsedTime(&total, G.ev[0], G.ev[2 * (G.spans - 1) + 1]));
  G.last.ms[ANNH_STAGES - 1] = total;
  G.spans = 0;
}

/* host wall-clock checkpoints (ANN_B200_HOSTPROF=1 prints them to stderr)                */
#include <time.h>
static double now_ms(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}
static int hostprof(void) {
  static int v = -1;
  if (v < 0) { const char *e = getenv("ANN_B200_HOSTPROF"); v = e && *e && *e != '0'; }
  return v;
}
#define HP(label)                                                                       \
  do {                                                                                  \
    if (hostprof()) {                                                                   \
      double t_ = now_ms();                                                             \
      fprintf(stderr, "[hostprof r%d] %-28s +%8.3f ms (total %8.3f)\n", annh_dist_rank(), label, t_ - hp_last, t_ - hp_start); \
      hp_last = t_;                                                                     \
    }                                                                                   \
  } while (0)

/* ------------------------------------------------------------------------------------ */
/* precomp_gpu                                                                            */

static void validate(size_t n, size_t k, size_t d, int tries, size_t len_b, size_t len_a,
                     size_t d_short) {
  if (n < 2 || k < 1 || d < 1 || tries < 1) annh_fatal("%s", "n >= 2, k >= 1, d >= 1, tries >= 1 required");
  if (n >= 0xFFFFFFFFull) annh_fatal("%s", "n must be below 2^32 - 1 (32-bit ids on the device)");
  if (k >= n) annh_fatal("%s", "k must be smaller than n");
  if (k > 256) annh_fatal("%s", "k > 256 is not supported");
  if (d_short > 28) annh_fatal("%s", "more than 2^28 buckets per try");
  if (2 * len_b > d) annh_fatal("%s", "2*rot_len_before must not exceed d (rand_pr.c:17-30)");
  if (2 * len_a > d_short) annh_fatal("%s", "2*rot_len_after must not exceed d_short (alg.c:66)");
}

/* save->which_par / par_maxes once every try's largest bucket is known (alg.c:270-271,376-377):
 * host arrays malloc()ed here, device tables (32-bit cells) in the query cache's own buffers
 * when it adopts this index, else in one library-owned block; dtab[t] receives the addresses.  */
static annh_tables *save_tables_begin(save_t *save, const annb_u32 *h_tm, size_t T, size_t buckets,
                                      int adopt, annb_u32 **dtab) {
  size_t *cells = malloc(sizeof(size_t) * T), total = 0;
  for (size_t t = 0; t < T; t++) {
    cells[t] = buckets * (size_t)h_tm[t];
    total += (cells[t] + 63) & ~(size_t)63;
    save->par_maxes[t] = h_tm[t];
    save->which_par[t] = malloc((cells[t] ? cells[t] : 1) * sizeof(size_t));
    if (!save->which_par[t]) annh_fatal("%s", "out of host memory for the bucket tables");
  }
  if (!adopt && total > G.table_cap) {
    if (G.table_buf) CK(cudaFree(G.table_buf));
    G.table_cap = total + total / 4 + 1024;
    CK(cudaMalloc((void **)&G.table_buf, G.table_cap * sizeof(annb_u32)));
  }
  size_t at = 0;
  for (size_t t = 0; t < T; t++) {
    dtab[t] = adopt ? annh_index_table_buffer((int)t, cells[t]) : G.table_buf + at;
    at += (cells[t] + 63) & ~(size_t)63;
  }
  annh_tables *tb = annh_tables_begin((int)T, cells, save->which_par, G.device);
  free(cells);
  return tb;
}

size_t *precomp_gpu(size_t n, size_t k, size_t d, const ftype *points, int tries,
                    size_t rots_before, size_t rot_len_before, size_t rots_after,
                    size_t rot_len_after, save_t *save, ftype **dists_o) {
  if (annh_multi_gpus() > 1 && !annh_multi_in_worker())
    return annh_multi_precomp(n, k, d, points, tries, rots_before, rot_len_before, rots_after, rot_len_after,
                              save, dists_o);
  return annh_precomp_impl(n, k, d, points, tries, rots_before, rot_len_before, rots_after, rot_len_after, save,
                           dists_o, NULL);
}

size_t *annh_precomp_impl(size_t n, size_t k, size_t d, const ftype *points, int tries,
                          size_t rots_before, size_t rot_len_before, size_t rots_after,
                          size_t rot_len_after, save_t *save, ftype **dists_o, const annh_call_ctx *ctx) {
  double hp_start = now_ms(), hp_last = hp_start;
  gpu_init();
  CK(cudaSetDevice(G.device));
  size_t d_short, d_max;
  annh_params(n, k, d, &d_short, &d_max);
  validate(n, k, d, tries, rot_len_before, rot_len_after, d_short);
  HP("init");
  const size_t T = (size_t)tries, buckets = (size_t)1 << d_short;
  const size_t planes = rots_before * rot_len_before + rots_after * rot_len_after;
  const size_t w = sizeof(ftype);
  cudaStream_t st = G.stream;

  /* sharding (ann_dist.c): rank r owns tries t = r, r+R, ... and the rows [row_lo, row_hi)  */
  const int single = ctx && ctx->force_single;
  const int R = single ? 1 : annh_dist_world(), rank = single ? 0 : annh_dist_rank();
  const int sharded = R > 1;
  size_t row_lo = 0, row_hi = n;
  if (sharded) annb200_dist_slice(n, rank, R, &row_lo, &row_hi);
  const size_t my_rows = row_hi - row_lo;
  size_t Tl = 0;                                               /* tries owned by this rank  */
  for (size_t t = 0; t < T; t++) Tl += annb200_dist_try_owner((int)t, R) == rank;
  if (sharded && T > 64) annh_fatal("%s", "more than 64 tries in sharded mode");
  const int full_result = !sharded || annh_dist_gather_results() || save != NULL;
  const size_t out_rows = full_result ? n : my_rows;
  annh_egress *eg = annh_egress_begin(out_rows, k, dists_o != NULL, save != NULL, G.device, ctx ? &ctx->out : NULL);

  HP("egress_begin");
  /* 1. transforms: ALL tries are drawn before any compute (alg.c:388-392), on every rank   */
  host_transform *tf = ctx && ctx->transforms ? ctx->transforms : malloc(sizeof(host_transform) * T);
  if (!(ctx && ctx->transforms))
    for (size_t t = 0; t < T; t++)
      draw_transform(tf + t, rots_before, rot_len_before, rots_after, rot_len_after, d_short, d, d_max);
  size_t *own = malloc(sizeof(size_t) * (Tl + 1));             /* global index of owned try j */
  for (size_t t = 0, j = 0; t < T; t++)
    if (annb200_dist_try_owner((int)t, R) == rank) own[j++] = t;
  annb_u32 *h_idx = malloc(sizeof(annb_u32) * (Tl * planes * 2 + 1));
  ftype *h_cs = malloc(sizeof(ftype) * (Tl * planes * 2 + 1));
  annb_u32 *h_permb = malloc(sizeof(annb_u32) * (Tl * d_max + 1));
  annb_u32 *h_pick = malloc(sizeof(annb_u32) * (Tl * d_short + 1));
  for (size_t j = 0; j < Tl; j++) {
    const host_transform *f = tf + own[j];
    for (size_t q = 0; q < planes; q++) {
      h_idx[(j * planes + q) * 2] = (annb_u32)f->ci[q];
      h_idx[(j * planes + q) * 2 + 1] = (annb_u32)f->cj[q];
      /* libm in double on the ftype-rounded angle, then rounded to ftype (ocl2c.h:10)  */
      h_cs[(j * planes + q) * 2] = cos(f->ang[q]);
      h_cs[(j * planes + q) * 2 + 1] = sin(f->ang[q]);
    }
    for (size_t y = 0; y < d_max; y++) {
      h_permb[j * d_max + y] = (annb_u32)f->perm_b[y];
      if (f->perm_ai[y] < d_short) h_pick[j * d_short + f->perm_ai[y]] = (annb_u32)y;
    }
  }

  HP("transforms");
  /* 2. device memory plan                                                              */
  annb_transform_desc desc;
  memset(&desc, 0, sizeof desc);
  desc.n = n; desc.d = d; desc.d_max = d_max; desc.d_short = d_short;
  desc.rots_before = rots_before; desc.rot_len_before = rot_len_before;
  desc.rots_after = rots_after; desc.rot_len_after = rot_len_after;
  desc.tries = (int)Tl;
  desc.inv_sqrt2 = 1 / sqrt(2.0);
  const size_t list_bytes = n * k * (4 + w);                   /* one per-try list      */
  const size_t np = sharded ? annh_dist_padded_rows(n) : n;    /* rows of all-gathered arrays */
  const size_t scratch_bytes = annb_leaf_scratch_bytes(n, d, d_short, k);
  size_t fixed = pad256(np * d * w) + pad256(n * d * w) + pad256(d * w) + pad256(Tl * n * 4 + 4) +
                 pad256(buckets * 4) + pad256((buckets + 1) * 4) + pad256(n * 4) * 2 +
                 pad256(T * 4) + pad256(annb_scan_tmp_bytes(buckets)) +
                 pad256(Tl * planes * 2 * 4 + 4) + pad256(Tl * planes * 2 * w + w) +
                 pad256(Tl * d_max * 4 + 4) + pad256(Tl * d_short * 4 + 4) +
                 pad256(annb_hash_scratch_bytes(&desc)) +
                 pad256(scratch_bytes) + 8192 + 256;
  const int s5_screened = annb_supercharge_screen_applies(d, k);
  if (s5_screened) fixed += pad256(n * d * 2) + pad256(n * 8); /* fp16 copy of the points (original order) + norms */
  /* ANN_B200_S5_LOCALITY=1: rows worked on in bucket order of one try.  Off by default: on iid
   * Gaussian points (the benchmark data) it changes neither the L2 hit rate nor the time (cfg3
   * 4.96 vs 4.92 ms, cfg4 112 vs 112 ms) — the candidates of a row come from the
   * other tries' buckets and share nothing with its bucket neighbours'.                         */
  int s5_local = 0;
  {
    const char *e = getenv("ANN_B200_S5_LOCALITY");
    if (e && *e && *e != '0') s5_local = s5_screened && Tl > 0;
  }
  const size_t local_bytes = s5_local ? annb_locality_scratch_bytes(my_rows, d_short, 8) : 0;
  if (s5_local) fixed += pad256(local_bytes) + pad256(my_rows * 4);
  if (sharded && save) fixed += pad256(T * n * 4);            /* every try's hashes, for the tables */
  if (sharded)   /* merged ids (all rows), merged dists + results (own rows), exchanged lists */
    fixed += pad256(np * k * 4) + pad256(my_rows * k * w) + pad256(my_rows * k * 4) + pad256(my_rows * k * w) +
             pad256(T * my_rows * k * 4) + pad256(T * my_rows * k * w) +
             (full_result ? pad256(np * k * 4) + pad256(np * k * w) : 0);
  else
    fixed += pad256(n * k * 4) * 3 + pad256(n * k * w) * 3;
  /* cutoff carried from try to try (annb200.h): only where the merged row is sorted as a whole,
   * i.e. k*tries is a power of two (no prefix cut, no corner rule), and every list is merged at once */
  int use_cut = Tl >= 2 && k * T >= 16 && ((k * T) & (k * T - 1)) == 0 && annb_cutoff_applies(d, d_short, k);
  if (use_cut) fixed += pad256(n * k * w) + pad256(n * w);
  /* S2 of try j+1 on a second stream while S3 of try j runs (ANN_B200_PIPE=0 switches it off):
   * a second set of the per-try buffers (offsets, order, sorted copy, leaf scratch)            */
  int pipe = Tl >= 2;
  {
    const char *e = getenv("ANN_B200_PIPE");
    if (e && *e && *e == '0') pipe = 0;
  }
  const size_t pipe_bytes = pad256(n * d * w) + pad256(scratch_bytes) + pad256((buckets + 1) * 4) + pad256(n * 4);
  size_t group;                                                /* lists kept before a merge */
  for (;;) {
    const size_t fx = fixed + (pipe ? pipe_bytes : 0);
    group = Tl ? Tl : 1;
    if (!sharded && fx + group * list_bytes + 512 > G.arena_bytes) {
      /* the arena has to grow: see what the device can give (cudaMemGetInfo costs milliseconds,
       * so it is not asked when the cached arena already fits the whole plan)                  */
      size_t free_b = 0, total_b = 0;
      CK(cudaMemGetInfo(&free_b, &total_b));
      free_b += G.arena_bytes;
      while (group > 1 && fx + group * list_bytes + 512 > free_b * 9 / 10) group--;
    }
    if (pipe && group < Tl) { pipe = 0; continue; }            /* no room for the second set */
    fixed = fx;
    break;
  }
  {
    /* ANN_B200_MERGE_GROUP=g forces the grouped merge (tests; otherwise only a list set that does
     * not fit the device triggers it)                                                          */
    const char *mg = getenv("ANN_B200_MERGE_GROUP");
    if (!sharded && mg && *mg && atoi(mg) >= 1 && (size_t)atoi(mg) < group) group = (size_t)atoi(mg);
  }
  if ((size_t)k * T < 16) group = Tl ? Tl : 1;
  if (!sharded && group < Tl) {
    /* A running merge sees the lists a group at a time, so rows with EXACT distance ties cannot be
     * redone with the reference's literal network and the prefix-corner rule is not applied.
     * Everything else is unchanged.  Say so: the contract elsewhere is bit-exactness.           */
    static int warned = 0;
    if (!warned) {
      fprintf(stderr, "approximatenn_b200: warning: the %zu per-try lists (%.1f GB) do not fit next to the rest of the "
              "plan; merging %zu at a time. Rows with exact distance ties between different points, and the "
              "prefix-corner rule (k*tries not a power of two), may differ from the reference in this mode.\n",
              Tl, Tl * list_bytes / 1e9, group);
      warned = 1;
    }
  }
  if (group < Tl) { use_cut = 0; pipe = 0; }
  annh_arena_reserve(fixed + group * list_bytes + 512);

  ftype *dX = annh_arena_take(np * d * w), *dXs = annh_arena_take(n * d * w), *dmean = annh_arena_take(d * w);
  annb_u32 *dhash = annh_arena_take(Tl * n * 4 + 4);
  annb_u32 *dcount = annh_arena_take(buckets * 4), *doffset = annh_arena_take((buckets + 1) * 4);
  annb_u32 *dorder_tmp = annh_arena_take(n * 4), *dorder = annh_arena_take(n * 4);
  annb_u32 *dtmax = annh_arena_take(T * 4);
  void *dscan = annh_arena_take(annb_scan_tmp_bytes(buckets));
  annb_u32 *d_idx = annh_arena_take(Tl * planes * 2 * 4 + 4);
  ftype *d_cs = annh_arena_take(Tl * planes * 2 * w + w);
  annb_u32 *d_permb = annh_arena_take(Tl * d_max * 4 + 4), *d_pick = annh_arena_take(Tl * d_short * 4 + 4);
  void *dhscratch = annh_arena_take(annb_hash_scratch_bytes(&desc));
  annb_u32 *dl_ids = annh_arena_take(group * n * k * 4);
  ftype *dl_dist = annh_arena_take(group * n * k * w);
  /* merged lists: ids for all rows (the graph S5 walks); distances for the owned rows only.
   * Single GPU: two sets, ping-ponged by the grouped merge.  out_*: result rows, indexed from
   * out_base (0 when the whole result is assembled here, row_lo when only the owned rows are). */
  const size_t out_base = (sharded && !full_result) ? row_lo : 0;
  const size_t out_cap = (sharded && !full_result) ? my_rows : np;
  annb_u32 *dm_ids = annh_arena_take(np * k * 4), *dm_ids2 = sharded ? NULL : annh_arena_take(n * k * 4);
  ftype *dm_dist = sharded ? NULL : annh_arena_take(n * k * w);
  ftype *dm_dist2 = annh_arena_take((sharded ? my_rows : n) * k * w);
  annb_u32 *dout_ids = annh_arena_take(out_cap * k * 4);
  ftype *dout_dist = annh_arena_take(out_cap * k * w);
  void *dscratch = annh_arena_take(scratch_bytes);
  int *dstatus = annh_arena_take(sizeof(int));
  unsigned *dscreen = annh_arena_take(256);                     /* scale word of the screened S3 path */
  const int screened = annb_screen_applies(d, d_short, k);
  void *dX16 = s5_screened ? annh_arena_take(n * d * 2) : NULL;
  void *dN16 = s5_screened ? annh_arena_take(n * 8) : NULL;
  void *dlocal = s5_local ? annh_arena_take(local_bytes) : NULL;
  annb_u32 *dperm = s5_local ? annh_arena_take(my_rows * 4) : NULL;
  /* the per-try buffers, twice when S2 is pipelined against S3 (set j & 1 serves try j) */
  ftype *Xs_set[2] = {dXs, pipe ? (ftype *)annh_arena_take(n * d * w) : dXs};
  void *scratch_set[2] = {dscratch, pipe ? annh_arena_take(scratch_bytes) : dscratch};
  annb_u32 *offset_set[2] = {doffset, pipe ? (annb_u32 *)annh_arena_take((buckets + 1) * 4) : doffset};
  annb_u32 *order_set[2] = {dorder, pipe ? (annb_u32 *)annh_arena_take(n * 4) : dorder};
  ftype *drun = use_cut ? annh_arena_take(n * k * w) : NULL;    /* k smallest distinct distances so far */
  ftype *dcut = use_cut ? annh_arena_take(n * w) : NULL;        /* their largest: the cutoff of a point  */
  annb_u32 *dhash_all = (sharded && save) ? annh_arena_take(T * n * 4) : NULL;
  annb_u32 *ds_ids = sharded ? annh_arena_take(T * my_rows * k * 4) : NULL;   /* [T][my_rows][k] */
  ftype *ds_dist = sharded ? annh_arena_take(T * my_rows * k * w) : NULL;
  CK(cudaMemsetAsync(dstatus, 0, sizeof(int), st));
  CK(cudaMemsetAsync(dtmax, 0, T * 4, st));

  HP("arena");
  /* 3. upload: the whole set, or this rank's rows followed by an all-gather over NVLink   */
  int sp = span_begin(0);
  if (my_rows) annh_ingest(dX + row_lo * d, points + row_lo * d, my_rows * d * w, st, G.device);
  if (sharded) annh_dist_allgather_rows(dX, n, d * w, st);
  if (planes && Tl) {
    CK(cudaMemcpyAsync(d_idx, h_idx, Tl * planes * 2 * 4, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(d_cs, h_cs, Tl * planes * 2 * w, cudaMemcpyHostToDevice, st));
  }
  if (Tl) CK(cudaMemcpyAsync(d_permb, h_permb, Tl * d_max * 4, cudaMemcpyHostToDevice, st));
  if (d_short && Tl) CK(cudaMemcpyAsync(d_pick, h_pick, Tl * d_short * 4, cudaMemcpyHostToDevice, st));
  desc.plane_idx = d_idx; desc.plane_cs = d_cs; desc.perm_b = d_permb; desc.pick = d_pick;
  desc.host_plane_idx = h_idx; desc.host_plane_cs = h_cs; desc.host_perm_b = h_permb; desc.host_pick = h_pick;

  HP("upload enqueued");
  /* 4. S0 column means (alg.c:367-368); the accumulator borrows the sorted-copy buffer */
  span_end(sp);
  sp = span_begin(1);
  {
    int left = floor_log2_sz(n);                   /* halvings until one row remains        */
    size_t len = n;
    const ftype *src = dX;
    while (left > 0) {
      int f = left < 4 ? left : 4;
      annb_fold_rows(src, dXs, len, d, f, src == dX, st);
      len >>= f;
      left -= f;
      src = dXs;
    }
  }
  annb_scale_means(dXs, n, d, dmean, st);
  if ((screened && Tl) || s5_screened) annb_screen_scale(dX, dmean, n, d, dscreen, st);
  if (s5_screened) annb_screen_prep_points(dX, dmean, n, d, dscreen, dX16, dN16, st);

  /* 5. S1 hashes of every owned try in one pass over the points                         */
  span_end(sp);
  sp = span_begin(2);
  if (Tl) annb_hash_points(dX, dmean, &desc, dhash, dhscratch, st);
  span_end(sp);

  int adopt = 0;
  annh_tables *tb = NULL;
  annb_u32 **dtab = NULL, *h_tm_all = NULL;
  if (save && !sharded) adopt = annh_index_adopt_begin(n, k, d_short, d, T);
  if (save) {
    save->tries = tries; save->n = n; save->k = k; save->d_short = d_short; save->d_long = d;
    save->row_means = malloc(w * d);
    save->which_par = malloc(sizeof(size_t *) * T);
    save->par_maxes = malloc(sizeof(size_t) * T);
    save->bases = malloc(w * T * d_short * d);
    dtab = malloc(sizeof(annb_u32 *) * T);
    h_tm_all = malloc(sizeof(annb_u32) * T);
    CK(cudaMemcpyAsync(save->row_means, dmean, d * w, cudaMemcpyDeviceToHost, st));
    if (!sharded) {
      /* every try's largest bucket first (a histogram each), ONE synchronisation, and from then on
       * the tables leave as 32-bit cells behind the tries without stopping the stream             */
      for (size_t t = 0; t < T; t++) annb_bucket_max(dhash + t * n, n, buckets, dcount, dtmax + t, st);
      CK(cudaMemcpyAsync(h_tm_all, dtmax, 4 * T, cudaMemcpyDeviceToHost, st));
    }
    for (size_t t = 0; t < T; t++)                 /* host work while the device hashes */
      projection_rows(tf + t, rots_before, rot_len_before, rots_after, rot_len_after, d_short, d,
                      d_max, save->bases + t * d_short * d);
    if (!sharded) {
      CK(cudaStreamSynchronize(st));
      tb = save_tables_begin(save, h_tm_all, T, buckets, adopt, dtab);
    }
  }

  /* 6. per owned try: S2 bucket tables, S3 k best; merge whenever `group` lists wait      */
  const size_t row_len = k * T;
  const int tiny_merge = row_len < 16;
  const size_t prefix = tiny_merge ? row_len : (size_t)1 << floor_log2_sz(row_len);
  int corner_list = -1, corner_pos = 0;
  if (!tiny_merge && prefix < row_len) {                         /* DESIGN.md "prefix corner" */
    corner_list = (int)(prefix / k);
    corner_pos = (int)(prefix % k);
  }
  int have_merged = 0;
  int admit[64];
  cudaStream_t sb = pipe ? G.stream2 : st;                     /* where S2 runs */
  if (pipe) {
    CK(cudaEventRecord(G.ev_ready, st));                        /* hashes (and the save preliminaries) done */
    CK(cudaStreamWaitEvent(sb, G.ev_ready, 0));
  }
  for (size_t j0 = 0; j0 < Tl; j0 += group) {
    size_t g = Tl - j0 < group ? Tl - j0 : group;
    if (g > 64) annh_fatal("%s", "more than 64 tries per merge group");
    for (size_t j = 0; j < g; j++) {
      size_t t = own[j0 + j];
      const annb_u32 *hash_t = dhash + (j0 + j) * n;
      /* S2 of try jj into buffer set jj & 1.  Unpipelined: try j, right here.  Pipelined (one group,
       * j0 = 0): try 0 before the first S3, and try j + 1 now, so that it runs beside S3 of try j;
       * its buffers were last read by S3 of try j - 1                                            */
      for (size_t jj = (pipe && j > 0) ? j + 1 : j; jj <= (pipe ? j + 1 : j) && jj < g; jj++) {
        const size_t tt = own[j0 + jj];
        const int set = pipe ? (int)(jj & 1) : 0;
        if (pipe && jj >= 2) CK(cudaStreamWaitEvent(sb, G.ev_leaf[set], 0));
        sp = span_begin_on(3, sb);
        annb_build_buckets(dhash + (j0 + jj) * n, n, buckets, dcount, offset_set[set], dorder_tmp, order_set[set],
                           dtmax + tt, dscan, sb);
        if (tb) {                                  /* padded table for save->which_par[t] */
          annb_export_table32(offset_set[set], order_set[set], n, buckets, h_tm_all[tt], dtab[tt], sb);
          annh_tables_submit(tb, (int)tt, dtab[tt], sb);
        }
        if (screened) annb_gather_rows_screen(dX, order_set[set], n, d, dmean, dscreen, Xs_set[set], scratch_set[set], sb);
        else annb_gather_rows(dX, order_set[set], n, d, Xs_set[set], sb);
        span_end_on(sp, sb);
        if (pipe) CK(cudaEventRecord(G.ev_s2[set], sb));
      }
      const int set = pipe ? (int)(j & 1) : 0;
      if (pipe) CK(cudaStreamWaitEvent(st, G.ev_s2[set], 0));
      sp = span_begin(4);
      annb_leaf_topk_cut(Xs_set[set], dmean, order_set[set], offset_set[set], hash_t, dtmax + t, n, d, d_short, k,
                         dl_ids + j * n * k, dl_dist + j * n * k, scratch_set[set], dstatus,
                         screened ? dscreen : NULL, screened, (use_cut && j > 0) ? dcut : NULL, st);
      if (use_cut && j + 1 < Tl) annb_cutoff_update(dl_dist + j * n * k, drun, dcut, n, k, j == 0, st);
      span_end(sp);
      if (pipe) CK(cudaEventRecord(G.ev_leaf[set], st));
      admit[j] = annb200_dist_admit(k, tries, (int)t);
    }
    if (!sharded) {
      int whole = group == Tl;                     /* corner + literal redo need every list */
      sp = span_begin(6);
      annb_merge_lists(dl_ids, dl_dist, (int)g, admit, whole ? corner_list : -1, corner_pos,
                       have_merged ? dm_ids : NULL, have_merged ? dm_dist : NULL, n, n, k, whole,
                       dm_ids2, dm_dist2, dscratch, scratch_bytes, dstatus, st);
      annb_u32 *ti = dm_ids; dm_ids = dm_ids2; dm_ids2 = ti;
      ftype *td = dm_dist; dm_dist = dm_dist2; dm_dist2 = td;
      have_merged = 1;
      span_end(sp);
    }
  }

  HP("tries enqueued");
  /* 6b. sharded: every list goes to the owner of its rows, who merges all T of them       */
  const ftype *own_dist_base = dm_dist;            /* indexed with GLOBAL row numbers       */
  if (sharded) {
    sp = span_begin(5);
    annh_dist_exchange_lists(dl_ids, ds_ids, n, k * 4, tries, st);
    annh_dist_exchange_lists(dl_dist, ds_dist, n, k * w, tries, st);
    span_end(sp);
    sp = span_begin(6);
    for (size_t t = 0; t < T; t++) admit[t] = annb200_dist_admit(k, tries, (int)t);
    if (my_rows)
      annb_merge_lists(ds_ids, ds_dist, tries, admit, corner_list, corner_pos, NULL, NULL, my_rows, n,
                       k, 1, dm_ids + row_lo * k, dm_dist2, dscratch, scratch_bytes, dstatus, st);
    span_end(sp);
    sp = span_begin(5);
    annh_dist_allgather_rows(dm_ids, n, k * 4, st);   /* neighbours' lists for supercharging */
    span_end(sp);
    own_dist_base = dm_dist2 - row_lo * k;
  }

  /* 7. S5 supercharging of the owned rows (alg.c:313-327); graph = the merged lists        */
  {
    annb_supercharge_opts s5;
    s5.points16 = dX16; s5.nrm = dN16; s5.scale_bits = dscreen; s5.row_perm = NULL; s5.perm_base = row_lo;
    int nch = full_result && sharded ? 1 : annh_egress_chunks(eg);
    if (s5_local && my_rows) {
      size_t lo[65];
      for (int c = 0; c < nch; c++) lo[c] = (my_rows * (size_t)c / nch) & ~(size_t)31;
      lo[nch] = my_rows;
      sp = span_begin(7);
      annb_locality_order(dhash + (Tl - 1) * n, row_lo, my_rows, d_short, nch, lo, dlocal, dperm, st);
      span_end(sp);
      s5.row_perm = dperm;
    }
    for (int c = 0; c < nch; c++) {
      size_t r0 = row_lo + ((my_rows * (size_t)c / nch) & ~(size_t)31);
      size_t r1 = c + 1 == nch ? row_hi : row_lo + ((my_rows * (size_t)(c + 1) / nch) & ~(size_t)31);
      sp = span_begin(7);
      annb_supercharge(dX, dX, dm_ids, own_dist_base, dm_ids, n, d, k, r0, r1, 1,
                       dout_ids + (r0 - out_base) * k, dout_dist + (r0 - out_base) * k, dscratch,
                       scratch_bytes, dstatus, s5_screened ? &s5 : NULL, st);
      span_end(sp);
      if (!(full_result && sharded))
        annh_egress_chunk(eg, r0 - row_lo, r1 - row_lo, dout_ids + (r0 - out_base) * k,
                          dout_dist + (r0 - out_base) * k, st);
    }
    if (full_result && sharded) {                  /* every rank ends up with all rows      */
      annh_dist_allgather_rows(dout_ids, n, k * 4, st);
      annh_dist_allgather_rows(dout_dist, n, k * w, st);
      annh_egress_chunk(eg, 0, n, dout_ids, dout_dist, st);
    }
  }

  if (sharded && save) {
    /* every rank gets every try's hashes, rebuilds the bucket tables and exports them     */
    for (size_t j = 0; j < Tl; j++)
      CK(cudaMemcpyAsync(dhash_all + own[j] * n, dhash + j * n, n * 4, cudaMemcpyDeviceToDevice, st));
    for (size_t t = 0; t < T; t++)
      annh_dist_broadcast(dhash_all + t * n, n * 4, annb200_dist_try_owner((int)t, R), st);
    for (size_t t = 0; t < T; t++) annb_bucket_max(dhash_all + t * n, n, buckets, dcount, dtmax + t, st);
    CK(cudaMemcpyAsync(h_tm_all, dtmax, 4 * T, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    tb = save_tables_begin(save, h_tm_all, T, buckets, 0, dtab);
    for (size_t t = 0; t < T; t++) {
      annb_build_buckets(dhash_all + t * n, n, buckets, dcount, doffset, dorder_tmp, dorder, dtmax + t, dscan, st);
      annb_export_table32(doffset, dorder, n, buckets, h_tm_all[t], dtab[t], st);
      annh_tables_submit(tb, (int)t, dtab[t], st);
    }
  }
  HP("supercharge enqueued");
  /* 8. results: the egress threads are already widening the first chunks                 */
  annb_u32 *h_tmax = malloc(4 * T);
  int h_status = 0;
  CK(cudaMemcpyAsync(h_tmax, dtmax, 4 * T, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&h_status, dstatus, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  HP("stream drained");
  collect_times();
  size_t *graph_copy = NULL;
  size_t *result = annh_egress_end(eg, dists_o, &graph_copy);
  HP("egress_end");
  if (h_status) annh_fatal("%s", "scratch too small for a literal candidate row (extremely unbalanced buckets)");
  for (size_t j = 0; j < Tl; j++)
    if ((d_short + 1) * (size_t)h_tmax[own[j]] < 16)
      annh_fatal("%s", "candidate rows shorter than 16 slots (n far too small for this k)");
  free(h_tmax);

  if (tb) annh_tables_end(tb);
  free(dtab); free(h_tm_all);
  if (save) save->graph = graph_copy;                           /* alg.c:428-432: two separate copies */
  if (adopt) annh_index_adopt_finish(save, points, dX, dmean, dout_ids);
  if (!(ctx && ctx->transforms)) {
    for (size_t t = 0; t < T; t++) free_transform(tf + t);
    free(tf);
  }
  free(own); free(h_idx); free(h_cs); free(h_permb); free(h_pick);
  return result;
}

/* ------------------------------------------------------------------------------------ */
/* query_gpu — see ann_query.c                                                            */

/* ------------------------------------------------------------------------------------ */
/* ann.h front door.  Weak, so that the reference's own ann.c wins when it is linked in.  */

__attribute__((weak)) void free_save(save_t *save) {            /* ann.c:25-34 */
  annh_forget_save(save);
  for (int i = 0; i < save->tries; i++) free(save->which_par[i]);
  free(save->which_par);
  free(save->par_maxes);
  free(save->graph);
  free(save->row_means);
  free(save->bases);
}

__attribute__((weak)) size_t *precomp(size_t n, size_t k, size_t d, const ftype *points, int tries,
                                      size_t rots_before, size_t rot_len_before, size_t rots_after,
                                      size_t rot_len_after, save_t *save, ftype **dists,
                                      char use_cpu) {
  if (use_cpu)
    annh_fatal("%s", "use_cpu != 0: this library contains no CPU path (link the reference's algc.c and ann.c for one)");
  return precomp_gpu(n, k, d, points, tries, rots_before, rot_len_before, rots_after,
                     rot_len_after, save, dists);
}

__attribute__((weak)) size_t *query(const save_t *save, const ftype *points, size_t ycnt,
                                    const ftype *y, ftype **dists, char use_cpu) {
  if (use_cpu)
    annh_fatal("%s", "use_cpu != 0: this library contains no CPU path (link the reference's algc.c and ann.c for one)");
  return query_gpu(save, points, ycnt, y, dists);
}
