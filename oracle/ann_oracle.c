/* ann_oracle.c — CPU restatement of approximateNN's randomized all-points kNN path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under approximatenn_b200/ may include, link, call or
 * execute this file; it exists so that tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline leg can check the CUDA path against the reference's algorithm.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_reference.py compares every output of this
 * restatement (graph ids, squared distances, save_t fields) bit for bit against the
 * reference's own pure-C path compiled from /root/reference into oracle/_ref (see
 * oracle/Makefile), and tests/golden/ holds vectors generated from that build.
 *
 * What is restated (all citations are into /root/reference):
 *   parameter derivation                 alg.c:347-357
 *   RNG draw order of the transforms     alg.c:37-74, rand_pr.c:6-30
 *   column means (stride-halving tree)   alg.c:122-128,367-369  compute.cl:15-49
 *   transform + sign hashing             alg.c:154-183  compute.cl:55-122,223-231
 *   projection rows for the index        alg.c:189-217
 *   bucket tables                        alg.c:252-267
 *   candidate rows, distances            alg.c:233-242,274-283  compute.cl:135-167,238-246
 *   sort / kill duplicates / sort        alg.c:137-144,224-230  compute.cl:181-217
 *   merge over tries + supercharging     alg.c:303-337  compute.cl:252-263
 *   query                                alg.c:438-519  compute.cl:268-275
 *
 * Unlike the reference, this file never materialises the n x L x d scratch array
 * (alg.c:237): candidate rows are produced, measured and sorted one point at a time, so
 * memory is O(n*k*tries) and the BASELINE configs with n >= 1e6 fit.  The arithmetic of
 * each row is the reference's (same operations, same association order, no fused
 * multiply-add: build with -ffp-contract=off).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include "ann.h"

#ifdef USE_FLOAT
typedef uint32_t fbits_t;
#else
typedef uint64_t fbits_t;
#endif

/* ------------------------------------------------------------------------------------ */
/* small helpers                                                                        */

static int floor_log2(size_t v) {            /* algc.c:13-22 computes the same quantity */
  int r = 0;
  while (v >>= 1) r++;
  return r;
}

static void *xmalloc(size_t bytes) {
  void *p = malloc(bytes ? bytes : 1);
  if (!p) abort();
  return p;
}

static fbits_t bits_of(ftype v) {
  fbits_t b;
  memcpy(&b, &v, sizeof b);
  return b;
}

/* uniform double in [0,1) from libc random(), rand_pr.c:6-8 */
static double unit_draw(void) {
  return (double)(unsigned long)random() / ((double)RAND_MAX + 1);
}

/* ------------------------------------------------------------------------------------ */
/* derived sizes, alg.c:347-357                                                          */

void orc_params(size_t n, size_t k, size_t d, size_t *d_short_o, size_t *d_max_o) {
  size_t d_short = ceil(log2((ftype)n / k));   /* ftype division, double log2: alg.c:347 */
  size_t d_max = 1;
  while (d_max < d) d_max <<= 1;               /* next power of two >= d, alg.c:348-355    */
  if (d_short > d_max) d_short = d_max;
  *d_short_o = d_short;
  *d_max_o = d_max;
}

/* ------------------------------------------------------------------------------------ */
/* transforms: drawn on the host from random(), alg.c:37-74 + rand_pr.c:10-30            */

typedef struct {
  size_t rots_b, len_b, rots_a, len_a;
  size_t *bi, *bj;       /* [rots_b][len_b] plane coordinates before the Walsh step       */
  ftype *bang;           /* [rots_b][len_b] angles                                        */
  size_t *ai, *aj;       /* [rots_a][len_a] after the Walsh step (coordinates < d_short)  */
  ftype *aang;
  size_t *perm_b;        /* [d_max]                                                       */
  size_t *perm_ai;       /* [d_max]                                                       */
} orc_transform;

/* partial Fisher-Yates over [0,range): `picks` draws, each random() % (range-i) + i     */
static size_t *draw_subperm(size_t picks, size_t range) {
  size_t *p = xmalloc(sizeof(size_t) * range);
  for (size_t i = 0; i < range; i++) p[i] = i;
  for (size_t i = 0; i < picks; i++) {
    size_t j = (unsigned long)random() % (range - i) + i;
    size_t t = p[i]; p[i] = p[j]; p[j] = t;
  }
  return p;
}

static void draw_sweeps(size_t sweeps, size_t planes, size_t range,
                        size_t *ci, size_t *cj, ftype *ang) {
  for (size_t s = 0; s < sweeps; s++) {
    size_t *p = draw_subperm(2 * planes, range);        /* coordinates first ...          */
    for (size_t q = 0; q < planes; q++) {
      ci[s * planes + q] = p[2 * q];
      cj[s * planes + q] = p[2 * q + 1];
      ang[s * planes + q] = unit_draw() * M_PI;         /* ... then the angles             */
    }
    free(p);
  }
}

static void orc_draw_transform(orc_transform *t, size_t rots_b, size_t len_b,
                               size_t rots_a, size_t len_a,
                               size_t d_short, size_t d, size_t d_max) {
  t->rots_b = rots_b; t->len_b = len_b; t->rots_a = rots_a; t->len_a = len_a;
  t->bi = xmalloc(sizeof(size_t) * rots_b * len_b);
  t->bj = xmalloc(sizeof(size_t) * rots_b * len_b);
  t->bang = xmalloc(sizeof(ftype) * rots_b * len_b);
  t->ai = xmalloc(sizeof(size_t) * rots_a * len_a);
  t->aj = xmalloc(sizeof(size_t) * rots_a * len_a);
  t->aang = xmalloc(sizeof(ftype) * rots_a * len_a);
  draw_sweeps(rots_b, len_b, d, t->bi, t->bj, t->bang);          /* alg.c:65 */
  draw_sweeps(rots_a, len_a, d_short, t->ai, t->aj, t->aang);    /* alg.c:66 */
  t->perm_b = draw_subperm(d, d_max);                            /* alg.c:67 */
  t->perm_ai = draw_subperm(d_short, d_max);                     /* alg.c:70 */
}

static void orc_free_transform(orc_transform *t) {
  free(t->bi); free(t->bj); free(t->bang);
  free(t->ai); free(t->aj); free(t->aang);
  free(t->perm_b); free(t->perm_ai);
}

/* ------------------------------------------------------------------------------------ */
/* column means: the reference's stride-halving tree, alg.c:122-128 + compute.cl:15-39   */

void orc_means(size_t n, size_t d, const ftype *points, ftype *means) {
  size_t half = n / 2;
  ftype *acc = xmalloc(sizeof(ftype) * (half ? half : 1) * d);
  for (size_t c = 0; c < d; c++) {
    for (size_t x = 0; x < half; x++) {
      ftype extra = (x == 0 && (n & 1)) ? points[(n - 1) * d + c] : 0;
      acc[x * d + c] = points[x * d + c] + points[(x + half) * d + c] + extra;
    }
  }
  for (size_t len = half; len >> 1; len >>= 1) {
    size_t h = len / 2;
    for (size_t c = 0; c < d; c++)
      for (size_t x = 0; x < h; x++) {
        ftype extra = (x == 0 && (len & 1)) ? acc[(len - 1) * d + c] : 0;
        acc[x * d + c] += acc[(x + h) * d + c] + extra;
      }
  }
  for (size_t c = 0; c < d; c++) means[c] = acc[c] / n;
  free(acc);
}

/* ------------------------------------------------------------------------------------ */
/* the orthogonal transform on ONE row, compute.cl:55-122                                */

static void givens_sweep(ftype *v, size_t planes, const size_t *ci, const size_t *cj,
                         const ftype *ang) {
  for (size_t q = 0; q < planes; q++) {
    ftype c = cos(ang[q]), s = sin(ang[q]);      /* double libm on the ftype angle, ocl2c.h:10 */
    ftype a = v[ci[q]], b = v[cj[q]];
    ftype na = a * c - b * s;
    ftype nb = a * s + b * c;
    v[ci[q]] = na;
    v[cj[q]] = nb;
  }
}

/* Orthonormal Walsh-Hadamard in natural (Sylvester) order: butterflies at stride
 * 1,2,4,...; halve on odd levels; one multiply by 1/sqrt(2) right after level 0 when
 * log2(len) is odd.  compute.cl:101-122, alg.c:112-120.                                  */
static void walsh_row(ftype *z, size_t len) {
  if (len == 1) return;
  int levels = floor_log2(len);
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

/* hash of one (already centred) row; scratch holds 2*d_max ftypes.  alg.c:154-183        */
static size_t hash_row(const ftype *centred, size_t d, size_t d_short, size_t d_max,
                       const orc_transform *t, ftype *scratch) {
  ftype *v = scratch, *z = scratch + d_max;
  memcpy(v, centred, sizeof(ftype) * d);
  for (size_t s = 0; s < t->rots_b; s++)
    givens_sweep(v, t->len_b, t->bi + s * t->len_b, t->bj + s * t->len_b, t->bang + s * t->len_b);
  for (size_t y = 0; y < d_max; y++)                       /* compute.cl:77-85 */
    z[y] = t->perm_b[y] < d ? v[t->perm_b[y]] : 0;
  walsh_row(z, d_max);
  for (size_t s = 0; s < t->rots_a; s++)
    givens_sweep(z, t->len_a, t->ai + s * t->len_a, t->aj + s * t->len_a, t->aang + s * t->len_a);
  for (size_t y = 0; y < d_max; y++)                       /* compute.cl:88-96 */
    if (t->perm_ai[y] < d_short) v[t->perm_ai[y]] = z[y];
  size_t h = 0;
  for (size_t i = 0; i < d_short; i++)                     /* compute.cl:223-231 */
    h = h << 1 | (size_t)(bits_of(v[i]) >> (sizeof(ftype) * 8 - 1));
  return h;
}

/* rows of the projection: push the identity through the inverse chain, alg.c:189-217     */
static void projection_rows(size_t d, size_t d_short, size_t d_max,
                            const orc_transform *t, ftype *out /* [d_short][d] */) {
  ftype *z = xmalloc(sizeof(ftype) * d_max);
  for (size_t r = 0; r < d_short; r++) {
    for (size_t y = 0; y < d_max; y++)                      /* embed e_r through perm_ai   */
      z[y] = t->perm_ai[y] < d_short ? (ftype)(t->perm_ai[y] == r) : 0;
    for (size_t s = t->rots_a; s-- > 0;)                     /* reversed, i/j swapped       */
      givens_sweep(z, t->len_a, t->aj + s * t->len_a, t->ai + s * t->len_a, t->aang + s * t->len_a);
    walsh_row(z, d_max);
    ftype *o = out + r * d;
    for (size_t y = 0; y < d_max; y++)
      if (t->perm_b[y] < d) o[t->perm_b[y]] = z[y];
    for (size_t s = t->rots_b; s-- > 0;)
      givens_sweep(o, t->len_b, t->bj + s * t->len_b, t->bi + s * t->len_b, t->bang + s * t->len_b);
  }
  free(z);
}

/* ------------------------------------------------------------------------------------ */
/* bucket table, alg.c:252-267: row b = ids hashing to b in DEcreasing order, pad = n      */

static size_t *bucket_table(size_t n, size_t d_short, const size_t *hash, size_t *tmax_o) {
  size_t buckets = (size_t)1 << d_short;
  size_t *fill = xmalloc(sizeof(size_t) * buckets);
  memset(fill, 0, sizeof(size_t) * buckets);
  for (size_t j = 0; j < n; j++) fill[hash[j]]++;
  size_t tmax = 0;
  for (size_t b = 0; b < buckets; b++) if (fill[b] > tmax) tmax = fill[b];
  size_t *tab = xmalloc(sizeof(size_t) * tmax * buckets);
  for (size_t i = 0; i < tmax * buckets; i++) tab[i] = n;
  for (size_t j = 0; j < n; j++) tab[hash[j] * tmax + --fill[hash[j]]] = j;
  free(fill);
  *tmax_o = tmax;
  return tab;
}

/* ------------------------------------------------------------------------------------ */
/* squared distance with the reference's summation tree, compute.cl:135-167               */

static ftype tree_sqdist(const ftype *a, const ftype *b, size_t d, ftype *tmp) {
  for (size_t z = 0; z < d; z++) {
    ftype diff = a[z] - b[z];
    tmp[z] = diff * diff;
  }
  for (size_t l = d; l >> 1; l >>= 1) {
    size_t h = l / 2;
    for (size_t z = 0; z < h; z++) {
      ftype extra = (z == 0 && (l & 1)) ? tmp[l - 1] : 0;
      tmp[z] += tmp[z + h] + extra;
    }
  }
  return tmp[0];
}

/* ------------------------------------------------------------------------------------ */
/* the sort network and the duplicate rule, compute.cl:181-217 + alg.c:137-144,224-230     */

static void network_sort(size_t *ids, ftype *key, size_t len) {
  int lk = floor_log2(len);
  size_t groups = (size_t)1 << (lk > 4 ? lk - 4 : 0);
  for (int stage = 0; stage < lk; stage++)
    for (int sub = stage; sub >= 0; sub--)
      for (size_t g = 0; g < groups; g++)
        for (size_t e = 0; e < 8; e++) {
          size_t w = g << 3 | e;
          size_t hi = (w >> sub) << sub, lo = w ^ hi;
          size_t pa = hi << 1 | lo;
          if (sub == stage) lo = ((size_t)1 << sub) - lo - 1;   /* the mirrored first pass */
          size_t pb = hi << 1 | (size_t)1 << sub | lo;
          if (pb < len && key[pa] > key[pb]) {
            ftype tk = key[pa]; key[pa] = key[pb]; key[pb] = tk;
            size_t ti = ids[pa]; ids[pa] = ids[pb]; ids[pb] = ti;
          }
        }
}

static void sort_and_uniq_row(size_t *ids, ftype *key, size_t len) {
  network_sort(ids, key, len);
  for (size_t y = 0; y + 1 < len; y++)             /* compute.cl:212-217 */
    if (ids[y] == ids[y + 1]) key[y] = INFINITY;
  network_sort(ids, key, len);
}

/* ------------------------------------------------------------------------------------ */
/* per-try candidate row of one point -> its k best, alg.c:274-288                        */

typedef struct { size_t *ids; ftype *key, *tmp; size_t cap; } row_buf;

static void row_reserve(row_buf *r, size_t len, size_t d) {
  if (len <= r->cap) return;
  free(r->ids); free(r->key); free(r->tmp);
  r->ids = xmalloc(sizeof(size_t) * len);
  r->key = xmalloc(sizeof(ftype) * len);
  r->tmp = xmalloc(sizeof(ftype) * (d ? d : 1));
  r->cap = len;
}

static void row_release(row_buf *r) { free(r->ids); free(r->key); free(r->tmp); }

/* distances of row r->ids[from..len) measured from `q`; `self` = id to exclude or n       */
static void row_distances(row_buf *r, size_t from, size_t len, const ftype *q, size_t self,
                          const ftype *points, size_t n, size_t d) {
  for (size_t y = from; y < len; y++) {
    size_t id = r->ids[y];
    if (id >= n || id == self) r->key[y] = INFINITY;   /* compute.cl:144-149 */
    else r->key[y] = tree_sqdist(q, points + id * d, d, r->tmp);
  }
}

static void try_row(size_t x, size_t n, size_t k, size_t d, size_t d_short,
                    const ftype *points, const size_t *hash, const size_t *tab, size_t tmax,
                    row_buf *r, size_t *ids_out, ftype *key_out) {
  size_t len = (d_short + 1) * tmax;
  row_reserve(r, len, d);
  for (size_t y = 0; y <= d_short; y++) {               /* compute.cl:238-246 */
    size_t b = hash[x] ^ (y ? (size_t)1 << (y - 1) : 0);
    memcpy(r->ids + y * tmax, tab + b * tmax, sizeof(size_t) * tmax);
  }
  row_distances(r, 0, len, points + x * d, x, points, n, d);
  sort_and_uniq_row(r->ids, r->key, len);
  memcpy(ids_out, r->ids, sizeof(size_t) * k);
  memcpy(key_out, r->key, sizeof(ftype) * k);
}

/* ------------------------------------------------------------------------------------ */
/* merge over tries + supercharging, alg.c:303-337                                        */

/* `rows_ids/rows_key`: [ycnt][len] candidate rows (keys already filled when have_keys).
 * `graph`: neighbour lists used for supercharging, row stride gstride; in precomp this is
 * rows_ids itself AFTER the first sort (alg.c:316).  Writes [ycnt][k] results.            */
static void finish_rows(size_t n, size_t k, size_t d, size_t ycnt, size_t len,
                        size_t *rows_ids, ftype *rows_key, int have_keys,
                        const size_t *graph, size_t gstride,
                        const ftype *y, const ftype *points,
                        size_t *ids_out, ftype *key_out) {
  int exclude_self = (y == points);                      /* compute.cl:145 pointer test */
  row_buf tmp = {0};
  for (size_t x = 0; x < ycnt; x++) {
    size_t *ids = rows_ids + x * len;
    ftype *key = rows_key + x * len;
    if (!have_keys) {
      row_buf view = {ids, key, NULL, len};
      view.tmp = xmalloc(sizeof(ftype) * (d ? d : 1));
      row_distances(&view, 0, len, y + x * d, exclude_self ? x : n, points, n, d);
      free(view.tmp);
    }
    sort_and_uniq_row(ids, key, len);
  }
  size_t wide = k * (k + 1);
  row_reserve(&tmp, wide, d);
  for (size_t x = 0; x < ycnt; x++) {
    const size_t *own = rows_ids + x * len;
    memcpy(tmp.ids, own, sizeof(size_t) * k);
    memcpy(tmp.key, rows_key + x * len, sizeof(ftype) * k);
    for (size_t j = 0; j < k; j++)                        /* compute.cl:252-263 */
      for (size_t z = 0; z < k; z++)
        tmp.ids[(j + 1) * k + z] = own[j] < n ? graph[own[j] * gstride + z] : n;
    row_distances(&tmp, k, wide, y + x * d, exclude_self ? x : n, points, n, d);
    sort_and_uniq_row(tmp.ids, tmp.key, wide);
    memcpy(ids_out + x * k, tmp.ids, sizeof(size_t) * k);
    if (key_out) memcpy(key_out + x * k, tmp.key, sizeof(ftype) * k);
  }
  row_release(&tmp);
}

/* ------------------------------------------------------------------------------------ */
/* stage-level entry points used by the parity tests                                      */

typedef struct {
  size_t n, k, d, d_short, d_max;
  int tries;
  orc_transform *tf;
  ftype *means;
  size_t **hash;     /* [tries][n]                 */
  size_t **tab;      /* [tries][2^d_short][tmax]   */
  size_t *tmax;      /* [tries]                    */
} orc_state;

/* Draws the transforms (consuming random() exactly like the reference), centres, hashes
 * and builds every bucket table.  Everything up to, not including, the distance work.    */
orc_state *orc_prepare(size_t n, size_t k, size_t d, const ftype *points, int tries,
                       size_t rots_b, size_t len_b, size_t rots_a, size_t len_a) {
  orc_state *s = xmalloc(sizeof *s);
  s->n = n; s->k = k; s->d = d; s->tries = tries;
  orc_params(n, k, d, &s->d_short, &s->d_max);
  s->means = xmalloc(sizeof(ftype) * d);
  orc_means(n, d, points, s->means);
  s->tf = xmalloc(sizeof(orc_transform) * tries);
  for (int t = 0; t < tries; t++)                          /* all tries first, alg.c:388-392 */
    orc_draw_transform(s->tf + t, rots_b, len_b, rots_a, len_a, s->d_short, d, s->d_max);
  s->hash = xmalloc(sizeof(size_t *) * tries);
  s->tab = xmalloc(sizeof(size_t *) * tries);
  s->tmax = xmalloc(sizeof(size_t) * tries);
  ftype *centred = xmalloc(sizeof(ftype) * d);
  ftype *scratch = xmalloc(sizeof(ftype) * 2 * s->d_max);
  for (int t = 0; t < tries; t++) s->hash[t] = xmalloc(sizeof(size_t) * n);
  for (size_t x = 0; x < n; x++) {
    for (size_t c = 0; c < d; c++) centred[c] = points[x * d + c] - s->means[c];
    for (int t = 0; t < tries; t++)
      s->hash[t][x] = hash_row(centred, d, s->d_short, s->d_max, s->tf + t, scratch);
  }
  free(centred); free(scratch);
  for (int t = 0; t < tries; t++)
    s->tab[t] = bucket_table(n, s->d_short, s->hash[t], s->tmax + t);
  return s;
}

void orc_release(orc_state *s, int keep_tables) {
  for (int t = 0; t < s->tries; t++) {
    orc_free_transform(s->tf + t);
    free(s->hash[t]);
    if (!keep_tables) free(s->tab[t]);
  }
  free(s->tf); free(s->hash); free(s->tab); free(s->tmax); free(s->means); free(s);
}

size_t orc_state_d_short(const orc_state *s) { return s->d_short; }
size_t orc_state_d_max(const orc_state *s) { return s->d_max; }
size_t orc_state_tmax(const orc_state *s, int t) { return s->tmax[t]; }
const size_t *orc_state_hash(const orc_state *s, int t) { return s->hash[t]; }
const size_t *orc_state_table(const orc_state *s, int t) { return s->tab[t]; }
const ftype *orc_state_means(const orc_state *s) { return s->means; }
void orc_state_projection(const orc_state *s, int t, ftype *out) {
  projection_rows(s->d, s->d_short, s->d_max, s->tf + t, out);
}

/* k best of try t for the points listed in rows[0..m): ids_out/key_out are [m][k].        */
void orc_try_lists(const orc_state *s, const ftype *points, int t,
                   const size_t *rows, size_t m, size_t *ids_out, ftype *key_out) {
  row_buf r = {0};
  for (size_t i = 0; i < m; i++)
    try_row(rows ? rows[i] : i, s->n, s->k, s->d, s->d_short, points, s->hash[t], s->tab[t],
            s->tmax[t], &r, ids_out + i * s->k, key_out + i * s->k);
  row_release(&r);
}

/* ------------------------------------------------------------------------------------ */
/* the public pair, same signatures as precomp_cpu / query_cpu (algc.h:5-11)              */

size_t *precomp_oracle(size_t n, size_t k, size_t d, const ftype *points, int tries,
                       size_t rots_before, size_t rot_len_before, size_t rots_after,
                       size_t rot_len_after, save_t *save, ftype **dists_o) {
  orc_state *s = orc_prepare(n, k, d, points, tries, rots_before, rot_len_before,
                             rots_after, rot_len_after);
  size_t len = k * (size_t)tries;
  size_t *rows_ids = xmalloc(sizeof(size_t) * n * len);
  ftype *rows_key = xmalloc(sizeof(ftype) * n * len);
  size_t *ids_k = xmalloc(sizeof(size_t) * k);
  ftype *key_k = xmalloc(sizeof(ftype) * k);
  row_buf r = {0};
  for (int t = 0; t < tries; t++)
    for (size_t x = 0; x < n; x++) {
      try_row(x, n, k, d, s->d_short, points, s->hash[t], s->tab[t], s->tmax[t], &r, ids_k, key_k);
      memcpy(rows_ids + x * len + k * t, ids_k, sizeof(size_t) * k);   /* alg.c:284-288 */
      memcpy(rows_key + x * len + k * t, key_k, sizeof(ftype) * k);
    }
  row_release(&r);
  free(ids_k); free(key_k);

  size_t *result = xmalloc(sizeof(size_t) * n * k);
  ftype *rdist = dists_o ? xmalloc(sizeof(ftype) * n * k) : NULL;
  finish_rows(n, k, d, n, len, rows_ids, rows_key, 1, rows_ids, len, points, points, result, rdist);
  free(rows_ids); free(rows_key);
  if (dists_o) *dists_o = rdist;

  if (save) {                                              /* alg.c:370-381,428-432 */
    save->tries = tries; save->n = n; save->k = k;
    save->d_short = s->d_short; save->d_long = d;
    save->row_means = xmalloc(sizeof(ftype) * d);
    memcpy(save->row_means, s->means, sizeof(ftype) * d);
    save->which_par = xmalloc(sizeof(size_t *) * tries);
    save->par_maxes = xmalloc(sizeof(size_t) * tries);
    save->bases = xmalloc(sizeof(ftype) * tries * s->d_short * d);
    for (int t = 0; t < tries; t++) {
      save->which_par[t] = s->tab[t];
      save->par_maxes[t] = s->tmax[t];
      projection_rows(d, s->d_short, s->d_max, s->tf + t, save->bases + (size_t)t * s->d_short * d);
    }
    save->graph = result;
    result = xmalloc(sizeof(size_t) * n * k);
    memcpy(result, save->graph, sizeof(size_t) * n * k);
  }
  orc_release(s, save != NULL);
  return result;
}

size_t *query_oracle(const save_t *save, const ftype *points, size_t ycnt, const ftype *y,
                     ftype **dists_o) {
  size_t n = save->n, k = save->k, d = save->d_long, ds = save->d_short;
  size_t T = save->tries;
  ftype *prod = xmalloc(sizeof(ftype) * (d ? d : 1));
  ftype *proj = xmalloc(sizeof(ftype) * (ds ? ds : 1));
  size_t *sign = xmalloc(sizeof(size_t) * T * ycnt);
  for (size_t x = 0; x < ycnt; x++)
    for (size_t t = 0; t < T; t++) {
      for (size_t i = 0; i < ds; i++) {                    /* compute.cl:268-275 + 160-167 */
        const ftype *b = save->bases + (t * ds + i) * d;
        for (size_t z = 0; z < d; z++) prod[z] = (y[x * d + z] - save->row_means[z]) * b[z];
        for (size_t l = d; l >> 1; l >>= 1) {
          size_t h = l / 2;
          for (size_t z = 0; z < h; z++) {
            ftype extra = (z == 0 && (l & 1)) ? prod[l - 1] : 0;
            prod[z] += prod[z + h] + extra;
          }
        }
        proj[i] = prod[0];
      }
      size_t h = 0;
      for (size_t i = 0; i < ds; i++)
        h = h << 1 | (size_t)(bits_of(proj[i]) >> (sizeof(ftype) * 8 - 1));
      sign[x * T + t] = h;                                 /* produced as [ycnt][tries] ... */
    }
  free(prod); free(proj);
  size_t total = 0;
  for (size_t t = 0; t < T; t++) total += save->par_maxes[t];
  size_t len = total * (ds + 1);
  size_t *rows_ids = xmalloc(sizeof(size_t) * ycnt * len);
  ftype *rows_key = xmalloc(sizeof(ftype) * ycnt * len);
  size_t off = 0;
  for (size_t t = 0; t < T; t++) {
    size_t w = save->par_maxes[t];
    for (size_t x = 0; x < ycnt; x++) {
      size_t h = sign[t * ycnt + x];                       /* ... consumed as [tries][ycnt], alg.c:495-499 */
      for (size_t f = 0; f <= ds; f++) {
        size_t b = h ^ (f ? (size_t)1 << (f - 1) : 0);
        memcpy(rows_ids + x * len + off * (ds + 1) + f * w, save->which_par[t] + b * w,
               sizeof(size_t) * w);
      }
    }
    off += w;
  }
  free(sign);
  size_t *result = xmalloc(sizeof(size_t) * ycnt * k);
  ftype *rdist = dists_o ? xmalloc(sizeof(ftype) * ycnt * k) : NULL;
  finish_rows(n, k, d, ycnt, len, rows_ids, rows_key, 0, save->graph, k, y, points, result, rdist);
  free(rows_ids); free(rows_key);
  if (dists_o) *dists_o = rdist;
  return result;
}

void free_save_oracle(save_t *save) {                      /* ann.c:25-34 */
  for (int t = 0; t < save->tries; t++) free(save->which_par[t]);
  free(save->which_par); free(save->par_maxes); free(save->graph);
  free(save->row_means); free(save->bases);
}

/* ------------------------------------------------------------------------------------ */
/* bounded CPU-baseline sample (bench.py cpu_baseline leg)                                */

/* Times the reference's per-point work at FULL problem size on `m` sampled points without
 * running all n of them: hashing and bucket tables for every point (cheap), then the
 * per-try candidate rows + merge for the sampled points and for every neighbour their
 * supercharging step reads, then supercharging of the sampled points.
 * secs[0] = hashing+tables for all n, secs[1] = per-try rows+merge per processed row,
 * secs[2] = supercharging per sampled point.  Returns the number of rows processed.       */
static size_t sampled_impl(size_t n, size_t k, size_t d, const ftype *points, int tries,
                           size_t rots_b, size_t len_b, size_t rots_a, size_t len_a,
                           const size_t *sample, size_t m, double secs[3], size_t *out_ids,
                           ftype *out_key) {
  struct timespec t0, t1;
  clock_gettime(CLOCK_MONOTONIC, &t0);
  orc_state *s = orc_prepare(n, k, d, points, tries, rots_b, len_b, rots_a, len_a);
  clock_gettime(CLOCK_MONOTONIC, &t1);
  secs[0] = (t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec);

  size_t len = k * (size_t)tries;
  size_t cap = m * (k + 1), rows = 0;
  size_t *row_of = xmalloc(sizeof(size_t) * cap);          /* processed point ids          */
  size_t *ids = xmalloc(sizeof(size_t) * cap * len);
  ftype *key = xmalloc(sizeof(ftype) * cap * len);
  row_buf r = {0};
  size_t *ids_k = xmalloc(sizeof(size_t) * k);
  ftype *key_k = xmalloc(sizeof(ftype) * k);

  clock_gettime(CLOCK_MONOTONIC, &t0);
  /* merged row of one point = per-try rows concatenated, then sort/uniq/sort            */
#define MERGED_ROW(pt)                                                                   \
  do {                                                                                   \
    for (int t = 0; t < tries; t++) {                                                    \
      try_row((pt), n, k, d, s->d_short, points, s->hash[t], s->tab[t], s->tmax[t], &r,  \
              ids_k, key_k);                                                             \
      memcpy(ids + rows * len + k * t, ids_k, sizeof(size_t) * k);                       \
      memcpy(key + rows * len + k * t, key_k, sizeof(ftype) * k);                        \
    }                                                                                    \
    sort_and_uniq_row(ids + rows * len, key + rows * len, len);                          \
    row_of[rows++] = (pt);                                                               \
  } while (0)
  for (size_t i = 0; i < m; i++) MERGED_ROW(sample[i]);
  for (size_t i = 0; i < m; i++)
    for (size_t j = 0; j < k; j++) {
      size_t nb = ids[i * len + j];
      if (nb < n) MERGED_ROW(nb);
    }
#undef MERGED_ROW
  clock_gettime(CLOCK_MONOTONIC, &t1);
  secs[1] = ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec)) / (double)rows;

  clock_gettime(CLOCK_MONOTONIC, &t0);
  size_t wide = k * (k + 1), next = m;
  row_buf w = {0};
  row_reserve(&w, wide, d);
  for (size_t i = 0; i < m; i++) {
    memcpy(w.ids, ids + i * len, sizeof(size_t) * k);
    memcpy(w.key, key + i * len, sizeof(ftype) * k);
    for (size_t j = 0; j < k; j++) {
      size_t nb = ids[i * len + j];
      for (size_t z = 0; z < k; z++) w.ids[(j + 1) * k + z] = nb < n ? ids[next * len + z] : n;
      if (nb < n) next++;
    }
    row_distances(&w, k, wide, points + sample[i] * d, sample[i], points, n, d);
    sort_and_uniq_row(w.ids, w.key, wide);
    if (out_ids) memcpy(out_ids + i * k, w.ids, sizeof(size_t) * k);
    if (out_key) memcpy(out_key + i * k, w.key, sizeof(ftype) * k);
  }
  clock_gettime(CLOCK_MONOTONIC, &t1);
  secs[2] = ((t1.tv_sec - t0.tv_sec) + 1e-9 * (t1.tv_nsec - t0.tv_nsec)) / (double)m;

  row_release(&w); row_release(&r);
  free(ids_k); free(key_k); free(row_of); free(ids); free(key);
  orc_release(s, 0);
  return rows;
}

size_t orc_sampled_cost(size_t n, size_t k, size_t d, const ftype *points, int tries,
                        size_t rots_b, size_t len_b, size_t rots_a, size_t len_a,
                        const size_t *sample, size_t m, double secs[3]) {
  return sampled_impl(n, k, d, points, tries, rots_b, len_b, rots_a, len_a, sample, m, secs, NULL, NULL);
}

/* Same walk, but the FINAL rows (ids and squared distances after supercharging, alg.c:328-335)
 * of the sampled points are returned: exact parity rows at sizes where a full run of the
 * restatement would take hours.  out_ids/out_key: [m][k].                                   */
size_t orc_sampled_rows(size_t n, size_t k, size_t d, const ftype *points, int tries,
                        size_t rots_b, size_t len_b, size_t rots_a, size_t len_a,
                        const size_t *sample, size_t m, double secs[3], size_t *out_ids,
                        ftype *out_key) {
  return sampled_impl(n, k, d, points, tries, rots_b, len_b, rots_a, len_a, sample, m, secs, out_ids, out_key);
}

/* ------------------------------------------------------------------------------------ */
/* the two halves of det_results as separate entry points (used by the sharding tests)   */

/* first sort_and_uniq of det_results (alg.c:312) on `rows` rows of `len` slots, in place  */
void orc_merge_rows(size_t *ids, ftype *key, size_t rows, size_t len) {
  for (size_t x = 0; x < rows; x++) sort_and_uniq_row(ids + x * len, key + x * len, len);
}

/* supercharge + compdists + second sort_and_uniq (alg.c:313-335) for query rows
 * [r0, r1): own lists are rows of own_ids/own_key (stride own_stride, indexed by x - r0),
 * neighbour lists are rows of `graph` (stride gstride, indexed by point id).               */
void orc_supercharge_rows(size_t n, size_t k, size_t d, const ftype *queries, const ftype *points,
                          size_t r0, size_t r1, const size_t *own_ids, const ftype *own_key,
                          size_t own_stride, const size_t *graph, size_t gstride, int exclude_self,
                          size_t *ids_out, ftype *key_out) {
  size_t wide = k * (k + 1);
  row_buf tmp = {0};
  row_reserve(&tmp, wide, d);
  for (size_t x = r0; x < r1; x++) {
    const size_t *own = own_ids + (x - r0) * own_stride;
    memcpy(tmp.ids, own, sizeof(size_t) * k);
    memcpy(tmp.key, own_key + (x - r0) * own_stride, sizeof(ftype) * k);
    for (size_t j = 0; j < k; j++)
      for (size_t z = 0; z < k; z++)
        tmp.ids[(j + 1) * k + z] = own[j] < n ? graph[own[j] * gstride + z] : n;
    row_distances(&tmp, k, wide, queries + x * d, exclude_self ? x : n, points, n, d);
    sort_and_uniq_row(tmp.ids, tmp.key, wide);
    memcpy(ids_out + (x - r0) * k, tmp.ids, sizeof(size_t) * k);
    memcpy(key_out + (x - r0) * k, tmp.key, sizeof(ftype) * k);
  }
  row_release(&tmp);
}
