/* ann_save_io.c — save_t <-> file (include/annb200_io.h). */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "annb200_io.h"

static const char MAGIC[8] = {'A', 'N', 'N', 'B', '2', 'S', '0', '1'};

static int put_ids(FILE *f, const size_t *ids, size_t count) {
  uint32_t buf[4096];
  for (size_t i = 0; i < count;) {
    size_t m = count - i < 4096 ? count - i : 4096;
    for (size_t j = 0; j < m; j++) buf[j] = (uint32_t)ids[i + j];
    if (fwrite(buf, 4, m, f) != m) return -1;
    i += m;
  }
  return 0;
}

static size_t *get_ids(FILE *f, size_t count) {
  size_t *ids = malloc(sizeof(size_t) * (count ? count : 1));
  uint32_t buf[4096];
  if (!ids) return NULL;
  for (size_t i = 0; i < count;) {
    size_t m = count - i < 4096 ? count - i : 4096;
    if (fread(buf, 4, m, f) != m) { free(ids); return NULL; }
    for (size_t j = 0; j < m; j++) ids[i + j] = buf[j];
    i += m;
  }
  return ids;
}

int ann_save_write(const save_t *s, const char *path) {
  if (s->n >= 0xFFFFFFFFull) return -1;
  FILE *f = fopen(path, "wb");
  if (!f) return -1;
  uint32_t w = sizeof(ftype), tries = (uint32_t)s->tries;
  uint64_t dims[4] = {s->n, s->k, s->d_short, s->d_long};
  int ok = fwrite(MAGIC, 1, 8, f) == 8 && fwrite(&w, 4, 1, f) == 1 && fwrite(&tries, 4, 1, f) == 1 &&
           fwrite(dims, 8, 4, f) == 4;
  for (uint32_t t = 0; ok && t < tries; t++) {
    uint64_t pm = s->par_maxes[t];
    ok = fwrite(&pm, 8, 1, f) == 1;
  }
  ok = ok && fwrite(s->row_means, sizeof(ftype), s->d_long, f) == s->d_long;
  size_t nb = (size_t)tries * s->d_short * s->d_long;
  ok = ok && fwrite(s->bases, sizeof(ftype), nb, f) == nb;
  ok = ok && put_ids(f, s->graph, s->n * s->k) == 0;
  for (uint32_t t = 0; ok && t < tries; t++)
    ok = put_ids(f, s->which_par[t], s->par_maxes[t] << s->d_short) == 0;
  if (fclose(f) != 0) ok = 0;
  return ok ? 0 : -1;
}

/* ids are point numbers or the sentinel n (pads, alg.c:261-266) */
static int ids_in_range(const size_t *ids, size_t count, size_t n) {
  for (size_t i = 0; i < count; i++)
    if (ids[i] > n) return 0;
  return 1;
}

int ann_save_read(save_t *s, const char *path) {
  FILE *f = fopen(path, "rb");
  if (!f) return -1;
  char magic[8];
  uint32_t w = 0, tries = 0;
  uint64_t dims[4];
  memset(s, 0, sizeof *s);
  if (fread(magic, 1, 8, f) != 8 || memcmp(magic, MAGIC, 8) || fread(&w, 4, 1, f) != 1 ||
      w != sizeof(ftype) || fread(&tries, 4, 1, f) != 1 || fread(dims, 8, 4, f) != 4) {
    fclose(f);
    return -1;
  }
  /* the header is not trusted: the limits are the library's own (ann_host.c validate()), the
   * products below cannot overflow inside them, and every id read later is range-checked      */
  if (tries < 1 || tries > 64 || dims[0] < 2 || dims[0] >= 0xFFFFFFFFull || dims[1] < 1 || dims[1] > 256 ||
      dims[1] >= dims[0] || dims[2] > 28 || dims[3] < 1 || dims[3] > ((uint64_t)1 << 20)) {
    fclose(f);
    return -1;
  }
  s->tries = (int)tries;
  s->n = dims[0]; s->k = dims[1]; s->d_short = dims[2]; s->d_long = dims[3];
  s->par_maxes = malloc(sizeof(size_t) * (tries ? tries : 1));
  s->which_par = calloc(tries ? tries : 1, sizeof(size_t *));
  s->row_means = malloc(sizeof(ftype) * (s->d_long ? s->d_long : 1));
  size_t nb = (size_t)tries * s->d_short * s->d_long;
  s->bases = malloc(sizeof(ftype) * (nb ? nb : 1));
  int ok = s->par_maxes && s->which_par && s->row_means && s->bases;
  for (uint32_t t = 0; ok && t < tries; t++) {
    uint64_t pm;
    ok = fread(&pm, 8, 1, f) == 1 && pm <= s->n;                /* a bucket holds at most n points */
    s->par_maxes[t] = ok ? pm : 0;
  }
  ok = ok && fread(s->row_means, sizeof(ftype), s->d_long, f) == s->d_long;
  ok = ok && fread(s->bases, sizeof(ftype), nb, f) == nb;
  if (ok) ok = (s->graph = get_ids(f, s->n * s->k)) != NULL && ids_in_range(s->graph, s->n * s->k, s->n);
  for (uint32_t t = 0; ok && t < tries; t++)
    ok = (s->which_par[t] = get_ids(f, s->par_maxes[t] << s->d_short)) != NULL &&
         ids_in_range(s->which_par[t], s->par_maxes[t] << s->d_short, s->n);
  fclose(f);
  if (!ok) {
    for (uint32_t t = 0; t < tries && s->which_par; t++) free(s->which_par[t]);
    free(s->which_par); free(s->par_maxes); free(s->graph); free(s->row_means); free(s->bases);
    memset(s, 0, sizeof *s);
    return -1;
  }
  return 0;
}
