/* annb200_io.h — on-disk form of save_t.  The reference keeps the
 * index in memory only (ann.h:8-12 "treat as opaque", no serialisation anywhere); this adds
 * "precomp once, query many processes later".  Not part of the drop-in surface.
 *
 * File layout (little endian): 8-byte magic "ANNB2S01", u32 sizeof(ftype), u32 tries,
 * u64 n, k, d_short, d_long, u64 par_maxes[tries], ftype row_means[d_long],
 * ftype bases[tries*d_short*d_long], u32 graph[n*k], then per try u32 table[2^d_short * par_max].
 * Ids are stored as 32 bits (n < 2^32 - 1, the library's own limit).                        */
#ifndef ANNB200_IO_H
#define ANNB200_IO_H
#include "ann.h"
#ifdef __cplusplus
extern "C" {
#endif
/* 0 on success, -1 on I/O error or a file that does not match this build's ftype          */
int ann_save_write(const save_t *save, const char *path);
/* fills *save with malloc()ed arrays (release with free_save); -1 also for a header outside the
 * library's limits, a truncated file, or ids beyond n                                      */
int ann_save_read(save_t *save, const char *path);

/* query_gpu keeps the index of the last save_t on the device, keyed on (save->graph, points,
 * shape) plus a fingerprint of SAMPLED graph / table / point cells — the reference re-ships
 * everything per call (alg.c:464-508).  A caller that rewrites `points` or the save_t arrays IN
 * PLACE (same addresses, same shape) must call this (or free_save) before the next query_gpu;
 * ANN_B200_QUERY_CACHE=0 disables the cache altogether.                                    */
void annb200_query_cache_invalidate(void);
#ifdef __cplusplus
}
#endif
#endif
