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
/* fills *save with malloc()ed arrays (release with free_save)                             */
int ann_save_read(save_t *save, const char *path);
#ifdef __cplusplus
}
#endif
#endif
