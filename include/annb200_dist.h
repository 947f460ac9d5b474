/* annb200_dist.h — multi-GPU control of libann_b200_*.so: one process per GPU, tries sharded
 * across ranks, lists exchanged over NCCL.  The reference has nothing of
 * the kind (one OpenCL device, alg.c:358-359); this is new surface, kept outside ann.h so
 * that the drop-in API stays byte-for-byte the reference's.
 *
 *   rank 0:      annb200_dist_unique_id(id)       -> ship the 128 bytes to every rank
 *   every rank:  annb200_dist_init(rank, world, id)
 *                srandom(seed); precomp_gpu(...)   identical arguments, identical seed, identical
 *                                                 points on every rank
 *   result:      by default rank r gets the rows [lo, hi) of annb200_dist_slice(n, r, world, ...)
 *                (a malloc()ed [hi-lo][k] array); annb200_dist_gather(1) makes every rank
 *                return all n rows instead.
 */
#ifndef ANNB200_DIST_H
#define ANNB200_DIST_H
#include <stddef.h>
#ifdef __cplusplus
extern "C" {
#endif
void annb200_dist_unique_id(char out[128]);
void annb200_dist_init(int rank, int world, const char id[128]);
void annb200_dist_shutdown(void);
void annb200_dist_gather(int on);
/* the partition, as pure functions */
int annb200_dist_try_owner(int t, int world);
void annb200_dist_slice(size_t n, int rank, int world, size_t *lo, size_t *hi);
int annb200_dist_admit(size_t k, int tries, int t);
#ifdef __cplusplus
}
#endif
#endif
