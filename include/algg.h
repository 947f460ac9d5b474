/* Entry points of the accelerator backend — the symbols the reference's ann.c binds
 * when use_cpu == 0 (/root/reference/algg.h:5-11, called from ann.c:11 and ann.c:21).
 * libann_b200_f32.so / libann_b200_f64.so export exactly these. */
#ifndef ALGGPU
#define ALGGPU
#include "ann.h"
#ifdef __cplusplus
extern "C" {
#endif
extern size_t *query_gpu(const save_t *save, const ftype *points,
                         size_t ycnt, const ftype *y, ftype **dists_o);
extern size_t *precomp_gpu(size_t n, size_t k, size_t d, const ftype *points,
                           int tries, size_t rots_before,
                           size_t rot_len_before, size_t rots_after,
                           size_t rot_len_after, save_t *save,
                           ftype **dists_o);
#ifdef __cplusplus
}
#endif
#endif
