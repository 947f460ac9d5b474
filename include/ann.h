/* Public C API of the randomized all-points k-nearest-neighbour path.
 *
 * This header is layout- and signature-compatible with the reference's ann.h
 * (/root/reference/ann.h:8-12 save_t, :46-49 precomp, :61-62 query, :65 free_save),
 * so the reference's own programs (time_results.c, compare_results.c,
 * test_correctness.c) compile against it unchanged.
 */
#ifndef ANN
#define ANN
#include <stddef.h>
#include "ftype.h"

#ifdef __cplusplus
extern "C" {
#endif

/* Index kept for later queries.  Same field order and types as the reference
 * (ann.h:8-12); every array is malloc()ed by the library and released by
 * free_save(); the struct itself belongs to the caller.
 *   which_par[t]  : [2^d_short][par_maxes[t]] bucket table of try t, each row holding the
 *                   point ids of that bucket in DEcreasing order, padded with n
 *   graph         : [n][k] neighbour ids (ascending squared distance)
 *   row_means     : [d_long] column means of the point set
 *   bases         : [tries][d_short][d_long] rows of the hashing projections          */
typedef struct {
  int tries;
  size_t n, k, d_short, d_long, **which_par, *par_maxes, *graph;
  ftype *row_means, *bases;
} save_t;

/* All-points approximate kNN.
 *   points : [n][d] row-major host array
 *   tries  : number of independent random transforms (hash tables)
 *   rots_before/rot_len_before : number of Givens sweeps before the Walsh-Hadamard step
 *                                and disjoint planes per sweep (2*rot_len_before <= d)
 *   rots_after/rot_len_after   : same, after the Walsh-Hadamard step (planes drawn
 *                                among the first d_short coordinates)
 *   save   : NULL, or receives the query index
 *   dists  : NULL, or *dists receives a malloc()ed [n][k] array of SQUARED distances
 *   use_cpu: nonzero selects the reference's single-core C path when that is linked in
 * Returns a malloc()ed [n][k] array of neighbour ids; the caller frees it.
 * The transform is drawn from libc random(): srandom(seed) before the call makes it
 * reproducible (compare_results.c:123-130 relies on this).                              */
extern size_t *precomp(size_t n, size_t k, size_t d, const ftype *points,
                       int tries, size_t rots_before, size_t rot_len_before,
                       size_t rots_after, size_t rot_len_after, save_t *save,
                       ftype **dists, char use_cpu);

/* kNN of ycnt query vectors y[ycnt][d_long] against the indexed point set. */
extern size_t *query(const save_t *save, const ftype *points,
                     size_t ycnt, const ftype *y, ftype **dists, char use_cpu);

/* Releases the arrays inside *save (not the struct). */
extern void free_save(save_t *save);

#ifdef __cplusplus
}
#endif
#endif
