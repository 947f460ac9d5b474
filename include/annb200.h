/* annb200.h — the thin C-ABI between the C host driver (approximatenn_b200/csrc/ann_host.c)
 * and the sm_100a kernels (approximatenn_b200/csrc/annb_kernels.cu).
 *
 * Every function enqueues work on `stream` and returns; all pointers are DEVICE pointers
 * unless named host_*.  One library is built per element type (ftype.h), so `ftype` below
 * is float in libann_b200_f32.so and double in libann_b200_f64.so.  Ids are 32-bit on the
 * device (n < 2^32 - 1); the sentinel "no point" is n, as in the reference.
 *
 * Each entry cites the part of the reference it replaces (paths under /root/reference).
 * The reference launches one OpenCL kernel per arrow of its dataflow through global
 * memory; the stages here are fused differently, but each
 * produces bit-identical values to the reference's pure-C path.
 */
#ifndef ANNB200_H
#define ANNB200_H
#include <stddef.h>
#include "ftype.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef unsigned int annb_u32;
typedef struct CUstream_st *annb_stream;      /* == cudaStream_t */

/* ---- S0: column means --------------------------------------------------------------
 * `levels` (1..4) levels of the reference's stride-halving row sum in one pass (alg.c:122-128;
 * compute.cl:15-31): each level maps rows [0,len) to rows [0,len/2) as
 * dst[x][c] = src[x][c] + src[x+len/2][c] (+ src[len-1][c] for x == 0 when len is odd), with the
 * reference's association order for the first level on the raw points (first != 0) and for
 * the in-place levels.  Writes (len >> levels) rows; dst may alias src when first == 0.
 * Needs len >> (levels-1) >= 2.                                                             */
void annb_fold_rows(const ftype *src, ftype *dst, size_t len, size_t d, int levels, int first,
                    annb_stream stream);
/* mean[c] = acc[c] / n   (compute.cl:36-39) */
void annb_scale_means(const ftype *acc, size_t n, size_t d, ftype *mean, annb_stream stream);

/* ---- S1: centre + random orthogonal transform + sign hash ---------------------------
 * Replaces subtract_off, apply_rotation, apply_permutation, apply_walsh_step,
 * apply_perm_inv, compute_signs (compute.cl:44-122,223-231; driver alg.c:154-183) for
 * `tries` transforms in one pass over the points.  Per try t the host supplies
 *   plane_idx[t][sweep][plane][2]   coordinates of each Givens plane (before sweeps first,
 *                                   then after sweeps), plane_cs the matching (cos, sin)
 *   perm_b[t][d_max]                sub-permutation embedding d -> d_max
 *   pick[t][d_short]                transformed coordinate that feeds hash bit i (MSB first)
 * hash[t][p] receives the d_short-bit bucket id of point p.                              */
typedef struct {
  size_t n, d, d_max, d_short;
  size_t rots_before, rot_len_before, rots_after, rot_len_after;
  int tries;
  ftype inv_sqrt2;                 /* (ftype)(1/sqrt(2.0)), computed on the host           */
  const annb_u32 *plane_idx;
  const ftype *plane_cs;
  const annb_u32 *perm_b;
  const annb_u32 *pick;
  /* the same four tables in HOST memory (optional, all or none): with them the register kernel
   * keeps the per-try tables in constant memory and never copies the tile per try             */
  const annb_u32 *host_plane_idx;
  const ftype *host_plane_cs;
  const annb_u32 *host_perm_b;
  const annb_u32 *host_pick;
} annb_transform_desc;
/* scratch: device workspace of annb_hash_scratch_bytes() bytes (may be NULL when 0)     */
size_t annb_hash_scratch_bytes(const annb_transform_desc *t);
void annb_hash_points(const ftype *points, const ftype *mean, const annb_transform_desc *t,
                      annb_u32 *hash, void *scratch, annb_stream stream);

/* ---- S2: bucket tables (replaces the serial host loop alg.c:252-267) ------------------
 * count[b]   = points hashing to b            (count has `buckets` entries, zeroed here)
 * offset[b]  = exclusive prefix sum, offset[buckets] = n
 * order[..]  = point ids grouped by bucket, DEcreasing id inside a bucket (the reference's
 *              table rows, without the padding)
 * *tmax      = largest bucket (the reference's par_maxes[t])
 * sorted_points[r][:] = points[order[r]][:]   (bucket-contiguous copy for S3)
 * scan_tmp needs annb_scan_tmp_bytes(buckets) bytes.                                      */
size_t annb_scan_tmp_bytes(size_t buckets);
void annb_build_buckets(const annb_u32 *hash, size_t n, size_t buckets,
                        annb_u32 *count, annb_u32 *offset, annb_u32 *order_tmp,
                        annb_u32 *order, annb_u32 *tmax, void *scan_tmp, annb_stream stream);
void annb_gather_rows(const ftype *points, const annb_u32 *order, size_t n, size_t d,
                      ftype *sorted_points, annb_stream stream);
/* padded table as save_t stores it (which_par[t], [buckets][tmax] size_t, pad = n)       */
void annb_export_table(const annb_u32 *offset, const annb_u32 *order, size_t n, size_t buckets,
                       size_t tmax, size_t *table, annb_stream stream);

/* the same table with 32-bit cells (the host widens them while the GPU goes on; the query
 * cache keeps this form), and the largest bucket of a try without building its table (count:
 * `buckets` scratch words) so that a caller can size every table before the tries run        */
void annb_export_table32(const annb_u32 *offset, const annb_u32 *order, size_t n, size_t buckets,
                         size_t tmax, annb_u32 *table, annb_stream stream);
void annb_bucket_max(const annb_u32 *hash, size_t n, size_t buckets, annb_u32 *count,
                     annb_u32 *tmax, annb_stream stream);

/* ---- exact ties ------------------------------------------------------------------------
 * S3, S4 and S5 each run a fast kernel that keeps the k best per row, then redo — with the
 * reference's literal row and sorting network — the rare rows in which two DIFFERENT ids
 * at exactly the same distance could influence the result (the reference's output then
 * depends on where its network leaves equal keys, and it can even keep an id twice).
 * `scratch` holds the per-row flags and the literal rows; *status (device int) is set to 1
 * if the scratch was too small for even one literal row.                                  */

/* ---- S3: candidate distances + per-point k best of one try ----------------------------
 * Replaces compute_which, compute_diffs_squared, add_cols_step, sort_two_step, rdups for
 * the per-try rows (compute.cl:135-217,238-246; alg.c:274-288).  A point's candidates
 * are the slots p < P of the virtual row [bucket h, h^1, h^2, h^4, ...] (tmax slots each),
 * P = 2^floor(log2((d_short+1)*tmax)) — the reference's prefix rule (SURVEY §8.A.3 rule 5).
 * list_ids/list_dist: [n][k], ascending squared distance, (n, +inf) where fewer than k
 * finite candidates exist.  `tmax` is read on the device.
 * `mean` (d entries, may be NULL) only recentres the low-precision copy the screened path
 * brackets distances with; it never enters a reported distance.
 * scale_bits / rows_prepared: see the screened-path helpers below (NULL / 0 = annb_leaf_topk
 * prepares everything itself).
 * scratch: annb_leaf_scratch_bytes(n, d, d_short, k) bytes.                                */
size_t annb_leaf_scratch_bytes(size_t n, size_t d, size_t d_short, size_t k);
void annb_leaf_topk(const ftype *sorted_points, const ftype *mean, const annb_u32 *order,
                    const annb_u32 *offset, const annb_u32 *hash, const annb_u32 *tmax, size_t n,
                    size_t d, size_t d_short, size_t k, annb_u32 *list_ids, ftype *list_dist,
                    void *scratch, int *status, const unsigned *scale_bits, int rows_prepared,
                    annb_stream stream);

/* Cutoff carried from try to try (screened path only).  The merge of the per-try lists keeps the
 * k smallest distinct ids, so an entry farther than the k-th best already found can never be
 * reported: with `cutoff` (one value per point, original order) a later try's list holds every
 * candidate at or below cutoff[x] at its position in the full list and may hold (n, +inf)
 * beyond.  annb_cutoff_update folds a finished list into the running k smallest distinct
 * distance values `run` ([n][k], first != 0: starts it) and writes cutoff[x] = their largest
 * (+inf while fewer than k are known).  Only valid where the merged row is sorted as a whole
 * (k*tries a power of two >= 16; alg.c:139 prefix rule) — the caller decides.  k <= 16.
 * Replaces nothing in the reference: it removes work whose result compute.cl:181-217 discards. */
int annb_cutoff_applies(size_t d, size_t d_short, size_t k);
void annb_leaf_cutoff_mode(int on);
void annb_cutoff_update(const ftype *new_dist, ftype *run, ftype *cutoff, size_t n, size_t k, int first,
                        annb_stream stream);
void annb_leaf_topk_cut(const ftype *sorted_points, const ftype *mean, const annb_u32 *order,
                        const annb_u32 *offset, const annb_u32 *hash, const annb_u32 *tmax, size_t n,
                        size_t d, size_t d_short, size_t k, annb_u32 *list_ids, ftype *list_dist,
                        void *scratch, int *status, const unsigned *scale_bits, int rows_prepared,
                        const ftype *cutoff, annb_stream stream);

/* Screened S3 path (float rows, d in {16, 32, 64}, k <= 16): fp16 tensor-core brackets of every
 * candidate distance decide which candidates go through the exact tree; the lists are the same
 * as the tiled kernel's, bit for bit.
 *   annb_screen_applies     1 if annb_leaf_topk will take it for this shape
 *   annb_screen_scale       once per point set: the power-of-two scale of the fp16 copy
 *                           (device word `scale_bits`), from max |x - mean|
 *   annb_gather_rows_screen annb_gather_rows fused with the per-try preparation (fp16 copy and
 *                           norms, written into the tail of `leaf_scratch`); pass the same
 *                           scale_bits and rows_prepared = 1 to annb_leaf_topk afterwards      */
int annb_screen_applies(size_t d, size_t d_short, size_t k);
void annb_screen_scale(const ftype *points, const ftype *mean, size_t n, size_t d,
                       unsigned *scale_bits, annb_stream stream);
void annb_gather_rows_screen(const ftype *points, const annb_u32 *order, size_t n, size_t d,
                             const ftype *mean, const unsigned *scale_bits, ftype *sorted_points,
                             void *leaf_scratch, annb_stream stream);

/* ---- S4: merge of the per-try lists (first sort_and_uniq of det_results, alg.c:312) ----
 * lists: [n_lists][n][k]; admit[i] = number of leading entries of list i that fall inside
 * the sorted prefix (k, or fewer for the try that straddles 2^floor(log2(k*tries))).
 * corner_list/corner_pos: the list entry sitting in the first slot OUTSIDE the prefix
 * (or corner_list < 0 when the row length is a power of two) — see DESIGN.md "prefix
 * corner".  merged: [n][k]; n = rows merged here, sentinel_n = size of the whole point set
 * (they differ when a rank merges only its row slice).  If merged_in != NULL it is treated as one more, fully
 * admitted list (running merge).  every_list != 0 says the call sees ALL lists of the row;
 * only then can tied rows be redone literally.  Rows shorter than 16
 * slots (k*n_lists < 16) are always done literally.
 * scratch: at least n + 512 bytes plus room for literal rows (64 MB is plenty).           */
void annb_merge_lists(const annb_u32 *lists_ids, const ftype *lists_dist, int n_lists,
                      const int *host_admit, int corner_list, int corner_pos,
                      const annb_u32 *merged_in_ids, const ftype *merged_in_dist,
                      size_t n, size_t sentinel_n, size_t k, int every_list, annb_u32 *merged_ids,
                      ftype *merged_dist, void *scratch, size_t scratch_bytes, int *status,
                      annb_stream stream);

/* ---- S5: supercharging (supercharge + compdists + second sort_and_uniq, alg.c:313-335) --
 * For each query row x in [row_begin, row_end): candidates = own list ++ the lists of its
 * neighbours (graph rows), prefix 2^floor(log2(k(k+1))); distances from queries[x] to
 * points[id]; out rows are relative to row_begin (32-bit ids; the host widens them).  exclude_self: the reference excludes
 * id == x only when the query set IS the point set (compute.cl:145).
 * graph may be the merged ids themselves (precomp) or save->graph (query).
 * scratch: at least (row_end-row_begin) + 512 bytes plus room for literal rows.           */
/* opts (may be NULL): extras of the precomp path, where the queries ARE the points.
 *   points16/nrm/scale_bits  fp16 copy of the points in original order, its per-row (norm bound,
 *                        squared fp16 norm) pairs (annb_screen_prep_points) and the scale word:
 *                        candidates are bracketed from them and only those that can reach the
 *                        row's k best are measured exactly (float, d in {16,32,64,128}, k <= 32;
 *                        same rows bit for bit; ANN_B200_S5_SCREEN=0 switches it off)
 *   row_perm/perm_base   the order in which the rows are worked on (annb_locality_order):
 *                        positions [row_begin, row_end) must map onto exactly those rows;
 *                        NULL = ascending                                                        */
typedef struct {
  const void *points16;
  const void *nrm;
  const unsigned *scale_bits;
  const annb_u32 *row_perm;   /* row worked on at position p of the call: perm_base + row_perm[p - perm_base] */
  size_t perm_base;
} annb_supercharge_opts;
void annb_supercharge(const ftype *queries, const ftype *points, const annb_u32 *own_ids,
                      const ftype *own_dist, const annb_u32 *graph, size_t n, size_t d,
                      size_t k, size_t row_begin, size_t row_end, int exclude_self,
                      annb_u32 *out_ids, ftype *out_dist, void *scratch, size_t scratch_bytes,
                      int *status, const annb_supercharge_opts *opts, annb_stream stream);
/* Locality order for the supercharge (the lever SURVEY 8.D names for its gather term): relative
 * rows [0, rows) of the slice starting at row_lo, grouped by chunk (chunk c = relative rows
 * [chunk_lo[c], chunk_lo[c+1]), host array of chunks+1 entries) and inside a chunk by the leading
 * bits of `hash` (one try's bucket ids, indexed by global row).  perm: rows entries.             */
size_t annb_locality_scratch_bytes(size_t rows, size_t d_short, int chunks);
void annb_locality_order(const annb_u32 *hash, size_t row_lo, size_t rows, size_t d_short, int chunks,
                         const size_t *chunk_lo, void *scratch, annb_u32 *perm, annb_stream stream);

/* 1 (default) / 0: the screened supercharge and the thread-per-point merge can be switched off at
 * run time (ANN_B200_S5_SCREEN=0, ANN_B200_THREAD_MERGE=0 do the same); results are identical  */
void annb_supercharge_screen_mode(int on);
void annb_merge_thread_mode(int on);
/* 1 if annb_supercharge will use points16 for this shape                                       */
int annb_supercharge_screen_applies(size_t d, size_t k);
/* fp16 copy of the points in ORIGINAL order for the screened supercharge: c' = fp16((x - mean)
 * * scale), scale from the word annb_screen_scale() wrote.  points16: n*d*2 bytes, nrm: n*8.    */
void annb_screen_prep_points(const ftype *points, const ftype *mean, size_t n, size_t d,
                             const unsigned *scale_bits, void *points16, void *nrm, annb_stream stream);
/* candidates the screened supercharge bracketed [0] and measured exactly [1] since the last reset */
void annb_supercharge_screen_stats(unsigned long long out[2], int reset);

/* ---- query path (alg.c:438-519) ----------------------------------------------------------
 * annb_query_hash: prods + add_up_cols + compute_signs (compute.cl:268-275,160-167,223-231):
 *   sign[x*tries + t] = sign bits of (y_x - mean) . bases[t][i], i < d_short.
 * annb_query_rows: shufcomp + compdists + first sort_and_uniq of det_results: candidate row
 *   of query x = for each try i the tables rows of h_i ^ flip_f, h_i = sign[i*ycnt + x] (the
 *   reference's own, transposed, read of the sign buffer), prefix 2^floor(log2(len));
 *   tables[i] is a HOST array entry holding a DEVICE pointer to the 32-bit padded table
 *   [2^d_short][par_maxes[i]]; par_maxes is a host array.  list_*: [ycnt][k].
 * annb_narrow_ids: size_t -> 32-bit ids on the device (host-format tables / graph).        */
void annb_query_hash(const ftype *y, const ftype *mean, const ftype *bases, size_t ycnt, size_t d,
                     size_t d_short, int tries, annb_u32 *sign, annb_stream stream);
/* screen (may be NULL; float, d in {16,32,64,128}, k <= 32): the fp16 copy of the indexed points
 * with its (norm bound, squared norm) pairs and scale word (annb_screen_scale +
 * annb_screen_prep_points) and the column means; candidates are then bracketed against the
 * running k-th best and only the survivors are measured exactly — same lists, bit for bit
 * (ANN_B200_QUERY_SCREEN=0 switches it off)                                                  */
typedef struct {
  const void *points16;
  const void *nrm;
  const ftype *mean;
  const unsigned *scale_bits;
} annb_query_screen;
int annb_query_screen_applies(size_t d, size_t k);
void annb_query_rows(const ftype *y, const ftype *points, const annb_u32 *const *tables,
                     const size_t *par_maxes, int tries, const annb_u32 *sign, size_t n,
                     size_t ycnt, size_t d, size_t d_short, size_t k, int exclude_self,
                     annb_u32 *list_ids, ftype *list_dist, void *scratch, size_t scratch_bytes,
                     int *status, const annb_query_screen *screen, annb_stream stream);
void annb_narrow_ids(const size_t *src, size_t count, annb_u32 *dst, annb_stream stream);

/* (point, real candidate) pairs whose distance S3 evaluated since the last reset (synchronous);
 * 3*d floating-point operations each (compute.cl:147-166) — the flop count of the leaf stage     */
unsigned long long annb_leaf_pairs(int reset);
/* of those, the pairs the screened path sent through the exact tree (0 when it is off)          */
unsigned long long annb_leaf_exact_pairs(int reset);
/* buckets the screened path handed to the tiled kernel (tables or survivor lists too small)     */
unsigned long long annb_leaf_overflow_buckets(int reset);
/* 1 (default): float rows with d in {16, 32, 64} and k <= 16 take the screened S3 path;
 * 0: the tiled kernel only.  The lists are identical either way (tests compare them).          */
void annb_leaf_screen_mode(int on);

/* rows redone by the literal kernels since the last reset: [0] S3, [1] S4, [2] S5 (synchronous)  */
void annb_literal_rows(unsigned long long out[3], int reset);

/* FP32-pipe probe (annb_probe.cu): TFLOP/s of dependent-chain streams of
 *   mode 0 FFMA (2 flops/instruction)   1 FMUL+FADD (separately rounded, 1 flop/instruction)
 *   mode 2 FFMA2 (packed, 4 flops)      3 FMUL2+FADD2 (packed, separately rounded, 2 flops)
 * best of `reps` launches, timed with CUDA events on `stream` (synchronous).  These are the
 * measured denominators of the S3 roofline (SURVEY 8.D asks for a measured FP32 figure).       */
double annb_probe_fp32(int mode, int reps, annb_stream stream);

/* number of kernels launched through this layer since the last reset (bench.py reports it) */
unsigned long annb_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif
