/* ann_host.h — internals shared by the C host files (not part of the public API). */
#ifndef ANN_HOST_H
#define ANN_HOST_H
#include <stddef.h>
#include "ann.h"
#include "annb200.h"

#ifdef __cplusplus
extern "C" {
#endif

/* prints "approximatenn_b200: <fmt % detail>" to stderr and exits(1): the reference's error
 * behaviour (gpu_comp.c:15-19), the API has no error channel                              */
void annh_fatal(const char *fmt, const char *detail);

/* d_short / d_max of a problem (alg.c:347-357) */
void annh_params(size_t n, size_t k, size_t d, size_t *d_short, size_t *d_max);

/* the library's CUDA stream (cudaStream_t) and device, after gpu_init() */
void *annh_stream(void);
int annh_device(void);

/* per-call device arena (one allocation reused across calls, grown on demand)           */
void annh_arena_reserve(size_t bytes);
void *annh_arena_take(size_t bytes);

/* Device time per stage of the last precomp_gpu call, summed over tries / chunks, measured
 * with CUDA events on the library stream when timing is on (annh_set_timing(1) or
 * ANN_B200_TIMING=1):
 *   0 upload (H2D + point all-gather)  1 means  2 hash  3 bucket tables + sorted copy
 *   4 leaf lists (S3 incl. literal redo)  5 list exchange (sharded only)  6 merge
 *   7 supercharge (incl. literal redo)  8 total from first to last event                    */
#define ANNH_STAGES 9
typedef struct { float ms[ANNH_STAGES]; } annh_stage_times;
const annh_stage_times *annh_last_times(void);
void annh_set_timing(int on);

/* result egress (ann_results.c): malloc()ed result arrays filled chunk by chunk through a
 * pinned staging buffer while later chunks are still being computed                        */
typedef struct annh_egress annh_egress;
/* into != NULL: the rows of this call go to caller-provided arrays (ids, and dist / ids2 when
 * wanted) instead of fresh malloc()s — the single-process multi-GPU mode hands every device
 * its slice of ONE result array this way                                                     */
typedef struct { size_t *ids; ftype *dist; size_t *ids2; } annh_egress_target;
annh_egress *annh_egress_begin(size_t rows, size_t k, int want_dist, int want_second_ids, int device,
                               const annh_egress_target *into);
int annh_egress_chunks(const annh_egress *e);          /* how many row chunks to produce    */
void annh_egress_chunk(annh_egress *e, size_t r0, size_t r1, const void *dev_ids_u32,
                       const void *dev_dist, void *producer_stream);
size_t *annh_egress_end(annh_egress *e, ftype **dists_o, size_t **second_ids_o);
void annh_prefault(void *ptr, size_t bytes);   /* touch the pages of a fresh allocation, 4 threads */
void annh_egress_release(void);

/* bucket tables of save_t leave as 32-bit cells and are widened by host threads (ann_results.c) */
typedef struct annh_tables annh_tables;
annh_tables *annh_tables_begin(int tries, const size_t *cells, size_t **host_tables, int device);
void annh_tables_submit(annh_tables *tb, int t, const void *dev_table32, void *producer_stream);
void annh_tables_end(annh_tables *tb);

/* sharded execution (ann_dist.c); world == 1 unless annb200_dist_init() was called          */
int annh_dist_rank(void);
int annh_dist_world(void);
int annh_dist_gather_results(void);
size_t annh_dist_padded_rows(size_t n);      /* capacity (rows) of arrays that get all-gathered */
void annh_dist_allgather_rows(void *base, size_t n, size_t row_bytes, void *stream);
void annh_dist_broadcast(void *buf, size_t bytes, int root, void *stream);
void annh_dist_exchange_lists(const void *local, void *slice, size_t n, size_t row_bytes, int tries,
                              void *stream);

/* host -> device upload of caller memory (ann_ingest.c): staged through pinned slots by a few
 * host threads when the source is pageable                                                 */
void annh_ingest(void *dst, const void *src, size_t bytes, void *consumer_stream, int device);
void annh_ingest_release(void);

/* precomp_gpu(save != NULL) hands its device-resident data to the query cache (ann_query.c) */
int annh_index_adopt_begin(size_t n, size_t k, size_t d_short, size_t d, size_t tries);
annb_u32 *annh_index_table_buffer(int t, size_t cells);   /* device buffer of try t's 32-bit table */
void annh_index_adopt_finish(const save_t *s, const ftype *host_points, const ftype *dev_points,
                             const ftype *dev_mean, const annb_u32 *dev_graph);

/* drops any device-resident copy of `save` kept for query_gpu (called by free_save)      */
void annh_forget_save(const save_t *save);

/* ---- single-process multi-GPU mode (ann_multi.c): ANN_B200_GPUS=N ---------------------------
 * One worker thread per device; all library state is per thread, so a worker is exactly what a
 * rank process is in the torchrun mode.  The public entry points forward to the workers.      */
typedef struct {
  void *transforms;            /* the tries' transforms, drawn once by the caller (annh_draw_transforms) */
  int force_single;            /* run the whole problem on this device although a world exists      */
  annh_egress_target out;      /* this rank's slice of the shared result arrays (ids == NULL: malloc) */
} annh_call_ctx;
void *annh_draw_transforms(size_t n, size_t k, size_t d, int tries, size_t rots_before, size_t rot_len_before,
                           size_t rots_after, size_t rot_len_after);
void annh_free_transforms(void *tf, int tries);
size_t *annh_precomp_impl(size_t n, size_t k, size_t d, const ftype *points, int tries, size_t rots_before,
                          size_t rot_len_before, size_t rots_after, size_t rot_len_after, save_t *save,
                          ftype **dists_o, const annh_call_ctx *ctx);
size_t *annh_query_impl(const save_t *save, const ftype *points, size_t ycnt, const ftype *y, ftype **dists_o);
void annh_forget_save_impl(const save_t *save);
void annh_gpu_init_on(int device);        /* gpu_init() of the calling thread on a given device */
void annh_gpu_cleanup_impl(void);
int annh_multi_gpus(void);                /* N of ANN_B200_GPUS (clamped to the devices present), 1 = off */
int annh_multi_in_worker(void);
void annh_multi_start(void);
void annh_multi_stop(void);
size_t *annh_multi_precomp(size_t n, size_t k, size_t d, const ftype *points, int tries, size_t rots_before,
                           size_t rot_len_before, size_t rots_after, size_t rot_len_after, save_t *save,
                           ftype **dists_o);
size_t *annh_multi_query(const save_t *save, const ftype *points, size_t ycnt, const ftype *y, ftype **dists_o);
void annh_multi_forget(const save_t *save);

#ifdef __cplusplus
}
#endif
#endif
