/* ann_query.c — query_gpu.  Placeholder until the device path
 * lands: it fails loudly instead of answering from the CPU. */
#include "ann.h"
#include "algg.h"
#include "ann_host.h"

void annh_forget_save(const save_t *save) { (void)save; }

size_t *query_gpu(const save_t *save, const ftype *points, size_t ycnt, const ftype *y,
                  ftype **dists_o) {
  (void)save; (void)points; (void)ycnt; (void)y; (void)dists_o;
  annh_fatal("%s", "query_gpu is not built yet");
  return NULL;
}
