/* ann_dist.c — one process per GPU: sharding of the independent tries across ranks and the
 * exchange steps between them, over NCCL.
 *
 * NCCL is bound at run time (dlopen of libnccl.so.2 — in a Python process that is the copy
 * torch already loaded), so the single-GPU library has no NCCL dependency.  Rendezvous is
 * the caller's: rank 0 asks for a unique id (annb200_dist_unique_id), ships the 128 bytes
 * to the other ranks by any means (bench.py: torch.distributed broadcast), and every rank
 * calls annb200_dist_init(rank, world, id).  From then on precomp_gpu() runs sharded:
 *
 *   rank r owns tries {t : t mod R == r} and the row slice [slice_lo(r), slice_hi(r))
 *   1. each rank uploads ITS slice of the points; all-gather(v) over NVLink
 *   2. means (replicated), hashes / bucket tables / per-try lists for the owned tries
 *   3. all-to-all: every per-try list is cut by row slice and sent to the slice owner, who
 *      now holds all T lists for its rows and merges them exactly like a single GPU
 *   4. all-gather(v) of the merged ids (supercharging reads the neighbours' lists)
 *   5. supercharging of the owned rows; the result rows stay with their owner unless
 *      annb200_dist_gather(1) asks every rank to end up with the full result
 *   save != NULL: the hashes of all tries are broadcast, every rank rebuilds all bucket tables
 *      (cheap) and gathers the full graph, so each rank ends up with a complete save_t
 */
#include <dlfcn.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <cuda_runtime_api.h>

#include "ann_host.h"
#include "gpu_comp.h"

typedef struct ncclComm *ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { NCCL_OK = 0, NCCL_UINT8 = 1 };

static __thread struct {
  void *lib;
  int (*GetUniqueId)(ncclUniqueId *);
  int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  int (*CommDestroy)(ncclComm_t);
  int (*Send)(const void *, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*Recv)(void *, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t);
  int (*Broadcast)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t);
  int (*GroupStart)(void);
  int (*GroupEnd)(void);
  const char *(*GetErrorString)(int);
  ncclComm_t comm;
  int rank, world, gather, hooked;
} D = {.world = 1};

#define NCK(call)                                                                       \
  do {                                                                                  \
    int r_ = (call);                                                                    \
    if (r_ != NCCL_OK) {                                                                \
      fprintf(stderr, "approximatenn_b200: %s failed: %s\n", #call,                     \
              D.GetErrorString ? D.GetErrorString(r_) : "?");                           \
      exit(1);                                                                          \
    }                                                                                   \
  } while (0)

static void load_nccl(void) {
  if (D.lib) return;
  D.lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!D.lib) D.lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
  if (!D.lib) annh_fatal("cannot load NCCL: %s", dlerror());
#define SYM(field, name)                                                                \
  do {                                                                                  \
    *(void **)(&D.field) = dlsym(D.lib, name);                                          \
    if (!D.field) annh_fatal("NCCL symbol missing: %s", name);                          \
  } while (0)
  SYM(GetUniqueId, "ncclGetUniqueId");
  SYM(CommInitRank, "ncclCommInitRank");
  SYM(CommDestroy, "ncclCommDestroy");
  SYM(Send, "ncclSend");
  SYM(Recv, "ncclRecv");
  SYM(AllGather, "ncclAllGather");
  SYM(Broadcast, "ncclBroadcast");
  SYM(GroupStart, "ncclGroupStart");
  SYM(GroupEnd, "ncclGroupEnd");
  SYM(GetErrorString, "ncclGetErrorString");
#undef SYM
}

void annb200_dist_unique_id(char out[128]) {
  load_nccl();
  ncclUniqueId id;
  NCK(D.GetUniqueId(&id));
  memcpy(out, id.internal, 128);
}

void annb200_dist_shutdown(void) {
  if (D.comm) {
    D.CommDestroy(D.comm);
    D.comm = NULL;
  }
  D.world = 1;
  D.rank = 0;
}

/* gpu_cleanup() pops its hook list, so the next annb200_dist_init() has to register again */
static void dist_cleanup_hook(void) {
  annb200_dist_shutdown();
  D.hooked = 0;
}

void annb200_dist_init(int rank, int world, const char id_bytes[128]) {
  if (world < 1 || rank < 0 || rank >= world) annh_fatal("%s", "annb200_dist_init: bad rank/world");
  annb200_dist_shutdown();
  if (world == 1) return;
  gpu_init();
  load_nccl();
  ncclUniqueId id;
  memcpy(id.internal, id_bytes, 128);
  NCK(D.CommInitRank(&D.comm, world, id, rank));
  D.rank = rank;
  D.world = world;
  if (!D.hooked) {
    register_cleanup(dist_cleanup_hook);
    D.hooked = 1;
  }
}

void annb200_dist_gather(int on) { D.gather = on != 0; }

int annh_dist_rank(void) { return D.rank; }
int annh_dist_world(void) { return D.world; }
int annh_dist_gather_results(void) { return D.gather; }

/* ---- the partition (pure functions, also exercised by the CPU tests) ------------------- */

int annb200_dist_try_owner(int t, int world) { return t % world; }

/* rows [lo, hi) owned by `rank`: equal slices of S = ceil(n/world) rounded up to 32 rows (the
 * last ones may be short or empty), so that row arrays can be all-gathered in place with one
 * ncclAllGather of S rows per rank into a buffer padded to world*S rows                      */
static size_t slice_rows(size_t n, int world) {
  size_t s = (n + (size_t)world - 1) / (size_t)world;
  return (s + 31) & ~(size_t)31;
}
void annb200_dist_slice(size_t n, int rank, int world, size_t *lo, size_t *hi) {
  size_t s = slice_rows(n, world);
  size_t a = s * (size_t)rank, b = a + s;
  *lo = a < n ? a : n;
  *hi = b < n ? b : n;
}
/* rows a buffer must hold to be all-gathered in place */
size_t annh_dist_padded_rows(size_t n) { return D.world > 1 ? slice_rows(n, D.world) * (size_t)D.world : n; }

/* entries of try t that fall inside the sorted prefix of the merged row */
int annb200_dist_admit(size_t k, int tries, int t) {
  size_t row = k * (size_t)tries, prefix = row;
  if (row >= 16) {
    prefix = 1;
    while (prefix * 2 <= row) prefix *= 2;
  }
  size_t first = k * (size_t)t;
  if (first >= prefix) return 0;
  return (int)(prefix - first < k ? prefix - first : k);
}

/* ---- exchange steps (all on `stream`) --------------------------------------------------- */

/* every rank contributes its slice of a row array (capacity annh_dist_padded_rows(n) rows);
 * afterwards all ranks hold all rows.  In-place ncclAllGather: NVLink/NVSwitch at full rate. */
void annh_dist_allgather_rows(void *base, size_t n, size_t row_bytes, void *stream) {
  if (D.world == 1) return;
  size_t s = slice_rows(n, D.world);
  NCK(D.AllGather((const char *)base + (size_t)D.rank * s * row_bytes, base, s * row_bytes, NCCL_UINT8,
                  D.comm, (cudaStream_t)stream));
}

/* in-place broadcast of `bytes` at `buf` from rank `root` */
void annh_dist_broadcast(void *buf, size_t bytes, int root, void *stream) {
  if (D.world == 1) return;
  NCK(D.Broadcast(buf, buf, bytes, NCCL_UINT8, root, D.comm, (cudaStream_t)stream));
}

/* all-to-all of per-try lists: `local` holds this rank's tries as [local_try][n][row_bytes];
 * afterwards `slice` holds, for the rows this rank owns, ALL tries: [try][rows][row_bytes]  */
void annh_dist_exchange_lists(const void *local, void *slice, size_t n, size_t row_bytes, int tries,
                              void *stream) {
  size_t mylo, myhi;
  annb200_dist_slice(n, D.rank, D.world, &mylo, &myhi);
  const size_t myrows = myhi - mylo;
  NCK(D.GroupStart());
  for (int t = 0; t < tries; t++) {
    int owner = annb200_dist_try_owner(t, D.world);
    char *dst = (char *)slice + (size_t)t * myrows * row_bytes;
    if (owner == D.rank) {
      const char *mine = (const char *)local + (size_t)(t / D.world) * n * row_bytes;
      for (int p = 0; p < D.world; p++) {
        size_t lo, hi;
        annb200_dist_slice(n, p, D.world, &lo, &hi);
        if (hi == lo) continue;
        if (p == D.rank) {
          if (cudaMemcpyAsync(dst, mine + lo * row_bytes, (hi - lo) * row_bytes, cudaMemcpyDeviceToDevice,
                              (cudaStream_t)stream) != cudaSuccess)
            annh_fatal("%s", "device copy failed in the list exchange");
        } else {
          NCK(D.Send(mine + lo * row_bytes, (hi - lo) * row_bytes, NCCL_UINT8, p, D.comm, (cudaStream_t)stream));
        }
      }
    } else if (myrows) {
      NCK(D.Recv(dst, myrows * row_bytes, NCCL_UINT8, owner, D.comm, (cudaStream_t)stream));
    }
  }
  NCK(D.GroupEnd());
}
