// annb_common.cuh — device helpers shared by the kernel files.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "annb200.h"

typedef ftype FT;
typedef annb_u32 u32;
#define FULL 0xffffffffu

extern unsigned long annb_g_launches;
// rows redone by the literal kernels: per-file device counters behind annb_literal_rows()
unsigned long long annb_leaf_literal_count(int reset);
void annb_finish_literal_counts(unsigned long long out[2], int reset);

// ANN_B200_DEBUG_SYNC=1: wait for every kernel right after its launch and name it on stderr
// (finds the kernel that faults or never returns; never set in timed runs)
static inline int annb_debug_sync() {
  static int v = -1;
  if (v < 0) { const char *e = getenv("ANN_B200_DEBUG_SYNC"); v = e && *e && *e != '0'; }
  return v;
}

#define LAUNCH_CHECK(what)                                                              \
  do {                                                                                  \
    __sync_fetch_and_add(&annb_g_launches, 1ul);                                        \
    cudaError_t e_ = cudaGetLastError();                                                \
    if (e_ == cudaSuccess && annb_debug_sync()) {                                       \
      fprintf(stderr, "[debug-sync] %s launched\n", what);                              \
      e_ = cudaDeviceSynchronize();                                                     \
      fprintf(stderr, "[debug-sync] %s finished: %s\n", what, cudaGetErrorString(e_));  \
    }                                                                                   \
    if (e_ != cudaSuccess) {                                                            \
      fprintf(stderr, "approximatenn_b200: launch of %s failed: %s\n", what,            \
              cudaGetErrorString(e_));                                                  \
      exit(1);                                                                          \
    }                                                                                   \
  } while (0)

// runtime calls made by the launch layer itself (memsets, attribute changes): fatal like launches
#define RT_CHECK(call)                                                                  \
  do {                                                                                  \
    cudaError_t e_ = (call);                                                            \
    if (e_ != cudaSuccess) {                                                            \
      fprintf(stderr, "approximatenn_b200: %s failed: %s\n", #call, cudaGetErrorString(e_)); \
      exit(1);                                                                          \
    }                                                                                   \
  } while (0)

// sqrt(kappa) of the fp16 distance brackets shared by the screened S3 and S5 kernels
// (kappa = 0.00104, DESIGN.md "screened leaf")
static constexpr float SCREEN_SQRT_KAPPA = 0.03226f;

static inline void fatal_config(const char *what) {
  fprintf(stderr, "approximatenn_b200: unsupported configuration: %s\n", what);
  exit(1);
}

__device__ __forceinline__ FT ft_inf() {
#ifdef USE_FLOAT
  return __int_as_float(0x7f800000);
#else
  return __longlong_as_double(0x7ff0000000000000LL);
#endif
}

__device__ __forceinline__ u32 sign_bit(FT v) {
#ifdef USE_FLOAT
  return ((u32)__float_as_int(v)) >> 31;
#else
  return (u32)(((unsigned long long)__double_as_longlong(v)) >> 63);
#endif
}

static inline unsigned grid_for(size_t items, unsigned block) {
  size_t g = (items + block - 1) / block;
  if (g == 0) g = 1;
  return (unsigned)g;
}

// =====================================================================================
// exact squared distance, warp-cooperative
// =====================================================================================
// The reference sums the d squared differences with a stride-halving tree
// (compute.cl:160-167): for l = d, d/2, ...: m[z] += m[z + l/2] (+ m[l-1] for z == 0 when l
// is odd).  Squares are never -0.0, so the "+ 0" the reference adds for even l is a no-op
// and is skipped.
//
// WarpRow<E>: d = 32*E (a power of two >= 32) — lane holds coordinates lane + 32*s, the
// first log2(E) tree levels are lane-local, the last five are xor-shuffles (a + b == b + a
// bit for bit, so both partners compute the same value).
// d == 16 uses E = 1 with coordinates >= 16 reading as zero distance contributions kept
// out of the tree (only xor 8,4,2,1 are applied).
// Any other d goes through a shared-memory tree (generic_sqdist).

template <int E>
struct WarpRow {
  FT x[E];
  __device__ __forceinline__ void load(const FT *__restrict__ row, int lane, int d) {
#pragma unroll
    for (int s = 0; s < E; s++) x[s] = (lane + 32 * s < d) ? row[lane + 32 * s] : (FT)0;
  }
};

template <int E>
__device__ __forceinline__ FT warp_sqdist(const WarpRow<E> &q, const WarpRow<E> &c, int d) {
  FT m[E];
#pragma unroll
  for (int s = 0; s < E; s++) {
    FT df = q.x[s] - c.x[s];
    m[s] = df * df;
  }
#pragma unroll
  for (int h = E / 2; h >= 1; h >>= 1)
#pragma unroll
    for (int s = 0; s < h; s++) m[s] = m[s] + m[s + h];
  FT v = m[0];
  if (d >= 32) v = v + __shfl_xor_sync(FULL, v, 16);
  v = v + __shfl_xor_sync(FULL, v, 8);
  v = v + __shfl_xor_sync(FULL, v, 4);
  v = v + __shfl_xor_sync(FULL, v, 2);
  v = v + __shfl_xor_sync(FULL, v, 1);
  return v;                                  // every lane (lane < 16 when d == 16) holds the sum
}

// Generic d: tmp is a per-warp shared buffer of d entries.  Result is warp-uniform.
__device__ __forceinline__ FT generic_sqdist(const FT *__restrict__ q, const FT *__restrict__ c,
                                             int d, FT *tmp, int lane) {
  for (int z = lane; z < d; z += 32) {
    FT df = q[z] - c[z];
    tmp[z] = df * df;
  }
  __syncwarp();
  for (int l = d; l >> 1; l >>= 1) {
    int h = l >> 1;
    for (int z = lane; z < h; z += 32) {
      FT add = tmp[z + h];
      if (z == 0 && (l & 1)) add = add + tmp[l - 1];
      tmp[z] = tmp[z] + add;
    }
    __syncwarp();
  }
  FT v = tmp[0];
  __syncwarp();
  return v;
}

// dispatch tag: 0 = generic, else E of WarpRow (d = 16 -> E = 1)
static int row_mode(size_t d) {
  if (d == 16 || d == 32) return 1;
  if (d == 64) return 2;
  if (d == 128) return 4;
  if (d == 256) return 8;
  return 0;
}

// =====================================================================================
// warp-resident sorted list of the k best (value, id) pairs
// =====================================================================================
// Position p lives in lane p % 32, register p / 32 (R registers per lane, k <= 32*R).
// Positions >= k are kept at (+inf, sentinel).  Inserting shifts the tail up by one with
// shfl_up; the value falling off position k-1 is dropped.

template <int R>
struct WarpList {
  FT v[R];
  u32 id[R];
  __device__ __forceinline__ void clear(u32 sentinel) {
#pragma unroll
    for (int r = 0; r < R; r++) { v[r] = ft_inf(); id[r] = sentinel; }
  }
  __device__ __forceinline__ FT kth(int k) const {
    int rr = (k - 1) >> 5;
    FT x = v[0];
#pragma unroll
    for (int r = 1; r < R; r++) if (r == rr) x = v[r];
    return __shfl_sync(FULL, x, (k - 1) & 31);
  }
  __device__ __forceinline__ bool contains(u32 key) const {
    bool f = false;
#pragma unroll
    for (int r = 0; r < R; r++) f |= (id[r] == key && v[r] != ft_inf());
    return __any_sync(FULL, f);
  }
  // (vn, idn) warp-uniform, vn < kth(k), idn not contained
  __device__ __forceinline__ void insert(FT vn, u32 idn, int k, u32 sentinel, int lane) {
    FT carry_v = 0;
    u32 carry_id = 0;
#pragma unroll
    for (int r = 0; r < R; r++) {
      FT pv = __shfl_up_sync(FULL, v[r], 1);
      u32 pid = __shfl_up_sync(FULL, id[r], 1);
      FT last_v = __shfl_sync(FULL, v[r], 31);
      u32 last_id = __shfl_sync(FULL, id[r], 31);
      bool prev_le;
      if (lane == 0) {
        pv = carry_v; pid = carry_id;
        prev_le = (r == 0) ? true : (carry_v <= vn);
      } else {
        prev_le = pv <= vn;
      }
      bool le = v[r] <= vn;
      FT nv = le ? v[r] : (prev_le ? vn : pv);
      u32 ni = le ? id[r] : (prev_le ? idn : pid);
      if (r * 32 + lane >= k) { nv = ft_inf(); ni = sentinel; }
      v[r] = nv; id[r] = ni;
      carry_v = last_v; carry_id = last_id;
    }
  }
  // remove `key` if present (prefix-corner rule); later entries move down, tail = (+inf, sentinel)
  __device__ __forceinline__ void remove(u32 key, u32 sentinel, int lane) {
    bool f = false;
    int mypos = 0;
#pragma unroll
    for (int r = 0; r < R; r++) if (id[r] == key) { f = true; mypos = r * 32 + lane; }
    unsigned who = __ballot_sync(FULL, f);
    if (!who) return;
    int pos = __shfl_sync(FULL, mypos, __ffs(who) - 1);
    FT first_v[R];
    u32 first_id[R];
#pragma unroll
    for (int r = 0; r < R; r++) { first_v[r] = __shfl_sync(FULL, v[r], 0); first_id[r] = __shfl_sync(FULL, id[r], 0); }
#pragma unroll
    for (int r = 0; r < R; r++) {
      FT nv = __shfl_down_sync(FULL, v[r], 1);
      u32 ni = __shfl_down_sync(FULL, id[r], 1);
      if (lane == 31) {
        if (r + 1 < R) { nv = first_v[r + 1 < R ? r + 1 : r]; ni = first_id[r + 1 < R ? r + 1 : r]; }
        else { nv = ft_inf(); ni = sentinel; }
      }
      if (r * 32 + lane >= pos) { v[r] = nv; id[r] = ni; }
    }
  }
};

// Offers (vn, idn) to the list.  `tie` is raised whenever an EXACT distance tie between two
// different ids could influence the k best: the reference's result then depends on the
// order its sorting network leaves equal keys in (it can even keep an id twice), so such
// rows are redone by the literal-network kernels (DESIGN.md "exact ties").
template <int R>
__device__ __forceinline__ void consider(WarpList<R> &L, FT &tau, FT vn, u32 idn, int k,
                                         u32 sentinel, int lane, bool &tie) {
  if (vn < tau) {
    if (!L.contains(idn)) {
      bool eq = false;
#pragma unroll
      for (int r = 0; r < R; r++) eq |= (L.v[r] == vn);
      if (__any_sync(FULL, eq)) tie = true;
      L.insert(vn, idn, k, sentinel, lane);
      tau = L.kth(k);
    }
  } else if (vn == tau && vn != ft_inf()) {
    if (!L.contains(idn)) tie = true;
  }
}

static int list_regs(size_t k) {
  if (k <= 32) return 1;
  if (k <= 64) return 2;
  if (k <= 128) return 4;
  if (k <= 256) return 8;
  return 0;
}

__device__ __forceinline__ int floor_log2_u(unsigned long long v) { return 63 - __clzll(v); }


// =====================================================================================
// rows to redo literally + the reference's sorting network, one CTA per row
// =====================================================================================
// Fast kernels append the rows in which an exact tie could matter to a TieList; the literal
// kernels walk that list (or all rows when rows == NULL: rows shorter than 16 slots).

struct TieList {
  u32 *count;     // [1]
  u32 *rows;      // [capacity] or NULL = every row
};

__device__ __forceinline__ void tie_report(const TieList &t, u32 row) {
  u32 i = atomicAdd(t.count, 1u);
  t.rows[i] = row;
}

// scratch = [count | rows[cap] | slabs ...]; carved identically on host and device
struct LiteralScratch {
  TieList list;
  unsigned char *slabs;
  size_t slab_bytes;
};
static inline LiteralScratch carve_literal_scratch(void *scratch, size_t scratch_bytes, size_t cap_rows) {
  LiteralScratch L;
  unsigned char *p = (unsigned char *)scratch;
  L.list.count = (u32 *)p;
  L.list.rows = (u32 *)(p + 256);
  size_t used = (256 + cap_rows * sizeof(u32) + 255) & ~(size_t)255;
  L.slabs = p + used;
  L.slab_bytes = scratch_bytes > used ? scratch_bytes - used : 0;
  return L;
}

// compute.cl:181-217 + alg.c:137-144,224-230.  ids/key may live in shared or global memory.
// Every (stage, sub) step is a set of disjoint compare-exchanges shared by the CTA's threads.
// For len < 16 the network degenerates exactly as the reference's does (one work item of 8
// comparators guarded by pb < len).
__device__ __forceinline__ void block_network_sort(u32 *ids, FT *key, int len) {
  const int lk = floor_log2_u((unsigned long long)len);
  const int comps = 8 << (lk > 4 ? lk - 4 : 0);
  for (int stage = 0; stage < lk; stage++)
    for (int sub = stage; sub >= 0; sub--) {
      for (int w = threadIdx.x; w < comps; w += blockDim.x) {
        int hi = (w >> sub) << sub, lo = w ^ hi;
        int pa = (hi << 1) | lo;
        if (sub == stage) lo = (1 << sub) - lo - 1;
        int pb = (hi << 1) | (1 << sub) | lo;
        if (pb < len) {
          FT ka = key[pa], kb = key[pb];
          if (ka > kb) {
            u32 ia = ids[pa], ib = ids[pb];
            key[pa] = kb; key[pb] = ka;
            ids[pa] = ib; ids[pb] = ia;
          }
        }
      }
      __syncthreads();
    }
}

__device__ __forceinline__ void block_sort_and_uniq(u32 *ids, FT *key, int len) {
  block_network_sort(ids, key, len);
  for (int y = threadIdx.x; y + 1 < len; y += blockDim.x)
    if (ids[y] == ids[y + 1]) key[y] = ft_inf();
  __syncthreads();
  block_network_sort(ids, key, len);
}

// Distances from one query row to `count` rows given by row_of(i), written by key_at(i):
// the CTA's warps share the rows, each warp keeps 8 candidate rows in flight.
template <int E, typename RowOf, typename Emit>
__device__ __forceinline__ void block_row_distances(const FT *__restrict__ qrow, const FT *__restrict__ base,
                                                    int d, u32 count, FT *tmp, RowOf row_of, Emit emit) {
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, nw = blockDim.x >> 5;
  if (E) {
    WarpRow<(E ? E : 1)> qr;
    qr.load(qrow, lane, d);
    for (u32 i0 = wib * 8; i0 < count; i0 += nw * 8) {
      WarpRow<(E ? E : 1)> cr[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        u32 i = min(i0 + u, count - 1);
        cr[u].load(base + row_of(i) * (size_t)d, lane, d);
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        FT dist = __shfl_sync(FULL, warp_sqdist<(E ? E : 1)>(qr, cr[u], d), 0);
        if (lane == 0 && i0 + u < count) emit(i0 + u, dist);
      }
    }
  } else {
    for (u32 i = wib; i < count; i += nw) {
      FT dist = generic_sqdist(qrow, base + row_of(i) * (size_t)d, d, tmp, lane);
      if (lane == 0) emit(i, dist);
    }
  }
}
