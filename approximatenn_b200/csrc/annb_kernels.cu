// annb_kernels.cu — sm_100a kernels of the randomized all-points kNN path and the thin
// extern "C" launch layer declared in include/annb200.h.
//
// Compiled once per element type (-DUSE_FLOAT => float, otherwise double) with
//   nvcc -gencode arch=compute_100a,code=sm_100a -fmad=false
// -fmad=false is part of the contract: the reference's C path evaluates every product and
// sum separately, and the kernels below reproduce its values bit for bit
// by following the same operation order, so no multiply-add may be fused.
//
// Stage map (DESIGN.md has the full picture; reference citations are on the C-ABI
// declarations in include/annb200.h):
//   S0 fold_rows / scale_means        column means, the reference's summation tree
//   S1 hash_points                    centre + Givens + sub-permutation + Walsh-Hadamard +
//                                     Givens + projection + sign bits, all tries per read
//   S2 histogram / scan / scatter / sort_buckets / gather_rows
//   S3 leaf_topk                      bucket ∪ Hamming-1 buckets distances + per-point k best
//   S4 merge_lists                    union of the per-try lists
//   S5 supercharge                    neighbours-of-neighbours gather-distance-top-k
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include "annb200.h"

typedef ftype FT;
typedef annb_u32 u32;
#define FULL 0xffffffffu

static unsigned long g_launches = 0;

#define LAUNCH_CHECK(what)                                                              \
  do {                                                                                  \
    g_launches++;                                                                       \
    cudaError_t e_ = cudaGetLastError();                                                \
    if (e_ != cudaSuccess) {                                                            \
      fprintf(stderr, "approximatenn_b200: launch of %s failed: %s\n", what,            \
              cudaGetErrorString(e_));                                                  \
      exit(1);                                                                          \
    }                                                                                   \
  } while (0)

extern "C" unsigned long annb_launch_count(int reset) {
  unsigned long v = g_launches;
  if (reset) g_launches = 0;
  return v;
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
// S0: column means
// =====================================================================================

__global__ void fold_rows_kernel(const FT *__restrict__ src, FT *dst, size_t half, size_t len,
                                 size_t d, int first) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= half * d) return;
  FT a = src[e], b = src[e + half * d];
  FT extra = (e < d && (len & 1)) ? src[(len - 1) * d + e] : (FT)0;
  dst[e] = first ? (a + b) + extra : a + (b + extra);
}

__global__ void scale_means_kernel(const FT *acc, size_t n, size_t d, FT *mean) {
  size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < d) mean[c] = acc[c] / (FT)n;
}

extern "C" void annb_fold_rows(const FT *src, FT *dst, size_t len, size_t d, int first,
                               annb_stream stream) {
  size_t half = len / 2;
  if (half == 0) return;
  fold_rows_kernel<<<grid_for(half * d, 256), 256, 0, stream>>>(src, dst, half, len, d, first);
  LAUNCH_CHECK("fold_rows");
}

extern "C" void annb_scale_means(const FT *acc, size_t n, size_t d, FT *mean, annb_stream stream) {
  scale_means_kernel<<<grid_for(d, 128), 128, 0, stream>>>(acc, n, d, mean);
  LAUNCH_CHECK("scale_means");
}

// =====================================================================================
// S1: transform + hash
// =====================================================================================
// One CTA owns TP consecutive points; thread t owns point t of the tile.  The tile lives
// transposed in memory, element (coordinate c, point t) at c*(TP+1)+t, so that every
// per-coordinate access of a warp is one conflict-free shared-memory wavefront while the
// coordinate index stays a run-time value (the sub-permutations are data).  Two planes:
//   V[d]     the centred row, rotated in place by the "before" sweeps
//   Z[d_max] the embedded row, Walsh-Hadamard and "after" sweeps in place
// The planes sit in shared memory when they fit, otherwise in a global scratch slab.

template <int TP>
__global__ void __launch_bounds__(TP)
hash_points_kernel(const FT *__restrict__ points, const FT *__restrict__ mean,
                   annb_transform_desc t, u32 *__restrict__ hash, FT *gscratch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int LD = TP + 1;
  const size_t d = t.d, dm = t.d_max;
  FT *V = gscratch ? gscratch + (size_t)blockIdx.x * (d + dm) * LD : reinterpret_cast<FT *>(smem_raw);
  FT *Z = V + d * LD;
  const int tid = threadIdx.x;
  const size_t tiles = (t.n + TP - 1) / TP;
  const size_t planes_b = t.rots_before * t.rot_len_before;
  const size_t planes_all = planes_b + t.rots_after * t.rot_len_after;
  int levels = 0;
  while (((size_t)1 << levels) < dm) levels++;

  for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const size_t p0 = tile * TP;
    const size_t rows = (t.n - p0 < (size_t)TP) ? t.n - p0 : (size_t)TP;
    for (int tr = 0; tr < t.tries; tr++) {
      __syncthreads();
      // coalesced load of the tile, centred (compute.cl:44-49), stored transposed
      for (size_t e = tid; e < rows * d; e += TP) {
        size_t r = e / d, c = e - r * d;
        V[c * LD + r] = points[p0 * d + e] - mean[c];
      }
      __syncthreads();
      if ((size_t)tid < rows) {
        const u32 *pidx = t.plane_idx + (size_t)tr * planes_all * 2;
        const FT *pcs = t.plane_cs + (size_t)tr * planes_all * 2;
        const u32 *permb = t.perm_b + (size_t)tr * dm;
        const u32 *pick = t.pick + (size_t)tr * t.d_short;
        // "before" Givens sweeps on the d-vector (compute.cl:55-68)
        for (size_t q = 0; q < planes_b; q++) {
          u32 i = pidx[2 * q], j = pidx[2 * q + 1];
          FT c = pcs[2 * q], s = pcs[2 * q + 1];
          FT a = V[i * LD + tid], b = V[j * LD + tid];
          V[i * LD + tid] = a * c - b * s;
          V[j * LD + tid] = a * s + b * c;
        }
        // embed through the sub-permutation (compute.cl:77-85) fused with butterfly level 0
        if (dm == 1) {
          Z[tid] = permb[0] < d ? V[permb[0] * LD + tid] : (FT)0;
        } else {
          const bool odd = levels & 1;
          for (size_t y = 0; y < dm; y += 2) {
            u32 pa = permb[y], pb = permb[y + 1];
            FT a = pa < d ? V[pa * LD + tid] : (FT)0;
            FT b = pb < d ? V[pb * LD + tid] : (FT)0;
            FT lo = a + b, hi = a - b;
            if (odd) { lo *= t.inv_sqrt2; hi *= t.inv_sqrt2; }   // compute.cl:117-121
            Z[y * LD + tid] = lo;
            Z[(y + 1) * LD + tid] = hi;
          }
          // remaining butterfly levels; halve on odd levels (compute.cl:107-116)
          for (int lev = 1; lev < levels; lev++) {
            const size_t stride = (size_t)1 << lev;
            const bool halve = lev & 1;
            for (size_t w = 0; w < dm / 2; w++) {
              size_t hi_part = (w >> lev) << lev, lo_part = w ^ hi_part;
              size_t ia = (hi_part << 1) | lo_part, ib = ia | stride;
              FT a = Z[ia * LD + tid], b = Z[ib * LD + tid];
              FT s = a + b, df = a - b;
              if (halve) { s *= (FT)0.5; df *= (FT)0.5; }
              Z[ia * LD + tid] = s;
              Z[ib * LD + tid] = df;
            }
          }
        }
        // "after" sweeps on the first d_short coordinates of the d_max-vector
        for (size_t q = planes_b; q < planes_all; q++) {
          u32 i = pidx[2 * q], j = pidx[2 * q + 1];
          FT c = pcs[2 * q], s = pcs[2 * q + 1];
          FT a = Z[i * LD + tid], b = Z[j * LD + tid];
          Z[i * LD + tid] = a * c - b * s;
          Z[j * LD + tid] = a * s + b * c;
        }
        // projection + sign bits, first hashed coordinate = most significant bit
        u32 h = 0;
        for (size_t i = 0; i < t.d_short; i++) h = (h << 1) | sign_bit(Z[pick[i] * LD + tid]);
        hash[(size_t)tr * t.n + p0 + tid] = h;
      }
    }
  }
}

static const int HASH_TP = 128;
static const size_t HASH_SMEM_LIMIT = 200 * 1024;
static const unsigned HASH_SCRATCH_GRID = 148 * 4;

static size_t hash_plane_bytes(const annb_transform_desc *t) {
  return (t->d + t->d_max) * (size_t)(HASH_TP + 1) * sizeof(FT);
}

extern "C" size_t annb_hash_scratch_bytes(const annb_transform_desc *t) {
  size_t need = hash_plane_bytes(t);
  return need <= HASH_SMEM_LIMIT ? 0 : need * HASH_SCRATCH_GRID;
}

extern "C" void annb_hash_points(const FT *points, const FT *mean, const annb_transform_desc *t,
                                 u32 *hash, void *scratch, annb_stream stream) {
  size_t tiles = (t->n + HASH_TP - 1) / HASH_TP;
  size_t need = hash_plane_bytes(t);
  if (need <= HASH_SMEM_LIMIT) {
    static size_t configured = 0;
    if (need > configured) {
      cudaFuncSetAttribute(hash_points_kernel<HASH_TP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)HASH_SMEM_LIMIT);
      configured = HASH_SMEM_LIMIT;
    }
    unsigned grid = (unsigned)(tiles < 148 * 16 ? tiles : 148 * 16);
    hash_points_kernel<HASH_TP><<<grid, HASH_TP, need, stream>>>(points, mean, *t, hash, nullptr);
  } else {
    unsigned grid = (unsigned)(tiles < HASH_SCRATCH_GRID ? tiles : HASH_SCRATCH_GRID);
    hash_points_kernel<HASH_TP><<<grid, HASH_TP, 0, stream>>>(points, mean, *t, hash, (FT *)scratch);
  }
  LAUNCH_CHECK("hash_points");
}

// =====================================================================================
// S2: bucket tables
// =====================================================================================

__global__ void histogram_kernel(const u32 *__restrict__ hash, size_t n, u32 *count) {
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p < n) atomicAdd(&count[hash[p]], 1u);
}

// Exclusive scan in three steps: per-block scan (+ block maximum for tmax), scan of the
// block totals by one CTA, and the add-back.  SCAN_ITEMS counts per block.
static const int SCAN_THREADS = 256;
static const int SCAN_PER_THREAD = 8;
static const int SCAN_ITEMS = SCAN_THREADS * SCAN_PER_THREAD;

__device__ __forceinline__ u32 block_exclusive_scan(u32 v, u32 *total, u32 *warp_sums) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  u32 inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    u32 up = __shfl_up_sync(FULL, inc, o);
    if (lane >= o) inc += up;
  }
  if (lane == 31) warp_sums[warp] = inc;
  __syncthreads();
  if (warp == 0) {
    u32 w = lane < (int)(blockDim.x >> 5) ? warp_sums[lane] : 0;
    u32 winc = w;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      u32 up = __shfl_up_sync(FULL, winc, o);
      if (lane >= o) winc += up;
    }
    warp_sums[lane] = winc - w;                 // exclusive warp offsets
    if (lane == 31) *total = winc;
  }
  __syncthreads();
  u32 res = warp_sums[warp] + inc - v;
  __syncthreads();
  return res;
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_blocks_kernel(const u32 *__restrict__ count, size_t buckets, u32 *offset, u32 *block_tot,
                   u32 *tmax) {
  __shared__ u32 warp_sums[32];
  __shared__ u32 total;
  size_t base = (size_t)blockIdx.x * SCAN_ITEMS + (size_t)threadIdx.x * SCAN_PER_THREAD;
  u32 v[SCAN_PER_THREAD], sum = 0, mx = 0;
#pragma unroll
  for (int i = 0; i < SCAN_PER_THREAD; i++) {
    v[i] = base + i < buckets ? count[base + i] : 0;
    sum += v[i];
    mx = max(mx, v[i]);
  }
  u32 ex = block_exclusive_scan(sum, &total, warp_sums);
#pragma unroll
  for (int i = 0; i < SCAN_PER_THREAD; i++) {
    if (base + i < buckets) offset[base + i] = ex;
    ex += v[i];
  }
  mx = __reduce_max_sync(FULL, mx);
  if ((threadIdx.x & 31) == 0 && mx) atomicMax(tmax, mx);
  if (threadIdx.x == 0) block_tot[blockIdx.x] = total;
}

__global__ void __launch_bounds__(1024) scan_totals_kernel(u32 *block_tot, size_t blocks) {
  __shared__ u32 warp_sums[32];
  __shared__ u32 total;
  u32 carry = 0;
  for (size_t base = 0; base < blocks; base += blockDim.x) {
    size_t i = base + threadIdx.x;
    u32 v = i < blocks ? block_tot[i] : 0;
    u32 ex = block_exclusive_scan(v, &total, warp_sums);
    if (i < blocks) block_tot[i] = ex + carry;
    carry += total;
    __syncthreads();
  }
}

__global__ void __launch_bounds__(SCAN_THREADS)
scan_addback_kernel(u32 *offset, size_t buckets, const u32 *__restrict__ block_tot, u32 n) {
  size_t base = (size_t)blockIdx.x * SCAN_ITEMS + (size_t)threadIdx.x * SCAN_PER_THREAD;
  u32 add = block_tot[blockIdx.x];
#pragma unroll
  for (int i = 0; i < SCAN_PER_THREAD; i++)
    if (base + i < buckets) offset[base + i] += add;
  if (blockIdx.x == 0 && threadIdx.x == 0) offset[buckets] = n;
}

// Any order inside the bucket; sort_buckets_kernel fixes it.  Consumes `count` (ends at 0).
__global__ void scatter_kernel(const u32 *__restrict__ hash, size_t n,
                               const u32 *__restrict__ offset, u32 *count, u32 *order_tmp) {
  size_t p = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  u32 h = hash[p];
  u32 slot = atomicSub(&count[h], 1u) - 1u;
  order_tmp[offset[h] + slot] = (u32)p;
}

// One warp per bucket: rank every id by the number of larger ids in the bucket, i.e. write
// the bucket in DEcreasing id order (the reference fills its rows back to front while
// scanning ids upwards, alg.c:265-266).
__global__ void sort_buckets_kernel(const u32 *__restrict__ order_tmp,
                                    const u32 *__restrict__ offset, size_t buckets, u32 *order) {
  size_t b = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (b >= buckets) return;
  const int lane = threadIdx.x & 31;
  u32 beg = offset[b], cnt = offset[b + 1] - beg;
  if (cnt <= 32) {
    u32 mine = lane < (int)cnt ? order_tmp[beg + lane] : 0;
    u32 rank = 0;
    for (u32 j = 0; j < cnt; j++) {
      u32 other = __shfl_sync(FULL, mine, j);
      rank += other > mine;
    }
    if (lane < (int)cnt) order[beg + rank] = mine;
  } else {
    for (u32 e = lane; e < cnt; e += 32) {
      u32 mine = order_tmp[beg + e], rank = 0;
      for (u32 j = 0; j < cnt; j++) rank += order_tmp[beg + j] > mine;
      order[beg + rank] = mine;
    }
  }
}

extern "C" size_t annb_scan_tmp_bytes(size_t buckets) {
  size_t blocks = (buckets + SCAN_ITEMS - 1) / SCAN_ITEMS;
  return (blocks + 1) * sizeof(u32);
}

extern "C" void annb_build_buckets(const u32 *hash, size_t n, size_t buckets, u32 *count,
                                   u32 *offset, u32 *order_tmp, u32 *order, u32 *tmax,
                                   void *scan_tmp, annb_stream stream) {
  u32 *block_tot = (u32 *)scan_tmp;
  size_t blocks = (buckets + SCAN_ITEMS - 1) / SCAN_ITEMS;
  cudaMemsetAsync(count, 0, buckets * sizeof(u32), stream);
  cudaMemsetAsync(tmax, 0, sizeof(u32), stream);
  histogram_kernel<<<grid_for(n, 256), 256, 0, stream>>>(hash, n, count);
  LAUNCH_CHECK("histogram");
  scan_blocks_kernel<<<(unsigned)blocks, SCAN_THREADS, 0, stream>>>(count, buckets, offset, block_tot, tmax);
  LAUNCH_CHECK("scan_blocks");
  scan_totals_kernel<<<1, 1024, 0, stream>>>(block_tot, blocks);
  LAUNCH_CHECK("scan_totals");
  scan_addback_kernel<<<(unsigned)blocks, SCAN_THREADS, 0, stream>>>(offset, buckets, block_tot, (u32)n);
  LAUNCH_CHECK("scan_addback");
  scatter_kernel<<<grid_for(n, 256), 256, 0, stream>>>(hash, n, offset, count, order_tmp);
  LAUNCH_CHECK("scatter");
  sort_buckets_kernel<<<grid_for(buckets * 32, 256), 256, 0, stream>>>(order_tmp, offset, buckets, order);
  LAUNCH_CHECK("sort_buckets");
}

// sorted_points[r] = points[order[r]]: a group of lanes per row, 16-byte pieces when rows allow
template <typename VEC>
__global__ void gather_rows_kernel(const VEC *__restrict__ src, const u32 *__restrict__ order,
                                   size_t n, u32 vec_per_row, u32 lanes_per_row, VEC *dst) {
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t r = gid / lanes_per_row;
  u32 sub = (u32)(gid - r * lanes_per_row);
  if (r >= n) return;
  const VEC *s = src + (size_t)order[r] * vec_per_row;
  VEC *o = dst + r * vec_per_row;
  for (u32 v = sub; v < vec_per_row; v += lanes_per_row) o[v] = s[v];
}

extern "C" void annb_gather_rows(const FT *points, const u32 *order, size_t n, size_t d,
                                 FT *sorted_points, annb_stream stream) {
  size_t row_bytes = d * sizeof(FT);
  if (row_bytes % 16 == 0) {
    u32 vpr = (u32)(row_bytes / 16);
    u32 lanes = 1;
    while (lanes < vpr && lanes < 32) lanes <<= 1;
    gather_rows_kernel<uint4><<<grid_for(n * lanes, 256), 256, 0, stream>>>(
        (const uint4 *)points, order, n, vpr, lanes, (uint4 *)sorted_points);
  } else {
    u32 vpr = (u32)(row_bytes / 4);
    u32 lanes = 1;
    while (lanes < vpr && lanes < 32) lanes <<= 1;
    gather_rows_kernel<u32><<<grid_for(n * lanes, 256), 256, 0, stream>>>(
        (const u32 *)points, order, n, vpr, lanes, (u32 *)sorted_points);
  }
  LAUNCH_CHECK("gather_rows");
}

__global__ void export_table_kernel(const u32 *__restrict__ offset, const u32 *__restrict__ order,
                                    size_t n, size_t buckets, size_t tmax, size_t *table) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= buckets * tmax) return;
  size_t b = e / tmax, z = e - b * tmax;
  u32 beg = offset[b], cnt = offset[b + 1] - beg;
  table[e] = z < cnt ? (size_t)order[beg + z] : n;
}

extern "C" void annb_export_table(const u32 *offset, const u32 *order, size_t n, size_t buckets,
                                  size_t tmax, size_t *table, annb_stream stream) {
  export_table_kernel<<<grid_for(buckets * tmax, 256), 256, 0, stream>>>(offset, order, n, buckets, tmax, table);
  LAUNCH_CHECK("export_table");
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
    for (int r = 0; r < R; r++) f |= (id[r] == key);
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

template <int R>
__device__ __forceinline__ void consider(WarpList<R> &L, FT &tau, FT vn, u32 idn, int k,
                                         u32 sentinel, int lane) {
  if (vn < tau) {
    if (!L.contains(idn)) {
      L.insert(vn, idn, k, sentinel, lane);
      tau = L.kth(k);
    }
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
// S3: per-try candidates -> k best per point  (generic, one warp per point)
// =====================================================================================

template <int E, int R>
__global__ void __launch_bounds__(256)
leaf_topk_warp_kernel(const FT *__restrict__ sp, const u32 *__restrict__ order,
                      const u32 *__restrict__ offset, const u32 *__restrict__ hash,
                      const u32 *__restrict__ tmax_p, size_t n, int d, int d_short, int k,
                      u32 *__restrict__ list_ids, FT *__restrict__ list_dist) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  size_t r = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;      // sorted position
  if (r >= n) return;
  FT *tmp = reinterpret_cast<FT *>(smem_raw) + (size_t)wib * (E == 0 ? d : 0);
  const u32 sentinel = (u32)n;
  const u32 me = order[r];
  const u32 h = hash[me];
  const unsigned long long tmax = *tmax_p;
  const unsigned long long L = (unsigned long long)(d_short + 1) * tmax;
  const unsigned long long P = 1ull << floor_log2_u(L);

  WarpRow<(E ? E : 1)> q;
  const FT *qrow = sp + r * (size_t)d;
  if (E) q.load(qrow, lane, d);
  WarpList<R> best;
  best.clear(sentinel);
  FT tau = ft_inf();

  for (int y = 0; y <= d_short; y++) {
    unsigned long long first_slot = (unsigned long long)y * tmax;
    if (first_slot >= P) break;
    u32 b = h ^ (y ? (1u << (y - 1)) : 0u);
    u32 beg = offset[b], cnt = offset[b + 1] - beg;
    unsigned long long room = P - first_slot;
    u32 inc = cnt < room ? cnt : (u32)room;
    for (u32 c = 0; c < inc; c++) {
      size_t row = (size_t)beg + c;
      if (row == r) continue;                                  // self (compute.cl:145)
      FT dist;
      if (E) {
        WarpRow<(E ? E : 1)> cr;
        cr.load(sp + row * (size_t)d, lane, d);
        dist = warp_sqdist<(E ? E : 1)>(q, cr, d);
        dist = __shfl_sync(FULL, dist, 0);
      } else {
        dist = generic_sqdist(qrow, sp + row * (size_t)d, d, tmp, lane);
      }
      consider<R>(best, tau, dist, order[row], k, sentinel, lane);
    }
  }
#pragma unroll
  for (int rr = 0; rr < R; rr++) {
    int p = rr * 32 + lane;
    if (p < k) {
      list_ids[(size_t)me * k + p] = best.id[rr];
      list_dist[(size_t)me * k + p] = best.v[rr];
    }
  }
}

template <int E>
static void launch_leaf_topk_r(int regs, dim3 grid, dim3 block, size_t smem, annb_stream stream,
                               const FT *sp, const u32 *order, const u32 *offset, const u32 *hash,
                               const u32 *tmax, size_t n, int d, int d_short, int k, u32 *ids,
                               FT *dist) {
  switch (regs) {
    case 1: leaf_topk_warp_kernel<E, 1><<<grid, block, smem, stream>>>(sp, order, offset, hash, tmax, n, d, d_short, k, ids, dist); break;
    case 2: leaf_topk_warp_kernel<E, 2><<<grid, block, smem, stream>>>(sp, order, offset, hash, tmax, n, d, d_short, k, ids, dist); break;
    case 4: leaf_topk_warp_kernel<E, 4><<<grid, block, smem, stream>>>(sp, order, offset, hash, tmax, n, d, d_short, k, ids, dist); break;
    default: leaf_topk_warp_kernel<E, 8><<<grid, block, smem, stream>>>(sp, order, offset, hash, tmax, n, d, d_short, k, ids, dist); break;
  }
}

static void fatal_config(const char *what) {
  fprintf(stderr, "approximatenn_b200: unsupported configuration: %s\n", what);
  exit(1);
}

extern "C" void annb_leaf_topk(const FT *sorted_points, const u32 *order, const u32 *offset,
                               const u32 *hash, const u32 *tmax, size_t n, size_t d,
                               size_t d_short, size_t k, u32 *list_ids, FT *list_dist,
                               annb_stream stream) {
  int regs = list_regs(k);
  if (!regs) fatal_config("k > 256");
  int mode = row_mode(d);
  dim3 block(256), grid(grid_for(n * 32, 256));
  size_t smem = mode ? 0 : 8 * d * sizeof(FT);
  if (smem > 200 * 1024) fatal_config("d too large for the generic distance path");
  switch (mode) {
    case 0:
      if (smem > 48 * 1024) {
        cudaFuncSetAttribute(leaf_topk_warp_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(leaf_topk_warp_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(leaf_topk_warp_kernel<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(leaf_topk_warp_kernel<0, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      }
      launch_leaf_topk_r<0>(regs, grid, block, smem, stream, sorted_points, order, offset, hash, tmax, n, (int)d, (int)d_short, (int)k, list_ids, list_dist);
      break;
    case 1: launch_leaf_topk_r<1>(regs, grid, block, smem, stream, sorted_points, order, offset, hash, tmax, n, (int)d, (int)d_short, (int)k, list_ids, list_dist); break;
    case 2: launch_leaf_topk_r<2>(regs, grid, block, smem, stream, sorted_points, order, offset, hash, tmax, n, (int)d, (int)d_short, (int)k, list_ids, list_dist); break;
    case 4: launch_leaf_topk_r<4>(regs, grid, block, smem, stream, sorted_points, order, offset, hash, tmax, n, (int)d, (int)d_short, (int)k, list_ids, list_dist); break;
    default: launch_leaf_topk_r<8>(regs, grid, block, smem, stream, sorted_points, order, offset, hash, tmax, n, (int)d, (int)d_short, (int)k, list_ids, list_dist); break;
  }
  LAUNCH_CHECK("leaf_topk");
}

// =====================================================================================
// S4: union of the per-try lists, one warp per point
// =====================================================================================
// The reference concatenates the per-try lists into a row of k*tries slots and runs
// sort / kill-adjacent-duplicate-ids / sort on the first 2^floor(log2(k*tries)) of them
// (SURVEY §8.A.3 rules 5-6).  Equal ids carry equal distances, so that is the k smallest
// distinct ids among the admitted entries.  One corner is reproduced as well: the largest
// admitted entry is dropped when its id equals the id in the first slot outside the
// sorted prefix and no admitted entry is infinite ("prefix corner").
struct MergeArgs {
  int n_lists;
  int admit[64];
  int corner_list, corner_pos;
};

template <int R>
__global__ void __launch_bounds__(256)
merge_lists_kernel(const u32 *__restrict__ lists_ids, const FT *__restrict__ lists_dist,
                   MergeArgs a, const u32 *__restrict__ prev_ids, const FT *__restrict__ prev_dist,
                   size_t n, int k, u32 *__restrict__ out_ids, FT *__restrict__ out_dist) {
  const int lane = threadIdx.x & 31;
  size_t x = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (x >= n) return;
  const u32 sentinel = (u32)n;
  WarpList<R> best;
  best.clear(sentinel);
  FT tau = ft_inf();
  FT max_v = -ft_inf();
  u32 max_id = sentinel;
  bool any_inf = false;

  for (int li = (prev_ids ? -1 : 0); li < a.n_lists; li++) {
    const u32 *ids = li < 0 ? prev_ids + x * (size_t)k : lists_ids + ((size_t)li * n + x) * k;
    const FT *dist = li < 0 ? prev_dist + x * (size_t)k : lists_dist + ((size_t)li * n + x) * k;
    const int admit = li < 0 ? k : a.admit[li];
    for (int base = 0; base < admit; base += 32) {
      int e = base + lane;
      FT mv = e < admit ? dist[e] : ft_inf();
      u32 mi = e < admit ? ids[e] : sentinel;
      int cnt = min(32, admit - base);
      for (int j = 0; j < cnt; j++) {
        FT vn = __shfl_sync(FULL, mv, j);
        u32 idn = __shfl_sync(FULL, mi, j);
        if (vn == ft_inf()) { any_inf = true; continue; }
        if (vn > max_v) { max_v = vn; max_id = idn; }
        consider<R>(best, tau, vn, idn, k, sentinel, lane);
      }
    }
  }
  if (a.corner_list >= 0 && !any_inf) {
    u32 cid = lists_ids[((size_t)a.corner_list * n + x) * k + a.corner_pos];
    if (cid == max_id) best.remove(cid, sentinel, lane);
  }
#pragma unroll
  for (int rr = 0; rr < R; rr++) {
    int p = rr * 32 + lane;
    if (p < k) {
      out_ids[x * (size_t)k + p] = best.id[rr];
      out_dist[x * (size_t)k + p] = best.v[rr];
    }
  }
}

extern "C" void annb_merge_lists(const u32 *lists_ids, const FT *lists_dist, int n_lists,
                                 const int *host_admit, int corner_list, int corner_pos,
                                 const u32 *merged_in_ids, const FT *merged_in_dist, size_t n,
                                 size_t k, u32 *merged_ids, FT *merged_dist, annb_stream stream) {
  int regs = list_regs(k);
  if (!regs) fatal_config("k > 256");
  if (n_lists > 64) fatal_config("more than 64 lists per merge call");
  MergeArgs a;
  a.n_lists = n_lists;
  for (int i = 0; i < n_lists; i++) a.admit[i] = host_admit[i];
  a.corner_list = corner_list;
  a.corner_pos = corner_pos;
  dim3 block(256), grid(grid_for(n * 32, 256));
  switch (regs) {
    case 1: merge_lists_kernel<1><<<grid, block, 0, stream>>>(lists_ids, lists_dist, a, merged_in_ids, merged_in_dist, n, (int)k, merged_ids, merged_dist); break;
    case 2: merge_lists_kernel<2><<<grid, block, 0, stream>>>(lists_ids, lists_dist, a, merged_in_ids, merged_in_dist, n, (int)k, merged_ids, merged_dist); break;
    case 4: merge_lists_kernel<4><<<grid, block, 0, stream>>>(lists_ids, lists_dist, a, merged_in_ids, merged_in_dist, n, (int)k, merged_ids, merged_dist); break;
    default: merge_lists_kernel<8><<<grid, block, 0, stream>>>(lists_ids, lists_dist, a, merged_in_ids, merged_in_dist, n, (int)k, merged_ids, merged_dist); break;
  }
  LAUNCH_CHECK("merge_lists");
}

// =====================================================================================
// S5: supercharging, one warp per query row
// =====================================================================================
// Row of the reference: [own k] ++ [list of own[0]] ++ ... ++ [list of own[k-1]], k(k+1)
// slots; the first P2 = 2^floor(log2(k(k+1))) compete.  Own distances are carried over,
// the others are measured here.  Same duplicate and prefix-corner rules as S4.

template <int E, int R>
__global__ void __launch_bounds__(256)
supercharge_kernel(const FT *__restrict__ queries, const FT *__restrict__ points,
                   const u32 *__restrict__ own_ids, const FT *__restrict__ own_dist,
                   const u32 *__restrict__ graph, size_t n, int d, int k, size_t row_begin,
                   size_t row_end, int exclude_self, size_t *__restrict__ out_ids,
                   FT *__restrict__ out_dist) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  size_t x = row_begin + (((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (x >= row_end) return;
  FT *tmp = reinterpret_cast<FT *>(smem_raw) + (size_t)wib * (E == 0 ? d : 0);
  const u32 sentinel = (u32)n;
  const int wide = k * (k + 1);
  const int P2 = 1 << floor_log2_u((unsigned long long)wide);

  WarpRow<(E ? E : 1)> q;
  const FT *qrow = queries + x * (size_t)d;
  if (E) q.load(qrow, lane, d);

  WarpList<R> best;
  FT max_v = -ft_inf();
  u32 max_id = sentinel;
  bool any_inf = false;
#pragma unroll
  for (int rr = 0; rr < R; rr++) {
    int p = rr * 32 + lane;
    best.v[rr] = p < k ? own_dist[x * (size_t)k + p] : ft_inf();
    best.id[rr] = p < k ? own_ids[x * (size_t)k + p] : sentinel;
    bool fin = best.v[rr] != ft_inf();
    if (p < k && !fin) any_inf = true;
    if (p < k && fin && best.v[rr] > max_v) { max_v = best.v[rr]; max_id = best.id[rr]; }
  }
  any_inf = __any_sync(FULL, any_inf);
  // the own list ascends, so its largest finite entry is the last finite one
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    FT ov = __shfl_xor_sync(FULL, max_v, o);
    u32 oi = __shfl_xor_sync(FULL, max_id, o);
    if (ov > max_v) { max_v = ov; max_id = oi; }
  }
  FT tau = best.kth(k);
  // own ids, kept for the neighbour walk
  u32 own_reg[R];
#pragma unroll
  for (int rr = 0; rr < R; rr++) own_reg[rr] = best.id[rr];

  const int cand = P2 - k;                     // slots k .. P2-1 of the row
  for (int base = 0; base < cand; base += 32) {
    int c = base + lane;
    u32 cid = sentinel;
    if (c < cand) {
      int j = c / k, z = c - j * k;
      // own id at position j
      u32 oj = sentinel;
      {
        int src_lane = j & 31, src_reg = j >> 5;
        // every lane needs a different j: go through shared-free shuffles per register
#pragma unroll
        for (int rr = 0; rr < R; rr++) {
          u32 got = __shfl_sync(FULL, own_reg[rr], src_lane);
          if (rr == src_reg) oj = got;
        }
      }
      cid = oj < sentinel ? graph[(size_t)oj * k + z] : sentinel;
    } else {
#pragma unroll
      for (int rr = 0; rr < R; rr++) (void)__shfl_sync(FULL, own_reg[rr], 0);
    }
    int cnt = min(32, cand - base);
    for (int i = 0; i < cnt; i++) {
      u32 idn = __shfl_sync(FULL, cid, i);
      if (idn >= sentinel || (exclude_self && idn == (u32)x)) { any_inf = true; continue; }
      FT dist;
      if (E) {
        WarpRow<(E ? E : 1)> cr;
        cr.load(points + (size_t)idn * d, lane, d);
        dist = warp_sqdist<(E ? E : 1)>(q, cr, d);
        dist = __shfl_sync(FULL, dist, 0);
      } else {
        dist = generic_sqdist(qrow, points + (size_t)idn * d, d, tmp, lane);
      }
      if (dist > max_v) { max_v = dist; max_id = idn; }
      consider<R>(best, tau, dist, idn, k, sentinel, lane);
    }
  }
  if (P2 < wide && !any_inf) {
    int c = P2 - k, j = c / k, z = c - j * k;
    u32 oj = sentinel;
#pragma unroll
    for (int rr = 0; rr < R; rr++) {
      u32 got = __shfl_sync(FULL, own_reg[rr], j & 31);
      if (rr == (j >> 5)) oj = got;
    }
    u32 cid = oj < sentinel ? graph[(size_t)oj * k + z] : sentinel;
    if (cid == max_id) best.remove(cid, sentinel, lane);
  }
  size_t orow = x - row_begin;
#pragma unroll
  for (int rr = 0; rr < R; rr++) {
    int p = rr * 32 + lane;
    if (p < k) {
      out_ids[orow * (size_t)k + p] = (size_t)best.id[rr];
      if (out_dist) out_dist[orow * (size_t)k + p] = best.v[rr];
    }
  }
}

template <int E>
static void launch_supercharge_r(int regs, dim3 grid, dim3 block, size_t smem, annb_stream stream,
                                 const FT *queries, const FT *points, const u32 *own_ids,
                                 const FT *own_dist, const u32 *graph, size_t n, int d, int k,
                                 size_t rb, size_t re, int ex, size_t *out_ids, FT *out_dist) {
  switch (regs) {
    case 1: supercharge_kernel<E, 1><<<grid, block, smem, stream>>>(queries, points, own_ids, own_dist, graph, n, d, k, rb, re, ex, out_ids, out_dist); break;
    case 2: supercharge_kernel<E, 2><<<grid, block, smem, stream>>>(queries, points, own_ids, own_dist, graph, n, d, k, rb, re, ex, out_ids, out_dist); break;
    case 4: supercharge_kernel<E, 4><<<grid, block, smem, stream>>>(queries, points, own_ids, own_dist, graph, n, d, k, rb, re, ex, out_ids, out_dist); break;
    default: supercharge_kernel<E, 8><<<grid, block, smem, stream>>>(queries, points, own_ids, own_dist, graph, n, d, k, rb, re, ex, out_ids, out_dist); break;
  }
}

extern "C" void annb_supercharge(const FT *queries, const FT *points, const u32 *own_ids,
                                 const FT *own_dist, const u32 *graph, size_t n, size_t d, size_t k,
                                 size_t row_begin, size_t row_end, int exclude_self,
                                 size_t *out_ids, FT *out_dist, annb_stream stream) {
  if (row_end <= row_begin) return;
  int regs = list_regs(k);
  if (!regs) fatal_config("k > 256");
  int mode = row_mode(d);
  dim3 block(256), grid(grid_for((row_end - row_begin) * 32, 256));
  size_t smem = mode ? 0 : 8 * d * sizeof(FT);
  if (smem > 200 * 1024) fatal_config("d too large for the generic distance path");
#define SC_ARGS regs, grid, block, smem, stream, queries, points, own_ids, own_dist, graph, n, (int)d, (int)k, row_begin, row_end, exclude_self, out_ids, out_dist
  switch (mode) {
    case 0:
      if (smem > 48 * 1024) {
        cudaFuncSetAttribute(supercharge_kernel<0, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(supercharge_kernel<0, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(supercharge_kernel<0, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        cudaFuncSetAttribute(supercharge_kernel<0, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      }
      launch_supercharge_r<0>(SC_ARGS); break;
    case 1: launch_supercharge_r<1>(SC_ARGS); break;
    case 2: launch_supercharge_r<2>(SC_ARGS); break;
    case 4: launch_supercharge_r<4>(SC_ARGS); break;
    default: launch_supercharge_r<8>(SC_ARGS); break;
  }
#undef SC_ARGS
  LAUNCH_CHECK("supercharge");
}

// =====================================================================================
// rows shorter than 16 slots: literal emulation of the reference's network, thread per row
// =====================================================================================
#define TINY_MAX 16

__device__ void tiny_network_sort(u32 *ids, FT *key, int len) {
  int lk = floor_log2_u((unsigned long long)len);
  for (int stage = 0; stage < lk; stage++)
    for (int sub = stage; sub >= 0; sub--)
      for (int w = 0; w < 8; w++) {                         // compute.cl:187-205, one work item
        int hi = (w >> sub) << sub, lo = w ^ hi;
        int pa = (hi << 1) | lo;
        if (sub == stage) lo = (1 << sub) - lo - 1;
        int pb = (hi << 1) | (1 << sub) | lo;
        if (pb < len && key[pa] > key[pb]) {
          FT tk = key[pa]; key[pa] = key[pb]; key[pb] = tk;
          u32 ti = ids[pa]; ids[pa] = ids[pb]; ids[pb] = ti;
        }
      }
}

__device__ void tiny_sort_and_uniq(u32 *ids, FT *key, int len) {
  tiny_network_sort(ids, key, len);
  for (int y = 0; y + 1 < len; y++)
    if (ids[y] == ids[y + 1]) key[y] = ft_inf();
  tiny_network_sort(ids, key, len);
}

__global__ void merge_lists_tiny_kernel(const u32 *__restrict__ lists_ids,
                                        const FT *__restrict__ lists_dist, int n_lists, size_t n,
                                        int k, u32 *out_ids, FT *out_dist) {
  size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n) return;
  u32 ids[TINY_MAX];
  FT key[TINY_MAX];
  int len = n_lists * k;
  for (int t = 0; t < n_lists; t++)
    for (int z = 0; z < k; z++) {
      ids[t * k + z] = lists_ids[((size_t)t * n + x) * k + z];
      key[t * k + z] = lists_dist[((size_t)t * n + x) * k + z];
    }
  tiny_sort_and_uniq(ids, key, len);
  for (int z = 0; z < k; z++) {
    out_ids[x * (size_t)k + z] = ids[z];
    out_dist[x * (size_t)k + z] = key[z];
  }
}

extern "C" void annb_merge_lists_tiny(const u32 *lists_ids, const FT *lists_dist, int n_lists,
                                      size_t n, size_t k, u32 *merged_ids, FT *merged_dist,
                                      annb_stream stream) {
  if ((size_t)n_lists * k >= TINY_MAX) fatal_config("tiny merge called with a row of 16+ slots");
  merge_lists_tiny_kernel<<<grid_for(n, 128), 128, 0, stream>>>(lists_ids, lists_dist, n_lists, n, (int)k, merged_ids, merged_dist);
  LAUNCH_CHECK("merge_lists_tiny");
}

// serial version of the reference's summation tree; sq must hold d entries (global scratch)
__device__ FT serial_sqdist(const FT *q, const FT *c, int d, FT *sq) {
  for (int z = 0; z < d; z++) {
    FT df = q[z] - c[z];
    sq[z] = df * df;
  }
  for (int l = d; l >> 1; l >>= 1) {
    int h = l >> 1;
    for (int z = 0; z < h; z++) {
      FT add = sq[z + h];
      if (z == 0 && (l & 1)) add = add + sq[l - 1];
      sq[z] = sq[z] + add;
    }
  }
  return sq[0];
}

__global__ void supercharge_tiny_kernel(const FT *__restrict__ queries, const FT *__restrict__ points,
                                        const u32 *__restrict__ own_ids, const FT *__restrict__ own_dist,
                                        const u32 *__restrict__ graph, size_t n, int d, int k,
                                        size_t row_begin, size_t row_end, int exclude_self,
                                        FT *scratch, size_t *out_ids, FT *out_dist) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t x = row_begin + i;
  if (x >= row_end) return;
  u32 ids[TINY_MAX];
  FT key[TINY_MAX];
  FT *sq = scratch + i * (size_t)d;
  int wide = k * (k + 1);
  for (int z = 0; z < k; z++) {
    ids[z] = own_ids[x * (size_t)k + z];
    key[z] = own_dist[x * (size_t)k + z];
  }
  for (int j = 0; j < k; j++)
    for (int z = 0; z < k; z++) {
      u32 oj = ids[j];
      u32 cid = oj < (u32)n ? graph[(size_t)oj * k + z] : (u32)n;
      ids[(j + 1) * k + z] = cid;
      if (cid >= (u32)n || (exclude_self && cid == (u32)x)) key[(j + 1) * k + z] = ft_inf();
      else key[(j + 1) * k + z] = serial_sqdist(queries + x * (size_t)d, points + (size_t)cid * d, d, sq);
    }
  tiny_sort_and_uniq(ids, key, wide);
  for (int z = 0; z < k; z++) {
    out_ids[i * (size_t)k + z] = (size_t)ids[z];
    if (out_dist) out_dist[i * (size_t)k + z] = key[z];
  }
}

static FT *g_tiny_scratch = nullptr;
static size_t g_tiny_scratch_elems = 0;

extern "C" void annb_supercharge_tiny(const FT *queries, const FT *points, const u32 *own_ids,
                                      const FT *own_dist, const u32 *graph, size_t n, size_t d,
                                      size_t k, size_t row_begin, size_t row_end, int exclude_self,
                                      size_t *out_ids, FT *out_dist, annb_stream stream) {
  if (k * (k + 1) >= TINY_MAX) fatal_config("tiny supercharge called with a row of 16+ slots");
  if (row_end <= row_begin) return;
  size_t rows = row_end - row_begin;
  if (rows * d > g_tiny_scratch_elems) {
    cudaStreamSynchronize(stream);
    cudaFree(g_tiny_scratch);
    if (cudaMalloc(&g_tiny_scratch, rows * d * sizeof(FT)) != cudaSuccess) fatal_config("out of device memory");
    g_tiny_scratch_elems = rows * d;
  }
  supercharge_tiny_kernel<<<grid_for(rows, 128), 128, 0, stream>>>(queries, points, own_ids, own_dist, graph, n, (int)d, (int)k, row_begin, row_end, exclude_self, g_tiny_scratch, out_ids, out_dist);
  LAUNCH_CHECK("supercharge_tiny");
}
