// annb_prepare.cu — S0 column means, S1 transform + sign hash, S2 bucket tables.
// See annb_common.cuh / include/annb200.h; compiled with -fmad=false (bit-exact contract).
#include "annb_common.cuh"

unsigned long annb_g_launches = 0;

extern "C" unsigned long annb_launch_count(int reset) {
  unsigned long v = annb_g_launches;
  if (reset) annb_g_launches = 0;
  return v;
}

extern "C" void annb_literal_rows(unsigned long long out[3], int reset) {
  out[0] = annb_leaf_literal_count(reset);
  annb_finish_literal_counts(out + 1, reset);
}

// =====================================================================================
// S0: column means
// =====================================================================================

// F levels of the reference's stride-halving row sum in one pass (alg.c:122-128):
//   V_{i+1}[x] = V_i[x] + (V_i[x + len_{i+1}] + extra),  len_{i+1} = len_i / 2,
//   extra = V_i[len_i - 1] for x == 0 when len_i is odd, else 0
// (the very first level, on the raw points, associates as (a + b) + extra, compute.cl:19-20).
// Fold<F>::get is the value of level +F at row x, evaluated as a compile-time recursion over the
// 2^F (plus the rare odd-length extras) source rows it depends on — same additions, same order.
template <int F>
struct Fold {
  static __device__ __forceinline__ FT get(const FT *__restrict__ src, size_t len, size_t d, size_t x,
                                           size_t c, bool first) {
    const size_t lp = len >> (F - 1), h = lp >> 1;
    FT a = Fold<F - 1>::get(src, len, d, x, c, first);
    FT b = Fold<F - 1>::get(src, len, d, x + h, c, first);
    FT e = (x == 0 && (lp & 1)) ? Fold<F - 1>::get(src, len, d, lp - 1, c, first) : (FT)0;
    return (F == 1 && first) ? (a + b) + e : a + (b + e);
  }
};
template <>
struct Fold<0> {
  static __device__ __forceinline__ FT get(const FT *__restrict__ src, size_t, size_t d, size_t x, size_t c, bool) {
    return src[x * d + c];
  }
};

template <int F>
__global__ void fold_rows_kernel(const FT *src, FT *dst, size_t len, size_t d, int first) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  const size_t out_rows = len >> F;
  if (e >= out_rows * d) return;
  size_t x = e / d, c = e - x * d;
  dst[e] = Fold<F>::get(src, len, d, x, c, first != 0);
}

__global__ void scale_means_kernel(const FT *acc, size_t n, size_t d, FT *mean) {
  size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c < d) mean[c] = acc[c] / (FT)n;
}

extern "C" void annb_fold_rows(const FT *src, FT *dst, size_t len, size_t d, int levels, int first,
                               annb_stream stream) {
  size_t out = (len >> levels) * d;
  if (out == 0) return;
  unsigned grid = grid_for(out, 256);
  switch (levels) {
    case 1: fold_rows_kernel<1><<<grid, 256, 0, stream>>>(src, dst, len, d, first); break;
    case 2: fold_rows_kernel<2><<<grid, 256, 0, stream>>>(src, dst, len, d, first); break;
    case 3: fold_rows_kernel<3><<<grid, 256, 0, stream>>>(src, dst, len, d, first); break;
    case 4: fold_rows_kernel<4><<<grid, 256, 0, stream>>>(src, dst, len, d, first); break;
    default: fatal_config("annb_fold_rows: 1..4 levels per pass");
  }
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
      for (size_t r = tid >> 5; r < rows; r += TP / 32) {
        const FT *src = points + (p0 + r) * d;
        for (size_t c = tid & 31; c < d; c += 32) V[c * LD + r] = src[c] - mean[c];
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

// Register variant for d_max in {16,32,64,128}: the Walsh-Hadamard butterflies run on a
// fully unrolled register array (compile-time indices), shared memory only carries the
// data-dependent accesses: the "before" sweeps, the sub-permutation gather, the "after"
// sweeps and the projection picks.  V0 keeps the centred tile for all tries (one global
// read of the points); W is the per-try working plane (rotated row, then transformed row).
template <int DM, int TP>
__global__ void __launch_bounds__(TP)
hash_points_reg_kernel(const FT *__restrict__ points, const FT *__restrict__ mean,
                       annb_transform_desc t, u32 *__restrict__ hash) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int LD = TP + 1;
  constexpr int LEVELS = DM == 16 ? 4 : DM == 32 ? 5 : DM == 64 ? 6 : 7;
  const int d = (int)t.d;
  FT *V0 = reinterpret_cast<FT *>(smem_raw);
  FT *W = V0 + (size_t)d * LD;
  const int tid = threadIdx.x;
  const size_t tiles = (t.n + TP - 1) / TP;
  const int planes_b = (int)(t.rots_before * t.rot_len_before);
  const int planes_all = planes_b + (int)(t.rots_after * t.rot_len_after);
  const int ds = (int)t.d_short;

  for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const size_t p0 = tile * TP;
    const size_t rows = (t.n - p0 < (size_t)TP) ? t.n - p0 : (size_t)TP;
    __syncthreads();
    // rows are handed to warps, a warp reads a row with coalesced 128-byte requests; consecutive
    // coordinates land in consecutive banks of the transposed tile (LD is odd)
    for (size_t r = tid >> 5; r < rows; r += TP / 32) {
      const FT *src = points + (p0 + r) * (size_t)d;
      for (int c = tid & 31; c < d; c += 32) V0[c * LD + r] = src[c] - mean[c];   // compute.cl:44-49
    }
    __syncthreads();
    if ((size_t)tid >= rows) continue;
    for (int tr = 0; tr < t.tries; tr++) {
      const u32 *pidx = t.plane_idx + (size_t)tr * planes_all * 2;
      const FT *pcs = t.plane_cs + (size_t)tr * planes_all * 2;
      const u32 *permb = t.perm_b + (size_t)tr * DM;
      const u32 *pick = t.pick + (size_t)tr * ds;
      for (int c = 0; c < d; c++) W[c * LD + tid] = V0[c * LD + tid];
      for (int q = 0; q < planes_b; q++) {                       // compute.cl:55-68
        u32 i = pidx[2 * q], j = pidx[2 * q + 1];
        FT c = pcs[2 * q], sn = pcs[2 * q + 1];
        FT a = W[i * LD + tid], b = W[j * LD + tid];
        W[i * LD + tid] = a * c - b * sn;
        W[j * LD + tid] = a * sn + b * c;
      }
      FT z[DM];
#pragma unroll
      for (int y = 0; y < DM; y++) {                             // compute.cl:77-85
        u32 src = permb[y];
        z[y] = src < (u32)d ? W[src * LD + tid] : (FT)0;
      }
#pragma unroll
      for (int lev = 0; lev < LEVELS; lev++) {                   // compute.cl:101-122
#pragma unroll
        for (int w = 0; w < DM / 2; w++) {
          const int hi_part = (w >> lev) << lev, lo_part = w ^ hi_part;
          const int ia = (hi_part << 1) | lo_part, ib = ia | (1 << lev);
          FT a = z[ia], b = z[ib];
          FT sm = a + b, df = a - b;
          if (lev & 1) { sm *= (FT)0.5; df *= (FT)0.5; }
          if (lev == 0 && (LEVELS & 1)) { sm *= t.inv_sqrt2; df *= t.inv_sqrt2; }
          z[ia] = sm;
          z[ib] = df;
        }
      }
#pragma unroll
      for (int y = 0; y < DM; y++) W[y * LD + tid] = z[y];
      for (int q = planes_b; q < planes_all; q++) {
        u32 i = pidx[2 * q], j = pidx[2 * q + 1];
        FT c = pcs[2 * q], sn = pcs[2 * q + 1];
        FT a = W[i * LD + tid], b = W[j * LD + tid];
        W[i * LD + tid] = a * c - b * sn;
        W[j * LD + tid] = a * sn + b * c;
      }
      u32 h = 0;
      for (int i = 0; i < ds; i++) h = (h << 1) | sign_bit(W[pick[i] * LD + tid]);
      hash[(size_t)tr * t.n + p0 + tid] = h;
    }
  }
}

// Second register variant (needs the tables on the host).  Differences to the kernel above:
//  * the per-try tables live in CONSTANT memory, re-written per launch: the index streams are
//    uniform across the warp, so they come through the constant cache instead of the LSU;
//  * the centred tile P is never copied per try: a "before" plane writes its two results to two
//    fresh patch rows behind the tile, and the host resolves, for every later read (following
//    planes, the sub-permutation gather), which row currently holds a coordinate;
//  * W holds the transformed row for the data-dependent reads after the butterflies only.
// Same operations in the same order on the same values: the hashes are bit-identical.
struct HashOp { unsigned short a, b; };                       // rows read by a plane
static constexpr int VE = 16 / sizeof(FT);                    // elements per 16-byte vector
struct __align__(16) HashVec { FT x[VE]; };
static const int HT_GSRC = 8192, HT_OPS = 1024, HT_PICK = 2048;
__constant__ unsigned short c_hash_gsrc[HT_GSRC];             // [try][d_max]  row of P feeding z[y], 0xffff = zero
__constant__ HashOp c_hash_ops[HT_OPS];                       // [try][planes] before (rows of P) then after (rows of W)
__constant__ FT c_hash_cs[2 * HT_OPS];                        // [try][planes][cos, sin]
__constant__ unsigned short c_hash_pick[HT_PICK];             // [try][d_short]

template <int DM, int TP>
__global__ void __launch_bounds__(TP)
hash_points_const_kernel(const FT *__restrict__ points, const FT *__restrict__ mean, size_t n, int d,
                         int planes_b, int planes_all, int ds, int tries, FT inv_sqrt2,
                         u32 *__restrict__ hash) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  constexpr int LD = TP + 1;
  constexpr int LEVELS = DM == 16 ? 4 : DM == 32 ? 5 : DM == 64 ? 6 : 7;
  FT *P = reinterpret_cast<FT *>(smem_raw);                   // [d + 2*planes_b][LD]
  FT *W = P + (size_t)(d + 2 * planes_b) * LD;                // [DM][LD]
  const int tid = threadIdx.x;
  const size_t tiles = (n + TP - 1) / TP;
  for (size_t tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const size_t p0 = tile * TP;
    const size_t rows = (n - p0 < (size_t)TP) ? n - p0 : (size_t)TP;
    __syncthreads();
    if ((d & (VE - 1)) == 0) {
      // the tile is one contiguous run of rows: 16-byte loads, four per thread in flight (the
      // pass is bound by this read of the points, so bytes in flight are what counts), then the
      // centred values go to the transposed tile (compute.cl:44-49)
      const int vpr = d / VE;                                  // vectors per row
      const int total = (int)rows * vpr;
      const HashVec *src = reinterpret_cast<const HashVec *>(points + p0 * (size_t)d);
      const HashVec *mu = reinterpret_cast<const HashVec *>(mean);
      for (int e0 = tid; e0 < total; e0 += 4 * TP) {
        HashVec v[4];
#pragma unroll
        for (int u = 0; u < 4; u++)
          if (e0 + u * TP < total) v[u] = src[e0 + u * TP];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const int e = e0 + u * TP;
          if (e < total) {
            const int r = e / vpr, cv = e - r * vpr;
            const HashVec m = mu[cv];
#pragma unroll
            for (int j = 0; j < VE; j++) P[(cv * VE + j) * LD + r] = v[u].x[j] - m.x[j];
          }
        }
      }
    } else {
      for (size_t r = tid >> 5; r < rows; r += TP / 32) {
        const FT *src = points + (p0 + r) * (size_t)d;
        for (int c = tid & 31; c < d; c += 32) P[c * LD + r] = src[c] - mean[c];
      }
    }
    __syncthreads();
    if ((size_t)tid >= rows) continue;
    for (int tr = 0; tr < tries; tr++) {
      const HashOp *ops = c_hash_ops + tr * planes_all;
      const FT *cs = c_hash_cs + 2 * tr * planes_all;
      for (int q = 0; q < planes_b; q++) {                     // compute.cl:55-68
        const FT c = cs[2 * q], sn = cs[2 * q + 1];
        const FT a = P[ops[q].a * LD + tid], b = P[ops[q].b * LD + tid];
        P[(d + 2 * q) * LD + tid] = a * c - b * sn;
        P[(d + 2 * q + 1) * LD + tid] = a * sn + b * c;
      }
      FT z[DM];
      const unsigned short *gsrc = c_hash_gsrc + tr * DM;
#pragma unroll
      for (int y = 0; y < DM; y++) {                           // compute.cl:77-85
        const unsigned src = gsrc[y];
        z[y] = src != 0xffffu ? P[src * LD + tid] : (FT)0;
      }
#pragma unroll
      for (int lev = 0; lev < LEVELS; lev++) {                 // compute.cl:101-122
#pragma unroll
        for (int w = 0; w < DM / 2; w++) {
          const int hi_part = (w >> lev) << lev, lo_part = w ^ hi_part;
          const int ia = (hi_part << 1) | lo_part, ib = ia | (1 << lev);
          FT a = z[ia], b = z[ib];
          FT sm = a + b, df = a - b;
          if (lev & 1) { sm *= (FT)0.5; df *= (FT)0.5; }
          if (lev == 0 && (LEVELS & 1)) { sm *= inv_sqrt2; df *= inv_sqrt2; }
          z[ia] = sm;
          z[ib] = df;
        }
      }
#pragma unroll
      for (int y = 0; y < DM; y++) W[y * LD + tid] = z[y];
      for (int q = planes_b; q < planes_all; q++) {
        const FT c = cs[2 * q], sn = cs[2 * q + 1];
        const int i = ops[q].a, j = ops[q].b;
        const FT a = W[i * LD + tid], b = W[j * LD + tid];
        W[i * LD + tid] = a * c - b * sn;
        W[j * LD + tid] = a * sn + b * c;
      }
      const unsigned short *pick = c_hash_pick + tr * ds;
      u32 h = 0;
      for (int i = 0; i < ds; i++) h = (h << 1) | sign_bit(W[pick[i] * LD + tid]);
      hash[(size_t)tr * n + p0 + tid] = h;
    }
  }
}

// tries are hashed in batches whose tables fit the constant arrays (one batch in every BASELINE
// configuration); returns false when the shape does not fit this kernel at all
template <int DM>
static bool launch_hash_const(const FT *points, const FT *mean, const annb_transform_desc *t, u32 *hash,
                              annb_stream stream) {
  constexpr int TP = 128;
  const size_t d = t->d, ds = t->d_short;
  const size_t planes_b = t->rots_before * t->rot_len_before;
  const size_t planes_all = planes_b + t->rots_after * t->rot_len_after;
  if (!t->host_perm_b || !t->host_pick || (planes_all && (!t->host_plane_idx || !t->host_plane_cs))) return false;
  const size_t smem = (d + 2 * planes_b + DM) * (size_t)(TP + 1) * sizeof(FT);
  if (smem > 200 * 1024 || d + 2 * planes_b >= 0xffff || planes_all > (size_t)HT_OPS || ds > (size_t)HT_PICK) return false;
  size_t batch = t->tries;
  while (batch > 1 && (batch * DM > (size_t)HT_GSRC || batch * planes_all > (size_t)HT_OPS || batch * ds > (size_t)HT_PICK)) batch--;
  static thread_local bool configured = false;   // per host thread = per device
  if (!configured) {
    RT_CHECK(cudaFuncSetAttribute(hash_points_const_kernel<DM, TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  static thread_local unsigned short *h_gsrc = NULL, *h_pick = NULL;
  static thread_local HashOp *h_ops = NULL;
  static thread_local FT *h_cs = NULL;
  static thread_local unsigned *latest = NULL;
  static thread_local size_t latest_cap = 0;
  if (!h_gsrc) {                                               // pinned: the async copies below read them later
    RT_CHECK(cudaMallocHost((void **)&h_gsrc, sizeof(unsigned short) * HT_GSRC));
    RT_CHECK(cudaMallocHost((void **)&h_pick, sizeof(unsigned short) * HT_PICK));
    RT_CHECK(cudaMallocHost((void **)&h_ops, sizeof(HashOp) * HT_OPS));
    RT_CHECK(cudaMallocHost((void **)&h_cs, sizeof(FT) * 2 * HT_OPS));
  }
  if (d > latest_cap) { free(latest); latest = (unsigned *)malloc(sizeof(unsigned) * d); latest_cap = d; }
  const size_t tiles = (t->n + TP - 1) / TP;
  const unsigned grid = (unsigned)(tiles < 148 * 8 ? tiles : 148 * 8);
  for (size_t t0 = 0; t0 < (size_t)t->tries; t0 += batch) {
    const size_t nb = (size_t)t->tries - t0 < batch ? (size_t)t->tries - t0 : batch;
    if (t0) RT_CHECK(cudaStreamSynchronize(stream));           // the staging arrays are about to be rewritten
    for (size_t j = 0; j < nb; j++) {
      const size_t tr = t0 + j;
      for (size_t c = 0; c < d; c++) latest[c] = (unsigned)c;  // row of P holding coordinate c right now
      for (size_t q = 0; q < planes_all; q++) {
        const u32 i = t->host_plane_idx[(tr * planes_all + q) * 2], jj = t->host_plane_idx[(tr * planes_all + q) * 2 + 1];
        HashOp &o = h_ops[j * planes_all + q];
        if (q < planes_b) {
          o.a = (unsigned short)latest[i]; o.b = (unsigned short)latest[jj];
          latest[i] = (unsigned)(d + 2 * q); latest[jj] = (unsigned)(d + 2 * q + 1);
        } else {
          o.a = (unsigned short)i; o.b = (unsigned short)jj;
        }
        h_cs[2 * (j * planes_all + q)] = t->host_plane_cs[(tr * planes_all + q) * 2];
        h_cs[2 * (j * planes_all + q) + 1] = t->host_plane_cs[(tr * planes_all + q) * 2 + 1];
      }
      for (size_t y = 0; y < (size_t)DM; y++) {
        const u32 src = t->host_perm_b[tr * DM + y];
        h_gsrc[j * DM + y] = src < d ? (unsigned short)latest[src] : (unsigned short)0xffff;
      }
      for (size_t i = 0; i < ds; i++) h_pick[j * ds + i] = (unsigned short)t->host_pick[tr * ds + i];
    }
    RT_CHECK(cudaMemcpyToSymbolAsync(c_hash_gsrc, h_gsrc, sizeof(unsigned short) * nb * DM, 0, cudaMemcpyHostToDevice, stream));
    if (ds) RT_CHECK(cudaMemcpyToSymbolAsync(c_hash_pick, h_pick, sizeof(unsigned short) * nb * ds, 0, cudaMemcpyHostToDevice, stream));
    if (planes_all) {
      RT_CHECK(cudaMemcpyToSymbolAsync(c_hash_ops, h_ops, sizeof(HashOp) * nb * planes_all, 0, cudaMemcpyHostToDevice, stream));
      RT_CHECK(cudaMemcpyToSymbolAsync(c_hash_cs, h_cs, sizeof(FT) * 2 * nb * planes_all, 0, cudaMemcpyHostToDevice, stream));
    }
    hash_points_const_kernel<DM, TP><<<grid, TP, smem, stream>>>(points, mean, t->n, (int)d, (int)planes_b, (int)planes_all,
                                                                (int)ds, (int)nb, t->inv_sqrt2, hash + t0 * t->n);
  }
  return true;
}

template <int DM>
static bool launch_hash_reg(const FT *points, const FT *mean, const annb_transform_desc *t, u32 *hash,
                            annb_stream stream) {
  constexpr int TP = 128;
  size_t smem = (t->d + DM) * (size_t)(TP + 1) * sizeof(FT);
  if (smem > 200 * 1024) return false;
  static thread_local bool configured = false;   // per host thread = per device
  if (!configured) {
    RT_CHECK(cudaFuncSetAttribute(hash_points_reg_kernel<DM, TP>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    configured = true;
  }
  size_t tiles = (t->n + TP - 1) / TP;
  unsigned grid = (unsigned)(tiles < 148 * 8 ? tiles : 148 * 8);
  hash_points_reg_kernel<DM, TP><<<grid, TP, smem, stream>>>(points, mean, *t, hash);
  return true;
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
  {
    const char *off = getenv("ANN_B200_NO_REG_HASH");
    bool done = false;
    const char *offc = getenv("ANN_B200_NO_CONST_HASH");
    if (!(off && *off && *off != '0') && !(offc && *offc && *offc != '0')) {
      switch (t->d_max) {
        case 16: done = launch_hash_const<16>(points, mean, t, hash, stream); break;
        case 32: done = launch_hash_const<32>(points, mean, t, hash, stream); break;
        case 64: done = launch_hash_const<64>(points, mean, t, hash, stream); break;
#ifdef USE_FLOAT
        case 128: done = launch_hash_const<128>(points, mean, t, hash, stream); break;
#endif
        default: break;
      }
      if (done) { LAUNCH_CHECK("hash_points_const"); return; }
    }
    if (!(off && *off && *off != '0')) {
      switch (t->d_max) {
        case 16: done = launch_hash_reg<16>(points, mean, t, hash, stream); break;
        case 32: done = launch_hash_reg<32>(points, mean, t, hash, stream); break;
        case 64: done = launch_hash_reg<64>(points, mean, t, hash, stream); break;
#ifdef USE_FLOAT
        case 128: done = launch_hash_reg<128>(points, mean, t, hash, stream); break;
#endif
        default: break;
      }
    }
    if (done) { LAUNCH_CHECK("hash_points_reg"); return; }
  }
  size_t tiles = (t->n + HASH_TP - 1) / HASH_TP;
  size_t need = hash_plane_bytes(t);
  if (need <= HASH_SMEM_LIMIT) {
    static thread_local size_t configured = 0;
    if (need > configured) {
      RT_CHECK(cudaFuncSetAttribute(hash_points_kernel<HASH_TP>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                           (int)HASH_SMEM_LIMIT));
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

extern "C" size_t annb_scan_tmp_bytes(size_t buckets);

// ---- S5 locality order ---------------------------------------------------------------------
// Rows of one supercharge chunk are worked on in the order of their leading hash bits of one
// try: neighbouring warps then gather overlapping candidate rows (points that share a bucket
// prefix share most of their neighbours' neighbours), which turns DRAM gathers into L2 hits.
// key = chunk * 2^bits + (hash >> shift); rows are grouped by key with the S2 machinery
// (histogram, scan, scatter); the order inside a key does not matter.
struct ChunkBounds { unsigned n; unsigned lo[65]; };          // chunk c = relative rows [lo[c], lo[c+1])

__global__ void locality_keys_kernel(const u32 *__restrict__ hash, size_t row_lo, size_t rows, int shift, int bits,
                                     ChunkBounds cb, u32 *__restrict__ keys) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows) return;
  unsigned c = 0;
  while (c + 1 < cb.n && (unsigned)i >= cb.lo[c + 1]) c++;
  keys[i] = (c << bits) | (hash[row_lo + i] >> shift);
}

static int locality_bits(size_t d_short) { return d_short < 17 ? (int)d_short : 17; }

extern "C" size_t annb_locality_scratch_bytes(size_t rows, size_t d_short, int chunks) {
  size_t nkeys = (size_t)chunks << locality_bits(d_short);
  return ((rows * 4 + 255) & ~(size_t)255) + ((nkeys * 4 + 255) & ~(size_t)255) + (((nkeys + 1) * 4 + 255) & ~(size_t)255) +
         ((annb_scan_tmp_bytes(nkeys) + 255) & ~(size_t)255) + 256;
}

extern "C" void annb_locality_order(const u32 *hash, size_t row_lo, size_t rows, size_t d_short, int chunks,
                                    const size_t *chunk_lo, void *scratch, u32 *perm, annb_stream stream) {
  if (rows == 0) return;
  if (chunks < 1 || chunks > 64) fatal_config("annb_locality_order: 1..64 chunks");
  const int bits = locality_bits(d_short), shift = (int)d_short - bits;
  const size_t nkeys = (size_t)chunks << bits;
  unsigned char *p = (unsigned char *)scratch;
  u32 *keys = (u32 *)p;
  p += (rows * 4 + 255) & ~(size_t)255;
  u32 *count = (u32 *)p;
  p += (nkeys * 4 + 255) & ~(size_t)255;
  u32 *offset = (u32 *)p;
  p += ((nkeys + 1) * 4 + 255) & ~(size_t)255;
  u32 *block_tot = (u32 *)p;
  ChunkBounds cb;
  cb.n = (unsigned)chunks;
  for (int c = 0; c <= chunks; c++) cb.lo[c] = (unsigned)chunk_lo[c];
  const size_t blocks = (nkeys + SCAN_ITEMS - 1) / SCAN_ITEMS;
  RT_CHECK(cudaMemsetAsync(count, 0, nkeys * sizeof(u32), stream));
  RT_CHECK(cudaMemsetAsync(block_tot + blocks, 0, sizeof(u32), stream));        // the unused "tmax" word
  locality_keys_kernel<<<grid_for(rows, 256), 256, 0, stream>>>(hash, row_lo, rows, shift, bits, cb, keys);
  LAUNCH_CHECK("locality_keys");
  histogram_kernel<<<grid_for(rows, 256), 256, 0, stream>>>(keys, rows, count);
  LAUNCH_CHECK("histogram");
  scan_blocks_kernel<<<(unsigned)blocks, SCAN_THREADS, 0, stream>>>(count, nkeys, offset, block_tot, block_tot + blocks);
  LAUNCH_CHECK("scan_blocks");
  scan_totals_kernel<<<1, 1024, 0, stream>>>(block_tot, blocks);
  LAUNCH_CHECK("scan_totals");
  scan_addback_kernel<<<(unsigned)blocks, SCAN_THREADS, 0, stream>>>(offset, nkeys, block_tot, (u32)rows);
  LAUNCH_CHECK("scan_addback");
  scatter_kernel<<<grid_for(rows, 256), 256, 0, stream>>>(keys, rows, offset, count, perm);
  LAUNCH_CHECK("scatter");
}

extern "C" size_t annb_scan_tmp_bytes(size_t buckets) {
  size_t blocks = (buckets + SCAN_ITEMS - 1) / SCAN_ITEMS;
  return (blocks + 2) * sizeof(u32);
}

extern "C" void annb_build_buckets(const u32 *hash, size_t n, size_t buckets, u32 *count,
                                   u32 *offset, u32 *order_tmp, u32 *order, u32 *tmax,
                                   void *scan_tmp, annb_stream stream) {
  u32 *block_tot = (u32 *)scan_tmp;
  size_t blocks = (buckets + SCAN_ITEMS - 1) / SCAN_ITEMS;
  RT_CHECK(cudaMemsetAsync(count, 0, buckets * sizeof(u32), stream));
  RT_CHECK(cudaMemsetAsync(tmax, 0, sizeof(u32), stream));
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

template <typename CELL>
__global__ void export_table_kernel(const u32 *__restrict__ offset, const u32 *__restrict__ order,
                                    size_t n, size_t buckets, size_t tmax, CELL *table) {
  size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= buckets * tmax) return;
  size_t b = e / tmax, z = e - b * tmax;
  u32 beg = offset[b], cnt = offset[b + 1] - beg;
  table[e] = z < cnt ? (CELL)order[beg + z] : (CELL)n;
}

extern "C" void annb_export_table(const u32 *offset, const u32 *order, size_t n, size_t buckets,
                                  size_t tmax, size_t *table, annb_stream stream) {
  export_table_kernel<size_t><<<grid_for(buckets * tmax, 256), 256, 0, stream>>>(offset, order, n, buckets, tmax, table);
  LAUNCH_CHECK("export_table");
}

extern "C" void annb_export_table32(const u32 *offset, const u32 *order, size_t n, size_t buckets,
                                    size_t tmax, u32 *table, annb_stream stream) {
  export_table_kernel<u32><<<grid_for(buckets * tmax, 256), 256, 0, stream>>>(offset, order, n, buckets, tmax, table);
  LAUNCH_CHECK("export_table32");
}

// largest bucket of a try without building its table: histogram + maximum
__global__ void max_count_kernel(const u32 *__restrict__ count, size_t buckets, u32 *tmax) {
  u32 mx = 0;
  for (size_t b = (size_t)blockIdx.x * blockDim.x + threadIdx.x; b < buckets; b += (size_t)gridDim.x * blockDim.x)
    mx = max(mx, count[b]);
  mx = __reduce_max_sync(FULL, mx);
  if ((threadIdx.x & 31) == 0 && mx) atomicMax(tmax, mx);
}

extern "C" void annb_bucket_max(const u32 *hash, size_t n, size_t buckets, u32 *count, u32 *tmax,
                                annb_stream stream) {
  RT_CHECK(cudaMemsetAsync(count, 0, buckets * sizeof(u32), stream));
  RT_CHECK(cudaMemsetAsync(tmax, 0, sizeof(u32), stream));
  histogram_kernel<<<grid_for(n, 256), 256, 0, stream>>>(hash, n, count);
  LAUNCH_CHECK("histogram");
  unsigned grid = grid_for(buckets, 256);
  if (grid > 148 * 8) grid = 148 * 8;
  max_count_kernel<<<grid, 256, 0, stream>>>(count, buckets, tmax);
  LAUNCH_CHECK("max_count");
}

