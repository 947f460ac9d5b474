// annb_leaf.cu — S3: candidate distances + per-point k best of one try.
//
// Three kernels behind annb_leaf_topk():
//   leaf_topk_tile_kernel   the fast path.  One warp per bucket: every point of a bucket
//                           has the same candidates (slots p < P of [b, b^1, b^2, b^4, ...]),
//                           so candidate rows are staged once per bucket in shared memory
//                           (cp.async, double buffered) and broadcast to lanes that each hold
//                           one query row in registers.  Distances follow the reference's
//                           summation tree exactly; each lane keeps its k best in registers
//                           and folds batches in with sorting networks.
//   leaf_topk_warp_kernel   any d, any k <= 256: one warp per point, cooperative distance.
//   leaf_literal_kernel     redoes the (rare) rows in which an exact distance tie could
//                           matter with the reference's literal row + sorting network.
#include "annb_common.cuh"
#include <cuda_fp16.h>

static __device__ unsigned long long leaf_literal_rows_dev;
static __device__ unsigned long long leaf_pairs_dev;      // (point, real candidate) pairs measured by S3
extern "C" unsigned long long annb_leaf_pairs(int reset) {
  unsigned long long v = 0, z = 0;
  cudaMemcpyFromSymbol(&v, leaf_pairs_dev, sizeof v);
  if (reset) cudaMemcpyToSymbol(leaf_pairs_dev, &z, sizeof z);
  return v;
}
unsigned long long annb_leaf_literal_count(int reset) {
  unsigned long long v = 0, z = 0;
  cudaMemcpyFromSymbol(&v, leaf_literal_rows_dev, sizeof v);
  if (reset) cudaMemcpyToSymbol(leaf_literal_rows_dev, &z, sizeof z);
  return v;
}

// =====================================================================================
// generic: one warp per point
// =====================================================================================

template <int E, int R>
__global__ void __launch_bounds__(256)
leaf_topk_warp_kernel(const FT *__restrict__ sp, const u32 *__restrict__ order,
                      const u32 *__restrict__ offset, const u32 *__restrict__ hash,
                      const u32 *__restrict__ tmax_p, size_t n, int d, int d_short, int k,
                      u32 *__restrict__ list_ids, FT *__restrict__ list_dist, TieList ties) {
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
  bool tie = false;
  unsigned npairs = 0;

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
        dist = __shfl_sync(FULL, warp_sqdist<(E ? E : 1)>(q, cr, d), 0);
      } else {
        dist = generic_sqdist(qrow, sp + row * (size_t)d, d, tmp, lane);
      }
      consider<R>(best, tau, dist, order[row], k, sentinel, lane, tie);
      npairs++;
    }
  }
  if (lane == 0) atomicAdd(&leaf_pairs_dev, (unsigned long long)npairs);
#pragma unroll
  for (int rr = 0; rr < R; rr++) {
    int p = rr * 32 + lane;
    if (p < k) {
      list_ids[(size_t)me * k + p] = best.id[rr];
      list_dist[(size_t)me * k + p] = best.v[rr];
    }
  }
  if (tie && lane == 0) tie_report(ties, me);
}

// =====================================================================================
// fast path: one warp per bucket, queries in registers, candidates broadcast from smem
// =====================================================================================

static constexpr int VW = 16 / sizeof(FT);          // elements per 16-byte vector
static constexpr int TILE_CH = 16;                  // candidate rows per staged chunk
static constexpr int TILE_WARPS = 1;            // one bucket per CTA: a finished bucket frees its registers at once

struct __align__(16) Vec16 { FT x[VW]; };

// Partial sums of the reference's tree (compute.cl:160-167), VW lanes wide.  Node (LEN, OFF)
// is elements OFF..OFF+VW-1 of the array at the stage where it has LEN entries:
// V_LEN[z] = V_2LEN[z] + V_2LEN[z + LEN]; the leaves (LEN == D) are squared differences.
template <int D, int LEN, int OFF>
struct TreeNode {
  static __device__ __forceinline__ void eval(const FT (&q)[D], const FT *crow, FT (&out)[VW]) {
    FT a[VW], b[VW];
    TreeNode<D, 2 * LEN, OFF>::eval(q, crow, a);
    TreeNode<D, 2 * LEN, OFF + LEN>::eval(q, crow, b);
#pragma unroll
    for (int w = 0; w < VW; w++) out[w] = a[w] + b[w];
  }
};
template <int D, int OFF>
struct TreeNode<D, D, OFF> {
  static __device__ __forceinline__ void eval(const FT (&q)[D], const FT *crow, FT (&out)[VW]) {
    Vec16 c = *reinterpret_cast<const Vec16 *>(crow + OFF);
#pragma unroll
    for (int w = 0; w < VW; w++) {
      FT df = q[OFF + w] - c.x[w];
      out[w] = df * df;
    }
  }
};

template <int D>
__device__ __forceinline__ FT tile_sqdist(const FT (&q)[D], const FT *crow) {
  FT v[VW];
  TreeNode<D, VW, 0>::eval(q, crow, v);
#pragma unroll
  for (int h = VW / 2; h >= 1; h >>= 1)
#pragma unroll
    for (int w = 0; w < h; w++) v[w] = v[w] + v[w + h];
  return v[0];
}

#ifdef USE_FLOAT
// Packed FP32x2 arithmetic (Blackwell FADD2 / FFMA2): two IEEE round-to-nearest results per
// instruction, i.e. the same bits as the scalar operations at half the issue slots.
// ptxas contracts mul.rn.f32x2 + add.rn.f32x2 into one FFMA2 even under -fmad=false, which
// would change the rounding; so the square is written as an explicit fma with a -0.0 addend
// (x*x + (-0.0) rounds exactly like x*x), leaving no multiply for the adds to absorb.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// nz must hold (-0.0f, -0.0f) and must NOT be a compile-time constant: ptxas would turn the
// fma back into a multiply and contract it with the next add.  It arrives as a kernel argument.
__device__ __forceinline__ f32x2 sqr2(f32x2 a, f32x2 nz) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %1, %2;" : "=l"(r) : "l"(a), "l"(nz));
  return r;
}

// Same tree as TreeNode/tile_sqdist on pairs: P_LEN[i] = P_2LEN[i] + P_2LEN[i + LEN/2],
// elements (2i, 2i+1) packed; the last level adds the two halves of the remaining pair.
template <int D, int LEN, int OFF>   // LEN, OFF in elements; node = one 16-byte vector (2 pairs)
struct PackedNode {
  static __device__ __forceinline__ void eval(const f32x2 (&q2)[D / 2], const FT *crow, f32x2 nz, f32x2 (&out)[2]) {
    f32x2 a[2], b[2];
    PackedNode<D, 2 * LEN, OFF>::eval(q2, crow, nz, a);
    PackedNode<D, 2 * LEN, OFF + LEN>::eval(q2, crow, nz, b);
    out[0] = add2(a[0], b[0]);
    out[1] = add2(a[1], b[1]);
  }
};
template <int D, int OFF>
struct PackedNode<D, D, OFF> {
  static __device__ __forceinline__ void eval(const f32x2 (&q2)[D / 2], const FT *crow, f32x2 nz, f32x2 (&out)[2]) {
    ulonglong2 c = *reinterpret_cast<const ulonglong2 *>(crow + OFF);
    out[0] = sqr2(sub2(q2[OFF / 2], c.x), nz);
    out[1] = sqr2(sub2(q2[OFF / 2 + 1], c.y), nz);
  }
};
template <int D>
__device__ __forceinline__ FT tile_sqdist_packed(const f32x2 (&q2)[D / 2], const FT *crow, f32x2 nz) {
  f32x2 v[2];
  PackedNode<D, 4, 0>::eval(q2, crow, nz, v);                      // V_4[0..3] as two pairs
  f32x2 h = add2(v[0], v[1]);                                       // V_2[0..1]
  float lo, hi;
  asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(h));
  return lo + hi;                                                   // V_1[0]
}
#endif

#define CE_ASC(da, ia, db, ib)                                    \
  {                                                               \
    bool sw_ = (da) > (db);                                       \
    FT lo_ = sw_ ? (db) : (da), hi_ = sw_ ? (da) : (db);          \
    u32 li_ = sw_ ? (ib) : (ia), hj_ = sw_ ? (ia) : (ib);         \
    (da) = lo_; (db) = hi_; (ia) = li_; (ib) = hj_;               \
  }

// list <- the KC smallest of list ∪ batch (B <= KC entries), ascending.  `tie` is raised when
// the smallest discarded value equals the largest kept one (an exact tie on the boundary).
template <int KC, int B>
__device__ __forceinline__ void fold_batch(FT (&ld)[KC], u32 (&li)[KC], FT (&bd)[B], u32 (&bi)[B],
                                           bool sort_batch, bool &tie) {
  if (sort_batch) {
#pragma unroll
    for (int kk = 2; kk <= B; kk <<= 1)
#pragma unroll
      for (int j = kk >> 1; j > 0; j >>= 1)
#pragma unroll
        for (int i = 0; i < B; i++) {
          int l = i ^ j;
          if (l > i) {
            if ((i & kk) == 0) CE_ASC(bd[i], bi[i], bd[l], bi[l])
            else CE_ASC(bd[l], bi[l], bd[i], bi[i])
          }
        }
  }
  // the batch, reversed, against the tail of the list: the minima are the KC smallest and
  // form a bitonic sequence with the untouched head of the list
  FT mdisc = ft_inf();
#pragma unroll
  for (int i = KC - B; i < KC; i++) {
    FT a = ld[i], b = bd[KC - 1 - i];
    bool tb = b < a;
    mdisc = fmin(mdisc, tb ? a : b);
    ld[i] = tb ? b : a;
    li[i] = tb ? bi[KC - 1 - i] : li[i];
  }
#pragma unroll
  for (int j = KC >> 1; j > 0; j >>= 1)
#pragma unroll
    for (int i = 0; i < KC; i++) {
      int l = i ^ j;
      if (l > i) CE_ASC(ld[i], li[i], ld[l], li[l])
    }
  if (mdisc == ld[KC - 1] && mdisc != ft_inf()) tie = true;
}

__device__ __forceinline__ void cp_async16(void *smem_dst, const void *gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async4(void *smem_dst, const void *gsrc) {
  unsigned s = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(s), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N)); }

template <int D, int KC, int B>
struct TileSmem {
  static constexpr int RS = D + VW;                                   // padded row stride
  static constexpr size_t rows_bytes = 2ull * TILE_CH * RS * sizeof(FT);
  static constexpr size_t cid_bytes = 2ull * TILE_CH * sizeof(u32);
  static constexpr size_t batch_bytes = (size_t)B * 32 * (sizeof(FT) + sizeof(u32));
  static constexpr size_t seg_bytes = (33 + 32) * sizeof(u32);
  static constexpr size_t per_warp = (rows_bytes + cid_bytes + batch_bytes + seg_bytes + 15) & ~(size_t)15;
};

template <int D, int KC, int B, int REGS>
__global__ void __maxnreg__(REGS)
leaf_topk_tile_kernel(const FT *__restrict__ sp, const u32 *__restrict__ order,
                      const u32 *__restrict__ offset, const u32 *__restrict__ tmax_p, size_t n,
                      size_t buckets, int d_short, int k, u32 *__restrict__ list_ids,
                      FT *__restrict__ list_dist, TieList ties, int pack_tries, int max_slices,
                      unsigned long long negzero2, u32 *__restrict__ ticket,
                      const u32 *__restrict__ bucket_list, const u32 *__restrict__ bucket_count) {
  typedef TileSmem<D, KC, B> SM;
  constexpr int RS = SM::RS;
  constexpr int PPR = D / VW;                                          // 16-byte pieces per row
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  unsigned char *mine = smem_raw + (size_t)wib * SM::per_warp;
  FT *rows = reinterpret_cast<FT *>(mine);                             // [2][CH][RS]
  u32 *cids = reinterpret_cast<u32 *>(mine + SM::rows_bytes);          // [2][CH]
  FT *bdist = reinterpret_cast<FT *>(mine + SM::rows_bytes + SM::cid_bytes);   // [B][32]
  u32 *bidx = reinterpret_cast<u32 *>(bdist + B * 32);                 // [B][32]
  u32 *segpos = bidx + B * 32;                                         // [33]
  u32 *segrow = segpos + 33;                                           // [32]

  const u32 sentinel = (u32)n;
  // persistent warps: buckets are handed out by an atomic ticket, so a warp that drew a light
  // bucket immediately takes another one
  for (;;) {
  size_t b = 0;
  if (lane == 0) b = atomicAdd(ticket, 1u);
  b = __shfl_sync(FULL, (u32)b, 0);
  // with a bucket list (the buckets the screened kernel handed over) tickets index the list
  if (b >= (bucket_list ? (size_t)*bucket_count : buckets)) return;
  if (bucket_list) b = bucket_list[b];
  const u32 beg = offset[b], Q = offset[b + 1] - beg;
  if (Q == 0) continue;
  const unsigned long long tmax = *tmax_p;
  const unsigned long long L = (unsigned long long)(d_short + 1) * tmax;
  const unsigned long long P = 1ull << floor_log2_u(L);

  // candidate segments: lane y describes bucket b ^ f_y (compute.cl:238-246 + prefix rule)
  {
    u32 inc = 0, srow = 0;
    unsigned long long first_slot = (unsigned long long)lane * tmax;
    if (lane <= d_short && first_slot < P) {
      u32 cb = (u32)b ^ (lane ? (1u << (lane - 1)) : 0u);
      u32 sb = offset[cb], cnt = offset[cb + 1] - sb;
      unsigned long long room = P - first_slot;
      inc = cnt < room ? cnt : (u32)room;
      srow = sb;
    }
    u32 incl = inc;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      u32 up = __shfl_up_sync(FULL, incl, o);
      if (lane >= o) incl += up;
    }
    segpos[lane] = incl - inc;
    segrow[lane] = srow;
    if (lane == 31) segpos[32] = incl;
  }
  __syncwarp();
  const u32 C = segpos[32];
  if (lane == 0) atomicAdd(&leaf_pairs_dev, (unsigned long long)Q * (C - 1));   // self excluded

  // Queries are packed as (query, slice): S slices of the candidate stream per query, so that
  // Qp*S lanes work.  The bucket is cut into `passes` groups of qpp queries where that
  // fills the warp better (cost ~ passes / S candidate scans).
  int passes = (int)((Q + 31) / 32), S = 1;
  {
    float best = 1e30f;
    const int t0 = passes;
    for (int t = t0; t < t0 + pack_tries; t++) {
      int qpp_t = (int)((Q + t - 1) / t);
      int s_t = min(max_slices, 32 / qpp_t);
      float cost = (float)t / (float)s_t;
      if (cost < best - 1e-6f) { best = cost; passes = t; S = s_t; }
    }
  }
  const u32 qpp = (Q + passes - 1) / passes;
  const int per = TILE_CH / S;                                     // candidates per lane per chunk
  const u32 CHS = (u32)(S * per);                                   // rows staged per chunk
  const u32 nchunks = (C + CHS - 1) / CHS;

  for (u32 qbase = 0; qbase < Q; qbase += qpp) {
    const u32 Qp = min(qpp, Q - qbase);
    const bool active = lane < (int)(Qp * S);
    const int s = active ? lane / (int)Qp : 0;
    const int qi = active ? lane - s * (int)Qp : 0;
    const size_t my_row = (size_t)beg + qbase + qi;
    const u32 my_id = order[my_row];

#ifdef USE_FLOAT
    f32x2 q[D / 2];                                                 // the query row, packed pairs
    {
      const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(sp + my_row * (size_t)D);
#pragma unroll
      for (int v = 0; v < PPR; v++) {
        ulonglong2 t = src[v];
        q[2 * v] = t.x;
        q[2 * v + 1] = t.y;
      }
    }
#else
    FT q[D];
    {
      const Vec16 *src = reinterpret_cast<const Vec16 *>(sp + my_row * (size_t)D);
#pragma unroll
      for (int v = 0; v < PPR; v++) {
        Vec16 t = src[v];
#pragma unroll
        for (int w = 0; w < VW; w++) q[v * VW + w] = t.x[w];
      }
    }
#endif
    FT ld[KC];
    u32 li[KC];
#pragma unroll
    for (int i = 0; i < KC; i++) { ld[i] = ft_inf(); li[i] = sentinel; }
    bool tie = false;
    int cnt = 0;
    int seg_y = 0;                                                  // this lane's segment cursor

    auto stage_chunk = [&](u32 c, int buf) {
      // lanes 0..15 locate the rows of this chunk; every lane then copies 16-byte pieces
      const u32 slot = lane & (TILE_CH - 1);
      u32 j = c * CHS + slot;
      const bool live = slot < CHS && j < C;
      u32 grow = 0;
      if (live) {
        while (j >= segpos[seg_y + 1]) seg_y++;                   // chunks advance monotonically
        grow = segrow[seg_y] + (j - segpos[seg_y]);
      }
      if (lane < TILE_CH) {
        if (live) cp_async4(&cids[buf * TILE_CH + lane], order + grow);
        else cids[buf * TILE_CH + lane] = sentinel;
      }
      FT *dst = rows + (size_t)buf * TILE_CH * RS;
#pragma unroll
      for (int it = 0; it < (TILE_CH * PPR + 31) / 32; it++) {
        int p = lane + 32 * it;
        int r = p / PPR, col = p - r * PPR;
        u32 gr = __shfl_sync(FULL, grow, r & (TILE_CH - 1));
        if (p < (int)CHS * PPR) cp_async16(dst + r * RS + col * VW, sp + (size_t)gr * D + col * VW);
      }
      cp_async_commit();
    };

    if (nchunks) stage_chunk(0, 0);
    for (u32 c = 0; c < nchunks; c++) {
      const int buf = c & 1;
      if (c + 1 < nchunks) { stage_chunk(c + 1, buf ^ 1); cp_async_wait<1>(); }
      else cp_async_wait<0>();
      __syncwarp();
      const FT *crows = rows + (size_t)buf * TILE_CH * RS;
      // runs of candidates between flushes: the inner loop is distance + append only
      for (int i = 0; i < per;) {
        const int run = min(per - i, B - cnt);
        for (int e = 0; e < run; e++, i++) {
          const int j = s + S * i;
          const u32 cid = cids[buf * TILE_CH + j];
#ifdef USE_FLOAT
          FT dist = tile_sqdist_packed<D>(q, crows + j * RS, negzero2);
#else
          FT dist = tile_sqdist<D>(q, crows + j * RS);
#endif
          if (cid == sentinel || cid == my_id) dist = ft_inf();      // pad / self (compute.cl:144-149)
          bdist[cnt * 32 + lane] = dist;
          bidx[cnt * 32 + lane] = cid;
          cnt++;
        }
        if (cnt == B) {                                              // cnt is warp-uniform
          FT bd[B];
          u32 bi[B];
#pragma unroll
          for (int e = 0; e < B; e++) { bd[e] = bdist[e * 32 + lane]; bi[e] = bidx[e * 32 + lane]; }
          fold_batch<KC, B>(ld, li, bd, bi, true, tie);
          cnt = 0;
        }
      }
      __syncwarp();
    }
    if (cnt > 0) {
      FT bd[B];
      u32 bi[B];
#pragma unroll
      for (int e = 0; e < B; e++) {
        bool have = e < cnt;
        bd[e] = have ? bdist[e * 32 + lane] : ft_inf();
        bi[e] = have ? bidx[e * 32 + lane] : sentinel;
      }
      fold_batch<KC, B>(ld, li, bd, bi, true, tie);
      cnt = 0;
    }
    // fold the slices of each query together (lists are sorted: no batch sort needed)
    for (int sc = S; sc > 1;) {
      const int hs = (sc + 1) >> 1;                                 // slices s < sc-hs take slice s+hs
      FT bd[KC];
      u32 bi[KC];
      int src = lane + hs * (int)Qp;
      src = src < 32 ? src : lane;
#pragma unroll
      for (int i = 0; i < KC; i++) {
        bd[i] = __shfl_sync(FULL, ld[i], src);
        bi[i] = __shfl_sync(FULL, li[i], src);
      }
      bool other_tie = __shfl_sync(FULL, (int)tie, src);
      if (s + hs < sc) {
        fold_batch<KC, KC>(ld, li, bd, bi, false, tie);
        tie |= other_tie;
      }
      sc = hs;
    }
    if (active && s == 0) {
#pragma unroll
      for (int i = 0; i < KC; i++)
        if (i < k) {
          list_ids[(size_t)my_id * k + i] = li[i];
          list_dist[(size_t)my_id * k + i] = ld[i];
          if (i + 1 < KC && ld[i] == ld[i + 1] && ld[i] != ft_inf()) tie = true;
        }
      if (tie) tie_report(ties, my_id);
    }
    __syncwarp();
  }
  }   // next ticket
}

#ifdef USE_FLOAT
#include "annb_screen_common.cuh"
#include "annb_leaf_screen.cuh"
#endif


// =====================================================================================
// literal rows for reported points (exact ties)
// =====================================================================================
// One CTA per reported point: builds the reference's candidate row (all (d_short+1)*tmax
// slots, pads and self at +inf), runs its sorting network / duplicate rule / sorting
// network, and rewrites the point's list.  Row storage is one slab per CTA in global
// scratch; CTAs without a slab do nothing, and if not even one fits, *status is set.

template <int E>
__global__ void __launch_bounds__(256)
leaf_literal_kernel(const FT *__restrict__ sp, const u32 *__restrict__ order,
                    const u32 *__restrict__ offset, const u32 *__restrict__ hash,
                    const u32 *__restrict__ rank_of, const u32 *__restrict__ tmax_p, size_t n, int d,
                    int d_short, int k, u32 *__restrict__ list_ids, FT *__restrict__ list_dist,
                    TieList ties, unsigned char *slabs, size_t slab_bytes, int *status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ u32 s_beg[33], s_cnt[33], s_pos[34];
  const int tid = threadIdx.x, wib = tid >> 5;
  FT *tmp = reinterpret_cast<FT *>(smem_raw) + (size_t)wib * (E == 0 ? d : 0);
  const u32 total = *ties.count;
  if (total == 0) return;
  const unsigned long long tmax = *tmax_p;
  const size_t L = (size_t)(d_short + 1) * tmax;
  const size_t slab = (L * (sizeof(FT) + 3 * sizeof(u32)) + 15) & ~(size_t)15;
  const size_t fit = slab ? slab_bytes / slab : 0;
  if (fit == 0) { if (blockIdx.x == 0 && tid == 0) *status = 1; return; }
  const u32 workers = (u32)(fit < gridDim.x ? fit : gridDim.x);
  if (blockIdx.x >= workers) return;
  if (blockIdx.x == 0 && tid == 0) atomicAdd(&leaf_literal_rows_dev, (unsigned long long)total);
  FT *key = reinterpret_cast<FT *>(slabs + (size_t)blockIdx.x * slab);
  u32 *ids = reinterpret_cast<u32 *>(key + L);
  u32 *crow = ids + L, *cslot = crow + L;
  const u32 sentinel = (u32)n;

  for (u32 it = blockIdx.x; it < total; it += workers) {
    const u32 x = ties.rows[it];
    const u32 h = hash[x];
    const u32 xr = rank_of[x];                                    // sorted position of x
    if (tid <= d_short) {
      u32 b = h ^ (tid ? (1u << (tid - 1)) : 0u);
      s_beg[tid] = offset[b];
      s_cnt[tid] = offset[b + 1] - offset[b];
    }
    __syncthreads();
    if (tid == 0) {
      u32 pos = 0;
      for (int y = 0; y <= d_short; y++) { s_pos[y] = pos; pos += s_cnt[y]; }
      s_pos[d_short + 1] = pos;
    }
    __syncthreads();
    for (size_t slot = tid; slot < L; slot += blockDim.x) {
      u32 y = (u32)(slot / tmax), z = (u32)(slot - (size_t)y * tmax);
      bool real = z < s_cnt[y];
      ids[slot] = real ? order[s_beg[y] + z] : sentinel;
      key[slot] = ft_inf();                                       // pads, and self below
      if (real) {
        u32 f = s_pos[y] + z;
        crow[f] = s_beg[y] + z;
        cslot[f] = (u32)slot;
      }
    }
    __syncthreads();
    block_row_distances<E>(sp + (size_t)xr * d, sp, d, s_pos[d_short + 1], tmp,
                           [&](u32 i) { return (size_t)crow[i]; },
                           [&](u32 i, FT dist) { if (crow[i] != xr) key[cslot[i]] = dist; });
    __syncthreads();
    block_sort_and_uniq(ids, key, (int)L);
    for (int i = tid; i < k; i += blockDim.x) {
      list_ids[(size_t)x * k + i] = ids[i];
      list_dist[(size_t)x * k + i] = key[i];
    }
    __syncthreads();
  }
}

// rank_of[order[r]] = r
__global__ void invert_order_kernel(const u32 *__restrict__ order, size_t n, u32 *rank_of) {
  size_t r = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r < n) rank_of[order[r]] = (u32)r;
}

// =====================================================================================
// launcher
// =====================================================================================

template <int E>
static void launch_warp_r(int regs, dim3 grid, dim3 block, size_t smem, annb_stream stream,
                          const FT *sp, const u32 *order, const u32 *offset, const u32 *hash,
                          const u32 *tmax, size_t n, int d, int d_short, int k, u32 *ids, FT *dist,
                          TieList flags) {
#define WARP_CASE(R)                                                                             \
  {                                                                                              \
    if (smem > 48 * 1024)                                                                        \
      RT_CHECK(cudaFuncSetAttribute(leaf_topk_warp_kernel<E, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    leaf_topk_warp_kernel<E, R><<<grid, block, smem, stream>>>(sp, order, offset, hash, tmax, n, d, d_short, k, ids, dist, flags); \
  }
  switch (regs) {
    case 1: WARP_CASE(1) break;
    case 2: WARP_CASE(2) break;
    case 4: WARP_CASE(4) break;
    default: WARP_CASE(8) break;
  }
#undef WARP_CASE
}

template <int D, int KC, int B, int REGS>
static void launch_tile(annb_stream stream, const FT *sp, const u32 *order, const u32 *offset,
                        const u32 *tmax, size_t n, size_t buckets, int d_short, int k, u32 *ids,
                        FT *dist, TieList flags, const u32 *bucket_list, const u32 *bucket_count) {
  size_t smem = TileSmem<D, KC, B>::per_warp * TILE_WARPS;
  static thread_local bool configured = false;   // per host thread = per device
  if (!configured) {
    RT_CHECK(cudaFuncSetAttribute(leaf_topk_tile_kernel<D, KC, B, REGS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    configured = true;
  }
  unsigned grid = (unsigned)((buckets + TILE_WARPS - 1) / TILE_WARPS);
  {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    unsigned resident = (unsigned)sms * 16;                 // more CTAs than can be resident: all SMs stay full
    if (grid > resident) grid = resident;
  }
  u32 *ticket = flags.count + (bucket_list ? 33 : 32);      // spare words of the tie-list header
  RT_CHECK(cudaMemsetAsync(ticket, 0, sizeof(u32), stream));
  static int pack_tries = -1, max_slices = 16;
  if (pack_tries < 0) {
    const char *e1 = getenv("ANN_B200_TILE_PASSES"), *e2 = getenv("ANN_B200_TILE_SLICES");
    pack_tries = e1 && *e1 ? atoi(e1) : 3;
    max_slices = e2 && *e2 ? atoi(e2) : 4;
    if (pack_tries < 1) pack_tries = 1;
    if (max_slices < 1) max_slices = 1;
    if (max_slices > 16) max_slices = 16;
  }
  leaf_topk_tile_kernel<D, KC, B, REGS><<<grid, TILE_WARPS * 32, smem, stream>>>(sp, order, offset, tmax, n, buckets, d_short, k, ids, dist, flags, pack_tries, max_slices, 0x8000000080000000ull, ticket, bucket_list, bucket_count);
}

// returns false when no tiled instantiation covers (d, k)
static bool try_launch_tile(annb_stream stream, const FT *sp, const u32 *order, const u32 *offset,
                            const u32 *tmax, size_t n, size_t buckets, size_t d, int d_short,
                            size_t k, u32 *ids, FT *dist, TieList flags,
                            const u32 *bucket_list = NULL, const u32 *bucket_count = NULL) {
  const char *off = getenv("ANN_B200_NO_TILE");
  if (off && *off && *off != '0' && !bucket_list) return false;
  if (d_short > 31) return false;
#define TILE_CASE(DD, KK, BB, MB) { launch_tile<DD, KK, BB, MB>(stream, sp, order, offset, tmax, n, buckets, d_short, (int)k, ids, dist, flags, bucket_list, bucket_count); return true; }
  if (k <= 16) {
    if (d == 16) TILE_CASE(16, 16, 16, 128)
    if (d == 32) TILE_CASE(32, 16, 16, 128)
#ifdef USE_FLOAT
    if (d == 64) TILE_CASE(64, 16, 16, 168)      // 12 resident warps/SM; lower caps spill and lose (profiles/)
#endif
  }
#ifdef USE_FLOAT
  else if (k <= 32) {
    if (d == 16) TILE_CASE(16, 32, 32, 168)
    if (d == 32) TILE_CASE(32, 32, 32, 200)
  }
#endif
#undef TILE_CASE
  return false;
}

static size_t literal_area_bytes(size_t n) { return 3 * n * sizeof(u32) + (64u << 20) + 1024; }
static size_t align256(size_t v) { return (v + 255) & ~(size_t)255; }
static bool screen_covers(size_t d, size_t d_short, size_t k) {
#ifdef USE_FLOAT
  return (d == 16 || d == 32 || d == 64) && k <= 16 && d_short >= 1 && d_short <= 31;
#else
  (void)d; (void)d_short; (void)k;
  return false;
#endif
}
#ifndef USE_FLOAT
extern "C" void annb_leaf_screen_mode(int on) { (void)on; }
extern "C" unsigned long long annb_leaf_exact_pairs(int reset) { (void)reset; return 0; }
extern "C" unsigned long long annb_leaf_overflow_buckets(int reset) { (void)reset; return 0; }
extern "C" int annb_screen_applies(size_t d, size_t d_short, size_t k) { (void)d; (void)d_short; (void)k; return 0; }
extern "C" void annb_screen_scale(const FT *, const FT *, size_t, size_t, unsigned *, annb_stream) {
  fatal_config("screened leaf path in the double build");
}
extern "C" void annb_gather_rows_screen(const FT *, const u32 *, size_t, size_t, const FT *, const unsigned *, FT *,
                                        void *, annb_stream) {
  fatal_config("screened leaf path in the double build");
}
extern "C" void annb_screen_prep_points(const FT *, const FT *, size_t, size_t, const unsigned *, void *, void *, annb_stream) {
  fatal_config("fp16 screen in the double build");
}
#endif
// fp16 copy, norms, scale word and the overflow bucket list of the screened path
static size_t screen_area_bytes(size_t n, size_t d, size_t d_short, size_t k) {
  if (!screen_covers(d, d_short, k)) return 0;
  return align256(n * d * 2) + align256(n * 8) + 256 + align256(((size_t)4 << d_short) + 4);
}
extern "C" size_t annb_leaf_scratch_bytes(size_t n, size_t d, size_t d_short, size_t k) {
  return align256(literal_area_bytes(n)) + screen_area_bytes(n, d, d_short, k);
}

#ifdef USE_FLOAT
// The screened path is the default where it applies; ANN_B200_SCREEN=0 (or annb_leaf_screen_mode(0))
// sends everything through the tiled kernel instead.  Both produce the same lists.
static int screen_mode = -1;
extern "C" void annb_leaf_screen_mode(int on) { screen_mode = on ? 1 : 0; }
static int screen_enabled() {
  if (screen_mode < 0) {
    const char *e = getenv("ANN_B200_SCREEN");
    screen_mode = (e && *e) ? (*e != '0') : 1;
    const char *nt = getenv("ANN_B200_NO_TILE");            // asking for the generic kernel switches both off
    if (nt && *nt && *nt != '0') screen_mode = 0;
  }
  return screen_mode;
}

struct ScreenArea {                                           // carved from the end of the leaf scratch
  unsigned short *sp16;                                       // [n][d] fp16
  float2 *nrm;                                                // [n]
  unsigned *words;                                            // [0] max |x - mean| when not supplied, [1] overflow count
  u32 *overflow_buckets;
};
static ScreenArea screen_area(void *scratch, size_t n, size_t d) {
  unsigned char *area = (unsigned char *)scratch + align256(literal_area_bytes(n));
  ScreenArea a;
  a.sp16 = reinterpret_cast<unsigned short *>(area);
  a.nrm = reinterpret_cast<float2 *>(area + align256(n * d * 2));
  a.words = reinterpret_cast<unsigned *>(area + align256(n * d * 2) + align256(n * 8));
  a.overflow_buckets = a.words + 64;
  return a;
}
static void launch_maxabs(annb_stream stream, const FT *points, const FT *mean, size_t n, size_t d, unsigned *bits) {
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  RT_CHECK(cudaMemsetAsync(bits, 0, sizeof(unsigned), stream));
  screen_maxabs_kernel<<<sms * 8, 256, 0, stream>>>(points, mean, n, (int)d, bits);
  LAUNCH_CHECK("screen_maxabs");
}

extern "C" int annb_screen_applies(size_t d, size_t d_short, size_t k) {
  return screen_enabled() && screen_covers(d, d_short, k);
}
// the fp16 scale of the whole set (the same for every try): once per precomp
extern "C" void annb_screen_scale(const FT *points, const FT *mean, size_t n, size_t d, unsigned *scale_bits,
                                  annb_stream stream) {
  launch_maxabs(stream, points, mean, n, d, scale_bits);
}
// S2's gather fused with the preparation of the screened path (fp16 copy + norms into `leaf_scratch`)
extern "C" void annb_gather_rows_screen(const FT *points, const u32 *order, size_t n, size_t d, const FT *mean,
                                        const unsigned *scale_bits, FT *sorted_points, void *leaf_scratch,
                                        annb_stream stream) {
  ScreenArea a = screen_area(leaf_scratch, n, d);
  screen_prep_kernel<<<grid_for(n * (d / 4), 256), 256, 0, stream>>>(points, order, sorted_points, mean, n, (int)d,
                                                                    scale_bits, a.sp16, a.nrm);
  LAUNCH_CHECK("gather_rows_screen");
}

// fp16 copy of the points in original order (the screened supercharge reads it)
extern "C" void annb_screen_prep_points(const FT *points, const FT *mean, size_t n, size_t d,
                                        const unsigned *scale_bits, void *points16, void *nrm, annb_stream stream) {
  if (d % 4 || d / 4 > 32 || ((d / 4) & (d / 4 - 1))) fatal_config("annb_screen_prep_points: d must be 4..128, a power of two");
  screen_prep_kernel<<<grid_for(n * (d / 4), 256), 256, 0, stream>>>(points, NULL, NULL, mean, n, (int)d, scale_bits,
                                                                    (unsigned short *)points16, (float2 *)nrm);
  LAUNCH_CHECK("screen_prep_points");
}

template <int D>
static void launch_screen(annb_stream stream, const FT *sp, const FT *mean, void *scratch,
                          const unsigned *scale_bits, int rows_prepared,
                          const u32 *order, const u32 *offset, const u32 *tmax, size_t n,
                          size_t buckets, int d_short, int k, u32 *ids, FT *dist, TieList flags,
                          ScreenOverflow *ovf_out, const FT *cutoff) {
  ScreenArea a = screen_area(scratch, n, D);
  unsigned short *sp16 = a.sp16;
  float2 *nrm = a.nrm;
  ScreenOverflow ovf;
  ovf.count = a.words + 1;
  ovf.buckets = a.overflow_buckets;
  RT_CHECK(cudaMemsetAsync(ovf.count, 0, sizeof(u32), stream));
  int dev = 0, sms = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (cutoff && !(rows_prepared && scale_bits)) fatal_config("a leaf cutoff needs the prepared rows and their scale word");
  if (!rows_prepared) {
    if (!scale_bits) {
      launch_maxabs(stream, sp, mean, n, D, a.words);
      scale_bits = a.words;
    }
    screen_prep_kernel<<<grid_for(n * (D / 4), 256), 256, 0, stream>>>(sp, NULL, NULL, mean, n, D, scale_bits, sp16, nrm);
    LAUNCH_CHECK("screen_prep");
  }
  // shared-memory tables sized a little above the expected candidate count (cfg3: 259 +- 16 -> 328);
  // the few bigger buckets go to the tiled kernel, and the smaller footprint buys two more warps per SM
  double expect = (double)(d_short + 1) * ((double)n / (double)buckets);
  size_t want = (size_t)(1.1 * expect) + 40;
  int ct = (int)(((want + 55) / 64) * 64 + 8);               // = 8 mod 64: conflict-free parking of the bounds
  if (ct > 1032) ct = 1032;
  if (ct < screen_min_ct(ScreenOverlay<D>::bytes)) ct = screen_min_ct(ScreenOverlay<D>::bytes);
  size_t smem = screen_smem_bytes(ct);
  static thread_local bool configured = false;   // per host thread = per device
  if (!configured) {
    RT_CHECK(cudaFuncSetAttribute(leaf_screen_kernel<D>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)screen_smem_bytes(1032)));
    configured = true;
  }
  unsigned grid = (unsigned)buckets;
  unsigned resident = (unsigned)sms * 16;
  if (grid > resident) grid = resident;
  u32 *ticket = flags.count + 32;
  RT_CHECK(cudaMemsetAsync(ticket, 0, sizeof(u32), stream));
  leaf_screen_kernel<D><<<grid, 32, smem, stream>>>(sp, sp16, nrm, order, offset, tmax, n, buckets, d_short, k, ids, dist, flags, 0x8000000080000000ull, ticket, ct, ovf, cutoff, scale_bits);
  LAUNCH_CHECK("leaf_screen");
  *ovf_out = ovf;
}

// returns false when the screened path does not cover (d, k) or is switched off
static bool try_launch_screen(annb_stream stream, const FT *sp, const FT *mean, void *scratch,
                              const unsigned *scale_bits, int rows_prepared,
                              const u32 *order, const u32 *offset, const u32 *tmax, size_t n,
                              size_t buckets, size_t d, int d_short, size_t k, u32 *ids, FT *dist,
                              TieList flags, const FT *cutoff) {
  if (!screen_enabled() || !screen_covers(d, (size_t)d_short, k)) return false;
  ScreenOverflow ovf;
#define SCR_ARGS stream, sp, mean, scratch, scale_bits, rows_prepared, order, offset, tmax, n, buckets, d_short, (int)k, ids, dist, flags, &ovf, cutoff
  if (d == 16) launch_screen<16>(SCR_ARGS);
  else if (d == 32) launch_screen<32>(SCR_ARGS);
  else launch_screen<64>(SCR_ARGS);
#undef SCR_ARGS
  // buckets the screen could not hold: the tiled kernel, driven by the list
  if (!try_launch_tile(stream, sp, order, offset, tmax, n, buckets, d, d_short, k, ids, dist, flags, ovf.buckets, ovf.count))
    fatal_config("screened leaf path without a tiled kernel for the same shape");
  return true;
}
#endif

// =====================================================================================
// cutoff carried from try to try
// =====================================================================================
// The merge keeps the k smallest distinct ids of a point over all tries, so an entry of a later
// try that is farther than the k-th best found so far can never be reported.  `run` holds, per
// point, the k smallest DISTINCT distance values seen in the lists so far (equal ids carry equal
// distances, so k distinct values belong to at least k distinct ids and their largest bounds the
// final k-th distance from above; counting values instead of ids keeps the ids out of this
// pass and only loosens the bound in rows with exact ties).  cutoff[x] = that value, +inf while
// fewer than k are known.  One thread per point, everything in registers: duplicates of the new
// list against the running one are turned into +inf, the new list is re-sorted by a bitonic
// network and the two ascending lists are merged by the usual reverse / min / bitonic-merge.
#define CUT_CE(a, b) { const FT lo_ = (a) < (b) ? (a) : (b), hi_ = (a) < (b) ? (b) : (a); (a) = lo_; (b) = hi_; }
template <int K>
__global__ void __launch_bounds__(128)
cutoff_update_kernel(const FT *__restrict__ new_dist, FT *__restrict__ run, FT *__restrict__ cutoff,
                     size_t n, int k, int first) {
  const size_t x = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n) return;
  const FT inf = ft_inf();
  FT v[K], r[K];
  constexpr int VW = 16 / sizeof(FT);
  if (k == K) {
#pragma unroll
    for (int i = 0; i < K / VW; i++) {
      FT tmp[VW];
      *reinterpret_cast<uint4 *>(tmp) = reinterpret_cast<const uint4 *>(new_dist + x * (size_t)K)[i];
#pragma unroll
      for (int e = 0; e < VW; e++) v[i * VW + e] = tmp[e];
      if (!first) *reinterpret_cast<uint4 *>(tmp) = reinterpret_cast<const uint4 *>(run + x * (size_t)K)[i];
#pragma unroll
      for (int e = 0; e < VW; e++) r[i * VW + e] = first ? inf : tmp[e];
    }
  } else {
#pragma unroll
    for (int i = 0; i < K; i++) {
      v[i] = i < k ? new_dist[x * (size_t)k + i] : inf;
      r[i] = (i < k && !first) ? run[x * (size_t)k + i] : inf;
    }
  }
  bool bad = false;                                   // NaN anywhere: no cutoff for this point
#pragma unroll
  for (int i = 0; i < K; i++) bad |= (v[i] != v[i]) || (r[i] != r[i]);
  // distinct values only: repeats inside the (ascending) new list, then repeats of the running list
  bool dup[K];
  dup[0] = false;
#pragma unroll
  for (int j = 1; j < K; j++) dup[j] = v[j] == v[j - 1];
#pragma unroll
  for (int j = 0; j < K; j++) {
#pragma unroll
    for (int i = 0; i < K; i++) dup[j] |= v[j] == r[i];
  }
#pragma unroll
  for (int j = 0; j < K; j++) if (dup[j]) v[j] = inf;
  // ascending sort of v (bitonic network, compile-time indices)
#pragma unroll
  for (int size = 2; size <= K; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride >= 1; stride >>= 1) {
#pragma unroll
      for (int i = 0; i < K; i++) {
        const int p = i ^ stride;
        if (p > i) {
          if ((i & size) == 0) CUT_CE(v[i], v[p]) else CUT_CE(v[p], v[i])
        }
      }
    }
  }
  // the K smallest of the two ascending lists: min against the reversed list is bitonic
#pragma unroll
  for (int i = 0; i < K; i++) r[i] = r[i] < v[K - 1 - i] ? r[i] : v[K - 1 - i];
#pragma unroll
  for (int stride = K >> 1; stride >= 1; stride >>= 1) {
#pragma unroll
    for (int i = 0; i < K; i++) {
      const int p = i ^ stride;
      if (p > i) CUT_CE(r[i], r[p])
    }
  }
  if (bad) {
#pragma unroll
    for (int i = 0; i < K; i++) r[i] = inf;
  }
  FT kth = inf;
#pragma unroll
  for (int i = 0; i < K; i++) if (i == k - 1) kth = r[i];
  cutoff[x] = kth;
  if (k == K) {
#pragma unroll
    for (int i = 0; i < K / VW; i++) {
      FT tmp[VW];
#pragma unroll
      for (int e = 0; e < VW; e++) tmp[e] = r[i * VW + e];
      reinterpret_cast<uint4 *>(run + x * (size_t)K)[i] = *reinterpret_cast<uint4 *>(tmp);
    }
  } else {
#pragma unroll
    for (int i = 0; i < K; i++) if (i < k) run[x * (size_t)k + i] = r[i];
  }
}
#undef CUT_CE

// 1 (default): later tries of the screened path stop their lists at the running cutoff; 0: full
// lists.  The merged result is the same either way (tests compare).  ANN_B200_CUTOFF=0 switches it off.
static int cutoff_mode = -1;
extern "C" void annb_leaf_cutoff_mode(int on) { cutoff_mode = on ? 1 : 0; }
extern "C" int annb_cutoff_applies(size_t d, size_t d_short, size_t k) {
  if (cutoff_mode < 0) {
    const char *e = getenv("ANN_B200_CUTOFF");
    cutoff_mode = (e && *e) ? (*e != '0') : 1;
  }
  return cutoff_mode && k <= 16 && annb_screen_applies(d, d_short, k);
}
extern "C" void annb_cutoff_update(const FT *new_dist, FT *run, FT *cutoff, size_t n, size_t k, int first,
                                   annb_stream stream) {
  if (k < 1 || k > 16) fatal_config("annb_cutoff_update: k must be 1..16");
  cutoff_update_kernel<16><<<grid_for(n, 128), 128, 0, stream>>>(new_dist, run, cutoff, n, (int)k, first);
  LAUNCH_CHECK("cutoff_update");
}

extern "C" void annb_leaf_topk(const FT *sorted_points, const FT *mean, const u32 *order,
                               const u32 *offset, const u32 *hash, const u32 *tmax, size_t n,
                               size_t d, size_t d_short, size_t k, u32 *list_ids, FT *list_dist,
                               void *scratch, int *status, const unsigned *scale_bits,
                               int rows_prepared, annb_stream stream) {
  annb_leaf_topk_cut(sorted_points, mean, order, offset, hash, tmax, n, d, d_short, k, list_ids, list_dist, scratch,
                     status, scale_bits, rows_prepared, NULL, stream);
}

// With `cutoff` (per point an upper bound of its final k-th distance, annb_cutoff_update) a list
// may stop early: every candidate at or below the cutoff is there, at the position it has in the
// full list, and whatever else the full list holds may be replaced by (n, +inf).
extern "C" void annb_leaf_topk_cut(const FT *sorted_points, const FT *mean, const u32 *order,
                                   const u32 *offset, const u32 *hash, const u32 *tmax, size_t n,
                                   size_t d, size_t d_short, size_t k, u32 *list_ids, FT *list_dist,
                                   void *scratch, int *status, const unsigned *scale_bits,
                                   int rows_prepared, const FT *cutoff, annb_stream stream) {
  int regs = list_regs(k);
  if (!regs) fatal_config("k > 256");
  // scratch layout: rank_of[n] | tie list (count, rows[n]) | literal-row slabs
  u32 *rank_of = (u32 *)scratch;
  const size_t rank_bytes = (n * sizeof(u32) + 255) & ~(size_t)255;
  LiteralScratch ls = carve_literal_scratch((unsigned char *)scratch + rank_bytes, literal_area_bytes(n) - rank_bytes, 2 * n);
  (void)mean; (void)scale_bits; (void)rows_prepared;
  TieList flags = ls.list;
  unsigned char *slabs = ls.slabs;
  size_t slab_bytes = ls.slab_bytes;
  RT_CHECK(cudaMemsetAsync(flags.count, 0, sizeof(u32), stream));
  const size_t buckets = (size_t)1 << d_short;
  int mode = row_mode(d);
  size_t gsmem = mode ? 0 : 8 * d * sizeof(FT);
  if (gsmem > 200 * 1024) fatal_config("d too large for the generic distance path");

  bool done = false;
#ifdef USE_FLOAT
  done = try_launch_screen(stream, sorted_points, mean, scratch, scale_bits, rows_prepared, order, offset, tmax,
                           n, buckets, d, (int)d_short, k, list_ids, list_dist, flags, cutoff);
#else
  (void)cutoff;
#endif
  if (!done && !try_launch_tile(stream, sorted_points, order, offset, tmax, n, buckets, d, (int)d_short, k,
                                list_ids, list_dist, flags)) {
    dim3 block(256), grid(grid_for(n * 32, 256));
#define W_ARGS regs, grid, block, gsmem, stream, sorted_points, order, offset, hash, tmax, n, (int)d, (int)d_short, (int)k, list_ids, list_dist, flags
    switch (mode) {
      case 0: launch_warp_r<0>(W_ARGS); break;
      case 1: launch_warp_r<1>(W_ARGS); break;
      case 2: launch_warp_r<2>(W_ARGS); break;
      case 4: launch_warp_r<4>(W_ARGS); break;
      default: launch_warp_r<8>(W_ARGS); break;
    }
#undef W_ARGS
  }
  LAUNCH_CHECK("leaf_topk");

  invert_order_kernel<<<grid_for(n, 256), 256, 0, stream>>>(order, n, rank_of);
  LAUNCH_CHECK("invert_order");
  dim3 lblock(256), lgrid(148 * 4);
#define L_ARGS sorted_points, order, offset, hash, rank_of, tmax, n, (int)d, (int)d_short, (int)k, list_ids, list_dist, flags, slabs, slab_bytes, status
#define L_CASE(EE)                                                                                \
  {                                                                                               \
    if (gsmem > 48 * 1024)                                                                        \
      RT_CHECK(cudaFuncSetAttribute(leaf_literal_kernel<EE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gsmem)); \
    leaf_literal_kernel<EE><<<lgrid, lblock, gsmem, stream>>>(L_ARGS);                            \
  }
  switch (mode) {
    case 0: L_CASE(0) break;
    case 1: L_CASE(1) break;
    case 2: L_CASE(2) break;
    case 4: L_CASE(4) break;
    default: L_CASE(8) break;
  }
#undef L_CASE
#undef L_ARGS
  LAUNCH_CHECK("leaf_literal");
}
