// annb_query.cu — the query path (alg.c:438-519): projection + sign hash of the query
// vectors, candidate rows out of the saved bucket tables, k best per query.  The final
// supercharging step is annb_supercharge() (annb_finish.cu) with graph = save->graph.
#include "annb_common.cuh"
#ifdef USE_FLOAT
#include <cuda_fp16.h>
#include "annb_screen_common.cuh"
#endif

static __device__ unsigned long long query_literal_rows_dev;
unsigned long long annb_query_literal_count(int reset) {
  unsigned long long v = 0, z = 0;
  cudaMemcpyFromSymbol(&v, query_literal_rows_dev, sizeof v);
  if (reset) cudaMemcpyToSymbol(query_literal_rows_dev, &z, sizeof z);
  return v;
}

// =====================================================================================
// Q1: hash of (query x, try t) = sign bits of (y_x - mean) . bases[t][i], i < d_short
// =====================================================================================
// prods + add_up_cols + compute_signs (compute.cl:268-275,160-167,223-231).  The products
// can be -0.0, so the "+ 0" of the reference's even tree levels is kept (it turns -0.0
// into +0.0 and thereby fixes the sign bit of an all-zero sum).  One warp per (x, t); the
// result is stored where the reference stores it: sign[x * tries + t].

__device__ __forceinline__ FT warp_dot_tree(const FT *__restrict__ yrow, const FT *__restrict__ mean,
                                            const FT *__restrict__ b, int d, FT *tmp, int lane) {
  for (int z = lane; z < d; z += 32) tmp[z] = (yrow[z] - mean[z]) * b[z];
  __syncwarp();
  for (int l = d; l >> 1; l >>= 1) {
    int h = l >> 1;
    for (int z = lane; z < h; z += 32) {
      FT extra = (z == 0 && (l & 1)) ? tmp[l - 1] : (FT)0;
      tmp[z] = tmp[z] + (tmp[z + h] + extra);
    }
    __syncwarp();
  }
  FT v = tmp[0];
  __syncwarp();
  return v;
}

__global__ void __launch_bounds__(256)
query_hash_kernel(const FT *__restrict__ y, const FT *__restrict__ mean, const FT *__restrict__ bases,
                  size_t ycnt, int d, int d_short, int tries, u32 *__restrict__ sign) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= ycnt * (size_t)tries) return;
  FT *tmp = reinterpret_cast<FT *>(smem_raw) + (size_t)wib * d;
  size_t x = w / tries;
  int t = (int)(w - x * tries);
  u32 h = 0;
  for (int i = 0; i < d_short; i++) {
    FT v = warp_dot_tree(y + x * (size_t)d, mean, bases + ((size_t)t * d_short + i) * d, d, tmp, lane);
    h = (h << 1) | sign_bit(v);
  }
  if (lane == 0) sign[w] = h;
}

// d in {16, 32, 64, 128, 256}: one warp per query for ALL tries, the centred query in registers
// (lane l holds coordinates l + 32 s), the basis vectors through the L1 (T*d_short*d values,
// 32 KB at cfg3), the product tree of the reference as lane-local levels followed by
// xor-shuffles.  Every level is m[z] = m[z] + (m[z + h] + 0) as above; a + (b + 0) gives the
// same bits on both partners of a shuffle (IEEE addition commutes, signed zeros included), so
// the lanes that hold the live half of a level hold exactly the reference's values.
template <int E>
__global__ void __launch_bounds__(256)
query_hash_warp_kernel(const FT *__restrict__ y, const FT *__restrict__ mean, const FT *__restrict__ bases,
                       size_t ycnt, int d, int d_short, int tries, u32 *__restrict__ sign) {
  const int lane = threadIdx.x & 31;
  const size_t x = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (x >= ycnt) return;
  FT c[E];
#pragma unroll
  for (int s = 0; s < E; s++) {
    const int z = lane + 32 * s;
    c[s] = z < d ? y[x * (size_t)d + z] - mean[z] : (FT)0;
  }
  for (int t = 0; t < tries; t++) {
    u32 h = 0;
    for (int i = 0; i < d_short; i++) {
      const FT *b = bases + ((size_t)t * d_short + i) * d;
      FT m[E];
#pragma unroll
      for (int s = 0; s < E; s++) {
        const int z = lane + 32 * s;
        m[s] = z < d ? c[s] * b[z] : (FT)0;
      }
#pragma unroll
      for (int hh = E / 2; hh >= 1; hh >>= 1)
#pragma unroll
        for (int s = 0; s < hh; s++) m[s] = m[s] + (m[s + hh] + (FT)0);
      FT v = m[0];
      if (d >= 32) v = v + (__shfl_xor_sync(FULL, v, 16) + (FT)0);
      v = v + (__shfl_xor_sync(FULL, v, 8) + (FT)0);
      v = v + (__shfl_xor_sync(FULL, v, 4) + (FT)0);
      v = v + (__shfl_xor_sync(FULL, v, 2) + (FT)0);
      v = v + (__shfl_xor_sync(FULL, v, 1) + (FT)0);
      h = (h << 1) | sign_bit(v);                      // lane 0 holds the reference's sum
    }
    if (lane == 0) sign[x * (size_t)tries + t] = h;
  }
}

extern "C" void annb_query_hash(const FT *y, const FT *mean, const FT *bases, size_t ycnt, size_t d,
                                size_t d_short, int tries, u32 *sign, annb_stream stream) {
  {
    const char *off = getenv("ANN_B200_NO_FAST_QUERY");
    const int E = row_mode(d);                         // 0 = generic, else d/32 (d = 16 -> 1)
    if (E && !(off && *off && *off != '0')) {
      dim3 block(256), grid(grid_for(ycnt * 32, 256));
      switch (E) {
        case 1: query_hash_warp_kernel<1><<<grid, block, 0, stream>>>(y, mean, bases, ycnt, (int)d, (int)d_short, tries, sign); break;
        case 2: query_hash_warp_kernel<2><<<grid, block, 0, stream>>>(y, mean, bases, ycnt, (int)d, (int)d_short, tries, sign); break;
        case 4: query_hash_warp_kernel<4><<<grid, block, 0, stream>>>(y, mean, bases, ycnt, (int)d, (int)d_short, tries, sign); break;
        default: query_hash_warp_kernel<8><<<grid, block, 0, stream>>>(y, mean, bases, ycnt, (int)d, (int)d_short, tries, sign); break;
      }
      LAUNCH_CHECK("query_hash_warp");
      return;
    }
  }
  size_t smem = 8 * d * sizeof(FT);
  if (smem > 200 * 1024) fatal_config("d too large for the query projection");
  if (smem > 48 * 1024)
    RT_CHECK(cudaFuncSetAttribute(query_hash_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  query_hash_kernel<<<grid_for(ycnt * tries * 32, 256), 256, smem, stream>>>(y, mean, bases, ycnt, (int)d, (int)d_short, tries, sign);
  LAUNCH_CHECK("query_hash");
}

// =====================================================================================
// Q2: candidate rows from the saved tables -> k best per query
// =====================================================================================
// Row of query x (alg.c:438-452,493-511): for try i, columns
// [(d_short+1)*off_i + f*w_i, +w_i) hold table_i[h_i ^ flip_f], off_i = sum of par_maxes
// before i, w_i = par_maxes[i]; h_i = sign[i*ycnt + x] — the reference reads the sign
// buffer with that layout although it was written as [x][try] (SURVEY "three facts" #3);
// reproduced as is.  The first 2^floor(log2(len)) columns compete.
struct QueryTables {
  const u32 *tab[64];      // [2^d_short][w_i] padded with n, ids descending
  u32 width[64];
  u32 offset[64];
  int tries;
  unsigned long long len, prefix;
  // hash of (query x, try t) = sign[t * s_try + x * s_row].  The reference writes the buffer as
  // [x][try] and reads it as [try][x] (alg.c:474-499): s_try = ycnt, s_row = 1 reproduces that;
  // s_try = 1, s_row = tries is the corrected read (ANN_B200_QUERY_LAYOUT=fixed, opt-in)
  size_t s_try, s_row;
};

template <int E, int R>
__global__ void __launch_bounds__(256)
query_rows_kernel(const FT *__restrict__ y, const FT *__restrict__ points, QueryTables q,
                  const u32 *__restrict__ sign, size_t n, size_t ycnt, int d, int d_short, int k,
                  int exclude_self, u32 *__restrict__ list_ids, FT *__restrict__ list_dist,
                  TieList ties) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  size_t x = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (x >= ycnt) return;
  FT *tmp = reinterpret_cast<FT *>(smem_raw) + (size_t)wib * (E == 0 ? d : 0);
  const u32 sentinel = (u32)n;
  WarpRow<(E ? E : 1)> qr;
  const FT *qrow = y + x * (size_t)d;
  if (E) qr.load(qrow, lane, d);
  WarpList<R> best;
  best.clear(sentinel);
  FT tau = ft_inf(), max_v = -ft_inf();
  u32 max_id = sentinel, corner_id = sentinel;
  bool any_inf = false, tie = false, done = false;

  for (int t = 0; t < q.tries && !done; t++) {
    const u32 h = sign[(size_t)t * q.s_try + x * q.s_row];
    const u32 w = q.width[t];
    for (int f = 0; f <= d_short; f++) {
      unsigned long long col = (unsigned long long)(d_short + 1) * q.offset[t] + (unsigned long long)f * w;
      const u32 *row = q.tab[t] + (size_t)(h ^ (f ? (1u << (f - 1)) : 0u)) * w;
      if (col >= q.prefix) {
        // first slot outside the prefix is in this segment iff col == prefix
        if (col == q.prefix && w) corner_id = row[0];
        done = true;
        break;
      }
      unsigned long long room = q.prefix - col;
      u32 take = w < room ? w : (u32)room;
      if (take < w) { corner_id = row[take]; done = true; }
      for (u32 z = 0; z < take; z++) {
        u32 id = row[z];
        if (id >= sentinel) { any_inf = true; break; }          // pads fill the rest of the row
        if (exclude_self && id == (u32)x) { any_inf = true; continue; }
        FT dist;
        if (E) {
          WarpRow<(E ? E : 1)> cr;
          cr.load(points + (size_t)id * d, lane, d);
          dist = __shfl_sync(FULL, warp_sqdist<(E ? E : 1)>(qr, cr, d), 0);
        } else {
          dist = generic_sqdist(qrow, points + (size_t)id * d, d, tmp, lane);
        }
        if (dist > max_v) { max_v = dist; max_id = id; }
        consider<R>(best, tau, dist, id, k, sentinel, lane, tie);
      }
      if (done) break;
    }
  }
  if (q.prefix < q.len && !any_inf && corner_id == max_id) best.remove(corner_id, sentinel, lane);
#pragma unroll
  for (int rr = 0; rr < R; rr++) {
    int p = rr * 32 + lane;
    if (p < k) {
      list_ids[x * (size_t)k + p] = best.id[rr];
      list_dist[x * (size_t)k + p] = best.v[rr];
    }
  }
  if (tie && lane == 0) tie_report(ties, (u32)x);
}

// Fast path (float, d in {16,32,64,128}, k <= 32): query_rows_kernel measures one candidate at a
// time (a dependent id load, then a dependent row load per candidate), which leaves the warp
// waiting on memory.  Here the ids of the row's segments are collected in a small per-warp
// buffer (coalesced loads from the tables, pads and the query itself dropped) and measured
// 2*CPR at a time: d/8 lanes per candidate, lane g holding coordinates 4g..4g+3 and
// d/2+4g..d/2+4g+3 (two 16-byte loads, whole 128-byte lines per instruction), the reference's
// summation tree as in supercharge_screen_kernel.  Same candidates, same tree, same list rules
// as the kernel above — only the order in which candidates are offered differs, and that order
// only matters in rows with exact ties, which both kernels hand to the literal kernel.
#ifdef USE_FLOAT
static constexpr int QBUF = 192;                                  // buffered candidate ids per warp

// With SCREEN (the index carries the fp16 copy of the points, annb_query_screen), every flush of
// the id buffer first brackets its candidates against the CURRENT k-th best — the bracket and
// the shuffle-free tensor-core operand layout of supercharge_screen_kernel, the query's fp16
// words computed on the fly from (y - mean) * scale — and only the survivors get the exact
// tree.  tau only shrinks, so a candidate dropped against the current tau could never have
// entered the final list; the prefix-corner rule can only fire when nothing was dropped at
// all (see annb_supercharge_screen.cuh), and then every candidate was measured exactly.
struct QueryScreenDev {
  const unsigned short *p16;
  const float2 *nrm;
  const float *mean;
  const unsigned *scale_bits;
};

template <int D, bool SCREEN>
__global__ void __launch_bounds__(256)
query_rows_fast_kernel(const float *__restrict__ y, const float *__restrict__ points, QueryTables q,
                       const u32 *__restrict__ sign, size_t n,
                       size_t ycnt, int d_short, int k, int exclude_self, u32 *__restrict__ list_ids,
                       float *__restrict__ list_dist, TieList ties, QueryScreenDev scr) {
  constexpr int LPC = D / 8, CPR = 32 / LPC;
  constexpr int KS2 = D >= 32 ? D / 32 : 1;
  __shared__ __align__(16) u32 s_buf[8][QBUF + 16];
  __shared__ u32 s_seg[8][29][32];                       // first 32 ids of the d_short+1 <= 29 table rows of a try
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  size_t x = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (x >= ycnt) return;
  u32 *buf = s_buf[wib];
  const u32 sentinel = (u32)n;
  const float inf = ft_inf();
  const int g = lane & (LPC - 1), grp = lane / LPC;
  const float4 qa = *reinterpret_cast<const float4 *>(y + x * (size_t)D + 4 * g);
  const float4 qb = *reinterpret_cast<const float4 *>(y + x * (size_t)D + D / 2 + 4 * g);
  WarpList<1> best;
  best.clear(sentinel);
  float tau = inf, max_v = -inf;
  u32 max_id = sentinel, corner_id = sentinel;
  bool any_inf = false, tie = false, done = false, dropped = false;
  int cnt = 0;

  // ---- screen operands of the query (SCREEN only)
  const int g8 = lane >> 2, t4 = lane & 3;
  u32 qw[KS2][2];
  float2 qn = make_float2(0.f, 0.f);
  float scale2 = 0.f;
  if (SCREEN) {
    float scale = 1.0f;
    {
      const float cmax = __uint_as_float(*scr.scale_bits);
      if (cmax > 0.f && cmax <= 3.0e38f) {
        const int e = ilogbf(cmax);
        scale = ldexpf(1.0f, 2 - max(-100, min(100, e)));               // the scale of the stored fp16 rows
        if (e >= -30 && e <= 30) scale2 = scale * scale;               // outside: everything is measured
      }
    }
    float ss = 0.f, n2 = 0.f;
    const float *yr = y + x * (size_t)D;
#pragma unroll
    for (int v = 0; v < KS2; v++) {
      constexpr int NC = D >= 32 ? 8 : 4;                             // coordinates of this lane's piece
      const int c0 = D >= 32 ? 32 * v + 8 * t4 : 4 * t4;
      u32 w[4] = {0u, 0u, 0u, 0u};
#pragma unroll
      for (int i = 0; i < NC; i += 2) {
        const float a = (yr[c0 + i] - scr.mean[c0 + i]) * scale, b = (yr[c0 + i + 1] - scr.mean[c0 + i + 1]) * scale;
        const __half2 h = __floats2half2_rn(a, b);
        const float ra = __low2float(h), rb = __high2float(h);
        ss += a * a + b * b;
        n2 += ra * ra + rb * rb;
        w[i >> 1] = *reinterpret_cast<const u32 *>(&h);
      }
      qw[v][0] = (g8 & 1) ? w[1] : w[0];
      qw[v][1] = D >= 32 ? ((g8 & 1) ? w[3] : w[2]) : 0u;
    }
    ss += __shfl_xor_sync(FULL, ss, 1); ss += __shfl_xor_sync(FULL, ss, 2);
    n2 += __shfl_xor_sync(FULL, n2, 1); n2 += __shfl_xor_sync(FULL, n2, 2);
    // the norm bound of screen_prep_kernel, with a little more slack for the different summation order
    qn.x = SCREEN_SQRT_KAPPA * (sqrtf(ss) * (1.0f + 1.0f / 2048.0f) + sqrtf((float)D) * (1.0f / 16384.0f));
    qn.y = n2;
  }

  auto flush = [&]() {
    __syncwarp();
    if (SCREEN && cnt > 0) {
      const int c16 = (cnt + 15) & ~15;
      if (lane < c16 - cnt) buf[cnt + lane] = buf[0];                   // pads: a valid row, dead by position
      __syncwarp();
      const float taus = scale2 > 0.f ? tau * scale2 : inf;
      int V = 0;
      // a round's rows are requested while the previous round is evaluated
      uint4 ra[KS2], rc[KS2];
      u32 ce = 0;
      float2 cn = make_float2(0.f, 0.f);
      auto load = [&](int base) {
        const u32 c0 = buf[base + g8], c1 = buf[base + g8 + 8];
        const unsigned short *p0 = scr.p16 + (size_t)c0 * D, *p1 = scr.p16 + (size_t)c1 * D;
#pragma unroll
        for (int v = 0; v < KS2; v++) {
          if (D >= 32) {
            ra[v] = *reinterpret_cast<const uint4 *>(p0 + 32 * v + 8 * t4);
            rc[v] = *reinterpret_cast<const uint4 *>(p1 + 32 * v + 8 * t4);
          } else {
            const uint2 w0 = *reinterpret_cast<const uint2 *>(p0 + 4 * t4), w1 = *reinterpret_cast<const uint2 *>(p1 + 4 * t4);
            ra[v] = make_uint4(w0.x, w0.y, 0u, 0u);
            rc[v] = make_uint4(w1.x, w1.y, 0u, 0u);
          }
        }
        ce = (t4 & 1) ? c1 : c0;
        cn = scr.nrm[ce];
      };
      load(0);
      for (int base = 0; base < c16; base += 16) {
        float ca[4] = {0.f, 0.f, 0.f, 0.f}, cc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int v = 0; v < KS2; v++) {
          mma_f16_16816(ca, ra[v].x, ra[v].y, ra[v].z, ra[v].w, qw[v][0], qw[v][1]);
          mma_f16_16816(cc, rc[v].x, rc[v].y, rc[v].z, rc[v].w, qw[v][0], qw[v][1]);
        }
        const u32 cur = ce;
        const float2 cnn = cn;
        if (base + 16 < c16) load(base + 16);               // ids of the next round are read before this round's survivors are written
        const float dot = (t4 & 1) ? cc[0] + cc[3] : ca[0] + ca[3];
        const float t = qn.x + cnn.x;
        const float lo = __fmaf_rn(-t, t, __fmaf_rn(-2.0f, dot, qn.y + cnn.y));
        const bool pass = t4 < 2 && base + g8 + 8 * (t4 & 1) < cnt && lo <= taus;
        const unsigned m = __ballot_sync(FULL, pass);
        __syncwarp();
        if (pass) buf[V + __popc(m & ((1u << lane) - 1))] = cur;
        V += __popc(m);
      }
      if (V < cnt) dropped = true;
      cnt = V;
      __syncwarp();
    }
    for (int base = 0; base < cnt; base += 2 * CPR) {
      u32 cid[2];
      float dist[2];
      float4 ca[2], cb[2];
#pragma unroll
      for (int h2 = 0; h2 < 2; h2++) {
        const int mine = base + CPR * h2 + grp;
        cid[h2] = buf[mine < cnt ? mine : base];
        const float *crow = points + (size_t)cid[h2] * D;
        ca[h2] = *reinterpret_cast<const float4 *>(crow + 4 * g);
        cb[h2] = *reinterpret_cast<const float4 *>(crow + D / 2 + 4 * g);
      }
#pragma unroll
      for (int h2 = 0; h2 < 2; h2++) {
        float m[4];
        {
          float d0 = qa.x - ca[h2].x, d1 = qa.y - ca[h2].y, d2 = qa.z - ca[h2].z, d3 = qa.w - ca[h2].w;
          float e0 = qb.x - cb[h2].x, e1 = qb.y - cb[h2].y, e2 = qb.z - cb[h2].z, e3 = qb.w - cb[h2].w;
          m[0] = d0 * d0 + e0 * e0; m[1] = d1 * d1 + e1 * e1;
          m[2] = d2 * d2 + e2 * e2; m[3] = d3 * d3 + e3 * e3;
        }
#pragma unroll
        for (int o = LPC / 2; o >= 1; o >>= 1) {
#pragma unroll
          for (int j = 0; j < 4; j++) m[j] = m[j] + __shfl_xor_sync(FULL, m[j], o);
        }
        const float tt = (m[0] + m[2]) + (m[1] + m[3]);
        const bool live = base + CPR * h2 + grp < cnt;
        dist[h2] = live ? tt : inf;
        if (live && tt > max_v) { max_v = tt; max_id = cid[h2]; }     // lane-local, reduced at the end
      }
      if (__any_sync(FULL, dist[0] <= tau || dist[1] <= tau)) {
#pragma unroll
        for (int h2 = 0; h2 < 2; h2++)
          for (int i = 0; i < CPR; i++) {
            float vn = __shfl_sync(FULL, dist[h2], LPC * i);
            u32 idn = __shfl_sync(FULL, cid[h2], LPC * i);
            if (vn != inf) consider<1>(best, tau, vn, idn, k, sentinel, lane, tie);
          }
      }
    }
    cnt = 0;
    __syncwarp();
  };

  for (int t = 0; t < q.tries && !done; t++) {
    const u32 h = sign[(size_t)t * q.s_try + x * q.s_row];
    const u32 w = q.width[t];
    // The first 32 ids of ALL d_short+1 table rows of the try are requested before any of them
    // is consumed and parked in shared memory: one dependent L2 round trip per segment (136 per
    // query at cfg3) was most of this kernel's time.  (Real ids are a prefix of a table row, so
    // only buckets with more than 32 points need a second, dependent load.)
    u32 (*seg)[32] = s_seg[wib];
    {
      u32 v[8];
      for (int f0 = 0; f0 <= d_short; f0 += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const int f = f0 + j;
          v[j] = sentinel;
          if (f <= d_short && (u32)lane < w) v[j] = q.tab[t][(size_t)(h ^ (f ? (1u << (f - 1)) : 0u)) * w + lane];
        }
#pragma unroll
        for (int j = 0; j < 8; j++)
          if (f0 + j <= d_short) seg[f0 + j][lane] = v[j];
      }
    }
    __syncwarp();
    for (int f = 0; f <= d_short; f++) {
      unsigned long long col = (unsigned long long)(d_short + 1) * q.offset[t] + (unsigned long long)f * w;
      const u32 *row = q.tab[t] + (size_t)(h ^ (f ? (1u << (f - 1)) : 0u)) * w;
      if (col >= q.prefix) {
        if (col == q.prefix && w) corner_id = seg[f][0];
        done = true;
        break;
      }
      unsigned long long room = q.prefix - col;
      u32 take = w < room ? w : (u32)room;
      if (take < w) { corner_id = take < 32 ? seg[f][take] : row[take]; done = true; }
      for (u32 z0 = 0; z0 < take; z0 += 32) {
        const u32 z = z0 + lane;
        const u32 id = z < take ? (z0 == 0 ? seg[f][lane] : row[z]) : sentinel;
        // pads fill the tail of a table row, so the real ids are a prefix of it
        const unsigned real = __ballot_sync(FULL, z < take && id < sentinel);
        if (__ballot_sync(FULL, z < take && id >= sentinel)) any_inf = true;
        const bool self = exclude_self && id == (u32)x;
        if (__ballot_sync(FULL, (real >> lane & 1) && self)) any_inf = true;
        const unsigned keep = __ballot_sync(FULL, (real >> lane & 1) && !self);
        if (keep >> lane & 1) buf[cnt + __popc(keep & ((1u << lane) - 1))] = id;
        cnt += __popc(keep);
        if (cnt > QBUF - 32) flush();
        if (real != (z0 + 32 <= take ? 0xffffffffu : ((1u << (take - z0)) - 1u))) break;   // a pad was seen
      }
      if (done) break;
    }
    __syncwarp();
  }
  flush();
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    float ov = __shfl_xor_sync(FULL, max_v, o);
    u32 oi = __shfl_xor_sync(FULL, max_id, o);
    if (ov > max_v || (ov == max_v && oi < max_id)) { max_v = ov; max_id = oi; }
  }
  if (q.prefix < q.len && !any_inf && !dropped && corner_id == max_id) best.remove(corner_id, sentinel, lane);
  if (lane < k) {
    list_ids[x * (size_t)k + lane] = best.id[0];
    list_dist[x * (size_t)k + lane] = best.v[0];
  }
  if (tie && lane == 0) tie_report(ties, (u32)x);
}
#endif

// literal row of a reported query: every column, the reference's network, first k out;
// one CTA per row
template <int E>
__global__ void __launch_bounds__(256)
query_literal_kernel(const FT *__restrict__ y, const FT *__restrict__ points, QueryTables q,
                     const u32 *__restrict__ sign, size_t n, size_t ycnt, int d, int d_short, int k,
                     int exclude_self, u32 *__restrict__ list_ids, FT *__restrict__ list_dist,
                     TieList ties, unsigned char *slabs, size_t slab_bytes, int *status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, wib = tid >> 5;
  FT *tmp = reinterpret_cast<FT *>(smem_raw) + (size_t)wib * (E == 0 ? d : 0);
  const u32 total = *ties.count;
  if (total == 0) return;
  const size_t L = (size_t)q.len;
  const size_t slab = (L * (sizeof(FT) + 2 * sizeof(u32)) + 15) & ~(size_t)15;
  const size_t fit = slab ? slab_bytes / slab : 0;
  if (fit == 0) { if (blockIdx.x == 0 && tid == 0) *status = 1; return; }
  const u32 workers = (u32)(fit < gridDim.x ? fit : gridDim.x);
  if (blockIdx.x >= workers) return;
  if (blockIdx.x == 0 && tid == 0) atomicAdd(&query_literal_rows_dev, (unsigned long long)total);
  FT *key = reinterpret_cast<FT *>(slabs + (size_t)blockIdx.x * slab);
  u32 *ids = reinterpret_cast<u32 *>(key + L);
  u32 *cslot = ids + L;
  __shared__ u32 s_live;
  const u32 sentinel = (u32)n;

  for (u32 it = blockIdx.x; it < total; it += workers) {
    const size_t x = ties.rows[it];
    if (tid == 0) s_live = 0;
    __syncthreads();
    for (int t = 0; t < q.tries; t++) {
      const u32 h = sign[(size_t)t * q.s_try + x * q.s_row];
      const u32 w = q.width[t];
      const size_t cols = (size_t)(d_short + 1) * w, col0 = (size_t)(d_short + 1) * q.offset[t];
      for (size_t c = tid; c < cols; c += blockDim.x) {
        u32 f = (u32)(c / w), z = (u32)(c - (size_t)f * w);
        u32 id = q.tab[t][(size_t)(h ^ (f ? (1u << (f - 1)) : 0u)) * w + z];
        ids[col0 + c] = id;
        key[col0 + c] = ft_inf();
        if (id < sentinel && !(exclude_self && id == (u32)x)) cslot[atomicAdd(&s_live, 1u)] = (u32)(col0 + c);
      }
    }
    __syncthreads();
    block_row_distances<E>(y + x * (size_t)d, points, d, s_live, tmp,
                           [&](u32 i) { return (size_t)ids[cslot[i]]; },
                           [&](u32 i, FT dist) { key[cslot[i]] = dist; });
    __syncthreads();
    block_sort_and_uniq(ids, key, (int)L);
    for (int i = tid; i < k; i += blockDim.x) {
      list_ids[x * (size_t)k + i] = ids[i];
      list_dist[x * (size_t)k + i] = key[i];
    }
    __syncthreads();
  }
}

template <int E>
static void launch_query_rows(int regs, size_t smem, annb_stream stream, const FT *y, const FT *points,
                              const QueryTables &q, const u32 *sign, size_t n, size_t ycnt, int d,
                              int d_short, int k, int ex, u32 *ids, FT *dist, const LiteralScratch &ls,
                              int *status, const annb_query_screen *screen) {
  (void)screen;
  dim3 block(256), grid(grid_for(ycnt * 32, 256));
  bool fast = false;
#ifdef USE_FLOAT
  {
    const char *off = getenv("ANN_B200_NO_FAST_QUERY");
    if (!(off && *off && *off != '0') && k <= 32 && (d == 16 || d == 32 || d == 64 || d == 128)) {
      QueryScreenDev sd = {NULL, NULL, NULL, NULL};
      const char *soff = getenv("ANN_B200_QUERY_SCREEN");
      const bool use_screen = screen && screen->points16 && screen->nrm && screen->mean && screen->scale_bits &&
                              !(soff && *soff == '0');
      if (use_screen) {
        sd.p16 = (const unsigned short *)screen->points16;
        sd.nrm = (const float2 *)screen->nrm;
        sd.mean = (const float *)screen->mean;
        sd.scale_bits = screen->scale_bits;
      }
#define QF_CASE(DD)                                                                                              \
  if (use_screen) query_rows_fast_kernel<DD, true><<<grid, block, 0, stream>>>(y, points, q, sign, n, ycnt, d_short, k, ex, ids, dist, ls.list, sd); \
  else query_rows_fast_kernel<DD, false><<<grid, block, 0, stream>>>(y, points, q, sign, n, ycnt, d_short, k, ex, ids, dist, ls.list, sd);
      switch (d) {
        case 16: QF_CASE(16) break;
        case 32: QF_CASE(32) break;
        case 64: QF_CASE(64) break;
        default: QF_CASE(128) break;
      }
#undef QF_CASE
      fast = true;
    }
  }
#endif
#define QR_CASE(R)                                                                                \
  {                                                                                               \
    if (smem > 48 * 1024)                                                                         \
      RT_CHECK(cudaFuncSetAttribute(query_rows_kernel<E, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    query_rows_kernel<E, R><<<grid, block, smem, stream>>>(y, points, q, sign, n, ycnt, d, d_short, k, ex, ids, dist, ls.list); \
  }
  if (!fast)
    switch (regs) {
      case 1: QR_CASE(1) break;
      case 2: QR_CASE(2) break;
      case 4: QR_CASE(4) break;
      default: QR_CASE(8) break;
    }
#undef QR_CASE
  LAUNCH_CHECK("query_rows");
  if (smem > 48 * 1024)
    RT_CHECK(cudaFuncSetAttribute(query_literal_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  query_literal_kernel<E><<<148 * 2, 256, smem, stream>>>(y, points, q, sign, n, ycnt, d, d_short, k, ex, ids, dist, ls.list, ls.slabs, ls.slab_bytes, status);
  LAUNCH_CHECK("query_literal");
}

extern "C" int annb_query_screen_applies(size_t d, size_t k) {
#ifdef USE_FLOAT
  return (d == 16 || d == 32 || d == 64 || d == 128) && k <= 32;
#else
  (void)d; (void)k;
  return 0;
#endif
}

extern "C" void annb_query_rows(const FT *y, const FT *points, const u32 *const *tables,
                                const size_t *par_maxes, int tries, const u32 *sign, size_t n,
                                size_t ycnt, size_t d, size_t d_short, size_t k, int exclude_self,
                                u32 *list_ids, FT *list_dist, void *scratch, size_t scratch_bytes,
                                int *status, const annb_query_screen *screen, annb_stream stream) {
  int regs = list_regs(k);
  if (!regs) fatal_config("k > 256");
  if (tries > 64) fatal_config("more than 64 tries in a saved index");
  QueryTables q;
  q.tries = tries;
  unsigned long long off = 0;
  for (int t = 0; t < tries; t++) {
    q.tab[t] = tables[t];
    q.width[t] = (u32)par_maxes[t];
    q.offset[t] = (u32)off;
    off += par_maxes[t];
  }
  {
    // opt-in corrected read of the sign buffer; the default reproduces the reference (bug included)
    const char *lay = getenv("ANN_B200_QUERY_LAYOUT");
    const bool fixed = lay && (lay[0] == 'f' || lay[0] == 'F');
    q.s_try = fixed ? 1 : ycnt;
    q.s_row = fixed ? (size_t)tries : 1;
  }
  q.len = off * (d_short + 1);
  if (q.len < 16) fatal_config("query candidate rows shorter than 16 slots");
  q.prefix = 1ull << (63 - __builtin_clzll(q.len));
  int mode = row_mode(d);
  size_t smem = mode ? 0 : 8 * d * sizeof(FT);
  if (smem > 200 * 1024) fatal_config("d too large for the generic distance path");
  LiteralScratch ls = carve_literal_scratch(scratch, scratch_bytes, ycnt);
  RT_CHECK(cudaMemsetAsync(ls.list.count, 0, sizeof(u32), stream));
#define Q_ARGS regs, smem, stream, y, points, q, sign, n, ycnt, (int)d, (int)d_short, (int)k, exclude_self, list_ids, list_dist, ls, status, screen
  switch (mode) {
    case 0: launch_query_rows<0>(Q_ARGS); break;
    case 1: launch_query_rows<1>(Q_ARGS); break;
    case 2: launch_query_rows<2>(Q_ARGS); break;
    case 4: launch_query_rows<4>(Q_ARGS); break;
    default: launch_query_rows<8>(Q_ARGS); break;
  }
#undef Q_ARGS
}

// size_t -> u32 narrowing of host-format tables / graphs already copied to the device
__global__ void narrow_ids_kernel(const size_t *__restrict__ src, size_t count, u32 *__restrict__ dst) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) dst[i] = (u32)src[i];
}

extern "C" void annb_narrow_ids(const size_t *src, size_t count, u32 *dst, annb_stream stream) {
  if (!count) return;
  narrow_ids_kernel<<<grid_for(count, 256), 256, 0, stream>>>(src, count, dst);
  LAUNCH_CHECK("narrow_ids");
}
