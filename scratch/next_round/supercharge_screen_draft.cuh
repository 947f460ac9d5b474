// DRAFT, NOT BUILT INTO THE LIBRARY — next step after the screened S3 path (DESIGN.md §4 S3 "Next").
//
// S5 (supercharge) gathers k*k - ... = 240 candidate rows of 256 B per point from all over the
// point array (69.6 GB per step on cfg3, the kernel runs at the L2/HBM gather rate) although only
// ~3 % of them beat the row's current k-th best.  Same idea as in S3: bracket every candidate
// with the fp16 copy (128 B per row, original order, prepared once per precomp with
// screen_prep_kernel(points, NULL, NULL, mean, ...)), and fetch the fp32 row only for candidates
// whose lower bound reaches tau0 = the row's k-th own distance (tau only shrinks afterwards, so
// tau0 is conservative).  Expected traffic per point: 240 x 128 B + ~10 x 256 B = 33 KB instead
// of 61 KB.  scratch/next_round/s5_screen_prototype.py replays the logic in numpy against the
// oracle's supercharge rows (no mismatch; 11 % survivors at n = 4096).
//
// Prefix corner (alg.c:313-327 via rdups): the fast kernel needs the id of the prefix's largest
// exact distance only to test it against the id in slot P2.  Here: if that id does not occur in
// the prefix, nothing happens; if some other prefix entry is provably farther (its lo exceeds
// the corner id's hi), nothing happens; otherwise the row is reported to the literal kernel.
//
// To integrate: (1) host: fp16 copy + norms of the points in ORIGINAL order once per precomp
// (arena + n*d*2 + n*8 bytes), scale word from annb_screen_scale; (2) annb_supercharge gains
// (points16, pnrm, scale_bits), NULL for query_gpu (its rows are not points); (3) launch this
// kernel instead of supercharge_fast_kernel<8> when d == 64 && k <= 32 && points16; (4) tests:
// tests/test_gpu_screen.py pattern (mode on/off bit-equal + oracle), data sets with ties and
// offsets; (5) generalise the 8-lanes-per-candidate screen to d = 16/32/128 (d/8 lanes).
#pragma once

template <int EPL>   // = d / 8; the draft assumes EPL == 8 (d = 64): one fp16 row = 8 lanes x 16 B
__global__ void __launch_bounds__(256)
supercharge_screen_kernel(const float *__restrict__ queries, const float *__restrict__ points,
                          const unsigned short *__restrict__ points16, const float2 *__restrict__ pnrm,
                          const unsigned *__restrict__ scale_bits,
                          const u32 *__restrict__ own_ids, const float *__restrict__ own_dist,
                          const u32 *__restrict__ graph, size_t n, int k, size_t row_begin,
                          size_t row_end, int exclude_self, u32 *__restrict__ out_ids,
                          float *__restrict__ out_dist, TieList ties) {
  constexpr int D = EPL * 8;
  static_assert(EPL == 8, "draft: d = 64 only");
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  size_t x = row_begin + (((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (x >= row_end) return;
  const u32 sentinel = (u32)n;
  const int wide = k * (k + 1);
  const int P2 = 1 << floor_log2_u((unsigned long long)wide);
  const int cand = P2 - k;
  u32 *uniq = reinterpret_cast<u32 *>(smem_raw) + (size_t)wib * ((size_t)k * k);
  const float inf = ft_inf();

  const int g = lane & 7, grp = lane >> 3;
  float q[EPL];
  {
    const float *qrow = queries + x * (size_t)D;
#pragma unroll
    for (int s = 0; s < EPL; s++) q[s] = qrow[g + 8 * s];
  }
  // scaled units of the fp16 copy: scale is a power of two, so the products below are exact
  float scale2;
  {
    float cmax = __uint_as_float(*scale_bits);
    float sc = 1.0f;
    if (cmax > 0.f && cmax <= 3.0e38f) sc = ldexpf(1.0f, 2 - max(-100, min(100, ilogbf(cmax))));
    scale2 = sc * sc;
  }

  WarpList<1> best;
  best.v[0] = lane < k ? own_dist[x * (size_t)k + lane] : inf;
  best.id[0] = lane < k ? own_ids[x * (size_t)k + lane] : sentinel;
  const u32 own_reg = best.id[0];
  const float own_v = best.v[0];
  bool tie = false;
  bool any_inf = lane < k && best.v[0] == inf;
  {
    float nxt = __shfl_down_sync(FULL, best.v[0], 1);
    if (lane + 1 < k && best.v[0] == nxt && nxt != inf) tie = true;
  }
  float tau = best.kth(k);
  const float tau0s = tau * scale2;                                   // +inf stays +inf

  // candidate ids -> uniq[0..U): pads and (in precomp) the point itself are dropped here
  int U = 0;
  for (int base = 0; base < cand; base += 32) {
    int c = base + lane;
    int j = c < cand ? c / k : 0;
    int z = c - j * k;
    u32 oj = __shfl_sync(FULL, own_reg, j);
    u32 cid = (c < cand && oj < sentinel) ? graph[(size_t)oj * k + z] : sentinel;
    bool keep = false;
    if (c < cand) {
      if (cid >= sentinel || (exclude_self && cid == (u32)x)) any_inf = true;
      else keep = true;
    }
    unsigned m = __ballot_sync(FULL, keep);
    if (keep) uniq[U + __popc(m & ((1u << lane) - 1))] = cid;
    U += __popc(m);
  }
  any_inf = __any_sync(FULL, any_inf);
  tie = __any_sync(FULL, tie);
  // the id in slot P2 of the row (the prefix corner), if the rule can apply at all
  u32 corner = sentinel;
  if (P2 < wide && !any_inf) {
    int c = P2 - k, j = c / k, z = c - j * k;
    u32 oj = __shfl_sync(FULL, own_reg, j);
    corner = oj < sentinel ? graph[(size_t)oj * k + z] : sentinel;
  }
  __syncwarp();

  // ---- screen: 4 candidates per round, 8 lanes each; lane g owns halves [8g, 8g+8) ----------
  float qf[8];
  {
    const uint4 v = *reinterpret_cast<const uint4 *>(points16 + x * (size_t)D + 8 * g);
    const __half2 *h = reinterpret_cast<const __half2 *>(&v);
#pragma unroll
    for (int i = 0; i < 4; i++) { float2 f = __half22float2(h[i]); qf[2 * i] = f.x; qf[2 * i + 1] = f.y; }
  }
  const float2 qn = pnrm[x];
  float far_lo = -inf;        // largest lower bound among prefix entries other than the corner id
  float corner_hi = -inf;     // upper bound of the corner id's distance, if it is in the prefix
  int V = 0;                  // survivors, compacted in place into uniq[0..V)
  for (int base = 0; base < U; base += 4) {
    const int mine = base + grp;
    const bool live = mine < U;
    const u32 cid = uniq[live ? mine : base];
    const uint4 v = *reinterpret_cast<const uint4 *>(points16 + (size_t)cid * D + 8 * g);
    const float2 cn = pnrm[cid];
    const __half2 *h = reinterpret_cast<const __half2 *>(&v);
    float dot = 0.f;
#pragma unroll
    for (int i = 0; i < 4; i++) {
      float2 f = __half22float2(h[i]);
      dot = __fmaf_rn(qf[2 * i], f.x, dot);
      dot = __fmaf_rn(qf[2 * i + 1], f.y, dot);
    }
    dot += __shfl_xor_sync(FULL, dot, 4);
    dot += __shfl_xor_sync(FULL, dot, 2);
    dot += __shfl_xor_sync(FULL, dot, 1);
    const float dp = (qn.y + cn.y) - 2.0f * dot;
    const float t = qn.x + cn.x, sl = t * t;
    const float lo = dp - sl, hi = dp + sl;
    if (live) {
      if (cid == corner) corner_hi = fmaxf(corner_hi, hi);
      else far_lo = fmaxf(far_lo, lo);
    }
    const bool pass = live && g == 0 && lo <= tau0s;
    const unsigned m = __ballot_sync(FULL, pass);
    __syncwarp();                                   // every group has read its uniq[] entry
    if (pass) uniq[V + __popc(m & ((1u << lane) - 1))] = cid;    // V + ... <= mine: in place is safe
    V += __popc(m);
    __syncwarp();
  }

  // ---- exact: the survivors, eight per iteration (as supercharge_fast_kernel) ---------------
  for (int base = 0; base < V; base += 8) {
    u32 cid[2];
    float v[2];
    float m[2][EPL];
#pragma unroll
    for (int h2 = 0; h2 < 2; h2++) {
      int mine = base + 4 * h2 + grp;
      cid[h2] = uniq[mine < V ? mine : base];
      const float *crow = points + (size_t)cid[h2] * D;
#pragma unroll
      for (int s = 0; s < EPL; s++) m[h2][s] = crow[g + 8 * s];
    }
#pragma unroll
    for (int h2 = 0; h2 < 2; h2++) {
#pragma unroll
      for (int s = 0; s < EPL; s++) {
        float df = q[s] - m[h2][s];
        m[h2][s] = df * df;
      }
#pragma unroll
      for (int h = EPL / 2; h >= 1; h >>= 1)
#pragma unroll
        for (int s = 0; s < h; s++) m[h2][s] = m[h2][s] + m[h2][s + h];
      float tt = m[h2][0];
      tt = tt + __shfl_xor_sync(FULL, tt, 4);
      tt = tt + __shfl_xor_sync(FULL, tt, 2);
      tt = tt + __shfl_xor_sync(FULL, tt, 1);
      v[h2] = (base + 4 * h2 + grp < V) ? tt : inf;
    }
    if (__any_sync(FULL, v[0] <= tau || v[1] <= tau)) {
#pragma unroll
      for (int h2 = 0; h2 < 2; h2++)
        for (int i = 0; i < 4; i++) {
          float vn = __shfl_sync(FULL, v[h2], 8 * i);
          u32 idn = __shfl_sync(FULL, cid[h2], 8 * i);
          if (vn <= tau && vn != inf && !best.contains(idn)) {
            if (vn < tau) {
              if (__any_sync(FULL, best.v[0] == vn)) tie = true;
              best.insert(vn, idn, k, sentinel, lane);
              tau = best.kth(k);
            } else {
              tie = true;
            }
          }
        }
    }
  }

  // ---- prefix corner from brackets ----------------------------------------------------------
  if (corner != sentinel) {
    if (lane < k && own_v != inf) {
      if (own_reg == corner) corner_hi = fmaxf(corner_hi, own_v * scale2);
      else far_lo = fmaxf(far_lo, own_v * scale2);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      corner_hi = fmaxf(corner_hi, __shfl_xor_sync(FULL, corner_hi, o));
      far_lo = fmaxf(far_lo, __shfl_xor_sync(FULL, far_lo, o));
    }
    // corner id in the prefix and nothing provably farther: the literal kernel decides
    if (corner_hi != -inf && !(far_lo > corner_hi)) tie = true;
  }
  size_t orow = x - row_begin;
  if (lane < k) {
    out_ids[orow * (size_t)k + lane] = best.id[0];
    if (out_dist) out_dist[orow * (size_t)k + lane] = best.v[0];
  }
  if (tie && lane == 0) tie_report(ties, (u32)orow);
}
