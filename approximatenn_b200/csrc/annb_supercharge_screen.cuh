// annb_supercharge_screen.cuh — S5 with an fp16 screen (float build; included by annb_finish.cu).
//
// supercharge_fast_kernel gathers every candidate's fp32 row (256 B at d = 64, as 32-byte pieces
// of eight different 128-byte lines per load instruction) although only a few per cent of the
// k*k candidates can beat the row's current k-th best (cfg3: 10 of 235).  Here every candidate
// is bracketed first from the fp16 copy of the points in ORIGINAL order, and only candidates
// whose lower bound reaches tau0 = the row's k-th own distance get the exact tree.  tau only
// shrinks while the row is processed, so tau0 is conservative; whatever is reported is computed
// by the exact tree (same operations, same order, same bits as the fast kernel).
//
// Bracket: the one of the screened S3 path (DESIGN.md "screened leaf" derives it,
// tests/test_screen_bounds.py replays it): c' = fp16((x - mean) * scale), per row
// (s, n2) = (sqrt(kappa) (||c|| (1 + 2^-12) + sqrt(d) 2^-14), sum c'^2) from the same preparation
// kernel, D' = n2q + n2c - 2 q'.c' with the dot product on the tensor cores (mma.m16n8k16, fp16
// in, fp32 accumulate), t = sq + sc, and |D - D'| <= t^2 for the exact path's value D in scaled
// units.  One warp owns one row: all 16 rows of the A operand are the query, the B operand is
// eight candidates, so after d/16 MMAs lane (g, t) holds the dot products of candidates 2t and
// 2t+1 of the round — no shuffles, ~4 instructions per candidate.
//
// Prefix corner (alg.c:313-327 via rdups): it removes the id in slot P2 from the k best only if
// that id is the largest entry of the prefix.  If at least one candidate was screened out it is
// strictly farther than tau0 >= every entry of the final list, so the prefix's largest entry is
// not in the list and the rule cannot fire.  If nothing was screened out every candidate went
// through the exact tree and the rule is applied exactly as in the fast kernel.
#pragma once

static __device__ unsigned long long s5_screen_stats_dev[2];        // candidates bracketed, candidates measured exactly
extern "C" void annb_supercharge_screen_stats(unsigned long long out[2], int reset) {
  unsigned long long z[2] = {0, 0};
  cudaMemcpyFromSymbol(out, s5_screen_stats_dev, sizeof z);
  if (reset) cudaMemcpyToSymbol(s5_screen_stats_dev, z, sizeof z);
}

template <int D>
__global__ void __launch_bounds__(256, D >= 128 ? 2 : 4)
supercharge_screen_kernel(const float *__restrict__ points, const unsigned short *__restrict__ points16,
                          const float2 *__restrict__ pnrm, const unsigned *__restrict__ scale_bits,
                          const u32 *__restrict__ own_ids,
                          const float *__restrict__ own_dist, const u32 *__restrict__ graph, size_t n,
                          int k, size_t row_begin, size_t row_end, const u32 *__restrict__ row_perm, size_t perm_base,
                          u32 *__restrict__ out_ids, float *__restrict__ out_dist, TieList ties) {
  constexpr int LPC = D / 8;                         // exact tree: lanes per candidate, 32 bytes of fp32 each
  constexpr int CPR = 32 / LPC;                      // exact tree: candidates per round
  constexpr int KS2 = D >= 32 ? D / 32 : 1;          // screen: MMAs per group of 8 candidates (32 coordinates each)
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned s_stats[2];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  if (threadIdx.x < 2) s_stats[threadIdx.x] = 0;
  __syncthreads();
  const size_t pos = row_begin + (((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (pos < row_end) {
    const size_t x = row_perm ? perm_base + (size_t)row_perm[pos - perm_base] : pos;
    const u32 sentinel = (u32)n;
    const float inf = ft_inf();
    const int wide = k * (k + 1);
    const int P2 = 1 << floor_log2_u((unsigned long long)wide);
    const int cand = P2 - k;
    u32 *uniq = reinterpret_cast<u32 *>(smem_raw) + (size_t)wib * ((size_t)k * k + 16);

    // scale of the fp16 copy (a power of two); outside a sane exponent range the screen passes
    // everything (the exact path's rounding is only relative away from underflow)
    float scale2 = 0.f;
    {
      const float cmax = __uint_as_float(*scale_bits);
      if (cmax > 0.f && cmax <= 3.0e38f) {
        const int e = ilogbf(cmax);
        if (e >= -30 && e <= 30) { const float sc = ldexpf(1.0f, 2 - e); scale2 = sc * sc; }
      }
    }

    WarpList<1> best;
    best.v[0] = lane < k ? own_dist[x * (size_t)k + lane] : inf;
    best.id[0] = lane < k ? own_ids[x * (size_t)k + lane] : sentinel;
    const u32 own_reg = best.id[0];
    bool tie = false;
    bool any_inf = lane < k && best.v[0] == inf;
    float max_v = (lane < k && best.v[0] != inf) ? best.v[0] : -inf;
    u32 max_id = best.id[0];
    {
      float nxt = __shfl_down_sync(FULL, best.v[0], 1);
      if (lane + 1 < k && best.v[0] == nxt && nxt != inf) tie = true;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
      float ov = __shfl_xor_sync(FULL, max_v, o);
      u32 oi = __shfl_xor_sync(FULL, max_id, o);
      if (ov > max_v || (ov == max_v && oi < max_id)) { max_v = ov; max_id = oi; }   // uniform on ties
    }
    float tau = best.kth(k);
    const float tau0s = scale2 > 0.f ? tau * scale2 : inf;            // +inf stays +inf: everything passes

    // candidate ids -> uniq[0..U): pads and the point itself are dropped here (compute.cl:145).
    // The loads of four batches are issued before any of them is consumed.
    int U = 0;
    for (int base0 = 0; base0 < cand; base0 += 128) {
      u32 cidv[4];
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int c = base0 + 32 * u + lane;
        const int j = c < cand ? c / k : 0;
        const int z = c - j * k;
        const u32 oj = __shfl_sync(FULL, own_reg, j);
        cidv[u] = (c < cand && oj < sentinel) ? graph[(size_t)oj * k + z] : sentinel;
      }
#pragma unroll
      for (int u = 0; u < 4; u++) {
        const int c = base0 + 32 * u + lane;
        bool keep = false;
        if (c < cand) {
          if (cidv[u] >= sentinel || cidv[u] == (u32)x) any_inf = true;
          else keep = true;
        }
        const unsigned m = __ballot_sync(FULL, keep);
        if (keep) uniq[U + __popc(m & ((1u << lane) - 1))] = cidv[u];
        U += __popc(m);
      }
    }
    any_inf = __any_sync(FULL, any_inf);
    tie = __any_sync(FULL, tie);
    __syncwarp();

    // ---- screen: rounds of 16 candidates.  The list is padded to whole rounds with the row's own
    // id (a valid row; dead by position).  Operand layout (no register shuffling at all): lane
    // (g8, t4) loads the 16 bytes at 64 v + 16 t4 of candidate g8's fp16 row straight into the four
    // A registers of k-step v.  The hardware reads them as (row g8, k 2t..2t+1), (row g8+8, same
    // k), (row g8, k 2t+8..), (row g8+8, ...), i.e. MMA rows g8 and g8+8 are the SAME candidate
    // with the words (x, z) and (y, w) of every 16-byte piece.  The B operand carries the query's
    // (x, z) words in its even columns and its (y, w) words in its odd columns, so
    // D[g8][even] + D[g8+8][odd] = accumulator 0 + accumulator 3 of every lane of the quad is the
    // full dot product.  Two such groups (candidates g8 and g8+8 of the round) per round; lanes
    // t4 = 0 / 1 finish one candidate each: one (s, n2) load, five flops, one vote.  Survivors are
    // compacted in place into uniq[0..V): a round's ids are read before the previous round's
    // survivors are written, and V never overtakes the round being evaluated. -------------------
    const int g8 = lane >> 2, t4 = lane & 3;
    const int U16 = (U + 15) & ~15;
    if (lane < U16 - U) uniq[U + lane] = (u32)x;
    __syncwarp();
    u32 qw[KS2][2];                                              // B operand: this lane's query words
    {
      const unsigned short *qrow = points16 + x * (size_t)D;
#pragma unroll
      for (int v = 0; v < KS2; v++) {
        if (D >= 32) {
          const uint4 w = *reinterpret_cast<const uint4 *>(qrow + 32 * v + 8 * t4);
          qw[v][0] = (g8 & 1) ? w.y : w.x;
          qw[v][1] = (g8 & 1) ? w.w : w.z;
        } else {                                                  // d = 16: 8 bytes per lane, one k half
          const uint2 w = *reinterpret_cast<const uint2 *>(qrow + 4 * t4);
          qw[v][0] = (g8 & 1) ? w.y : w.x;
          qw[v][1] = 0u;
        }
      }
    }
    const float2 qn = pnrm[x];
    int V = 0;
    {
      uint4 ra[KS2], rc[KS2];                                    // pieces of candidates g8 and g8+8
      u32 ce = 0;
      float2 ne = make_float2(0.f, 0.f);
      auto load = [&](int b) {
        const u32 c0 = uniq[b + g8], c1 = uniq[b + g8 + 8];
        const unsigned short *p0 = points16 + (size_t)c0 * D, *p1 = points16 + (size_t)c1 * D;
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
        ce = (t4 & 1) ? c1 : c0;                                  // the candidate this lane finishes
        ne = pnrm[ce];
      };
      if (U16 > 0) load(0);
      for (int base = 0; base < U16; base += 16) {
        float ca[4] = {0.f, 0.f, 0.f, 0.f}, cc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int v = 0; v < KS2; v++) {
          mma_f16_16816(ca, ra[v].x, ra[v].y, ra[v].z, ra[v].w, qw[v][0], qw[v][1]);
          mma_f16_16816(cc, rc[v].x, rc[v].y, rc[v].z, rc[v].w, qw[v][0], qw[v][1]);
        }
        const u32 cur = ce;
        const float2 cn = ne;
        if (base + 16 < U16) load(base + 16);                     // next round's loads fly during the epilogue
        const float dot = (t4 & 1) ? cc[0] + cc[3] : ca[0] + ca[3];
        const float t = qn.x + cn.x;
        const float lo = __fmaf_rn(-t, t, __fmaf_rn(-2.0f, dot, qn.y + cn.y));
        const bool pass = t4 < 2 && base + g8 + 8 * (t4 & 1) < U && lo <= tau0s;
        const unsigned m = __ballot_sync(FULL, pass);
        if (m) {
          if (pass) uniq[V + __popc(m & ((1u << lane) - 1))] = cur;
          V += __popc(m);
        }
      }
    }
    __syncwarp();

    // ---- the query's fp32 pieces for the exact tree ---------------------------------------------
    const int g = lane & (LPC - 1), grp = lane / LPC;
    const float4 qa = *reinterpret_cast<const float4 *>(points + x * (size_t)D + 4 * g);
    const float4 qb = *reinterpret_cast<const float4 *>(points + x * (size_t)D + D / 2 + 4 * g);

    // ---- exact tree for the survivors, two rounds of CPR candidates per iteration: lane g of a
    // candidate's LPC lanes holds coordinates 4g..4g+3 and d/2+4g..d/2+4g+3 (two 16-byte loads,
    // whole lines per instruction).  The reference's tree (compute.cl:160-167): stride d/2 is
    // lane-local, strides d/4..4 are xor-shuffles LPC/2..1 (a + b == b + a bit for bit, so
    // partners agree), strides 2 and 1 are lane-local again. --------------------------------
    float lmax_v = -inf;
    u32 lmax_id = sentinel;
    for (int base = 0; base < V; base += 2 * CPR) {
      u32 cid[2];
      float dist[2];
      float4 ca[2], cb[2];
#pragma unroll
      for (int h2 = 0; h2 < 2; h2++) {
        const int mine = base + CPR * h2 + grp;
        cid[h2] = uniq[mine < V ? mine : base];
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
        const bool live = base + CPR * h2 + grp < V;
        dist[h2] = live ? tt : inf;
        if (live && tt > lmax_v) { lmax_v = tt; lmax_id = cid[h2]; }
      }
      if (__any_sync(FULL, dist[0] <= tau || dist[1] <= tau)) {
#pragma unroll
        for (int h2 = 0; h2 < 2; h2++)
          for (int i = 0; i < CPR; i++) {
            float vn = __shfl_sync(FULL, dist[h2], LPC * i);
            u32 idn = __shfl_sync(FULL, cid[h2], LPC * i);
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
    // prefix corner: only when nothing was screened out (see the header)
    if (P2 < wide && !any_inf && V == U) {
#pragma unroll
      for (int o = 16; o >= 1; o >>= 1) {
        float ov = __shfl_xor_sync(FULL, lmax_v, o);
        u32 oi = __shfl_xor_sync(FULL, lmax_id, o);
        if (ov > lmax_v || (ov == lmax_v && oi < lmax_id)) { lmax_v = ov; lmax_id = oi; }   // uniform on ties
      }
      if (lmax_v > max_v) { max_v = lmax_v; max_id = lmax_id; }
      int c = P2 - k, j = c / k, z = c - j * k;
      u32 oj = __shfl_sync(FULL, own_reg, j);
      u32 cid = oj < sentinel ? graph[(size_t)oj * k + z] : sentinel;
      if (cid == max_id) best.remove(cid, sentinel, lane);
    }
    const size_t orow = x - row_begin;
    if (lane < k) {
      out_ids[orow * (size_t)k + lane] = best.id[0];
      if (out_dist) out_dist[orow * (size_t)k + lane] = best.v[0];
    }
    if (tie && lane == 0) tie_report(ties, (u32)orow);
    if (lane == 0) { atomicAdd(&s_stats[0], (unsigned)U); atomicAdd(&s_stats[1], (unsigned)V); }
  }
  __syncthreads();
  if (threadIdx.x < 2 && s_stats[threadIdx.x])
    atomicAdd(&s5_screen_stats_dev[threadIdx.x], (unsigned long long)s_stats[threadIdx.x]);
}
