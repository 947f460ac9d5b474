// annb_supercharge_screen.cuh — S5 with an fp16 screen (float build; included by annb_finish.cu).
//
// supercharge_fast_kernel gathers every candidate's fp32 row (256 B at d = 64, as 32-byte pieces
// of eight different 128-byte lines per load instruction) although only a few per cent of the
// k*k candidates can beat the row's current k-th best.  Here every candidate is bracketed first
// from the fp16 copy of the points in ORIGINAL order (one 128-byte line per row at d = 64, read
// by d/8 lanes with one 16-byte load each, i.e. whole lines per instruction), and only
// candidates whose lower bound reaches tau0 = the row's k-th own distance get the exact tree.
// tau only shrinks while the row is processed, so tau0 is conservative; whatever is reported
// is computed by the exact tree (same operations, same order, same bits as the fast kernel).
//
// Bracket (DESIGN.md "screened leaf" derives it; tests/test_screen_bounds.py replays it): with
// c' = fp16((x - mean) * scale), D' = n2q + n2c - 2 q'.c' in fp32 and t = sq + sc,
// |D - D'| <= t^2 for the exact path's value D in scaled units.  This kernel has no norm table:
// n2 = sum c'^2 is accumulated from the row it has just loaded (FHFMA: fp16 x fp16 + fp32, one
// instruction, products exact), and the norm bound is taken from it:
//   ||c|| <= ||c'|| (1 + 2^-10) + sqrt(d) 2^-24   (fp16 rounding incl. subnormals)
//   s' = sqrt(kappa) (sqrt(n2) (1 + 2^-9) + sqrt(d) 2^-13) >= sqrt(kappa) (||c|| (1 + 2^-12) + sqrt(d) 2^-14) = s,
// the extra 2^-10 covering the rounding of n2 (<= d 2^-24 relative) and of sqrtf.  A larger t
// only widens the bracket.
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

// dot += a . b and nn += b . b over the 8 fp16 values of two 16-byte pieces (FHFMA)
__device__ __forceinline__ void half8_dot_norm(const uint4 &a, const uint4 &b, float &dot, float &nn) {
  const unsigned as[4] = {a.x, a.y, a.z, a.w}, bs[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 4; i++) {
    asm("{\n\t.reg .f16 a0, a1, b0, b1;\n\t"
        "mov.b32 {a0, a1}, %2;\n\tmov.b32 {b0, b1}, %3;\n\t"
        "fma.rn.f32.f16 %0, a0, b0, %0;\n\tfma.rn.f32.f16 %0, a1, b1, %0;\n\t"
        "fma.rn.f32.f16 %1, b0, b0, %1;\n\tfma.rn.f32.f16 %1, b1, b1, %1;\n\t}"
        : "+f"(dot), "+f"(nn) : "r"(as[i]), "r"(bs[i]));
  }
}

template <int D>
__global__ void __launch_bounds__(256)
supercharge_screen_kernel(const float *__restrict__ points, const unsigned short *__restrict__ points16,
                          const unsigned *__restrict__ scale_bits, const u32 *__restrict__ own_ids,
                          const float *__restrict__ own_dist, const u32 *__restrict__ graph, size_t n,
                          int k, size_t row_begin, size_t row_end, const u32 *__restrict__ row_perm,
                          u32 *__restrict__ out_ids, float *__restrict__ out_dist, TieList ties) {
  constexpr int LPC = D / 8;                         // lanes per candidate: 16 bytes of fp16 / 32 bytes of fp32 each
  constexpr int CPR = 32 / LPC;                      // candidates per round
  constexpr int SR = 4;                              // screen rounds in flight
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned s_stats[2];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  if (threadIdx.x < 2) s_stats[threadIdx.x] = 0;
  __syncthreads();
  const size_t pos = row_begin + (((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (pos < row_end) {
    const size_t x = row_perm ? (size_t)row_perm[pos] : pos;
    const u32 sentinel = (u32)n;
    const float inf = ft_inf();
    const int wide = k * (k + 1);
    const int P2 = 1 << floor_log2_u((unsigned long long)wide);
    const int cand = P2 - k;
    u32 *uniq = reinterpret_cast<u32 *>(smem_raw) + (size_t)wib * ((size_t)k * k);
    const int g = lane & (LPC - 1), grp = lane / LPC;

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
      if (ov > max_v) { max_v = ov; max_id = oi; }
    }
    float tau = best.kth(k);
    const float tau0s = scale2 > 0.f ? tau * scale2 : inf;            // +inf stays +inf: everything passes

    // candidate ids -> uniq[0..U): pads and the point itself are dropped here (compute.cl:145)
    int U = 0;
    for (int base = 0; base < cand; base += 32) {
      int c = base + lane;
      int j = c < cand ? c / k : 0;
      int z = c - j * k;
      u32 oj = __shfl_sync(FULL, own_reg, j);
      u32 cid = (c < cand && oj < sentinel) ? graph[(size_t)oj * k + z] : sentinel;
      bool keep = false;
      if (c < cand) {
        if (cid >= sentinel || cid == (u32)x) any_inf = true;
        else keep = true;
      }
      unsigned m = __ballot_sync(FULL, keep);
      if (keep) uniq[U + __popc(m & ((1u << lane) - 1))] = cid;
      U += __popc(m);
    }
    any_inf = __any_sync(FULL, any_inf);
    tie = __any_sync(FULL, tie);
    __syncwarp();

    // ---- the query: fp16 piece for the screen, fp32 pieces for the exact tree -----------------
    const uint4 q16 = *reinterpret_cast<const uint4 *>(points16 + x * (size_t)D + 8 * g);
    float qn2 = 0.f;
    {
      float dummy = 0.f;
      half8_dot_norm(q16, q16, dummy, qn2);
#pragma unroll
      for (int o = LPC / 2; o >= 1; o >>= 1) qn2 += __shfl_xor_sync(FULL, qn2, o);
    }
    const float root_d = sqrtf((float)D);
    const float sq = SCREEN_SQRT_KAPPA * (sqrtf(qn2) * (1.0f + 1.0f / 512.0f) + root_d * (1.0f / 8192.0f));
    const float4 qa = *reinterpret_cast<const float4 *>(points + x * (size_t)D + 4 * g);
    const float4 qb = *reinterpret_cast<const float4 *>(points + x * (size_t)D + D / 2 + 4 * g);

    // ---- screen: CPR candidates per round, SR rounds in flight; survivors are compacted in
    // place into uniq[0..V) (V never overtakes the round being evaluated) ----------------------
    int V = 0;
    {
      uint4 v[SR];
      u32 cid[SR];
      auto load = [&](int b, uint4 &vv, u32 &cc) {
        const int mine = b + grp;
        cc = uniq[mine < U ? mine : U - 1];
        vv = *reinterpret_cast<const uint4 *>(points16 + (size_t)cc * D + 8 * g);
      };
      auto eval = [&](int b, const uint4 &vv, u32 cc) {
        float dot = 0.f, cn2 = 0.f;
        half8_dot_norm(q16, vv, dot, cn2);
#pragma unroll
        for (int o = LPC / 2; o >= 1; o >>= 1) {
          dot += __shfl_xor_sync(FULL, dot, o);
          cn2 += __shfl_xor_sync(FULL, cn2, o);
        }
        const float sc = SCREEN_SQRT_KAPPA * (sqrtf(cn2) * (1.0f + 1.0f / 512.0f) + root_d * (1.0f / 8192.0f));
        const float t = sq + sc;
        const float lo = ((qn2 + cn2) - 2.0f * dot) - t * t;
        const bool pass = (b + grp < U) && g == 0 && lo <= tau0s;
        const unsigned m = __ballot_sync(FULL, pass);
        if (pass) uniq[V + __popc(m & ((1u << lane) - 1))] = cc;
        V += __popc(m);
      };
      if (U > 0) {
#pragma unroll
        for (int r = 0; r < SR; r++)
          if (r * CPR < U) load(r * CPR, v[r], cid[r]);
        for (int base = 0; base < U; base += SR * CPR) {
#pragma unroll
          for (int r = 0; r < SR; r++) {
            const int b = base + r * CPR;
            if (b < U) {
              eval(b, v[r], cid[r]);
              if (b + SR * CPR < U) load(b + SR * CPR, v[r], cid[r]);
            }
          }
        }
      }
    }
    __syncwarp();

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
        if (ov > lmax_v) { lmax_v = ov; lmax_id = oi; }
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
