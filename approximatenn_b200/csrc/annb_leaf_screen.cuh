// annb_leaf_screen.cuh — S3, screened path (float build; included by annb_leaf.cu).
//
// The fp32 work of S3 is (points) x (candidates in the row prefix) exact distances, but only
// the k smallest of each row survive.  This path brackets every distance with a tensor-core
// approximation first and sends only the candidates whose bracket reaches the k-th best
// through the exact tree (same operations, same order, same bits as the tiled kernel):
//
//   scale     once per point set: a power of two that puts max |x - mean| in [4, 8).
//   prep      once per try, fused with S2's gather: c = (x - mean) * scale, c' = fp16(c);
//             sp16 row = c', nrm = (s, n2) with
//             s = sqrt(kappa) * (||c|| (1 + 2^-12) + sqrt(d) 2^-14), n2 = sum c'^2.
//   pass 1    16 queries of a bucket against the candidate stream, 8 candidates per
//             mma.m16n8k16 (fp16 in, fp32 accumulate): D' = n2q + n2c - 2 q'.c'.  With
//             t = sq + sc the exact-path value D (in scaled units) satisfies |D - D'| <= t^2
//             (DESIGN.md "screened leaf": fp16 rounding of both rows incl. subnormals, the
//             centring subtraction, the accumulation and the exact path's own rounding;
//             tests/test_screen_bounds.py checks the bracket on the CPU).
//             lo = D' - t^2 is parked in shared memory as fp16 (rounded towards zero, negative
//             values clamped to 0: still a lower bound); each lane keeps, per query and column
//             parity, the 4 smallest upper bounds hi = D' + t^2 it has seen (as fp16, rounded
//             to nearest: monotone, undone by a factor at the end).  The four lanes that share
//             a query then hold 32 upper bounds of 32 different candidates, and Theta = the
//             16th smallest of them bounds the 16th (hence k-th, k <= 16) smallest exact
//             distance of the row from above.
//   scan      lo > Theta means strictly farther than the k-th best (not even a tie): dropped.
//             Lane (query, half of the stream) collects the others, about k + 2 per query.
//   exact     the tile's (query, survivor) pairs as one flat list: 4 lanes per pair reading
//             64 contiguous bytes per instruction, three rounds of 8 pairs in flight, query
//             rows from shared memory (which reuses the parked-bounds area).
//   rank      position = number of strictly smaller survivors of the same query; equal
//             counts below k are exact ties and send the row to the literal kernel.
// Buckets the screen cannot hold (more candidates than the shared-memory tables, more than 32
// survivors in one half of a query's stream, or more than 416 pairs in a tile) are appended to
// a list and done by the tiled kernel afterwards, so the result never depends on how well the
// screen did.

static __device__ unsigned long long leaf_exact_pairs_dev;   // pairs that reached the exact tree
static __device__ unsigned long long leaf_overflow_dev;      // buckets handed to the tiled kernel
extern "C" unsigned long long annb_leaf_exact_pairs(int reset) {
  unsigned long long v = 0, z = 0;
  cudaMemcpyFromSymbol(&v, leaf_exact_pairs_dev, sizeof v);
  if (reset) cudaMemcpyToSymbol(leaf_exact_pairs_dev, &z, sizeof z);
  return v;
}
extern "C" unsigned long long annb_leaf_overflow_buckets(int reset) {
  unsigned long long v = 0, z = 0;
  cudaMemcpyFromSymbol(&v, leaf_overflow_dev, sizeof v);
  if (reset) cudaMemcpyToSymbol(leaf_overflow_dev, &z, sizeof z);
  return v;
}


// max |x - mean| over the set, as the bits of a non-negative float
__global__ void __launch_bounds__(256)
screen_maxabs_kernel(const float *__restrict__ sp, const float *__restrict__ mean, size_t n, int d,
                     unsigned *__restrict__ maxbits) {
  const size_t total4 = n * (size_t)d / 4;
  const int d4 = d >> 2;
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total4; i += (size_t)gridDim.x * blockDim.x) {
    float4 x = reinterpret_cast<const float4 *>(sp)[i];
    float4 mu = make_float4(0, 0, 0, 0);
    if (mean) mu = reinterpret_cast<const float4 *>(mean)[i % d4];
    m = fmaxf(m, fmaxf(fmaxf(fabsf(x.x - mu.x), fabsf(x.y - mu.y)), fmaxf(fabsf(x.z - mu.z), fabsf(x.w - mu.w))));
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) m = fmaxf(m, __shfl_xor_sync(FULL, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(maxbits, __float_as_uint(m));
}

__device__ __forceinline__ float screen_scale(unsigned maxbits) {
  float cmax = __uint_as_float(maxbits);
  if (!(cmax > 0.f) || cmax > 3.0e38f) return 1.0f;
  int e = ilogbf(cmax);                                  // cmax in [2^e, 2^(e+1))
  e = max(-100, min(100, e));
  return ldexpf(1.0f, 2 - e);
}

// c = (x - mean) * scale, fp16 copy and the two norms of one row; d/4 lanes per row.  With
// `order` the row is gathered from the original array (src[order[row]]) and also written to
// the bucket-ordered copy `sorted` — S2's gather and this preparation in one pass.
__global__ void __launch_bounds__(256)
screen_prep_kernel(const float *__restrict__ src, const u32 *__restrict__ order, float *__restrict__ sorted,
                   const float *__restrict__ mean, size_t n, int d, const unsigned *__restrict__ maxbits,
                   unsigned short *__restrict__ sp16, float2 *__restrict__ nrm) {
  const int lpr = d >> 2;
  size_t gid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t row = gid / lpr;
  int l = (int)(gid - row * lpr);
  const bool live = row < n;
  const float scale = screen_scale(*maxbits);
  float4 x = make_float4(0, 0, 0, 0), m = make_float4(0, 0, 0, 0);
  if (live) {
    const size_t from = order ? (size_t)order[row] : row;
    x = reinterpret_cast<const float4 *>(src + from * (size_t)d)[l];
    if (sorted) reinterpret_cast<float4 *>(sorted + row * (size_t)d)[l] = x;
    if (mean) m = reinterpret_cast<const float4 *>(mean)[l];
  }
  float c0 = (x.x - m.x) * scale, c1 = (x.y - m.y) * scale, c2 = (x.z - m.z) * scale, c3 = (x.w - m.w) * scale;
  __half2 p01 = __floats2half2_rn(c0, c1), p23 = __floats2half2_rn(c2, c3);
  float r0 = __low2float(p01), r1 = __high2float(p01), r2 = __low2float(p23), r3 = __high2float(p23);
  float ss = (c0 * c0 + c1 * c1) + (c2 * c2 + c3 * c3);
  float n2 = (r0 * r0 + r1 * r1) + (r2 * r2 + r3 * r3);
  for (int o = lpr >> 1; o >= 1; o >>= 1) {             // lpr is a power of two <= 32
    ss += __shfl_xor_sync(FULL, ss, o);
    n2 += __shfl_xor_sync(FULL, n2, o);
  }
  if (live) {
    uint2 packed;
    packed.x = *reinterpret_cast<unsigned *>(&p01);
    packed.y = *reinterpret_cast<unsigned *>(&p23);
    reinterpret_cast<uint2 *>(sp16 + row * (size_t)d)[l] = packed;
    if (l == 0 && nrm) {
      float s = sqrtf(ss) * (1.0f + 1.0f / 4096.0f) + sqrtf((float)d) * (1.0f / 16384.0f);
      nrm[row] = make_float2(s * SCREEN_SQRT_KAPPA, n2);
    }
  }
}

// (lo, hi) -> fp16x2 rounded towards zero with negative inputs clamped to 0: a value that is
// never above the input, i.e. still a lower bound.  +inf stays +inf.
__device__ __forceinline__ u32 pack_lower_bounds(float lo_half, float hi_half) {
  u32 r;
  asm("cvt.rz.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_half), "f"(lo_half));
  return r;
}

// packed fp32x2 helpers for the bounds (plain IEEE operations; contraction does not matter here)
__device__ __forceinline__ f32x2 pack2(float lo, float hi) {
  f32x2 r;
  asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(f32x2 v, float &lo, float &hi) {
  asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) {
  f32x2 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
  f32x2 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
// (lo, hi) -> fp16x2 round-to-nearest with negative inputs clamped to 0 (monotone)
__device__ __forceinline__ u32 pack_upper_bounds(float lo_half, float hi_half) {
  u32 r;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi_half), "f"(lo_half));
  return r;
}
__device__ __forceinline__ u32 hmin2u(u32 a, u32 b) {
  u32 r;
  asm("min.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
__device__ __forceinline__ u32 hmax2u(u32 a, u32 b) {
  u32 r;
  asm("max.f16x2 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(b));
  return r;
}
// two ascending lists of 4 side by side (low half: even columns, high half: odd columns)
__device__ __forceinline__ void keep_smallest4x2(u32 (&h)[4], u32 x) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    u32 lo = hmin2u(h[i], x);
    x = hmax2u(h[i], x);
    h[i] = lo;
  }
}

#define SCR_CE(a, b) { float lo_ = fminf(a, b), hi_ = fmaxf(a, b); (a) = lo_; (b) = hi_; }
// ascending sort of a bitonic sequence of 8
__device__ __forceinline__ void bitonic_merge8(float (&m)[8]) {
#pragma unroll
  for (int i = 0; i < 4; i++) SCR_CE(m[i], m[i + 4])
  SCR_CE(m[0], m[2]) SCR_CE(m[1], m[3]) SCR_CE(m[4], m[6]) SCR_CE(m[5], m[7])
  SCR_CE(m[0], m[1]) SCR_CE(m[2], m[3]) SCR_CE(m[4], m[5]) SCR_CE(m[6], m[7])
}

// The four lanes of a quad each hold two ascending lists of 4 (the halves of h); returns (to
// all four) the 16th smallest of the 32 values.
__device__ __forceinline__ float quad_sixteenth_smallest(const u32 (&h)[4], int t) {
  float a[4], b[4];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    __half2 v = *reinterpret_cast<const __half2 *>(&h[i]);
    a[i] = __low2float(v);
    b[i] = __high2float(v);
  }
  float m[8] = {a[0], a[1], a[2], a[3], b[3], b[2], b[1], b[0]};
  bitonic_merge8(m);                                              // own 8, ascending
  float c[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    float other = __shfl_xor_sync(FULL, m[7 - i], 1);
    c[i] = (t & 1) ? fmaxf(m[i], other) : fminf(m[i], other);   // even lane: the 8 smallest of the pair's 16
  }
  bitonic_merge8(c);                                              // lanes (t, t^1): 16 ascending, low half in the even lane
  float w = -ft_inf();
#pragma unroll
  for (int r = 0; r < 8; r++) {
    float other = __shfl_xor_sync(FULL, c[7 - r], 3);            // A_i against B_(15-i)
    w = fmaxf(w, fminf(c[r], other));
  }
  return fmaxf(w, __shfl_xor_sync(FULL, w, 1));
}

struct ScreenOverflow {
  u32 *count;      // [1]
  u32 *buckets;    // [number of buckets]
};

static constexpr int SCREEN_HALF = 32;                 // survivor slots per (query, half of the stream)

static constexpr int SCREEN_PAIRS = 416;               // (query, survivor) pairs per tile of 16 queries (cfg3: 300 +- 20)
template <int D> struct ScreenOverlay {                // what replaces the parked bounds after the scan
  static constexpr int QROW = D + 4;                   // padded query row (floats)
  static constexpr size_t bytes = 16 * QROW * 4 + SCREEN_PAIRS * (4 + 4 + 2 + 2 + 1 + 1);
};
static int screen_min_ct(size_t overlay_bytes) {       // smallest table size (= 8 mod 64) whose bounds area holds the overlay
  size_t ct = (overlay_bytes + 31) / 32;
  return (int)(((ct + 55) / 64) * 64 + 8);
}
static size_t screen_smem_bytes(int ct) {
  return (size_t)ct * (4 + 4 + 4 + 32) + 32 * SCREEN_HALF * 2 + (16 * 5 + 33 + 32) * 4 + 64;
}

// less += (a < b), as exactly two instructions
__device__ __forceinline__ void count_less(u32 &less, float a, float b) {
  asm("{\n\t.reg .pred p;\n\tsetp.lt.f32 p, %1, %2;\n\t@p add.u32 %0, %0, 1;\n\t}" : "+r"(less) : "f"(a), "f"(b));
}

// Build-time knobs of the kernel (tools/variants.py builds and times alternatives side by side;
// cfg3, leaf stage per step: depth 3 / tables 0: 9.99 ms, 4 / 0: 9.95, 3 / 1: 9.82, 4 / 1: 9.77):
//   SCR_DEPTH    candidate tiles of pass 1 in flight (3 or 4)
//   SCR_TABLES4  1: the candidate tables are filled four stream positions per lane at a time (the
//                norm loads of a bucket overlap) and the next bucket's ticket is drawn one bucket ahead
//   SCR_REGCAP   register cap
#ifndef SCR_DEPTH
#define SCR_DEPTH 4
#endif
#ifndef SCR_TABLES4
#define SCR_TABLES4 1
#endif
#ifndef SCR_REGCAP
#define SCR_REGCAP 160
#endif
template <int D>
__global__ void __maxnreg__(SCR_REGCAP)
leaf_screen_kernel(const float *__restrict__ sp, const unsigned short *__restrict__ sp16,
                   const float2 *__restrict__ nrm, const u32 *__restrict__ order,
                   const u32 *__restrict__ offset, const u32 *__restrict__ tmax_p, size_t n,
                   size_t buckets, int d_short, int k, u32 *__restrict__ list_ids,
                   float *__restrict__ list_dist, TieList ties, unsigned long long negzero2,
                   u32 *__restrict__ ticket, int CT, ScreenOverflow ovf,
                   const float *__restrict__ cutoff, const unsigned *__restrict__ scale_bits) {
  constexpr int KS = D / 16;                                           // k-steps per candidate tile
  constexpr int NV = D / 16;                                           // 16-byte pieces per lane in the exact tree
  constexpr int QROW = ScreenOverlay<D>::QROW;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  // shared memory: candidate tables, parked lower bounds (later the tile's pair tables), lists
  float *cand_s = reinterpret_cast<float *>(smem_raw);                 // [CT]  sqrt(kappa) * norm bound
  float *cand_n2 = cand_s + CT;                                        // [CT]  squared norm of the fp16 row
  u32 *cand_row = reinterpret_cast<u32 *>(cand_n2 + CT);               // [CT]  bucket-ordered row
  unsigned short *lo_h = reinterpret_cast<unsigned short *>(cand_row + CT);   // [16][CT] fp16
  float *theta = reinterpret_cast<float *>(lo_h + 16 * (size_t)CT);    // [16]
  u32 *qoffs = reinterpret_cast<u32 *>(theta + 16);                    // [16] first pair of a query
  u32 *qcnt = qoffs + 16;                                              // [16] survivors of a query
  u32 *qids = qcnt + 16;                                               // [16] point id of a query
  u32 *qtie = qids + 16;                                               // [16]
  u32 *segpos = qtie + 16;                                             // [33]
  u32 *segrow = segpos + 33;                                           // [32]
  unsigned short *slist = reinterpret_cast<unsigned short *>(segrow + 32);    // [32][SCREEN_HALF] candidate index
  const u32 sentinel = (u32)n;
  const float inf = ft_inf();
  // cutoff (optional): per point an upper bound of its FINAL k-th distance, known from earlier
  // tries (cutoff_update_kernel).  In the units of the brackets it is cutoff * scale^2; outside
  // the trusted range of the scale it is ignored.
  double cut_scale2 = 0.0;
  if (cutoff) {
    const float sc = screen_scale(*scale_bits);
    if (sc >= 1.0f / 268435456.0f && sc <= 4294967296.0f) cut_scale2 = (double)sc * (double)sc;
  }

#if SCR_TABLES4
  u32 next_ticket = 0;
  if (lane == 0) next_ticket = atomicAdd(ticket, 1u);
#endif
  for (;;) {
    size_t b = 0;
#if SCR_TABLES4
    b = __shfl_sync(FULL, next_ticket, 0);
    if (b >= buckets) return;
    if (lane == 0) next_ticket = atomicAdd(ticket, 1u);               // waited for one bucket later
#else
    if (lane == 0) b = atomicAdd(ticket, 1u);
    b = __shfl_sync(FULL, (u32)b, 0);
    if (b >= buckets) return;
#endif
    const u32 beg = offset[b], Q = offset[b + 1] - beg;
    if (Q == 0) continue;
    const unsigned long long tmax = *tmax_p;
    const unsigned long long L = (unsigned long long)(d_short + 1) * tmax;
    const unsigned long long P = 1ull << floor_log2_u(L);
    __syncwarp();
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
    const u32 C16 = (C + 15) & ~15u;                                   // both halves of the scan in 8-column steps
    bool overflow = C16 > (u32)CT;
    const u32 own_len = segpos[1];                                     // candidates taken from the own bucket

    if (!overflow) {
      int seg = 0;
#if SCR_TABLES4
      for (u32 j0 = lane; j0 < C16; j0 += 128) {
        u32 rows[4];
        float2 nr[4];
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const u32 j = j0 + 32 * u;
          rows[u] = beg;
          if (j < C) {
            while (j >= segpos[seg + 1]) seg++;
            rows[u] = segrow[seg] + (j - segpos[seg]);
          }
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          nr[u] = make_float2(0.f, inf);                               // tail of the last tile: D' = +inf
          if (j0 + 32 * u < C) nr[u] = nrm[rows[u]];
        }
#pragma unroll
        for (int u = 0; u < 4; u++) {
          const u32 j = j0 + 32 * u;
          if (j < C16) {
            cand_row[j] = rows[u];
            cand_s[j] = nr[u].x;
            cand_n2[j] = nr[u].y;
          }
        }
      }
#else
      for (u32 j = lane; j < C16; j += 32) {
        u32 row = beg;
        float2 nr = make_float2(0.f, inf);                             // tail of the last tile: D' = +inf
        if (j < C) {
          while (j >= segpos[seg + 1]) seg++;
          row = segrow[seg] + (j - segpos[seg]);
          nr = nrm[row];
        }
        cand_row[j] = row;
        cand_s[j] = nr.x;
        cand_n2[j] = nr.y;
      }
#endif
      __syncwarp();
    }

    for (u32 qbase = 0; qbase < Q && !overflow; qbase += 16) {
      const u32 Qp = min(16u, Q - qbase);
      // the cutoff of query (lane & 15), requested now and used by the scan
      float cut_s = inf;
      if (cut_scale2 != 0.0 && (u32)(lane & 15) < Qp) {
        const float c = cutoff[order[(size_t)beg + qbase + (lane & 15)]];
        cut_s = __double2float_ru((double)c * cut_scale2);          // exact product, rounded up
      }
      // ---- pass 1: bounds for query rows g and g + 8 of this tile ----------------------
      {
        const bool qv0 = (u32)g < Qp, qv1 = (u32)g + 8 < Qp;
        const u32 qrow0 = beg + qbase + (qv0 ? g : 0), qrow1 = beg + qbase + (qv1 ? g + 8 : 0);
        // the query itself sits in the own bucket's part of the stream
        const u32 self0 = (qv0 && qbase + g < own_len) ? qbase + g : 0xffffffffu;
        const u32 self1 = (qv1 && qbase + g + 8 < own_len) ? qbase + g + 8 : 0xffffffffu;
        u32 af[KS][4];                                                  // A fragments in mma operand order
        {
          ScreenRow<D> qa, qb;
          qa.load(sp16 + (size_t)qrow0 * D, t);
          qb.load(sp16 + (size_t)qrow1 * D, t);
#pragma unroll
          for (int ks = 0; ks < KS; ks++) {
            af[ks][0] = qa.r[2 * ks]; af[ks][1] = qb.r[2 * ks];
            af[ks][2] = qa.r[2 * ks + 1]; af[ks][3] = qb.r[2 * ks + 1];
          }
        }
        const float2 qn0 = nrm[qrow0], qn1 = nrm[qrow1];
        const f32x2 qs0 = pack2(qn0.x, qn0.x), qs1 = pack2(qn1.x, qn1.x);
        const f32x2 qm0 = pack2(qn0.y, qn0.y), qm1 = pack2(qn1.y, qn1.y);
        const f32x2 minus2 = pack2(-2.0f, -2.0f);
        u32 h0[4], h1[4];                                               // upper bounds kept: query g, query g + 8
#pragma unroll
        for (int i = 0; i < 4; i++) { h0[i] = 0x7c007c00u; h1[i] = 0x7c007c00u; }   // (+inf, +inf)

        // one candidate tile (8 stream positions from j0): products, bounds, lists, parking
        auto tile = [&](const ScreenRow<D> &rb, u32 j0) {
          float c[4] = {0.f, 0.f, 0.f, 0.f}, c2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int ks = 0; ks < KS; ks++) {
            if (ks & 1) mma_f16_16816(c2, af[ks][0], af[ks][1], af[ks][2], af[ks][3], rb.r[2 * ks], rb.r[2 * ks + 1]);
            else mma_f16_16816(c, af[ks][0], af[ks][1], af[ks][2], af[ks][3], rb.r[2 * ks], rb.r[2 * ks + 1]);
          }
          f32x2 p0 = pack2(c[0], c[1]), p1 = pack2(c[2], c[3]);         // columns (cA, cA + 1) of query g, g + 8
          if (KS > 1) {
            p0 = add2(p0, pack2(c2[0], c2[1]));
            p1 = add2(p1, pack2(c2[2], c2[3]));
          }
          const u32 cA = j0 + 2 * t;                                    // this lane's columns: cA and cA + 1
          const f32x2 cs = *reinterpret_cast<const f32x2 *>(&cand_s[cA]);
          const f32x2 cm = *reinterpret_cast<const f32x2 *>(&cand_n2[cA]);
          const f32x2 t0 = add2(qs0, cs), t1 = add2(qs1, cs);
          const f32x2 s0 = mul2(t0, t0), s1 = mul2(t1, t1);
          const f32x2 d0 = fma2(p0, minus2, add2(qm0, cm)), d1 = fma2(p1, minus2, add2(qm1, cm));
          float hiA0, hiB0, hiA1, hiB1, loA0, loB0, loA1, loB1;
          unpack2(add2(d0, s0), hiA0, hiB0);
          unpack2(add2(d1, s1), hiA1, hiB1);
          unpack2(sub2(d0, s0), loA0, loB0);
          unpack2(sub2(d1, s1), loA1, loB1);
          if (j0 < qbase + 16 && j0 + 8 > qbase) {                       // warp-uniform: tiles that can hold a query
            if (cA == self0) { hiA0 = inf; loA0 = inf; }
            if (cA + 1 == self0) { hiB0 = inf; loB0 = inf; }
            if (cA == self1) { hiA1 = inf; loA1 = inf; }
            if (cA + 1 == self1) { hiB1 = inf; loB1 = inf; }
          }
          keep_smallest4x2(h0, pack_upper_bounds(hiA0, hiB0));
          keep_smallest4x2(h1, pack_upper_bounds(hiA1, hiB1));
          *reinterpret_cast<u32 *>(&lo_h[(size_t)g * CT + cA]) = pack_lower_bounds(loA0, loB0);
          *reinterpret_cast<u32 *>(&lo_h[(size_t)(g + 8) * CT + cA]) = pack_lower_bounds(loA1, loB1);
        };
        // three candidate tiles in flight, buffers in fixed roles (no register rotation: a load
        // is only waited for three tiles after it was issued)
        const u32 C8 = (C + 7) & ~7u;
        constexpr u32 STEP = 8 * SCR_DEPTH;
        ScreenRow<D> r0, r1, r2;
        r0.load(sp16 + (size_t)cand_row[g] * D, t);
        if (8 < C8) r1.load(sp16 + (size_t)cand_row[8 + g] * D, t);
        if (16 < C8) r2.load(sp16 + (size_t)cand_row[16 + g] * D, t);
#if SCR_DEPTH == 4
        ScreenRow<D> r3;
        if (24 < C8) r3.load(sp16 + (size_t)cand_row[24 + g] * D, t);
#endif
        for (u32 j0 = 0; j0 < C8; j0 += STEP) {
          tile(r0, j0);
          if (j0 + STEP < C8) r0.load(sp16 + (size_t)cand_row[j0 + STEP + g] * D, t);
          if (j0 + 8 < C8) {
            tile(r1, j0 + 8);
            if (j0 + STEP + 8 < C8) r1.load(sp16 + (size_t)cand_row[j0 + STEP + 8 + g] * D, t);
          }
          if (j0 + 16 < C8) {
            tile(r2, j0 + 16);
            if (j0 + STEP + 16 < C8) r2.load(sp16 + (size_t)cand_row[j0 + STEP + 16 + g] * D, t);
          }
#if SCR_DEPTH == 4
          if (j0 + 24 < C8) {
            tile(r3, j0 + 24);
            if (j0 + STEP + 24 < C8) r3.load(sp16 + (size_t)cand_row[j0 + STEP + 24 + g] * D, t);
          }
#endif
        }
        if (C8 < C16) {                                                  // columns the scan reads beyond the last tile
          *reinterpret_cast<u32 *>(&lo_h[(size_t)g * CT + C8 + 2 * t]) = 0x7c007c00u;
          *reinterpret_cast<u32 *>(&lo_h[(size_t)(g + 8) * CT + C8 + 2 * t]) = 0x7c007c00u;
        }
        // the kept bounds were rounded to nearest fp16 (monotone): undo at most that much
        // (2^-11 relative in the normal range, 2^-25 absolute below it)
        const float th0 = quad_sixteenth_smallest(h0, t) * (1.0f + 1.0f / 1024.0f) + 1.0f / 16777216.0f;
        const float th1 = quad_sixteenth_smallest(h1, t) * (1.0f + 1.0f / 1024.0f) + 1.0f / 16777216.0f;
        if (t == 0) { theta[g] = th0; theta[g + 8] = th1; }
      }
      __syncwarp();

      // ---- survivors of all 16 queries at once: lane = (query, half of the stream) -----
      const int sq = lane & 15, sh = lane >> 4;
      u32 my_cnt = 0;
      {
        const u32 half_cols = C16 >> 1;                                 // multiple of 8
        // thresholds as fp16 rounded up; +inf (fewer than 16 candidates) becomes the largest
        // finite value so that the +inf of pads and of the query itself never passes
        // a candidate with D <= cutoff has lo <= D <= cut_s, so it passes; what is dropped beyond
        // Theta could not enter this try's k best, what is dropped beyond the cutoff is farther
        // than the point's final k-th neighbour and could not survive the merge (fminf ignores NaN)
        const float thf = fminf(theta[sq], cut_s);
        __half thh = __float2half_ru(thf);
        if (__hisinf(thh)) thh = __ushort_as_half((unsigned short)0x7bff);
        const __half2 th2 = __half2half2(thh);
        const uint4 *src = reinterpret_cast<const uint4 *>(lo_h + (size_t)sq * CT + sh * half_cols);
        unsigned short *mine = slist + lane * SCREEN_HALF;
        const u32 jbase = sh * half_cols;
        for (u32 i = 0; i < half_cols; i += 8) {
          const uint4 v = src[i >> 3];
          const u32 w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
          for (int e = 0; e < 4; e++) {
            const u32 m = __hle2_mask(*reinterpret_cast<const __half2 *>(&w[e]), th2);
            // appends wrap inside the lane's list: harmless, a list that wrapped abandons the bucket
            const bool pa = (m & 0xffffu) != 0, pb = (m >> 16) != 0;
            if (pa) mine[my_cnt & (SCREEN_HALF - 1)] = (unsigned short)(jbase + i + 2 * e);
            my_cnt += pa;
            if (pb) mine[my_cnt & (SCREEN_HALF - 1)] = (unsigned short)(jbase + i + 2 * e + 1);
            my_cnt += pb;
          }
        }
        if (__any_sync(FULL, my_cnt > SCREEN_HALF && (u32)sq < Qp)) overflow = true;
      }
      __syncwarp();                                                     // lo_h is dead from here on
      if (overflow) break;

      // ---- flat pair list of the tile (query-major, each query padded to 4) -------------
      // overlay on the parked bounds: query rows, distances, ids, pair tables
      float *qrows = reinterpret_cast<float *>(lo_h);                                    // [16][QROW]
      float *sdist = qrows + 16 * QROW;                                                  // [PAIRS]
      u32 *sid = reinterpret_cast<u32 *>(sdist + SCREEN_PAIRS);                          // [PAIRS]
      unsigned short *pj = reinterpret_cast<unsigned short *>(sid + SCREEN_PAIRS);       // [PAIRS] candidate index
      unsigned short *slot = pj + SCREEN_PAIRS;                                          // [PAIRS] rank -> pair
      unsigned char *lessn = reinterpret_cast<unsigned char *>(slot + SCREEN_PAIRS);     // [PAIRS] rank of a pair
      unsigned char *pq = lessn + SCREEN_PAIRS;                                          // [PAIRS] query, 0xff = padding
      const u32 other_cnt = __shfl_xor_sync(FULL, my_cnt, 16);
      const u32 Sq = ((u32)sq < Qp) ? my_cnt + other_cnt : 0;          // survivors of query sq (both halves agree)
      const u32 Sq4 = (Sq + 3) & ~3u;
      u32 incl = sh == 0 ? Sq4 : 0;
#pragma unroll
      for (int o = 1; o < 16; o <<= 1) {
        u32 up = __shfl_up_sync(FULL, incl, o);
        if (sq >= o) incl += up;
      }
      const u32 qoff = __shfl_sync(FULL, incl, sq) - Sq4;               // lanes >= 16 take their query's offset
      const u32 Ptot = __shfl_sync(FULL, incl, 15);
      if (Ptot > SCREEN_PAIRS) { overflow = true; break; }
      // query rows (and ids) on their way while the pair list is built
      {
#pragma unroll
        for (int it = 0; it < (16 * (D / 4) + 31) / 32; it++) {
          const int piece = lane + 32 * it;
          const int r = piece / (D / 4), col = piece - r * (D / 4);
          if (piece < 16 * (D / 4) && (u32)r < Qp)
            cp_async16(qrows + r * QROW + col * 4, sp + ((size_t)beg + qbase + r) * D + col * 4);
        }
        cp_async_commit();
      }
      if ((u32)sq < Qp) {
        const unsigned short *mine = slist + lane * SCREEN_HALF;
        const u32 dst = qoff + (sh ? other_cnt : 0);
        for (u32 i = 0; i < my_cnt; i++) { pj[dst + i] = mine[i]; pq[dst + i] = (unsigned char)sq; }
        if (sh == 0) {
          for (u32 i = Sq; i < Sq4; i++) { pq[qoff + i] = 0xff; pj[qoff + i] = 0; sdist[qoff + i] = inf; }
          qoffs[sq] = qoff;
          qcnt[sq] = Sq;
          qids[sq] = order[(size_t)beg + qbase + sq];
          qtie[sq] = 0;
        }
      }
      cp_async_wait<0>();
      __syncwarp();

      // ---- exact distances, 4 lanes per pair, 8 pairs per round ------------------------
      {
        const int gq = lane >> 2, l4 = lane & 3;
        const u32 rounds = (Ptot + 7) >> 3;
        auto fetch = [&](u32 r, ulonglong2 (&cb)[NV], u32 &cid) {
          const u32 p = min(8 * r + gq, Ptot - 1);
          const u32 row = cand_row[pj[p]];
          const ulonglong2 *src = reinterpret_cast<const ulonglong2 *>(sp + (size_t)row * D);
#pragma unroll
          for (int v = 0; v < NV; v++) cb[v] = src[4 * v + l4];
          cid = order[row];
        };
        auto finish = [&](u32 r, const ulonglong2 (&cb)[NV], u32 cid) {
          const u32 p = 8 * r + gq;
          const u32 qq = pq[min(p, Ptot - 1)];
          const float *qr = qrows + (qq & 15) * QROW;
          f32x2 x[NV][2];
#pragma unroll
          for (int v = 0; v < NV; v++) {
            const ulonglong2 qv = *reinterpret_cast<const ulonglong2 *>(qr + 16 * v + 4 * l4);
            x[v][0] = sqr2(sub2(qv.x, cb[v].x), negzero2);
            x[v][1] = sqr2(sub2(qv.y, cb[v].y), negzero2);
          }
#pragma unroll
          for (int h = NV / 2; h >= 1; h >>= 1)
#pragma unroll
            for (int v = 0; v < h; v++) {
              x[v][0] = add2(x[v][0], x[v + h][0]);
              x[v][1] = add2(x[v][1], x[v + h][1]);
            }
          // lane l4 now holds V16[4*l4 .. 4*l4+3]; V8 = V16[z] + V16[z+8], V4 = V8[z] + V8[z+4]
          f32x2 y0 = add2(x[0][0], __shfl_xor_sync(FULL, x[0][0], 2));
          f32x2 y1 = add2(x[0][1], __shfl_xor_sync(FULL, x[0][1], 2));
          y0 = add2(y0, __shfl_xor_sync(FULL, y0, 1));
          y1 = add2(y1, __shfl_xor_sync(FULL, y1, 1));
          f32x2 hh = add2(y0, y1);                                       // (V2[0], V2[1])
          float lo, hi;
          unpack2(hh, lo, hi);
          if (l4 == 0 && p < Ptot && qq != 0xff) { sdist[p] = lo + hi; sid[p] = cid; }
        };
        // three rounds in flight, buffers in fixed roles
        ulonglong2 b0[NV], b1[NV], b2[NV];
        u32 i0 = 0, i1 = 0, i2 = 0;
        if (0 < rounds) fetch(0, b0, i0);
        if (1 < rounds) fetch(1, b1, i1);
        if (2 < rounds) fetch(2, b2, i2);
#define SCR_STEP(o, bb, ii)                                 \
  if (r + o < rounds) {                                     \
    finish(r + o, bb, ii);                                  \
    if (r + o + 3 < rounds) fetch(r + o + 3, bb, ii);       \
  }
        for (u32 r = 0; r < rounds; r += 3) {
          SCR_STEP(0, b0, i0) SCR_STEP(1, b1, i1) SCR_STEP(2, b2, i2)
        }
#undef SCR_STEP
        if (lane == 0) atomicAdd(&leaf_exact_pairs_dev, (unsigned long long)Ptot);
      }
      __syncwarp();

      // ---- ranks by counting: position = number of strictly smaller survivors of the same
      // query; two survivors with the same count are an exact tie, which matters below k ----
      for (u32 e0 = 0; e0 < Ptot; e0 += 32) {
        const u32 p = e0 + lane;
        const u32 qq = p < Ptot ? pq[p] : 0xffu;
        if (qq != 0xff) {
          const u32 off = qoffs[qq], cnt4 = (qcnt[qq] + 3) & ~3u;
          const float dme = sdist[p];
          u32 less = 0;
          for (u32 j = 0; j < cnt4; j += 4) {
            const float4 dj = *reinterpret_cast<const float4 *>(&sdist[off + j]);
            count_less(less, dj.x, dme); count_less(less, dj.y, dme);
            count_less(less, dj.z, dme); count_less(less, dj.w, dme);
          }
          lessn[p] = (unsigned char)less;
          slot[off + less] = (unsigned short)p;
        }
      }
      __syncwarp();
      for (u32 e0 = 0; e0 < Ptot; e0 += 32) {
        const u32 p = e0 + lane;
        const u32 qq = p < Ptot ? pq[p] : 0xffu;
        if (qq != 0xff) {
          const u32 less = lessn[p];
          if (less < (u32)k) {
            if (slot[qoffs[qq] + less] != p) qtie[qq] = 1;
            const size_t o = (size_t)qids[qq] * k + less;
            list_ids[o] = sid[p];
            list_dist[o] = sdist[p];
          }
        }
      }
      __syncwarp();
      // fewer than k survivors in a row (the rule with a cutoff): (n, +inf) behind them.  Slot
      // (query e >> 4, position e & 15), so that a store instruction covers whole rows (k <= 16)
      for (u32 e = lane; e < 256; e += 32) {
        const u32 q = e >> 4, i = e & 15;
        if (q < Qp && i < (u32)k && i >= qcnt[q]) {
          const size_t o = (size_t)qids[q] * k + i;
          list_ids[o] = sentinel;
          list_dist[o] = inf;
        }
      }
      if ((u32)lane < Qp && qtie[lane]) tie_report(ties, qids[lane]);
      __syncwarp();
    }

    if (lane == 0) {
      if (overflow) {
        ovf.buckets[atomicAdd(ovf.count, 1u)] = (u32)b;
        atomicAdd(&leaf_overflow_dev, 1ull);
      } else {
        atomicAdd(&leaf_pairs_dev, (unsigned long long)Q * (C - 1));   // self excluded
      }
    }
    __syncwarp();
  }
}
