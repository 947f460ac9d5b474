// annb_finish.cu — S4 union of the per-try lists, S5 supercharging, and their literal
// (exact-tie / short-row) counterparts.
#include "annb_common.cuh"

static __device__ unsigned long long finish_literal_rows_dev[2];
void annb_finish_literal_counts(unsigned long long out[2], int reset) {
  unsigned long long z[2] = {0, 0};
  cudaMemcpyFromSymbol(out, finish_literal_rows_dev, sizeof z);
  if (reset) cudaMemcpyToSymbol(finish_literal_rows_dev, z, sizeof z);
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
                   size_t n, u32 sentinel, int k, u32 *__restrict__ out_ids,
                   FT *__restrict__ out_dist, TieList ties) {
  const int lane = threadIdx.x & 31;
  size_t x = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (x >= n) return;
  WarpList<R> best;
  best.clear(sentinel);
  FT tau = ft_inf();
  FT max_v = -ft_inf();
  u32 max_id = sentinel;
  bool any_inf = false, tie = false;

  // Each list ascends, so only its leading entries can beat the current k-th best: one ballot
  // finds how many are worth offering (ties with the k-th best included, they raise `tie`).
  for (int li = (prev_ids ? -1 : 0); li < a.n_lists; li++) {
    const u32 *ids = li < 0 ? prev_ids + x * (size_t)k : lists_ids + ((size_t)li * n + x) * k;
    const FT *dist = li < 0 ? prev_dist + x * (size_t)k : lists_dist + ((size_t)li * n + x) * k;
    const int admit = li < 0 ? k : a.admit[li];
    for (int base = 0; base < admit; base += 32) {
      int e = base + lane;
      FT mv = e < admit ? dist[e] : ft_inf();
      u32 mi = e < admit ? ids[e] : sentinel;
      unsigned fin = __ballot_sync(FULL, e < admit && mv != ft_inf());
      if (__ballot_sync(FULL, e < admit && mv == ft_inf())) any_inf = true;
      if (fin) {                                               // largest finite entry of this run
        int last = 31 - __clz(fin);
        FT lv = __shfl_sync(FULL, mv, last);
        u32 lid = __shfl_sync(FULL, mi, last);
        if (lv > max_v) { max_v = lv; max_id = lid; }
      }
      unsigned pass = __ballot_sync(FULL, e < admit && mv <= tau && mv != ft_inf());
      while (pass) {
        int j = __ffs(pass) - 1;
        pass &= pass - 1;
        FT vn = __shfl_sync(FULL, mv, j);
        u32 idn = __shfl_sync(FULL, mi, j);
        consider<R>(best, tau, vn, idn, k, sentinel, lane, tie);
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
  if (tie && ties.rows && lane == 0) tie_report(ties, (u32)x);
}

// Fast union for k <= 16: the running result sits ascending in lanes 0-15, the next list is
// loaded DEscending into lanes 16-31, so the warp holds a bitonic sequence and five xor-shuffle
// compare-exchange steps sort all 32 entries; equal ids (always adjacent: equal ids have equal
// distances) are dropped and the first k distinct entries are compacted back into lanes 0-15.
__global__ void __launch_bounds__(256)
merge_lists_fast_kernel(const u32 *__restrict__ lists_ids, const FT *__restrict__ lists_dist,
                        MergeArgs a, size_t n, u32 sentinel, int k, u32 *__restrict__ out_ids,
                        FT *__restrict__ out_dist, TieList ties) {
  const int lane = threadIdx.x & 31;
  size_t x = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (x >= n) return;
  const int e = lane & 15;
  FT v = ft_inf();
  u32 id = sentinel;
  FT max_v = -ft_inf();
  u32 max_id = sentinel;
  bool any_inf = false, tie = false;

  // list li: ascending into lanes 0-15 for li == 0, descending into lanes 16-31 afterwards;
  // the loads of list li+1 are issued before list li is merged (one list of latency hidden)
  auto fetch = [&](int li, FT &fv, u32 &fi) {
    fv = ft_inf();
    fi = sentinel;
    if (li >= a.n_lists) return;
    const bool mine = li == 0 ? lane < 16 : lane >= 16;
    const int src = li == 0 ? e : 15 - e;
    if (mine && src < a.admit[li]) {
      fv = lists_dist[((size_t)li * n + x) * k + src];
      fi = lists_ids[((size_t)li * n + x) * k + src];
    }
  };
  FT nxt_v, nxt_v2;
  u32 nxt_i, nxt_i2;
  fetch(0, nxt_v, nxt_i);
  fetch(1, nxt_v2, nxt_i2);
  for (int li = 0; li < a.n_lists; li++) {
    const int admit = a.admit[li];
    const bool mine = li == 0 ? lane < 16 : lane >= 16;
    const int src = li == 0 ? e : 15 - e;
    FT nv = nxt_v;
    u32 ni = nxt_i;
    nxt_v = nxt_v2; nxt_i = nxt_i2;
    fetch(li + 2, nxt_v2, nxt_i2);
    if (a.corner_list >= 0) {                      // bookkeeping of the prefix-corner rule only
      if (__ballot_sync(FULL, mine && src < admit && nv == ft_inf())) any_inf = true;
      // the list is sorted, so its largest finite admitted entry is the last finite one:
      // highest such lane for the ascending list 0, lowest for the descending ones
      unsigned fin = __ballot_sync(FULL, mine && src < admit && nv != ft_inf());
      if (fin) {
        int at = li == 0 ? 31 - __clz(fin) : __ffs(fin) - 1;
        FT cv = __shfl_sync(FULL, nv, at);
        u32 ci = __shfl_sync(FULL, ni, at);
        if (cv > max_v) { max_v = cv; max_id = ci; }
      }
    }
    if (mine) { v = nv; id = ni; }
    if (li == 0) {
      if (lane >= 16) { v = ft_inf(); id = sentinel; }
      if (a.n_lists > 1) continue;
    }
    // bitonic merge of the 32 lanes (ascending)
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
      FT pv = __shfl_xor_sync(FULL, v, j);
      u32 pi = __shfl_xor_sync(FULL, id, j);
      bool low = (lane & j) == 0;
      bool take = low ? (pv < v) : (pv > v);
      if (take) { v = pv; id = pi; }
    }
    // duplicates, ties, compaction of the first k distinct entries
    FT pv = __shfl_up_sync(FULL, v, 1);
    u32 pi = __shfl_up_sync(FULL, id, 1);
    bool finite = v != ft_inf();
    bool dup = lane > 0 && finite && id == pi;
    unsigned uniq = __ballot_sync(FULL, finite && !dup);
    int rank = __popc(uniq & ((1u << lane) - 1));                 // distinct entries before this lane
    // an equal-distance pair of different ids whose earlier member is kept (rank-1 < k)
    bool eqpair = lane > 0 && finite && !dup && v == pv && rank >= 1 && rank - 1 < k;
    if (__any_sync(FULL, eqpair)) tie = true;
    int want = lane;                                              // lane r takes the r-th distinct entry
    int from = (want < k && want < __popc(uniq)) ? (int)__fns(uniq, 0, want + 1) : -1;
    FT gv = __shfl_sync(FULL, v, from < 0 ? 0 : from);
    u32 gi = __shfl_sync(FULL, id, from < 0 ? 0 : from);
    v = from < 0 ? ft_inf() : gv;
    id = from < 0 ? sentinel : gi;
  }
  if (a.corner_list >= 0 && !any_inf) {
    u32 cid = lists_ids[((size_t)a.corner_list * n + x) * k + a.corner_pos];
    if (cid == max_id) {                                          // drop it, close the gap
      unsigned hit = __ballot_sync(FULL, id == cid && v != ft_inf());
      if (hit) {
        int pos = __ffs(hit) - 1;
        FT nv = __shfl_down_sync(FULL, v, 1);
        u32 ni = __shfl_down_sync(FULL, id, 1);
        if (lane >= pos) { v = lane == 31 ? ft_inf() : nv; id = lane == 31 ? sentinel : ni; }
      }
    }
  }
  if (lane < k) {
    out_ids[x * (size_t)k + lane] = id;
    out_dist[x * (size_t)k + lane] = v;
  }
  if (tie && ties.rows && lane == 0) tie_report(ties, (u32)x);
}

// One THREAD per point (any k, up to ML lists with at least one admitted entry).  The warp
// kernels above spend ~260 warp instructions per point on shuffles; a sequential ML-way merge
// of the sorted lists needs ~25 steps of ~40 instructions per THREAD, so the stage becomes a
// plain streaming pass over the lists.  A CTA stages the lists of its PT points in shared
// memory with coalesced 16-byte loads (row of a point: [list][k] distances, [list][k] ids, the
// output row, padded to 16 mod 32 words), each thread then walks its own row: the heads of the
// lists live in registers (static indexing), the smallest head is emitted unless it repeats the
// id emitted last (equal ids carry equal distances, so duplicates are adjacent in the merged
// order), and an equal distance with a DIFFERENT id among the kept entries (or across the
// keep/drop boundary) reports the row to the literal kernel, as in the warp kernels.
struct MergeThreadArgs {
  int n_src;                 // lists staged: those with admit > 0, then (if not among them) the corner list
  int src[17];               // original list index
  int admit[17];
  int corner_slot, corner_pos;   // staged slot of the corner list, or -1
};

template <int ML>
__global__ void __launch_bounds__(64)
merge_lists_thread_kernel(const u32 *__restrict__ lists_ids, const FT *__restrict__ lists_dist,
                          MergeThreadArgs a, size_t n, u32 sentinel, int k, int row_words,
                          u32 *__restrict__ out_ids, FT *__restrict__ out_dist, TieList ties) {
  constexpr int PT = 64;
  constexpr int FW = sizeof(FT) / 4;                              // words per distance
  extern __shared__ __align__(16) unsigned char smem_raw[];
  u32 *sm = reinterpret_cast<u32 *>(smem_raw);
  const int tid = threadIdx.x;
  const size_t x0 = (size_t)blockIdx.x * PT;
  const int pts = (int)min((size_t)PT, n - x0);
  const int NS = a.n_src;
  // Only the DISTANCES are staged (row of a point: [NS][k] distances, then k 16-bit codes of
  // the emitted entries): the order of the merge is decided by them alone.  An id is needed
  // (a) for the k entries that are kept — fetched after the walk, k independent loads — and
  // (b) when a head has exactly the distance emitted last: the same id again (a duplicate,
  // equal ids carry equal distances) or a different one (an exact tie -> literal kernel).
  // Half the shared memory per point doubles the resident warps of this latency-bound kernel.
  const int off_code = NS * k * FW;
  for (int s = 0; s < NS; s++) {
    const size_t g0 = ((size_t)a.src[s] * n + x0) * k;
    const int cells = pts * k;
    if ((k & 3) == 0 && FW == 1) {
      const uint4 *gd = reinterpret_cast<const uint4 *>(lists_dist + g0);
      for (int e = tid; e < cells / 4; e += PT) {
        const int p = (4 * e) / k, z = 4 * e - p * k;
        *reinterpret_cast<uint4 *>(sm + (size_t)p * row_words + s * k + z) = gd[e];
      }
    } else {
      for (int e = tid; e < cells; e += PT) {
        const int p = e / k, z = e - p * k;
        *reinterpret_cast<FT *>(sm + (size_t)p * row_words + (s * k + z) * FW) = lists_dist[g0 + e];
      }
    }
  }
  __syncthreads();
  if (tid >= pts) return;
  const size_t x = x0 + tid;
  u32 *row = sm + (size_t)tid * row_words;
  const FT *rd = reinterpret_cast<const FT *>(row);
  unsigned short *code = reinterpret_cast<unsigned short *>(row + off_code);
  auto id_of = [&](int slot, int pos) { return lists_ids[((size_t)a.src[slot] * n + x) * k + pos]; };
  const FT inf = ft_inf();
  FT hv[ML];
  int hp[ML];
#pragma unroll
  for (int s = 0; s < ML; s++) {
    const bool on = s < NS && a.admit[s] > 0;
    hv[s] = on ? rd[s * k] : inf;
    hp[s] = 0;
  }
  int out = 0, last_l = 0, last_p = 0;
  FT last_v = -inf;
  bool tie = false;
  for (;;) {
    FT bv = hv[0];
    int bl = 0;
#pragma unroll
    for (int s = 1; s < ML; s++)
      if (hv[s] < bv) { bv = hv[s]; bl = s; }
    if (bv == inf) break;
    int bp = hp[0];
#pragma unroll
    for (int s = 1; s < ML; s++)
      if (s == bl) bp = hp[s];
    bool dup = false;
    if (out > 0 && bv == last_v) {                                // same id again, or an exact tie
      dup = id_of(bl, bp) == id_of(last_l, last_p);
      if (!dup) tie = true;                                       // earlier one kept (or the k-th vs the first dropped)
    }
    if (!dup) {
      if (out == k) break;                                        // the first dropped entry has been looked at
      out_dist[x * (size_t)k + out] = bv;
      code[out] = (unsigned short)((bl << 8) | bp);
      out++;
      last_l = bl; last_p = bp; last_v = bv;
    }
    // advance list bl
    bp++;
    FT nv = inf;
    if (bp < a.admit[bl]) nv = rd[bl * k + bp];
#pragma unroll
    for (int s = 0; s < ML; s++)
      if (s == bl) { hv[s] = nv; hp[s] = bp; }
  }
  // prefix corner: the largest admitted entry dies if its id sits in the first slot
  // outside the prefix and no admitted entry is infinite; as the largest of everything it can
  // only be the last entry kept
  if (a.corner_slot >= 0 && out > 0) {
    bool any_inf = false;
    FT max_v = -inf;
    int max_l = 0, max_p = 0;
    for (int s = 0; s < NS; s++) {
      int e = a.admit[s] - 1;
      while (e >= 0 && rd[s * k + e] == inf) { any_inf = true; e--; }
      if (e >= 0 && rd[s * k + e] > max_v) { max_v = rd[s * k + e]; max_l = s; max_p = e; }
    }
    if (!any_inf) {
      const u32 cid = id_of(a.corner_slot, a.corner_pos);
      if (cid == id_of(max_l, max_p) && id_of(last_l, last_p) == cid) out--;
    }
  }
  for (int e = 0; e < out; e++) out_ids[x * (size_t)k + e] = id_of(code[e] >> 8, code[e] & 255);
  for (int e = out; e < k; e++) { out_dist[x * (size_t)k + e] = inf; out_ids[x * (size_t)k + e] = sentinel; }
  if (tie && ties.rows) tie_report(ties, (u32)x);
}

// Literal row: the n_lists lists of a point side by side (k*n_lists slots), the reference's
// network, first k slots out.  One CTA per reported row (ties.rows == NULL: every row — rows
// shorter than 16 slots).
__global__ void __launch_bounds__(256)
merge_literal_kernel(const u32 *__restrict__ lists_ids, const FT *__restrict__ lists_dist,
                     int n_lists, size_t n, int k, u32 *__restrict__ out_ids,
                     FT *__restrict__ out_dist, TieList ties, unsigned char *slabs,
                     size_t slab_bytes, int *status) {
  const int tid = threadIdx.x;
  const u32 total = ties.rows ? *ties.count : (u32)n;
  if (total == 0) return;
  const int len = n_lists * k;
  const size_t slab = ((size_t)len * (sizeof(u32) + sizeof(FT)) + 15) & ~(size_t)15;
  const size_t fit = slab_bytes / slab;
  if (fit == 0) { if (blockIdx.x == 0 && tid == 0) *status = 1; return; }
  const u32 workers = (u32)(fit < gridDim.x ? fit : gridDim.x);
  if (blockIdx.x >= workers) return;
  if (blockIdx.x == 0 && tid == 0) atomicAdd(&finish_literal_rows_dev[0], (unsigned long long)total);
  FT *key = reinterpret_cast<FT *>(slabs + (size_t)blockIdx.x * slab);
  u32 *ids = reinterpret_cast<u32 *>(key + len);
  for (u32 it = blockIdx.x; it < total; it += workers) {
    const size_t x = ties.rows ? ties.rows[it] : it;
    for (int e = tid; e < len; e += blockDim.x) {
      int t = e / k, z = e - t * k;
      ids[e] = lists_ids[((size_t)t * n + x) * k + z];
      key[e] = lists_dist[((size_t)t * n + x) * k + z];
    }
    __syncthreads();
    block_sort_and_uniq(ids, key, len);
    for (int i = tid; i < k; i += blockDim.x) {
      out_ids[x * (size_t)k + i] = ids[i];
      out_dist[x * (size_t)k + i] = key[i];
    }
    __syncthreads();
  }
}

// 1 (default): the thread-per-point merge where it applies; 0: the warp kernels (same rows)
static int merge_thread_mode = -1;
extern "C" void annb_merge_thread_mode(int on) { merge_thread_mode = on ? 1 : 0; }
static int merge_thread_enabled() {
  if (merge_thread_mode < 0) { const char *e = getenv("ANN_B200_THREAD_MERGE"); merge_thread_mode = (e && *e) ? (*e != '0') : 1; }
  return merge_thread_mode;
}

extern "C" void annb_merge_lists(const u32 *lists_ids, const FT *lists_dist, int n_lists,
                                 const int *host_admit, int corner_list, int corner_pos,
                                 const u32 *merged_in_ids, const FT *merged_in_dist, size_t n,
                                 size_t sentinel_n, size_t k, int every_list, u32 *merged_ids, FT *merged_dist, void *scratch,
                                 size_t scratch_bytes, int *status, annb_stream stream) {
  int regs = list_regs(k);
  if (!regs) fatal_config("k > 256");
  if (n_lists > 64) fatal_config("more than 64 lists per merge call");
  LiteralScratch ls = carve_literal_scratch(scratch, scratch_bytes, n);
  const bool whole_row = merged_in_ids == NULL && every_list;   // the literal redo needs every list of the row
  TieList all_rows = {ls.list.count, NULL};
  if ((size_t)n_lists * k < 16) {
    if (!whole_row) fatal_config("rows shorter than 16 slots cannot be merged incrementally");
    merge_literal_kernel<<<148 * 4, 256, 0, stream>>>(lists_ids, lists_dist, n_lists, n, (int)k, merged_ids, merged_dist, all_rows, ls.slabs, ls.slab_bytes, status);
    LAUNCH_CHECK("merge_literal");
    return;
  }
  MergeArgs a;
  a.n_lists = n_lists;
  for (int i = 0; i < n_lists; i++) a.admit[i] = host_admit[i];
  a.corner_list = corner_list;
  a.corner_pos = corner_pos;
  if (whole_row) RT_CHECK(cudaMemsetAsync(ls.list.count, 0, sizeof(u32), stream));
  TieList f = whole_row ? ls.list : all_rows;        // rows == NULL: ties are not recorded
  dim3 block(256), grid(grid_for(n * 32, 256));
  const char *nofast = getenv("ANN_B200_NO_FAST_MERGE");
  bool threaded = false;
  if (merged_in_ids == NULL && merge_thread_enabled()) {
    MergeThreadArgs ta;
    ta.n_src = 0;
    ta.corner_slot = -1;
    ta.corner_pos = corner_pos;
    bool fits = true;
    for (int i = 0; i < n_lists && fits; i++)
      if (host_admit[i] > 0 || i == corner_list) {
        if (ta.n_src == 17) { fits = false; break; }
        if (i == corner_list) ta.corner_slot = ta.n_src;
        ta.src[ta.n_src] = i;
        ta.admit[ta.n_src] = host_admit[i];
        ta.n_src++;
      }
    int live = 0;
    for (int i = 0; i < ta.n_src; i++) live += ta.admit[i] > 0;
    const int FW = (int)(sizeof(FT) / 4);
    int row_words = ta.n_src * (int)k * FW + ((int)k + 1) / 2;     // distances + 16-bit codes of the kept entries
    row_words = (row_words + 3) & ~3;
    row_words += (16 - (row_words & 31) + 32) & 31;                // = 16 mod 32: two points never share a bank phase
    const size_t tsmem = (size_t)row_words * 4 * 64;
    if (fits && live >= 1 && live <= 16 && ta.n_src <= 17 && tsmem <= 200 * 1024) {
      // the corner list (admit 0) is staged but never a merge source; sources must sit in slots < ML
      const bool big = live > 8 || ta.n_src > 8;
      static thread_local size_t configured[2] = {0, 0};
      if (tsmem > configured[big]) {
        if (big) RT_CHECK(cudaFuncSetAttribute(merge_lists_thread_kernel<17>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        else RT_CHECK(cudaFuncSetAttribute(merge_lists_thread_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        configured[big] = 200 * 1024;
      }
      unsigned tgrid = (unsigned)((n + 63) / 64);
      if (big) merge_lists_thread_kernel<17><<<tgrid, 64, tsmem, stream>>>(lists_ids, lists_dist, ta, n, (u32)sentinel_n, (int)k, row_words, merged_ids, merged_dist, f);
      else merge_lists_thread_kernel<8><<<tgrid, 64, tsmem, stream>>>(lists_ids, lists_dist, ta, n, (u32)sentinel_n, (int)k, row_words, merged_ids, merged_dist, f);
      threaded = true;
    }
  }
  if (threaded) {
  } else if (k <= 16 && merged_in_ids == NULL && !(nofast && *nofast && *nofast != '0')) {
    merge_lists_fast_kernel<<<grid, block, 0, stream>>>(lists_ids, lists_dist, a, n, (u32)sentinel_n, (int)k, merged_ids, merged_dist, f);
  } else
  switch (regs) {
    case 1: merge_lists_kernel<1><<<grid, block, 0, stream>>>(lists_ids, lists_dist, a, merged_in_ids, merged_in_dist, n, (u32)sentinel_n, (int)k, merged_ids, merged_dist, f); break;
    case 2: merge_lists_kernel<2><<<grid, block, 0, stream>>>(lists_ids, lists_dist, a, merged_in_ids, merged_in_dist, n, (u32)sentinel_n, (int)k, merged_ids, merged_dist, f); break;
    case 4: merge_lists_kernel<4><<<grid, block, 0, stream>>>(lists_ids, lists_dist, a, merged_in_ids, merged_in_dist, n, (u32)sentinel_n, (int)k, merged_ids, merged_dist, f); break;
    default: merge_lists_kernel<8><<<grid, block, 0, stream>>>(lists_ids, lists_dist, a, merged_in_ids, merged_in_dist, n, (u32)sentinel_n, (int)k, merged_ids, merged_dist, f); break;
  }
  LAUNCH_CHECK("merge_lists");
  if (whole_row) {
    merge_literal_kernel<<<148 * 2, 256, 0, stream>>>(lists_ids, lists_dist, n_lists, n, (int)k, merged_ids, merged_dist, ls.list, ls.slabs, ls.slab_bytes, status);
    LAUNCH_CHECK("merge_literal");
  }
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
                   size_t row_end, int exclude_self, u32 *__restrict__ out_ids,
                   FT *__restrict__ out_dist, TieList ties) {
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
  bool any_inf = false, tie = false;
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
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    FT ov = __shfl_xor_sync(FULL, max_v, o);
    u32 oi = __shfl_xor_sync(FULL, max_id, o);
    if (ov > max_v || (ov == max_v && oi < max_id)) { max_v = ov; max_id = oi; }   // uniform on ties
  }
  FT tau = best.kth(k);
  // equal neighbours inside the own list: the network decides their final order
#pragma unroll
  for (int rr = 0; rr < R; rr++) {
    FT nxt = __shfl_down_sync(FULL, best.v[rr], 1);
    FT wrap = __shfl_sync(FULL, best.v[rr + 1 < R ? rr + 1 : rr], 0);
    if (lane == 31) nxt = rr + 1 < R ? wrap : ft_inf();
    if (rr * 32 + lane + 1 < k && best.v[rr] == nxt && nxt != ft_inf()) tie = true;
  }
  tie = __any_sync(FULL, tie);
  u32 own_reg[R];
#pragma unroll
  for (int rr = 0; rr < R; rr++) own_reg[rr] = best.id[rr];

  const int cand = P2 - k;                     // slots k .. P2-1 of the row
  for (int base = 0; base < cand; base += 32) {
    int c = base + lane;
    int j = c < cand ? c / k : 0;
    int z = c - j * k;
    u32 oj = sentinel;
#pragma unroll
    for (int rr = 0; rr < R; rr++) {
      u32 got = __shfl_sync(FULL, own_reg[rr], j & 31);
      if (rr == (j >> 5)) oj = got;
    }
    u32 cid = (c < cand && oj < sentinel) ? graph[(size_t)oj * k + z] : sentinel;
    int cnt = min(32, cand - base);
    for (int i = 0; i < cnt; i++) {
      u32 idn = __shfl_sync(FULL, cid, i);
      if (idn >= sentinel || (exclude_self && idn == (u32)x)) { any_inf = true; continue; }
      FT dist;
      if (E) {
        WarpRow<(E ? E : 1)> cr;
        cr.load(points + (size_t)idn * d, lane, d);
        dist = __shfl_sync(FULL, warp_sqdist<(E ? E : 1)>(q, cr, d), 0);
      } else {
        dist = generic_sqdist(qrow, points + (size_t)idn * d, d, tmp, lane);
      }
      if (dist > max_v) { max_v = dist; max_id = idn; }
      consider<R>(best, tau, dist, idn, k, sentinel, lane, tie);
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
      out_ids[orow * (size_t)k + p] = best.id[rr];
      if (out_dist) out_dist[orow * (size_t)k + p] = best.v[rr];
    }
  }
  if (tie && lane == 0) tie_report(ties, (u32)orow);
}

// ---- fast path (k <= 32, d in {16,32,64,128}) -------------------------------------------
// One warp per row.  The candidate ids of the row (the neighbours' lists) are staged in
// shared memory with coalesced loads; candidates are then measured eight at a time, 8 lanes
// per candidate: lane g holds coordinates g, g+8, g+16, ... so that the first levels of the
// reference's summation tree are lane-local and the last three are xor-shuffles inside the
// 8-lane group.  After 8 tries hardly any candidate beats the current k-th best (3 % at
// cfg3), so the per-lane list is only touched when a ballot says some candidate passes;
// repeated ids are caught there (equal ids have equal distances).
template <int EPL>
__global__ void __launch_bounds__(256)
supercharge_fast_kernel(const FT *__restrict__ queries, const FT *__restrict__ points,
                        const u32 *__restrict__ own_ids, const FT *__restrict__ own_dist,
                        const u32 *__restrict__ graph, size_t n, int k, size_t row_begin,
                        size_t row_end, int exclude_self,
                        u32 *__restrict__ out_ids, FT *__restrict__ out_dist, TieList ties) {
  constexpr int D = EPL * 8;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  size_t x = row_begin + (((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (x >= row_end) return;
  const u32 sentinel = (u32)n;
  const int wide = k * (k + 1);
  const int P2 = 1 << floor_log2_u((unsigned long long)wide);
  const int cand = P2 - k;
  u32 *uniq = reinterpret_cast<u32 *>(smem_raw) + (size_t)wib * ((size_t)k * k);

  const int g = lane & 7, grp = lane >> 3;
  FT q[EPL];
  {
    const FT *qrow = queries + x * (size_t)D;
#pragma unroll
    for (int s = 0; s < EPL; s++) q[s] = qrow[g + 8 * s];
  }

  WarpList<1> best;
  best.v[0] = lane < k ? own_dist[x * (size_t)k + lane] : ft_inf();
  best.id[0] = lane < k ? own_ids[x * (size_t)k + lane] : sentinel;
  const u32 own_reg = best.id[0];
  bool tie = false;
  bool any_inf = lane < k && best.v[0] == ft_inf();
  FT max_v = (lane < k && best.v[0] != ft_inf()) ? best.v[0] : -ft_inf();
  u32 max_id = best.id[0];
  {
    FT nxt = __shfl_down_sync(FULL, best.v[0], 1);
    if (lane + 1 < k && best.v[0] == nxt && nxt != ft_inf()) tie = true;
  }
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    FT ov = __shfl_xor_sync(FULL, max_v, o);
    u32 oi = __shfl_xor_sync(FULL, max_id, o);
    if (ov > max_v || (ov == max_v && oi < max_id)) { max_v = ov; max_id = oi; }   // uniform on ties
  }
  FT tau = best.kth(k);
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
  __syncwarp();

  // Eight candidates per iteration (two groups of four in flight).  After 8 tries almost no
  // candidate beats the k-th best, so the lists are only touched when some lane's candidate
  // passes a ballot; the running (max, id) for the prefix-corner rule stays lane-local.
  FT lmax_v = -ft_inf();
  u32 lmax_id = sentinel;
  for (int base = 0; base < U; base += 8) {
    u32 cid[2];
    FT v[2];
    FT m[2][EPL];
#pragma unroll
    for (int h2 = 0; h2 < 2; h2++) {
      int mine = base + 4 * h2 + grp;
      cid[h2] = uniq[mine < U ? mine : base];
      const FT *crow = points + (size_t)cid[h2] * D;
#pragma unroll
      for (int s = 0; s < EPL; s++) m[h2][s] = crow[g + 8 * s];
    }
#pragma unroll
    for (int h2 = 0; h2 < 2; h2++) {
#pragma unroll
      for (int s = 0; s < EPL; s++) {
        FT df = q[s] - m[h2][s];
        m[h2][s] = df * df;
      }
#pragma unroll
      for (int h = EPL / 2; h >= 1; h >>= 1)
#pragma unroll
        for (int s = 0; s < h; s++) m[h2][s] = m[h2][s] + m[h2][s + h];
      FT t = m[h2][0];
      t = t + __shfl_xor_sync(FULL, t, 4);
      t = t + __shfl_xor_sync(FULL, t, 2);
      t = t + __shfl_xor_sync(FULL, t, 1);
      bool live = base + 4 * h2 + grp < U;
      v[h2] = live ? t : ft_inf();
      if (live && t > lmax_v) { lmax_v = t; lmax_id = cid[h2]; }
    }
    if (__any_sync(FULL, v[0] <= tau || v[1] <= tau)) {
#pragma unroll
      for (int h2 = 0; h2 < 2; h2++)
        for (int i = 0; i < 4; i++) {
          FT vn = __shfl_sync(FULL, v[h2], 8 * i);
          u32 idn = __shfl_sync(FULL, cid[h2], 8 * i);
          if (vn <= tau && vn != ft_inf() && !best.contains(idn)) {
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
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    FT ov = __shfl_xor_sync(FULL, lmax_v, o);
    u32 oi = __shfl_xor_sync(FULL, lmax_id, o);
    if (ov > lmax_v || (ov == lmax_v && oi < lmax_id)) { lmax_v = ov; lmax_id = oi; }   // uniform on ties
  }
  if (lmax_v > max_v) { max_v = lmax_v; max_id = lmax_id; }
  if (P2 < wide && !any_inf) {
    int c = P2 - k, j = c / k, z = c - j * k;
    u32 oj = __shfl_sync(FULL, own_reg, j);
    u32 cid = oj < sentinel ? graph[(size_t)oj * k + z] : sentinel;
    if (cid == max_id) best.remove(cid, sentinel, lane);
  }
  size_t orow = x - row_begin;
  if (lane < k) {
    out_ids[orow * (size_t)k + lane] = best.id[0];
    if (out_dist) out_dist[orow * (size_t)k + lane] = best.v[0];
  }
  if (tie && lane == 0) tie_report(ties, (u32)orow);
}

// Literal row of k(k+1) slots (supercharge + compdists + sort_and_uniq, alg.c:313-327), one
// CTA per reported row (ties.rows == NULL: every row of the range).
template <int E>
__global__ void __launch_bounds__(256)
supercharge_literal_kernel(const FT *__restrict__ queries, const FT *__restrict__ points,
                           const u32 *__restrict__ own_ids, const FT *__restrict__ own_dist,
                           const u32 *__restrict__ graph, size_t n, int d, int k, size_t row_begin,
                           size_t row_end, int exclude_self, u32 *__restrict__ out_ids,
                           FT *__restrict__ out_dist, TieList ties, unsigned char *slabs,
                           size_t slab_bytes, int *status) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x, wib = tid >> 5;
  FT *tmp = reinterpret_cast<FT *>(smem_raw) + (size_t)wib * (E == 0 ? d : 0);
  const size_t rows = row_end - row_begin;
  const u32 total = ties.rows ? *ties.count : (u32)rows;
  if (total == 0) return;
  const int wide = k * (k + 1);
  const size_t slab = ((size_t)wide * (2 * sizeof(u32) + sizeof(FT)) + 15) & ~(size_t)15;
  const size_t fit = slab_bytes / slab;
  if (fit == 0) { if (blockIdx.x == 0 && tid == 0) *status = 1; return; }
  const u32 workers = (u32)(fit < gridDim.x ? fit : gridDim.x);
  if (blockIdx.x >= workers) return;
  if (blockIdx.x == 0 && tid == 0) atomicAdd(&finish_literal_rows_dev[1], (unsigned long long)total);
  FT *key = reinterpret_cast<FT *>(slabs + (size_t)blockIdx.x * slab);
  u32 *ids = reinterpret_cast<u32 *>(key + wide);
  u32 *cslot = ids + wide;                                        // slots that need a distance
  __shared__ u32 s_live;
  const u32 sentinel = (u32)n;

  for (u32 it = blockIdx.x; it < total; it += workers) {
    const size_t orow = ties.rows ? ties.rows[it] : it, x = row_begin + orow;
    if (tid == 0) s_live = 0;
    for (int e = tid; e < k; e += blockDim.x) {
      ids[e] = own_ids[x * (size_t)k + e];
      key[e] = own_dist[x * (size_t)k + e];
    }
    __syncthreads();
    for (int e = tid; e < k * k; e += blockDim.x) {               // compute.cl:252-263
      int j = e / k, z = e - j * k;
      u32 oj = ids[j];
      u32 cid = oj < sentinel ? graph[(size_t)oj * k + z] : sentinel;
      ids[k + e] = cid;
      key[k + e] = ft_inf();
      if (cid < sentinel && !(exclude_self && cid == (u32)x)) cslot[atomicAdd(&s_live, 1u)] = (u32)(k + e);
    }
    __syncthreads();
    block_row_distances<E>(queries + x * (size_t)d, points, d, s_live, tmp,
                           [&](u32 i) { return (size_t)ids[cslot[i]]; },
                           [&](u32 i, FT dist) { key[cslot[i]] = dist; });
    __syncthreads();
    block_sort_and_uniq(ids, key, wide);
    for (int i = tid; i < k; i += blockDim.x) {
      out_ids[orow * (size_t)k + i] = ids[i];
      if (out_dist) out_dist[orow * (size_t)k + i] = key[i];
    }
    __syncthreads();
  }
}

template <int E>
static void launch_supercharge(int regs, size_t smem, annb_stream stream, const FT *queries,
                               const FT *points, const u32 *own_ids, const FT *own_dist,
                               const u32 *graph, size_t n, int d, int k, size_t rb, size_t re, int ex,
                               u32 *out_ids, FT *out_dist, const LiteralScratch &ls, int *status,
                               bool all_literal, bool skip_main) {
  dim3 block(256), grid(grid_for((re - rb) * 32, 256));
#define SC_CASE(R)                                                                               \
  {                                                                                              \
    if (smem > 48 * 1024)                                                                        \
      RT_CHECK(cudaFuncSetAttribute(supercharge_kernel<E, R>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    supercharge_kernel<E, R><<<grid, block, smem, stream>>>(queries, points, own_ids, own_dist, graph, n, d, k, rb, re, ex, out_ids, out_dist, ls.list); \
  }
  if (!all_literal && !skip_main) {
    switch (regs) {
      case 1: SC_CASE(1) break;
      case 2: SC_CASE(2) break;
      case 4: SC_CASE(4) break;
      default: SC_CASE(8) break;
    }
    LAUNCH_CHECK("supercharge");
  }
#undef SC_CASE
  if (smem > 48 * 1024)
    RT_CHECK(cudaFuncSetAttribute(supercharge_literal_kernel<E>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TieList which = ls.list;
  if (all_literal) which.rows = NULL;
  supercharge_literal_kernel<E><<<all_literal ? 148 * 4 : 148 * 2, 256, smem, stream>>>(
      queries, points, own_ids, own_dist, graph, n, d, k, rb, re, ex, out_ids, out_dist, which,
      ls.slabs, ls.slab_bytes, status);
  LAUNCH_CHECK("supercharge_literal");
}

#ifdef USE_FLOAT
#include "annb_screen_common.cuh"
#include "annb_supercharge_screen.cuh"
static int s5_screen_mode = -1;
extern "C" void annb_supercharge_screen_mode(int on) { s5_screen_mode = on ? 1 : 0; }
static int s5_screen_enabled() {
  if (s5_screen_mode < 0) { const char *e = getenv("ANN_B200_S5_SCREEN"); s5_screen_mode = (e && *e) ? (*e != '0') : 1; }
  return s5_screen_mode;
}
extern "C" int annb_supercharge_screen_applies(size_t d, size_t k) {
  return s5_screen_enabled() && (d == 16 || d == 32 || d == 64 || d == 128) && k <= 32 && k * (k + 1) >= 16;
}
#else
extern "C" int annb_supercharge_screen_applies(size_t d, size_t k) { (void)d; (void)k; return 0; }
extern "C" void annb_supercharge_screen_mode(int on) { (void)on; }
extern "C" void annb_supercharge_screen_stats(unsigned long long out[2], int reset) { (void)reset; out[0] = out[1] = 0; }
#endif

extern "C" void annb_supercharge(const FT *queries, const FT *points, const u32 *own_ids,
                                 const FT *own_dist, const u32 *graph, size_t n, size_t d, size_t k,
                                 size_t row_begin, size_t row_end, int exclude_self,
                                 u32 *out_ids, FT *out_dist, void *scratch, size_t scratch_bytes,
                                 int *status, const annb_supercharge_opts *opts, annb_stream stream) {
  if (row_end <= row_begin) return;
  int regs = list_regs(k);
  if (!regs) fatal_config("k > 256");
  int mode = row_mode(d);
  size_t smem = mode ? 0 : 8 * d * sizeof(FT);
  if (smem > 200 * 1024) fatal_config("d too large for the generic distance path");
  const size_t rows = row_end - row_begin;
  LiteralScratch ls = carve_literal_scratch(scratch, scratch_bytes, rows);
  const bool all_literal = k * (k + 1) < 16;      // the network degenerates: literal rows only
  bool all_literal_redo_only = false;             // fast kernel already ran: only the reported redo
  RT_CHECK(cudaMemsetAsync(ls.list.count, 0, sizeof(u32), stream));
  const u32 *row_perm = opts ? opts->row_perm : NULL;
#ifdef USE_FLOAT
  // screened path: precomp only (the queries are the points, so their fp16 rows exist)
  if (opts && opts->points16 && opts->nrm && opts->scale_bits && exclude_self && queries == points &&
      annb_supercharge_screen_applies(d, k)) {
    const size_t ssmem = (k * k + 16) * sizeof(u32) * 8;
    dim3 block(256), grid(grid_for(rows * 32, 256));
    const unsigned short *p16 = (const unsigned short *)opts->points16;
    const float2 *pn = (const float2 *)opts->nrm;
#define SCREEN_CASE(DD) supercharge_screen_kernel<DD><<<grid, block, ssmem, stream>>>(points, p16, pn, opts->scale_bits, own_ids, own_dist, graph, n, (int)k, row_begin, row_end, row_perm, opts->perm_base, out_ids, out_dist, ls.list)
    switch (d) {
      case 16: SCREEN_CASE(16); break;
      case 32: SCREEN_CASE(32); break;
      case 64: SCREEN_CASE(64); break;
      default: SCREEN_CASE(128); break;
    }
#undef SCREEN_CASE
    LAUNCH_CHECK("supercharge_screen");
    all_literal_redo_only = true;
  } else
#endif
  // fast path
  {
    const char *off = getenv("ANN_B200_NO_FAST_SUPERCHARGE");
    bool allow = !(off && *off && *off != '0');
    int epl = (d == 16 || d == 32 || d == 64 || d == 128) ? (int)(d / 8) : 0;
    if (row_perm) fatal_config("row_perm without the screened supercharge");
    if (allow && !all_literal && k <= 32 && epl) {
      size_t fsmem = k * k * sizeof(u32) * 8;
      dim3 block(256), grid(grid_for(rows * 32, 256));
#define FAST_CASE(EE)                                                                              \
  {                                                                                                \
    if (fsmem > 48 * 1024)                                                                         \
      RT_CHECK(cudaFuncSetAttribute(supercharge_fast_kernel<EE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fsmem)); \
    supercharge_fast_kernel<EE><<<grid, block, fsmem, stream>>>(queries, points, own_ids, own_dist, graph, n, (int)k, row_begin, row_end, exclude_self, out_ids, out_dist, ls.list); \
  }
      switch (epl) {
        case 2: FAST_CASE(2) break;
        case 4: FAST_CASE(4) break;
        case 8: FAST_CASE(8) break;
        default: FAST_CASE(16) break;
      }
#undef FAST_CASE
      LAUNCH_CHECK("supercharge_fast");
      all_literal_redo_only = true;
    }
  }
#define SC_ARGS regs, smem, stream, queries, points, own_ids, own_dist, graph, n, (int)d, (int)k, row_begin, row_end, exclude_self, out_ids, out_dist, ls, status, all_literal, all_literal_redo_only
  switch (mode) {
    case 0: launch_supercharge<0>(SC_ARGS); break;
    case 1: launch_supercharge<1>(SC_ARGS); break;
    case 2: launch_supercharge<2>(SC_ARGS); break;
    case 4: launch_supercharge<4>(SC_ARGS); break;
    default: launch_supercharge<8>(SC_ARGS); break;
  }
#undef SC_ARGS
}
