// knn_join_kernels.cuh — ivpq_search_in (kNN-join over an inverted multi-index + PQ),
// reference: ivpq_search_in.c:61-721, index_utils.c:252-443 (cell selection),
// :673-682 (getConfidenceHyp), :477-498 (postverify).
//
// Per round of the reference's retry loop (ivpq_search_in.c:299) two kernels run:
//   ivpq_select_kernel   per active query: the 2 x Kc sub-distances, the Kc^2 cell sums, cells
//                        in ascending sum order, cumulative cell frequency, and the first prefix
//                        whose getConfidenceHyp reaches `confidence` (the
//                        reference's priority-queue traversal visits exactly that prefix)
//   ivpq_scan_kernel     per active query: the target rows whose cell was selected, in table
//                        order (= the reference's candidate arrival order) -> PQ / exact
//                        distances -> the reference's selection:
//                          method 0/1  strict-admission top-k replayed literally over the
//                                      rows that can matter (fact B, tests/test_topk_semantics.py)
//                          method 2    the k*pvf smallest PQ distances (what the (200 + k*pvf)
//                                      buffer + qsort keeps), then exact distances on those and
//                                      the literal top-k in that order (postverify)
// Where the reference's qsort (not stable on current glibc) decides between equal PQ distances,
// this code orders ties by arrival; DESIGN.md §3 states that deviation.
#pragma once
#include "exact_kernels.cuh"

namespace fb {

constexpr int kJoinThreads = 256;
constexpr int kJoinSortN = 2048;     // keys sorted in shared memory at once
constexpr int kJoinMaxP = 1024;      // max k * pvf (and max k)

struct JoinParams {
  int d, m, K, Kc;
  int k, pvf, method;                // 0 PQ, 1 exact, 2 PQ + post verification
  int n_targets_sql;                 // |targets| as passed in SQL (getConfidenceHyp `size`)
  int min_target;                    // k * alpha of this round
  float confidence;
  int stat_total;                    // (int)statistics[cells]
  int skip_below;                    // target-list mode: k * alpha_original, else 0
  int last_iteration;                // every active query exhausted all cells
  int pair_sums;                     // alpha*k > long_codes_threshold: distances as sums of pair sums (index_utils.c:457-475)
};

// Pair-LUT variant (getPrecomputedDistancesDouble, index_utils.c:457-475 + computePQDistanceInt16): the table entry
// of a code pair is d(pos 2l, c0) + d(pos 2l+1, c1); the row distance is the sequential sum of the m/2 pair entries
// (a trailing odd position is never looked at).  The pair sums are formed on the fly: same additions, same order.
__device__ __forceinline__ float adc_row_pairs(const CodeTableDev& tab, int blk, int lane_in_blk,
                                               const float* __restrict__ lut, int K) {
  float acc = 0.0f;
  const uint2* up = tab.units + ((size_t)blk * tab.U) * 32 + lane_in_blk;
  const int pairs = tab.m / 2;
  for (int u = 0; u < tab.U; u++) {
    const uint2 v = up[u * 32];
    const int p = 4 * u;
    if (p / 2 < pairs)
      acc = xadd(acc, xadd(lut[(size_t)(p + 0) * K + ((v.x & 0xFFFFu) >> 2)], lut[(size_t)(p + 1) * K + (v.x >> 18)]));
    if (p / 2 + 1 < pairs)
      acc = xadd(acc, xadd(lut[(size_t)(p + 2) * K + ((v.y & 0xFFFFu) >> 2)], lut[(size_t)(p + 3) * K + (v.y >> 18)]));
  }
  return acc;
}

// index_utils.c:673-682, float/double mix as written there
__device__ __forceinline__ float confidence_hyp(int expect, int size, float p, int stat_size) {
  if (expect > size) return 0.0f;
  const float mu = (float)size * p;
  const float sig = (float)(sqrt((double)((float)size * p) * (1.0 - (double)p)) *
                            (double)(((float)stat_size - (float)size) / ((double)(float)stat_size - 1.0)));
  return (float)(1.0 - 0.5 * (1.0 + erf(((double)(float)expect - 0.5 - (double)mu) / ((double)sig * sqrt(2.0)))));
}

// ---------------------------------------------------------------------------------------
// cell selection: one CTA (1024 threads) per active query
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(1024)
ivpq_select_kernel(const float* __restrict__ queries, const int32_t* __restrict__ active, int d, int Kc,
                   const float* __restrict__ coarse_multi,    // [2][Kc][d/2]
                   const float* __restrict__ stats,           // [Kc*Kc + 1]
                   JoinParams prm,
                   uint32_t* __restrict__ bitmaps,            // [n_active][32]
                   int32_t* __restrict__ n_cells) {           // [n_active]
  __shared__ float sd[64];
  __shared__ u64 keys[1024];
  __shared__ float prob[1025];
  __shared__ int first_ok;
  const int x = blockIdx.x, q = active[x], tid = threadIdx.x;
  const int cells = Kc * Kc, half = d / 2;
  const float* qv = queries + (size_t)q * d;
  if (tid < 2 * Kc) {                                                             // index_utils.c:296-305
    const int h = tid / Kc, j = tid % Kc;
    const float* c = coarse_multi + ((size_t)h * Kc + j) * half;
    float acc = 0.0f;
    for (int i = 0; i < half; i++) {
      const float t = xsub(qv[h * half + i], c[i]);
      acc = xadd(acc, xmul(t, t));
    }
    sd[tid] = acc;
  }
  if (tid == 0) first_ok = cells;
  __syncthreads();
  u64 key = kKeyInf;
  if (tid < cells) key = make_key(xadd(xadd(0.0f, sd[tid % Kc]), sd[Kc + tid / Kc]), (uint32_t)tid);   // :306-313
  keys[tid] = key;
  __syncthreads();
  for (int size = 2; size <= 1024; size <<= 1) {                                  // ascending bitonic sort
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (tid < 512) {
        const int lo = 2 * tid - (tid & (stride - 1)), hi = lo + stride;
        const bool asc = (lo & size) == 0;
        const u64 a = keys[lo], b = keys[hi];
        if ((a > b) == asc) { keys[lo] = b; keys[hi] = a; }
      }
      __syncthreads();
    }
  }
  if (tid == 0) {                                                                 // prob += statistics[next.id], in visit order
    float p = 0.0f;
    prob[0] = 0.0f;
    for (int n = 0; n < cells; n++) { p = xadd(p, stats[key_t(keys[n])]); prob[n + 1] = p; }
  }
  __syncthreads();
  // the loop `while (conf(prob) < confidence && visited < cells)` stops at the first n with conf(prob_n) >= confidence
  for (int n = tid; n <= cells; n += 1024)
    if (n < cells && !(confidence_hyp(prm.min_target, prm.n_targets_sql, prob[n], prm.stat_total) < prm.confidence))
      atomicMin(&first_ok, n);
  __syncthreads();
  const int n_sel = first_ok;
  if (tid < 32) bitmaps[(size_t)x * 32 + tid] = 0u;
  __syncthreads();
  if (tid < n_sel) {
    const uint32_t cell = key_t(keys[tid]);
    atomicOr(&bitmaps[(size_t)x * 32 + (cell >> 5)], 1u << (cell & 31));
  }
  if (tid == 0) n_cells[x] = n_sel;
}

// ---------------------------------------------------------------------------------------
// block-level helpers over a key array in global memory (keys unique in their low word)
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ void block_bitonic_sort(u64* s, int n_pad) {
  for (int size = 2; size <= n_pad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < n_pad / 2; i += kJoinThreads) {
        const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
        const bool asc = (lo & size) == 0;
        const u64 a = s[lo], b = s[hi];
        if ((a > b) == asc) { s[lo] = b; s[hi] = a; }
      }
      __syncthreads();
    }
  }
}

// Leaves in sbuf[0 .. n_out) the `want` smallest keys of keys[0 .. n) (all of them if n < want),
// ascending by (distance bits, arrival); if also_ties, every further key whose distance equals the
// want-th smallest distance is kept too, up to `tie_cap` of them (earliest arrivals first).
// Returns n_out.  All threads of the CTA call it.
__device__ int block_select_smallest(const u64* __restrict__ keys, int n, int want, bool also_ties, int tie_cap,
                                     u64* sbuf, unsigned* hist, int* misc) {
  const int tid = threadIdx.x;
  if (n <= kJoinSortN) {
    int n_pad = 32;
    while (n_pad < n) n_pad <<= 1;
    for (int i = tid; i < n_pad; i += kJoinThreads) sbuf[i] = (i < n) ? keys[i] : kKeyInf;
    __syncthreads();
    block_bitonic_sort(sbuf, n_pad);
    int n_out = min(n, want);
    if (also_ties && n > want) {
      const uint32_t v = key_dbits(sbuf[want - 1]);
      while (n_out < n && n_out < want + tie_cap && key_dbits(sbuf[n_out]) == v) n_out++;
    }
    return n_out;
  }
  // radix select of the want-th smallest distance (4 x 8 bits)
  uint32_t prefix = 0;
  int remaining = want;
  for (int pass = 0; pass < 4; pass++) {
    const int shift = 24 - 8 * pass;
    __syncthreads();
    for (int i = tid; i < 256; i += kJoinThreads) hist[i] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kJoinThreads) {
      const uint32_t db = key_dbits(keys[i]);
      if (pass == 0 || (db >> (shift + 8)) == prefix) atomicAdd(&hist[(db >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      int cum = 0, bin = 0;
      for (; bin < 255; bin++) {
        const int c = (int)hist[bin];
        if (cum + c >= remaining) break;
        cum += c;
      }
      misc[0] = bin;
      misc[1] = remaining - cum;
    }
    __syncthreads();
    prefix = (prefix << 8) | (uint32_t)misc[0];
    remaining = misc[1];
  }
  const uint32_t v = prefix;
  // keys below v (fewer than `want`) first, then the ties at v by arrival
  __syncthreads();
  if (tid == 0) { misc[2] = 0; misc[3] = 0; }
  __syncthreads();
  for (int i = tid; i < n; i += kJoinThreads) {
    const u64 key = keys[i];
    if (key_dbits(key) < v) sbuf[atomicAdd(&misc[2], 1)] = key;
  }
  __syncthreads();
  const int n_less = misc[2];
  const int room = kJoinSortN - n_less;
  // ties: collect up to `room`; if more exist keep the earliest (smallest low word) by radix select on arrival
  uint32_t t_cut = 0xFFFFFFFFu;
  for (int attempt = 0; attempt < 2; attempt++) {
    __syncthreads();
    if (tid == 0) misc[3] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += kJoinThreads) {
      const u64 key = keys[i];
      if (key_dbits(key) == v && key_t(key) <= t_cut) {
        const int slot = atomicAdd(&misc[3], 1);
        if (slot < room) sbuf[n_less + slot] = key;
      }
    }
    __syncthreads();
    if (misc[3] <= room) break;
    uint32_t tp = 0;
    int rem = room;
    for (int pass = 0; pass < 4; pass++) {
      const int shift = 24 - 8 * pass;
      __syncthreads();
      for (int i = tid; i < 256; i += kJoinThreads) hist[i] = 0;
      __syncthreads();
      for (int i = tid; i < n; i += kJoinThreads) {
        const u64 key = keys[i];
        const uint32_t t = key_t(key);
        if (key_dbits(key) == v && (pass == 0 || (t >> (shift + 8)) == tp)) atomicAdd(&hist[(t >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (tid == 0) {
        int cum = 0, bin = 0;
        for (; bin < 255; bin++) {
          const int c = (int)hist[bin];
          if (cum + c >= rem) break;
          cum += c;
        }
        misc[0] = bin;
        misc[1] = rem - cum;
      }
      __syncthreads();
      tp = (tp << 8) | (uint32_t)misc[0];
      rem = misc[1];
    }
    t_cut = tp;
  }
  const int n_eq = min(misc[3], room);
  const int n_s = n_less + n_eq;
  int n_pad = 32;
  while (n_pad < n_s) n_pad <<= 1;
  __syncthreads();
  for (int i = n_s + tid; i < n_pad; i += kJoinThreads) sbuf[i] = kKeyInf;
  __syncthreads();
  block_bitonic_sort(sbuf, n_pad);
  int n_out = min(n_s, want);
  if (also_ties) n_out = min(n_s, want + tie_cap);
  return n_out;
}

// Reference top-k (strict `<` admission, insert before equal entries; index_utils.c:19-33) of a
// candidate stream given as keys (distance bits << 32 | arrival) in global memory, arrival = position in
// the stream's order.  Fills tk_d / tk_t (shared, k entries; tk_t = arrival of the winner or 0xFFFFFFFF).
//   no tie across the k-th place  -> the k smallest keys, equal-distance runs reversed (fact A)
//   otherwise                     -> literal replay, in arrival order, of the rows with d <= v and of the
//                                    k earliest rows tied at v (fact B)
// All threads of the CTA (kJoinThreads) call it.
__device__ void block_reference_topk(const u64* __restrict__ keys, int n, int k, float sentinel,
                                     u64* sbuf, unsigned* hist, int* misc, float* tk_d, uint32_t* tk_t) {
  const int tid = threadIdx.x;
  for (int i = tid; i < k; i += kJoinThreads) { tk_d[i] = sentinel; tk_t[i] = 0xFFFFFFFFu; }
  __syncthreads();
  if (n <= 0) return;
  const int n_s = block_select_smallest(keys, n, k, true, k, sbuf, hist, misc);
  __syncthreads();
  const bool boundary_tie = n_s > k && key_dbits(sbuf[k]) == key_dbits(sbuf[k - 1]);
  if (!boundary_tie) {
    const int n_out = min(n_s, k);
    for (int i = tid; i < n_out; i += kJoinThreads) {
      const uint32_t db = key_dbits(sbuf[i]);
      if (!(__uint_as_float(db) < sentinel)) continue;     // strict admission below the sentinel: the slot stays (-1, sentinel)
      int s0 = i, e0 = i;
      while (s0 > 0 && key_dbits(sbuf[s0 - 1]) == db) s0--;
      while (e0 + 1 < n_out && key_dbits(sbuf[e0 + 1]) == db) e0++;
      const int pos = s0 + (e0 - i);                       // later arrival first inside a run
      tk_d[pos] = __uint_as_float(db);
      tk_t[pos] = key_t(sbuf[i]);
    }
    __syncthreads();
    return;
  }
  int n_pad = 32;
  while (n_pad < n_s) n_pad <<= 1;
  for (int i = tid; i < n_pad; i += kJoinThreads) {
    const u64 e = sbuf[i];
    sbuf[i] = (i < n_s) ? (((u64)key_t(e) << 32) | key_dbits(e)) : kKeyInf;   // re-key by arrival
  }
  __syncthreads();
  block_bitonic_sort(sbuf, n_pad);
  if (tid == 0) {
    float max_dist = sentinel;
    for (int i = 0; i < n_s; i++) {
      const float dist = __uint_as_float((uint32_t)sbuf[i]);
      if (dist < max_dist) {
        int slot = k;
        while (slot > 0 && !(tk_d[slot - 1] < dist)) slot--;
        if (slot < k) {
          for (int j = k - 1; j > slot; j--) { tk_d[j] = tk_d[j - 1]; tk_t[j] = tk_t[j - 1]; }
          tk_d[slot] = dist;
          tk_t[slot] = (uint32_t)(sbuf[i] >> 32);
        }
        max_dist = tk_d[k - 1];
      }
    }
  }
  __syncthreads();
}

// top-k of a materialised key stream per query (the large-k form of ivfadc_search / pq_search*):
// keys[q] = key_base + q * stride, n[q] keys; arrival values are table rows -> ids[]
__global__ void __launch_bounds__(kJoinThreads)
topk_from_keys_kernel(const u64* __restrict__ key_base, size_t stride, const int32_t* __restrict__ n_keys, int k,
                      float sentinel, const int32_t* __restrict__ ids, const uint32_t* __restrict__ qflags,
                      int32_t* __restrict__ out_ids, float* __restrict__ out_dists) {
  __shared__ u64 sbuf[kJoinSortN];
  __shared__ float tk_d[kJoinMaxP];
  __shared__ uint32_t tk_t[kJoinMaxP];
  __shared__ unsigned s_hist[256];
  __shared__ int s_misc[8];
  const int q = blockIdx.x;
  if (qflags != nullptr && (qflags[q] & kFlagExact)) return;   // re-done by the general kernel
  block_reference_topk(key_base + (size_t)q * stride, n_keys[q], k, sentinel, sbuf, s_hist, s_misc, tk_d, tk_t);
  for (int i = threadIdx.x; i < k; i += kJoinThreads) {
    out_ids[(size_t)q * k + i] = (tk_t[i] == 0xFFFFFFFFu) ? -1 : ids[tk_t[i]];
    out_dists[(size_t)q * k + i] = tk_d[i];
  }
}

// ---------------------------------------------------------------------------------------
// candidate scan + selection: one CTA per active query
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kJoinThreads)
ivpq_scan_kernel(const float* __restrict__ queries, const int32_t* __restrict__ active, JoinParams prm,
                 // target rows of this call (table order): blocked codes, cell, word-vector row (-1 = none), id
                 CodeTableDev ttab, int n_trows, const int32_t* __restrict__ t_cell, const int32_t* __restrict__ t_vrow,
                 const int32_t* __restrict__ t_id,
                 const float* __restrict__ vT,                 // word vectors, dimension-major 32-row blocks
                 const float* __restrict__ luts,               // [nq][m][K] indexed by query (methods 0, 2)
                 const uint32_t* __restrict__ bitmaps, int32_t* __restrict__ target_counts,   // [nq] accumulated, as the reference
                 u64* __restrict__ key_scratch,                // [gridDim.x][n_trows]
                 int32_t* __restrict__ out_ids, float* __restrict__ out_dists,                // [nq][k]
                 int32_t* __restrict__ filled) {               // [n_active] 1 if the k-th slot is filled
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint32_t s_bits[32];
  __shared__ int s_warp_cnt[kJoinThreads / 32];
  __shared__ int s_base;
  __shared__ unsigned s_hist[256];
  __shared__ int s_misc[8];
  u64* sbuf = reinterpret_cast<u64*>(smem_raw);                                   // [kJoinSortN]
  float* tk_d = reinterpret_cast<float*>(sbuf + kJoinSortN);                      // [k]
  int32_t* tk_id = reinterpret_cast<int32_t*>(tk_d + kJoinMaxP);                  // [k]
  float* pv_d = reinterpret_cast<float*>(tk_id + kJoinMaxP);                      // [P] exact distances (method 2)
  float* slut = pv_d + kJoinMaxP;                                                 // [m*K] (methods 0, 2)
  const int x = blockIdx.x, q = active[x], tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = prm.d, m = prm.m, K = prm.K, k = prm.k;
  const float MAX_DIST = 1000.0f;
  const float* qv = queries + (size_t)q * d;
  u64* keys = key_scratch + (size_t)x * n_trows;

  if (tid < 32) s_bits[tid] = bitmaps[(size_t)x * 32 + tid];
  if (tid == 0) s_base = 0;
  if (prm.method != 1) {
    const float4* src = reinterpret_cast<const float4*>(luts + (size_t)q * m * K);
    for (int i = tid; i < m * K / 4; i += kJoinThreads) reinterpret_cast<float4*>(slut)[i] = src[i];
  }
  __syncthreads();

  // ---- candidates in arrival order: target rows whose cell is selected (and, for methods 1/2,
  //      that join to a word vector) ----
  for (int t0 = 0; t0 < n_trows; t0 += kJoinThreads) {
    const int t = t0 + tid;
    bool cand = false;
    if (t < n_trows) {
      const int cell = t_cell[t];
      cand = (s_bits[cell >> 5] >> (cell & 31)) & 1u;
      if (cand && prm.method != 0 && t_vrow[t] < 0) cand = false;                 // INNER JOIN vecs
    }
    const unsigned bal = __ballot_sync(0xffffffffu, cand);
    if (lane == 0) s_warp_cnt[warp] = __popc(bal);
    __syncthreads();
    int off = s_base;
    for (int wv = 0; wv < warp; wv++) off += s_warp_cnt[wv];
    if (cand) {
      const int slot = off + __popc(bal & ((1u << lane) - 1u));
      float dist;
      if (prm.method == 1) {
        const int vr = t_vrow[t];
        const float* vp = vT + ((size_t)(vr >> 5) * d) * 32 + (vr & 31);
        float acc = 0.0f;
        for (int i = 0; i < d; i++) {
          const float tt = xsub(qv[i], vp[(size_t)i * 32]);
          acc = xadd(acc, xmul(tt, tt));
        }
        dist = acc;
      } else {
        dist = prm.pair_sums ? adc_row_pairs(ttab, t >> 5, t & 31, slut, K) : adc_row_global(ttab, t >> 5, t & 31, slut, K);
      }
      keys[slot] = make_key(dist, (uint32_t)t);                                   // t = arrival order
    }
    __syncthreads();
    if (tid == 0) {
      int tot = 0;
      for (int wv = 0; wv < kJoinThreads / 32; wv++) tot += s_warp_cnt[wv];
      s_base += tot;
    }
    __syncthreads();
  }
  const int n_cand = s_base;
  __threadfence_block();
  __syncthreads();

  // ---- the reference's bookkeeping around the round ----
  bool skipped = false;
  if (tid == 0) {
    const int tc = target_counts[q] + n_cand;                                     // ivpq_search_in.c:461
    if (prm.skip_below > 0 && tc < prm.skip_below && !prm.last_iteration) { target_counts[q] = 0; s_misc[7] = 1; }   // :553-557
    else { target_counts[q] = tc; s_misc[7] = 0; }
  }
  for (int i = tid; i < k; i += kJoinThreads) { tk_d[i] = MAX_DIST; tk_id[i] = -1; }
  __syncthreads();
  skipped = s_misc[7] != 0;

  if (!skipped && n_cand > 0) {
    if (prm.method != 2) {
      // strict-admission top-k over the stream in arrival order: only rows with d <= v (v = k-th
      // smallest) matter, of the ties at v only the k earliest; replay them literally.
      const int n_s = block_select_smallest(keys, n_cand, k, true, k, sbuf, s_hist, s_misc);
      __syncthreads();
      // re-key by arrival and sort
      int n_pad = 32;
      while (n_pad < n_s) n_pad <<= 1;
      for (int i = tid; i < n_pad; i += kJoinThreads) {
        const u64 e = sbuf[i];
        sbuf[i] = (i < n_s) ? (((u64)key_t(e) << 32) | key_dbits(e)) : kKeyInf;
      }
      __syncthreads();
      block_bitonic_sort(sbuf, n_pad);
      if (tid == 0) {
        float max_dist = MAX_DIST;
        for (int i = 0; i < n_s; i++) {
          const float dist = __uint_as_float((uint32_t)sbuf[i]);
          if (dist < max_dist) {                                                  // :523-530 / :531-541
            int slot = k;
            while (slot > 0 && !(tk_d[slot - 1] < dist)) slot--;
            if (slot < k) {
              for (int j = k - 1; j > slot; j--) { tk_d[j] = tk_d[j - 1]; tk_id[j] = tk_id[j - 1]; }
              tk_d[slot] = dist;
              tk_id[slot] = t_id[(uint32_t)(sbuf[i] >> 32)];
            }
            max_dist = tk_d[k - 1];
          }
        }
      }
    } else {
      // PQ + post verification: the k*pvf smallest PQ distances survive the buffer (ties by arrival),
      // exact distances on them in that order, literal top-k (postverify, index_utils.c:477-498)
      const int P = k * prm.pvf;
      const int n_s = block_select_smallest(keys, n_cand, P, false, 0, sbuf, s_hist, s_misc);
      __syncthreads();
      for (int j = tid; j < n_s; j += kJoinThreads) {
        if (!(key_dist(sbuf[j]) < MAX_DIST)) { pv_d[j] = MAX_DIST; continue; }   // never entered the PV buffer (ivpq_search_in.c:505)
        const int t = (int)key_t(sbuf[j]);
        const int vr = t_vrow[t];
        const float* vp = vT + ((size_t)(vr >> 5) * d) * 32 + (vr & 31);
        float acc = 0.0f;
        for (int i = 0; i < d; i++) {
          const float tt = xsub(qv[i], vp[(size_t)i * 32]);
          acc = xadd(acc, xmul(tt, tt));
        }
        pv_d[j] = acc;
      }
      __syncthreads();
      if (tid == 0) {
        float max_dist = MAX_DIST;
        for (int j = 0; j < n_s; j++) {
          const float dist = pv_d[j];
          if (dist < max_dist) {
            int slot = k;
            while (slot > 0 && !(tk_d[slot - 1] < dist)) slot--;
            if (slot < k) {
              for (int jj = k - 1; jj > slot; jj--) { tk_d[jj] = tk_d[jj - 1]; tk_id[jj] = tk_id[jj - 1]; }
              tk_d[slot] = dist;
              tk_id[slot] = t_id[key_t(sbuf[j])];
            }
            max_dist = tk_d[k - 1];
          }
        }
      }
    }
  }
  __syncthreads();
  for (int i = tid; i < k; i += kJoinThreads) {
    out_ids[(size_t)q * k + i] = tk_id[i];
    out_dists[(size_t)q * k + i] = tk_d[i];
  }
  if (tid == 0) filled[x] = (tk_d[k - 1] != MAX_DIST) ? 1 : 0;                    // :643
}

}  // namespace fb
