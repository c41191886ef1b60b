// knn_join_kernels.cuh — ivpq_search_in (kNN-join over an inverted multi-index + PQ),
// reference: ivpq_search_in.c:61-721, index_utils.c:252-443 (cell selection),
// :673-682 (getConfidenceHyp), :477-498 (postverify).
//
// Per round of the reference's retry loop (ivpq_search_in.c:299) two kernels run:
//   ivpq_select_kernel   per active query: the 2 x Kc sub-distances, the Kc^2 cell sums, cells
//                        in ascending sum order, cumulative cell frequency, and the first prefix
//                        whose getConfidenceHyp reaches `confidence` (the
//                        reference's priority-queue traversal visits exactly that prefix)
//   ivpq_scan_cells_kernel  per active query: the target rows of its selected cells (the per-call target
//                        set is laid out cell-major; keys carry the rows' table order = the reference's
//                        candidate arrival order) -> PQ / exact distances -> the reference's selection:
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
                   uint16_t* __restrict__ sel_cells,          // [n_active][1024] selected cells, ascending cell sum
                   int32_t* __restrict__ n_cells,             // [n_active]
                   int32_t* __restrict__ round_state) {       // [0]: AND over the active queries of "all cells selected"
  __shared__ float sd[64];
  __shared__ u64 keys[1024];
  const int x = blockIdx.x, q = active[x], tid = threadIdx.x, lane = tid & 31;
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
  __syncthreads();
  u64 key = kKeyInf;
  if (tid < cells) key = make_key(xadd(xadd(0.0f, sd[tid % Kc]), sd[Kc + tid / Kc]), (uint32_t)tid);   // :306-313
  // every warp sorts its 32 keys in registers (ascending over the lanes); the 32 sorted runs go to shared memory
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      const u64 other = shfl_xor_u64(key, stride);
      const bool keep_min = ((lane & stride) == 0) == ((lane & size) == 0);
      key = keep_min ? (key < other ? key : other) : (key > other ? key : other);
    }
  }
  keys[tid] = key;
  __syncthreads();
  if (tid >= 32) return;
  // Warp 0 merges the runs lazily: cells are visited in ascending sum order only as far as the confidence test needs
  // (typically a handful of the 1024).  Lane l owns run l.  Per batch of 32 visited cells: prob[n] — the sequential
  // fp32 sum of the statistics of the cells visited before n, in visit order — then one confidence test per lane.
  // The loop `while (conf(prob) < confidence && visited < cells)` stops at the first n with conf(prob[n]) >= confidence.
  int head = 0;
  u64 cur = keys[32 * lane];
  int n_done = 0, n_sel = cells;
  float p = 0.0f;
  bool found = false;
  while (n_done < cells && !found) {
    const int batch = min(32, cells - n_done);
    u64 mine = kKeyInf;                                   // lane j: the j-th cell of this batch
    for (int j = 0; j < batch; j++) {
      u64 m = cur;
#pragma unroll
      for (int sft = 16; sft >= 1; sft >>= 1) { const u64 o = shfl_xor_u64(m, sft); m = o < m ? o : m; }
      if (cur == m) { head++; cur = head < 32 ? keys[32 * lane + head] : kKeyInf; }   // keys are unique (cell id in the low word)
      if (lane == j) mine = m;
    }
    const float sj = lane < batch ? stats[key_t(mine)] : 0.0f;
    float pj = p;                                         // prob[n_done + lane]
    for (int i = 0; i < batch; i++) {
      const float si = __shfl_sync(0xffffffffu, sj, i);
      if (lane > i) pj = xadd(pj, si);
    }
    const float p_end = xadd(__shfl_sync(0xffffffffu, pj, batch - 1), __shfl_sync(0xffffffffu, sj, batch - 1));
    const bool ok = lane < batch && !(confidence_hyp(prm.min_target, prm.n_targets_sql, pj, prm.stat_total) < prm.confidence);
    const unsigned bal = __ballot_sync(0xffffffffu, ok);
    const int upto = bal ? __ffs(bal) - 1 : batch;        // cells of this batch that are visited
    if (lane < upto) sel_cells[(size_t)x * 1024 + n_done + lane] = (uint16_t)key_t(mine);
    if (bal) { n_sel = n_done + upto; found = true; }
    n_done += batch;
    p = p_end;
  }
  if (lane == 0) {
    n_cells[x] = n_sel;
    if (n_sel < cells) atomicAnd(round_state, 0);                                 // index_utils.c:404-406 (lastIteration)
  }
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

// =======================================================================================
// Second draft of the join's data path: the per-call target set as a CELL-MAJOR compact table.
//   join_rows_kernel        per selected target row t (table order = arrival order): cell, id, word-vector row
//   join_cell_count/scan/scatter   stable counting sort of the rows by cell (arrival order kept inside a
//                           cell), gathering the code units into a compact blocked table in that order
//   ivpq_scan_cells_kernel  per active query: only the rows of ITS selected cells are touched
// The selection that follows is key based ((distance, arrival t)), so the order in which a query's
// candidates are generated does not matter.
// =======================================================================================
constexpr int kJoinChunk = 512;      // rows per counting-sort chunk (one warp walks a chunk in arrival order)
constexpr int kJoinCells = 1024;

__global__ void join_rows_kernel(const int32_t* __restrict__ sel_rows, const int32_t* __restrict__ n_sel,
                                 const int32_t* __restrict__ cells, const int32_t* __restrict__ ids,
                                 const int32_t* __restrict__ vec_sorted_ids, const int32_t* __restrict__ vec_sorted_rows, int n_vec,
                                 int need_vec, int32_t* __restrict__ t_cell, int32_t* __restrict__ t_id, int32_t* __restrict__ t_vrow) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= *n_sel) return;
  const int row = sel_rows[t];
  const int id = ids[row];
  t_cell[t] = cells[row];
  t_id[t] = id;
  int vr = -1;
  if (need_vec && id >= 0) {                                        // `INNER JOIN vecs ON fq.id = vecs.id`: first row carrying the id
    int lo = 0, hi = n_vec;
    while (lo < hi) { const int mid = (lo + hi) >> 1; if (vec_sorted_ids[mid] < id) lo = mid + 1; else hi = mid; }
    if (lo < n_vec && vec_sorted_ids[lo] == id) vr = vec_sorted_rows[lo];
  }
  t_vrow[t] = vr;
}

// counts[chunk][cell]: one warp per chunk, shared-memory histogram
__global__ void __launch_bounds__(32)
join_cell_count_kernel(const int32_t* __restrict__ t_cell, const int32_t* __restrict__ n_sel, int32_t* __restrict__ counts) {
  __shared__ int hist[kJoinCells];
  const int chunk = blockIdx.x, lane = threadIdx.x;
  const int n = *n_sel;
  for (int c = lane; c < kJoinCells; c += 32) hist[c] = 0;
  __syncwarp();
  const int t0 = chunk * kJoinChunk, t1 = min(n, t0 + kJoinChunk);
  for (int t = t0 + lane; t < t1; t += 32) atomicAdd(&hist[t_cell[t]], 1);
  __syncwarp();
  for (int c = lane; c < kJoinCells; c += 32) counts[(size_t)chunk * kJoinCells + c] = hist[c];
}

// cell_start[c] (exclusive scan over cells of the per-cell totals, cell_start[1024] = n) and, in place,
// counts[chunk][c] -> first slot of that chunk's rows inside cell c.  One CTA, thread = cell.
__global__ void __launch_bounds__(kJoinCells)
join_cell_scan_kernel(int32_t* __restrict__ counts, int n_chunks, int32_t* __restrict__ cell_start) {
  __shared__ int s_part[kJoinCells];
  const int c = threadIdx.x;
  int tot = 0;
  for (int ch = 0; ch < n_chunks; ch++) tot += counts[(size_t)ch * kJoinCells + c];
  s_part[c] = tot;
  __syncthreads();
  for (int off = 1; off < kJoinCells; off <<= 1) {
    const int v = (c >= off) ? s_part[c - off] : 0;
    __syncthreads();
    s_part[c] += v;
    __syncthreads();
  }
  int base = s_part[c] - tot;
  cell_start[c] = base;
  if (c == kJoinCells - 1) cell_start[kJoinCells] = s_part[c];
  for (int ch = 0; ch < n_chunks; ch++) {
    const int v = counts[(size_t)ch * kJoinCells + c];
    counts[(size_t)ch * kJoinCells + c] = base;
    base += v;
  }
}

// stable scatter: row t goes to slot pos = (first slot of its chunk in its cell) + (rows of the same cell earlier in
// the chunk); perm[pos] = t and the row's code units are gathered from the pinned table into the compact one
__global__ void __launch_bounds__(32)
join_cell_scatter_kernel(const int32_t* __restrict__ t_cell, const int32_t* __restrict__ n_sel, const int32_t* __restrict__ chunk_base,
                         const int32_t* __restrict__ sel_rows, const uint2* __restrict__ src_units, int U,
                         int32_t* __restrict__ perm, uint2* __restrict__ dst_units) {
  __shared__ int cursor[kJoinCells];
  const int chunk = blockIdx.x, lane = threadIdx.x;
  const int n = *n_sel;
  for (int c = lane; c < kJoinCells; c += 32) cursor[c] = chunk_base[(size_t)chunk * kJoinCells + c];
  __syncwarp();
  const int t0 = chunk * kJoinChunk, t1 = min(n, t0 + kJoinChunk);
  for (int tb = t0; tb < t1; tb += 32) {
    const int t = tb + lane;
    const bool live = t < t1;
    const int cell = live ? t_cell[t] : -1 - lane;                  // dead lanes match nobody
    const unsigned peers = __match_any_sync(0xffffffffu, cell);
    const int rank = __popc(peers & ((1u << lane) - 1u));
    int pos = -1;
    if (live) pos = cursor[cell] + rank;
    __syncwarp();
    if (live && rank == 0) cursor[cell] += __popc(peers);
    __syncwarp();
    if (live) {
      perm[pos] = t;
      const int r = sel_rows[t];
      const int sb = r >> 5, sl = r & 31, db = pos >> 5, dl = pos & 31;
      for (int u = 0; u < U; u++) dst_units[((size_t)db * U + u) * 32 + dl] = src_units[((size_t)sb * U + u) * 32 + sl];
    }
  }
}

// next round's active list: the queries of this round whose k-th slot stayed empty (ivpq_search_in.c:639-666)
__global__ void join_next_active_kernel(const int32_t* __restrict__ active, const int32_t* __restrict__ filled, int n_active,
                                        int32_t* __restrict__ next_active, int32_t* __restrict__ round_state) {   // [1]: count
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  if (x >= n_active) return;
  if (!filled[x]) next_active[atomicAdd(round_state + 1, 1)] = active[x];
}

// ---------------------------------------------------------------------------------------
// candidate scan + selection over the cell-major target table: persistent CTAs, one active query at a time
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kJoinThreads)
ivpq_scan_cells_kernel(const float* __restrict__ queries, const int32_t* __restrict__ active, int n_active, JoinParams prm,
                       CodeTableDev ctab,                     // compact cell-major code table (slot = position in perm)
                       const int32_t* __restrict__ cell_start, const int32_t* __restrict__ perm,
                       const int32_t* __restrict__ t_vrow, const int32_t* __restrict__ t_id,
                       const float* __restrict__ vR, const float* __restrict__ luts,
                       const uint16_t* __restrict__ sel_cells, const int32_t* __restrict__ n_cells,
                       const int32_t* __restrict__ round_state,   // [0] = last iteration (set by the select kernel of this round)
                       int32_t* __restrict__ target_counts,
                       u64* __restrict__ key_scratch, size_t scratch_stride,   // [gridDim.x][scratch_stride] for queries beyond the shared buffer
                       int32_t* __restrict__ out_ids, float* __restrict__ out_dists, int32_t* __restrict__ filled,
                       int32_t* __restrict__ work_counter, u64* __restrict__ pair_counter) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ unsigned s_hist[256];
  __shared__ int s_misc[8];
  __shared__ int s_x, s_n;
  u64* sbuf = reinterpret_cast<u64*>(smem_raw);                                   // [kJoinSortN]
  u64* kbuf = sbuf + kJoinSortN;                                                  // [kJoinSortN] candidate keys of small queries
  float* tk_d = reinterpret_cast<float*>(kbuf + kJoinSortN);                      // [k]
  int32_t* tk_id = reinterpret_cast<int32_t*>(tk_d + kJoinMaxP);                  // [k]
  float* pv_d = reinterpret_cast<float*>(tk_id + kJoinMaxP);                      // [P] exact distances (method 2)
  float* slut = pv_d + kJoinMaxP;                                                 // [m*K] (methods 0, 2)
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int d = prm.d, m = prm.m, K = prm.K, k = prm.k;
  const float MAX_DIST = 1000.0f;
  const int last_iteration = round_state[0];

  for (;;) {
    __syncthreads();
    if (tid == 0) { s_x = atomicAdd(work_counter, 1); s_n = 0; }
    __syncthreads();
    const int x = s_x;
    if (x >= n_active) break;
    const int q = active[x];
    const float* qv = queries + (size_t)q * d;
    const int nc = n_cells[x];
    const uint16_t* cl = sel_cells + (size_t)x * 1024;
    if (prm.method != 1) {
      const float4* src = reinterpret_cast<const float4*>(luts + (size_t)q * m * K);
      for (int i = tid; i < m * K / 4; i += kJoinThreads) reinterpret_cast<float4*>(slut)[i] = src[i];
    }
    // rows this query will look at (an upper bound of its candidates): decides where the keys live
    int cap_rows = 0;
    for (int ci = tid; ci < nc; ci += kJoinThreads) { const int c = cl[ci]; cap_rows += cell_start[c + 1] - cell_start[c]; }
    for (int o = 16; o >= 1; o >>= 1) cap_rows += __shfl_xor_sync(0xffffffffu, cap_rows, o);
    if (lane == 0 && cap_rows) atomicAdd(&s_n, cap_rows);
    __syncthreads();
    const int rows_upper = s_n;
    __syncthreads();
    if (tid == 0) s_n = 0;
    __syncthreads();
    u64* keys = (rows_upper <= kJoinSortN) ? kbuf : key_scratch + (size_t)blockIdx.x * scratch_stride;

    // ---- candidates: the rows of the selected cells (and, for methods 1/2, that join to a word vector) ----
    for (int ci = warp; ci < nc; ci += kJoinThreads / 32) {
      const int c = cl[ci];
      const int s0 = cell_start[c], s1 = cell_start[c + 1];
      for (int pb = s0; pb < s1; pb += 32) {
        const int pos = pb + lane;
        bool cand = pos < s1;
        int t = 0;
        if (cand) {
          t = perm[pos];
          if (prm.method != 0 && t_vrow[t] < 0) cand = false;                     // INNER JOIN vecs
        }
        float dist = 0.0f;
        if (cand) {
          if (prm.method == 1) {
            const int vr = t_vrow[t];
            const float* vp = vR + (size_t)vr * d;               // row-major image: one row = d contiguous floats
            float acc = 0.0f;
            for (int i = 0; i < d; i++) {
              const float tt = xsub(qv[i], __ldg(vp + i));
              acc = xadd(acc, xmul(tt, tt));
            }
            dist = acc;
          } else {
            dist = prm.pair_sums ? adc_row_pairs(ctab, pos >> 5, pos & 31, slut, K) : adc_row_global(ctab, pos >> 5, pos & 31, slut, K);
          }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, cand);
        int base = 0;
        if (lane == 0 && bal) base = atomicAdd(&s_n, __popc(bal));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (cand) keys[base + __popc(bal & ((1u << lane) - 1u))] = make_key(dist, (uint32_t)t);   // t = arrival order
      }
    }
    __threadfence_block();
    __syncthreads();
    const int n_cand = s_n;

    // ---- the reference's bookkeeping around the round ----
    if (tid == 0) {
      const int tc = target_counts[q] + n_cand;                                   // ivpq_search_in.c:461
      if (prm.skip_below > 0 && tc < prm.skip_below && !last_iteration) { target_counts[q] = 0; s_misc[7] = 1; }   // :553-557
      else { target_counts[q] = tc; s_misc[7] = 0; }
      if (pair_counter) atomicAdd(pair_counter, (u64)n_cand);
    }
    for (int i = tid; i < k; i += kJoinThreads) { tk_d[i] = MAX_DIST; tk_id[i] = -1; }
    __syncthreads();
    const bool skipped = s_misc[7] != 0;

    if (!skipped && n_cand > 0) {
      if (prm.method != 2) {
        // strict-admission top-k over the stream in arrival order: only rows with d <= v (v = k-th
        // smallest) matter, of the ties at v only the k earliest; replay them literally.
        const int n_s = block_select_smallest(keys, n_cand, k, true, k, sbuf, s_hist, s_misc);
        __syncthreads();
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
            if (dist < max_dist) {                                                // :523-530 / :531-541
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
          if (!(key_dist(sbuf[j]) < MAX_DIST)) { pv_d[j] = MAX_DIST; continue; }   // never entered the PV buffer (:505)
          const int vr = t_vrow[key_t(sbuf[j])];
          const float* vp = vR + (size_t)vr * d;
          float acc = 0.0f;
          for (int i = 0; i < d; i++) {
            const float tt = xsub(qv[i], __ldg(vp + i));
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
    if (tid == 0) filled[x] = (tk_d[k - 1] != MAX_DIST) ? 1 : 0;                   // :643
  }
}

}  // namespace fb
