// exact_kernels.cuh — the general ("exact path") kernels.
//
// The streaming kernels in ivfadc_kernels.cuh reproduce the reference's top-k
// whenever its result is a pure function of the (distance, arrival) keys.  Three
// rare situations are not, and finalize_kernel / coarse_select_kernel flag them:
//   (1) a distance tie straddling the k-th place: updateTopK admits with strict
//       `<` and inserts before equal entries (index_utils.c:19-33), so which of
//       the tied rows survive depends on arrival order;
//   (2) the same for the w-th coarse centroid (freddy.c:272-283);
//   (3) the probed lists hold fewer than k rows, so the reference's
//       `while (foundInstances < k)` loop re-probes with a blacklist
//       (freddy.c:262-293, :377).
// Flagged queries are re-done here from scratch, one CTA per query, following the
// reference's control flow literally.  The only shortcut is provably neutral:
// with v = the k-th smallest distance among (current top-k ∪ this round's rows),
// rows with distance > v can never end up in, or change the arrangement of, the
// entries <= v (they sit behind them in the sorted array and are evicted first),
// so only S = {rows with d <= v} is replayed, in arrival (table) order, through a
// literal updateTopK.  v is found by a 4-pass radix select on the distance bits;
// of the rows with d == v only the k earliest can ever be admitted, which bounds
// S by 2k even for duplicate-heavy tables.
#pragma once
#include "ivfadc_kernels.cuh"

namespace fb {

constexpr int kExactThreads = 1024;
constexpr int kExactMaxK = 1024;
constexpr int kExactECap = 1024;  // capacity for rows with d == v (must be >= kExactMaxK)
constexpr int kExactSortN = 2048;
constexpr uint32_t kNoRow = 0xFFFFFFFFu;

struct ExactShared {
  float* tk_d;      // [k]   current top-k distances (ascending)
  uint32_t* tk_t;   // [k]   their table rows (kNoRow = empty slot)
  u64* lbuf;        // [kExactMaxK]  rows with d <  v : (t << 32 | dbits)
  u64* ebuf;        // [kExactECap]  rows with d == v
  u64* sbuf;        // [kExactSortN] merged, sorted by arrival
  unsigned* hist;   // [256]
  int* misc;        // [8]
  float* slut;      // [m*K] staging for the LUT of the list being scanned, or nullptr (then global reads)
};

// ADC distance of one row (codes pre-scaled by 4) against a LUT in global memory:
// the reference's left-to-right sum (index_utils.c:1126-1133).
__device__ __forceinline__ float adc_row_global(const CodeTableDev& tab, int blk, int lane_in_blk,
                                                const float* __restrict__ lut, int K) {
  float acc = 0.0f;
  const uint2* up = tab.units + ((size_t)blk * tab.U) * 32 + lane_in_blk;
  for (int u = 0; u < tab.U; u++) {
    uint2 v = up[u * 32];
    int p = 4 * u;
    if (p + 0 < tab.m) acc = xadd(acc, lut[(size_t)(p + 0) * K + ((v.x & 0xFFFFu) >> 2)]);
    if (p + 1 < tab.m) acc = xadd(acc, lut[(size_t)(p + 1) * K + (v.x >> 18)]);
    if (p + 2 < tab.m) acc = xadd(acc, lut[(size_t)(p + 2) * K + ((v.y & 0xFFFFu) >> 2)]);
    if (p + 3 < tab.m) acc = xadd(acc, lut[(size_t)(p + 3) * K + (v.y >> 18)]);
  }
  return acc;
}

// visit every row of one list, two rows in flight per thread (the loads of both are
// issued before either is consumed); f(row_in_list, adc_distance)
template <typename F>
__device__ __forceinline__ void for_rows2(const CodeTableDev& tab, int blk0, int len,
                                          const float* __restrict__ lut, int K, F f) {
  for (int r = threadIdx.x; r < len; r += 2 * kExactThreads) {
    const int r1 = r + kExactThreads;
    const bool v1 = r1 < len;
    const float a0 = adc_row_global(tab, blk0 + (r >> 5), r & 31, lut, K);
    const float a1 = v1 ? adc_row_global(tab, blk0 + (r1 >> 5), r1 & 31, lut, K) : 0.0f;
    f(r, a0);
    if (v1) f(r1, a1);
  }
}

// literal updateTopK (index_utils.c:19-33) on the shared top-k arrays
__device__ __forceinline__ void update_topk_literal(float* tk_d, uint32_t* tk_t, float dist, uint32_t t, int k) {
  int slot = k;
  while (slot > 0 && !(tk_d[slot - 1] < dist)) slot--;
  if (slot >= k) return;
  for (int j = k - 1; j > slot; j--) { tk_d[j] = tk_d[j - 1]; tk_t[j] = tk_t[j - 1]; }
  tk_d[slot] = dist;
  tk_t[slot] = t;
}

// block-wide 8-bit radix-select step helper: thread 0 locates the bucket
__device__ __forceinline__ void radix_pick(ExactShared& sh, uint32_t& prefix, int& remaining) {
  __syncthreads();
  if (threadIdx.x == 0) {
    int cum = 0, bin = 0;
    for (; bin < 255; bin++) {
      int c = (int)sh.hist[bin];
      if (cum + c >= remaining) break;
      cum += c;
    }
    sh.misc[0] = bin;
    sh.misc[1] = remaining - cum;
  }
  __syncthreads();
  prefix = (prefix << 8) | (uint32_t)sh.misc[0];
  remaining = sh.misc[1];
  __syncthreads();
}

// copy the LUT of pair j into shared memory (all threads) when a staging buffer exists
#define FB_EXACT_LUT(j)                                                                         \
  const float* lut = luts + (size_t)(j) * lut_stride;                                           \
  if (sh.slut != nullptr) {                                                                     \
    __syncthreads();                                                                            \
    for (int x_ = tid; x_ < lut_n4; x_ += kExactThreads)                                        \
      reinterpret_cast<float4*>(sh.slut)[x_] = reinterpret_cast<const float4*>(lut)[x_];        \
    __syncthreads();                                                                            \
    lut = sh.slut;                                                                              \
  }

// One pass of the reference's row loop (freddy.c:347-373 and its siblings) over
// n_pairs (list, LUT) pairs, continuing from the top-k state in sh.tk_*.
// If the k-th smallest distance v is already known (the streaming pass computed it:
// finalize_kernel's kth_key), pass have_v = true and skip the radix select.
__device__ void exact_round(const CodeTableDev& tab, const int* lists, int n_pairs,
                            const float* __restrict__ luts, size_t lut_stride, int K, int k,
                            ExactShared& sh, bool have_v = false, uint32_t v_known = 0) {
  const int tid = threadIdx.x;
  const int lut_n4 = tab.m * K / 4;
  long long n_rows = 0;
  for (int j = 0; j < n_pairs; j++) n_rows += tab.list_len[lists[j]];
  int carried = 0;
  for (int i = 0; i < k; i++) carried += (sh.tk_t[i] != kNoRow);
  const long long total = n_rows + carried;

  // ---- v = k-th smallest distance bits among carried entries and rows ----
  uint32_t vbits = 0xFFFFFFFFu;
  if (have_v) {
    vbits = v_known;
  } else if (total >= k) {
    uint32_t prefix = 0;
    int remaining = k;
    for (int pass = 0; pass < 4; pass++) {
      const int shift = 24 - 8 * pass;
      __syncthreads();
      for (int i = tid; i < 256; i += kExactThreads) sh.hist[i] = 0;
      __syncthreads();
      for (int i = tid; i < k; i += kExactThreads) {
        if (sh.tk_t[i] != kNoRow) {
          uint32_t db = __float_as_uint(sh.tk_d[i]);
          if (pass == 0 || (db >> (shift + 8)) == prefix) atomicAdd(&sh.hist[(db >> shift) & 255u], 1u);
        }
      }
      for (int j = 0; j < n_pairs; j++) {
        const int list = lists[j], blk0 = tab.list_blk[list], len = tab.list_len[list];
        FB_EXACT_LUT(j)
        for_rows2(tab, blk0, len, lut, K, [&](int, float a) {
          const uint32_t db = __float_as_uint(a);
          if (pass == 0 || (db >> (shift + 8)) == prefix) atomicAdd(&sh.hist[(db >> shift) & 255u], 1u);
        });
      }
      radix_pick(sh, prefix, remaining);
    }
    vbits = prefix;
  }

  // ---- collect S = rows with d < v  (< k of them)  and rows with d == v ----
  uint32_t t_cut = 0xFFFFFFFFu;
  for (int attempt = 0; attempt < 2; attempt++) {
    __syncthreads();
    if (tid == 0) { sh.misc[2] = 0; sh.misc[3] = 0; }
    __syncthreads();
    for (int j = 0; j < n_pairs; j++) {
      const int list = lists[j], blk0 = tab.list_blk[list], len = tab.list_len[list];
      FB_EXACT_LUT(j)
      for_rows2(tab, blk0, len, lut, K, [&](int r, float a) {
        const uint32_t db = __float_as_uint(a);
        if (db > vbits) return;
        const uint32_t t = (uint32_t)tab.rowno[(size_t)(blk0 + (r >> 5)) * 32 + (r & 31)];
        if (db < vbits) {
          int slot = atomicAdd(&sh.misc[2], 1);
          if (slot < kExactMaxK) sh.lbuf[slot] = ((u64)t << 32) | db;
        } else if (t <= t_cut) {
          int slot = atomicAdd(&sh.misc[3], 1);
          if (slot < kExactECap) sh.ebuf[slot] = ((u64)t << 32) | db;
        }
      });
    }
    __syncthreads();
    if (sh.misc[3] <= kExactECap) break;
    // more rows tie at v than fit: only the k earliest of them can be admitted;
    // radix-select the k-th smallest arrival among them and collect again.
    uint32_t prefix = 0;
    int remaining = k;
    for (int pass = 0; pass < 4; pass++) {
      const int shift = 24 - 8 * pass;
      __syncthreads();
      for (int i = tid; i < 256; i += kExactThreads) sh.hist[i] = 0;
      __syncthreads();
      for (int j = 0; j < n_pairs; j++) {
        const int list = lists[j], blk0 = tab.list_blk[list], len = tab.list_len[list];
        FB_EXACT_LUT(j)
        for_rows2(tab, blk0, len, lut, K, [&](int r, float a) {
          if (__float_as_uint(a) != vbits) return;
          const uint32_t t = (uint32_t)tab.rowno[(size_t)(blk0 + (r >> 5)) * 32 + (r & 31)];
          if (pass == 0 || (t >> (shift + 8)) == prefix) atomicAdd(&sh.hist[(t >> shift) & 255u], 1u);
        });
      }
      radix_pick(sh, prefix, remaining);
    }
    t_cut = prefix;
  }
  __syncthreads();
  const int n_less = min(sh.misc[2], kExactMaxK);
  const int n_eq = min(sh.misc[3], kExactECap);
  const int n_s = n_less + n_eq;
  int n_pad = 32;
  while (n_pad < n_s) n_pad <<= 1;
  for (int i = tid; i < n_pad; i += kExactThreads)
    sh.sbuf[i] = (i < n_less) ? sh.lbuf[i] : (i < n_s ? sh.ebuf[i - n_less] : kKeyInf);
  __syncthreads();
  // bitonic sort by (arrival, dbits)
  for (int size = 2; size <= n_pad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = tid; i < n_pad / 2; i += kExactThreads) {
        int lo = 2 * i - (i & (stride - 1));
        int hi = lo + stride;
        bool asc = (lo & size) == 0;
        u64 a = sh.sbuf[lo], b = sh.sbuf[hi];
        if ((a > b) == asc) { sh.sbuf[lo] = b; sh.sbuf[hi] = a; }
      }
      __syncthreads();
    }
  }
  // ---- replay S in arrival order through the literal gate + updateTopK ----
  if (tid == 0) {
    float max_dist = sh.tk_d[k - 1];
    for (int i = 0; i < n_s; i++) {
      u64 e = sh.sbuf[i];
      float dist = __uint_as_float((uint32_t)e);
      if (dist < max_dist) {                       // freddy.c:369-372
        update_topk_literal(sh.tk_d, sh.tk_t, dist, (uint32_t)(e >> 32), k);
        max_dist = sh.tk_d[k - 1];
      }
    }
  }
  __syncthreads();
}

__device__ __forceinline__ ExactShared exact_carve(unsigned char* p, int k, int lut_stage_floats) {
  ExactShared sh;
  sh.lbuf = reinterpret_cast<u64*>(p); p += sizeof(u64) * kExactMaxK;
  sh.ebuf = reinterpret_cast<u64*>(p); p += sizeof(u64) * kExactECap;
  sh.sbuf = reinterpret_cast<u64*>(p); p += sizeof(u64) * kExactSortN;
  sh.tk_d = reinterpret_cast<float*>(p); p += sizeof(float) * kExactMaxK;
  sh.tk_t = reinterpret_cast<uint32_t*>(p); p += sizeof(uint32_t) * kExactMaxK;
  sh.hist = reinterpret_cast<unsigned*>(p); p += sizeof(unsigned) * 256;
  sh.misc = reinterpret_cast<int*>(p); p += sizeof(int) * 8;
  sh.slut = lut_stage_floats > 0 ? reinterpret_cast<float*>(p) : nullptr;
  (void)k;
  return sh;
}
constexpr size_t kExactFixedSmem = sizeof(u64) * (kExactMaxK + kExactECap + kExactSortN) +
                                   sizeof(float) * kExactMaxK + sizeof(uint32_t) * kExactMaxK +
                                   sizeof(unsigned) * 256 + sizeof(int) * 8;  // multiple of 16; the LUT stage follows

__device__ __forceinline__ void exact_write_result(const ExactShared& sh, int k, const int32_t* ids,
                                                   int32_t* out_ids, float* out_dists) {
  for (int i = threadIdx.x; i < k; i += kExactThreads) {
    out_ids[i] = (sh.tk_t[i] == kNoRow) ? -1 : ids[sh.tk_t[i]];
    out_dists[i] = sh.tk_d[i];
  }
}

// ---------------------------------------------------------------------------
// ivfadc_search, general path: the reference's first-call body (freddy.c:247-378)
// for each flagged query, including the re-probe loop and its blacklist.
// Persistent CTAs pull flagged queries from exact_list.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kExactThreads)
ivfadc_exact_kernel(const float* __restrict__ queries, int d,
                    const float* __restrict__ coarse,    // [C][d]
                    const float* __restrict__ coarseT,   // [d][Cs]
                    int C, int Cs,
                    const float* __restrict__ cbT,       // [m][sub][K]
                    int K, int sub,
                    CodeTableDev tab, int w, int k,
                    const int32_t* __restrict__ exact_list, const int32_t* __restrict__ exact_count,
                    int32_t* __restrict__ work_counter,
                    float* __restrict__ lut_scratch,     // [gridDim.x][w][m*K]
                    // products of the streaming pass, reused when the only reason is a scan tie:
                    const uint32_t* __restrict__ qflags, const int32_t* __restrict__ probes,   // [nq][w]
                    const u64* __restrict__ kth_key,
                    int32_t* __restrict__ out_ids, float* __restrict__ out_dists,
                    int32_t* __restrict__ error_flag, int lut_stage_floats,
                    float sentinel,                      // 1000.0 (freddy.c:184) / 100.0 (freddy.c:823-827)
                    u64* __restrict__ rows_counter,      // statistics (rows whose ADC distance was computed), or nullptr
                    int batch_mode) {                    // ivfadc_batch_search's loop (freddy.c:838-981), see below
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ExactShared sh = exact_carve(smem_raw, k, lut_stage_floats);
  unsigned char* p = smem_raw + kExactFixedSmem + sizeof(float) * (size_t)lut_stage_floats;
  float* cdist = reinterpret_cast<float*>(p); p += sizeof(float) * Cs;
  float* qv = reinterpret_cast<float*>(p); p += sizeof(float) * ((d + 3) & ~3);
  float* sel_d = reinterpret_cast<float*>(p); p += sizeof(float) * ((w + 3) & ~3);
  int* sel = reinterpret_cast<int*>(p); p += sizeof(int) * ((w + 3) & ~3);
  unsigned char* black = p;
  __shared__ int s_item;
  const int tid = threadIdx.x;
  const int m = tab.m;
  const size_t lut_stride = (size_t)m * K;
  float* my_luts = lut_scratch + (size_t)blockIdx.x * w * lut_stride;
  const float MAX_DIST = sentinel;

  while (true) {
    __syncthreads();
    if (tid == 0) s_item = atomicAdd(work_counter, 1);
    __syncthreads();
    const int item = s_item;
    if (item >= *exact_count) break;
    const int q = exact_list[item];
    for (int i = tid; i < d; i += kExactThreads) qv[i] = queries[(size_t)q * d + i];
    for (int i = tid; i < C; i += kExactThreads) black[i] = 0;
    for (int i = tid; i < k; i += kExactThreads) { sh.tk_d[i] = MAX_DIST; sh.tk_t[i] = kNoRow; }  // freddy.c:258-260
    __syncthreads();
    // LUTs of the lists in sel[] (freddy.c:296-314) into this CTA's scratch
    auto build_luts = [&]() {
      for (int idx = tid; idx < w * m * K; idx += kExactThreads) {
        int j = idx / (m * K), rem = idx % (m * K);
        int pos = rem / K, code = rem % K;
        const float* cvec = coarse + (size_t)sel[j] * d + pos * sub;
        float acc = 0.0f;
        for (int i = 0; i < sub; i++) {
          float r = xsub(qv[pos * sub + i], cvec[i]);
          float t = xsub(r, cbT[((size_t)pos * sub + i) * K + code]);
          acc = xadd(acc, xmul(t, t));
        }
        my_luts[(size_t)j * lut_stride + rem] = acc;
      }
      __syncthreads();
    };
    const uint32_t why = qflags ? qflags[q] : kWhyForced;
    if ((why & ~(kFlagExact | kWhyScanTie)) == 0) {
      // Only a distance tie across the k-th place: the coarse selection and the k-th
      // distance of the streaming pass stand (one round, freddy.c:262 exits after it);
      // rebuild the LUTs and replay the tied neighbourhood literally.
      for (int j = tid; j < w; j += kExactThreads) sel[j] = probes[(size_t)q * w + j];
      __syncthreads();
      build_luts();
      exact_round(tab, sel, w, my_luts, lut_stride, K, k, sh, true, key_dbits(kth_key[q]));
      exact_write_result(sh, k, tab.ids, out_ids + (size_t)q * k, out_dists + (size_t)q * k);
      continue;
    }
    long long found = 0;
    int n_black = 0;
    bool failed = false;
    // ivfadc_search counts the rows it fetched (freddy.c:377).  ivfadc_batch_search counts ADMISSIONS
    // (`foundInstances++` inside `if (distance < maxDist)`, freddy.c:966-972) and probes on while a query has
    // fewer than k of them, i.e. exactly while its top-k still has an empty slot (every admission fills one
    // until the array is full); its per-round list is the plain arg-min below 1000 (freddy.c:854-866).
    while (batch_mode ? (sh.tk_t[k - 1] == kNoRow) : (found < k)) {   // freddy.c:262 / :977-981
      if (C - n_black < w) { failed = true; break; }     // reference would index cq[-1] / reuse a stale list
      for (int c = tid; c < C; c += kExactThreads) {     // freddy.c:272-278
        float acc = 0.0f;
        for (int i = 0; i < d; i++) {
          float t = xsub(qv[i], coarseT[(size_t)i * Cs + c]);
          acc = xadd(acc, xmul(t, t));
        }
        cdist[c] = acc;
      }
      __syncthreads();
      if (tid == 0) {                                    // freddy.c:266-283, literal
        float min_dist = 1000.0f;
        for (int j = 0; j < w; j++) { sel_d[j] = 100.0f; sel[j] = -1; }
        int bad = 0;
        for (int i = 0; i < C; i++) {
          if (black[i]) continue;
          float dist = cdist[i];
          if (batch_mode) {                               // freddy.c:854-866: first minimum below 1000
            if (dist < min_dist) { min_dist = dist; sel[0] = i; sel_d[0] = dist; }
            continue;
          }
          if (dist < min_dist) {
            if (!(dist < 100.0f)) { bad = 1; break; }     // reference would write sel[w]
            int slot = w;
            while (slot > 0 && !(sel_d[slot - 1] < dist)) slot--;
            for (int j = w - 1; j > slot; j--) { sel_d[j] = sel_d[j - 1]; sel[j] = sel[j - 1]; }
            sel_d[slot] = dist;
            sel[slot] = i;
            min_dist = sel_d[w - 1];
          }
        }
        if (batch_mode && sel[0] < 0) bad = 1;           // nothing below 1000 left: the reference re-reads a stale list
        if (!bad) for (int j = 0; j < w; j++) black[sel[j]] = 1;   // freddy.c:289-293
        sh.misc[4] = bad;
      }
      __syncthreads();
      if (sh.misc[4]) { failed = true; break; }
      n_black += w;
      build_luts();
      exact_round(tab, sel, w, my_luts, lut_stride, K, k, sh);
      for (int j = 0; j < w; j++) found += tab.list_len[sel[j]];   // freddy.c:377
      __syncthreads();
    }
    if (tid == 0 && rows_counter != nullptr) atomicAdd(rows_counter, (u64)found);
    if (failed) {
      if (tid == 0) atomicExch(error_flag, 1);
      for (int i = tid; i < k; i += kExactThreads) { sh.tk_d[i] = MAX_DIST; sh.tk_t[i] = kNoRow; }
      __syncthreads();
    }
    exact_write_result(sh, k, tab.ids, out_ids + (size_t)q * k, out_dists + (size_t)q * k);
  }
}

// ---------------------------------------------------------------------------
// flat PQ, general path: one LUT per query, every list of the table
// (pq_search freddy.c:74-134; pq_search_in[_batch] freddy.c:514-631, :1070-1143).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(kExactThreads)
pq_exact_kernel(CodeTableDev tab, const int32_t* __restrict__ all_lists,  // [n_lists] 0..n_lists-1
                const float* __restrict__ lut, int K, int k, float sentinel,
                const int32_t* __restrict__ exact_list, const int32_t* __restrict__ exact_count,
                int32_t* __restrict__ work_counter,
                const u64* __restrict__ kth_key,          // from the streaming pass, or nullptr
                int32_t* __restrict__ out_ids, float* __restrict__ out_dists, int lut_stage_floats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  ExactShared sh = exact_carve(smem_raw, k, lut_stage_floats);
  __shared__ int s_item;
  const int tid = threadIdx.x;
  const size_t lut_stride = (size_t)tab.m * K;
  while (true) {
    __syncthreads();
    if (tid == 0) s_item = atomicAdd(work_counter, 1);
    __syncthreads();
    const int item = s_item;
    if (item >= *exact_count) break;
    const int q = exact_list[item];
    for (int i = tid; i < k; i += kExactThreads) { sh.tk_d[i] = sentinel; sh.tk_t[i] = kNoRow; }
    __syncthreads();
    // every pair uses the same LUT: stride 0
    exact_round(tab, all_lists, tab.n_lists, lut + (size_t)q * lut_stride, 0, K, k, sh, kth_key != nullptr,
                kth_key ? key_dbits(kth_key[q]) : 0u);
    exact_write_result(sh, k, tab.ids, out_ids + (size_t)q * k, out_dists + (size_t)q * k);
  }
}

}  // namespace fb
