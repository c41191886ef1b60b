// build_kernels.cuh — pinning a code table (fb_load_fine / fb_load_pq / fb_load_ivpq) on the device.
//
// What the reference does per query through SPI — `SELECT id, vector FROM fine_quantization WHERE coarse_id IN
// (...)` (freddy.c:455-517) — is done here once, when the table is pinned: rows are grouped by list (CSR), each
// list is cut into 32-row blocks, and inside a list rows are placed so that the shared-memory LUT gather of a
// block sees few bank conflicts (the rule is stated at place_rows_of_list in engine.cu; this file holds the same
// rule as a kernel, one CTA per list, one thread per candidate of the window).  The host sends the raw
// (id, coarse_id, codes) columns once; counting, ordering, placement and packing run here.
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"

namespace fb {

// diag[0] = first row whose list id is out of range, diag[1] = first row with a code out of range (INT_MAX: none),
// diag[2] = max(0, largest id).  len[c] = rows of list c (only counted when rows carry a list id).
__global__ void table_count_kernel(const int32_t* __restrict__ list_of_row, int64_t N, int n_lists, const int16_t* __restrict__ codes,
                                   int m, int K, const int32_t* __restrict__ ids, int32_t* __restrict__ len, int32_t* __restrict__ diag) {
  int32_t my_max = 0;
  for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < N; r += (int64_t)gridDim.x * blockDim.x) {
    if (list_of_row != nullptr) {
      const int c = list_of_row[r];
      if (c < 0 || c >= n_lists) atomicMin(diag + 0, (int32_t)r);
      else atomicAdd(len + c, 1);
    }
    const int16_t* cr = codes + (size_t)r * m;
    bool bad = false;
    for (int p = 0; p < m; p++) { const int code = cr[p]; bad |= (code < 0) | (code >= K); }
    if (bad) atomicMin(diag + 1, (int32_t)r);
    my_max = max(my_max, ids[r]);
  }
  for (int s = 16; s >= 1; s >>= 1) my_max = max(my_max, __shfl_xor_sync(0xffffffffu, my_max, s));
  if ((threadIdx.x & 31) == 0 && my_max > 0) atomicMax(diag + 2, my_max);
}

__global__ void iota_i32_kernel(int32_t* __restrict__ out, int64_t n) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = (int32_t)i;
}

// Conflict-aware placement of n rows into 32-row blocks by one CTA (thread j owns one candidate of the window).
// code_of(a, p) = code of the row with arrival index a at position p; emit(slot, a) records the row placed in `slot`.
// The window is the `window` earliest-arrived unplaced rows; per slot the candidate with the lowest
// (cost, arrival index) wins, cost = 4096 x (positions whose per-bank maximum of distinct codes it raises) + (codes
// already in its banks) — the same choice the host loop makes (first candidate with the smallest cost; all costs are
// equal in an empty block, so slot 0 takes the earliest row).
template <typename CodeOf, typename Emit>
__device__ __forceinline__ void place_rows_cta(int n, int m, int K, int window, unsigned char* smem, CodeOf code_of, Emit emit) {
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int kw = (K + 31) >> 5;
  uint32_t* seen = reinterpret_cast<uint32_t*>(smem);                        // [m][kw] bit per code present in the open block
  unsigned long long* red = reinterpret_cast<unsigned long long*>(seen + (size_t)m * kw + ((m * kw) & 1));   // [32]
  int16_t* wcode = reinterpret_cast<int16_t*>(red + 32);                     // [window][m]
  uint8_t* cnt = reinterpret_cast<uint8_t*>(wcode + (size_t)window * m);     // [m][32] distinct codes per bank
  uint8_t* mx = cnt + (size_t)m * 32;                                        // [m]
  const int lane = tid & 31, warp = tid >> 5, nwarps = nthr >> 5;

  int mine = (tid < window && tid < n) ? tid : -1;     // arrival index of my candidate
  int next = min(window, n);                           // next arrival index to enter the window (uniform)
  if (mine >= 0)
    for (int p = 0; p < m; p++) wcode[(size_t)tid * m + p] = (int16_t)code_of(mine, p);
  // rows enter the window in arrival order, so the entrant of the NEXT step is known: thread p < m keeps its code of
  // position p in a register, fetched one step ahead (the loop is a chain of n dependent steps; nothing in it may wait
  // for global memory)
  int pre = (tid < m && next < n) ? code_of(next, tid) : 0;
  for (int placed = 0; placed < n; placed++) {
    if ((placed & 31) == 0) {
      for (int i = tid; i < m * kw; i += nthr) seen[i] = 0u;
      for (int i = tid; i < m * 32 + m; i += nthr) cnt[i] = 0;   // cnt and mx are contiguous
      __syncthreads();
    }
    unsigned long long key = ~0ull;
    if (mine >= 0) {
      int raises = 0, load = 0;
      const int16_t* wc = wcode + (size_t)tid * m;
#pragma unroll 4
      for (int p = 0; p < m; p++) {                       // branch-free: the loads of all positions overlap
        const int code = wc[p];
        const int fresh = 1 - (int)((seen[p * kw + (code >> 5)] >> (code & 31)) & 1u);   // 0: same address as a placed row, merged
        const int l = cnt[p * 32 + (code & 31)];
        raises += fresh & (int)(l + 1 > mx[p]);
        load += fresh * l;
      }
      key = ((unsigned long long)(raises * 4096 + load) << 42) | ((unsigned long long)(uint32_t)mine << 10) | (unsigned long long)tid;
    }
    for (int s = 16; s >= 1; s >>= 1) {
      const unsigned long long o = __shfl_xor_sync(0xffffffffu, key, s);
      key = o < key ? o : key;
    }
    if (lane == 0) red[warp] = key;
    __syncthreads();
    unsigned long long best = red[0];
    for (int wv = 1; wv < nwarps; wv++) { const unsigned long long o = red[wv]; best = o < best ? o : best; }
    const int win_thread = (int)(best & 1023u);
    const int win_arrival = (int)((best >> 10) & 0xffffffffu);
    // threads p < m record the winner's code of position p and hand its window slot to the entrant
    if (tid < m) {
      const int code = wcode[(size_t)win_thread * m + tid];
      uint32_t& wd = seen[tid * kw + (code >> 5)];
      if (!((wd >> (code & 31)) & 1u)) {
        wd |= 1u << (code & 31);
        const uint8_t l = ++cnt[tid * 32 + (code & 31)];
        if (l > mx[tid]) mx[tid] = l;
      }
      if (next < n) wcode[(size_t)win_thread * m + tid] = (int16_t)pre;
    }
    if (tid == 0) emit(placed, win_arrival);
    if (tid == win_thread) mine = next < n ? next : -1;
    if (next < n) next++;
    if (tid < m && next < n) pre = code_of(next, tid);     // in flight during the next step
    __syncthreads();
  }
}

// one CTA per list: arrival[row_start[c] ..] = the list's rows in arrival order; order[row_start[c] + s] = the row in slot s
__global__ void __launch_bounds__(1024)
place_rows_kernel(const int16_t* __restrict__ codes, int m, int K, const int32_t* __restrict__ arrival, const int32_t* __restrict__ row_start,
                  const int32_t* __restrict__ len, int window, int32_t* __restrict__ order) {
  extern __shared__ __align__(16) unsigned char place_smem[];
  const int c = blockIdx.x;
  const int n = len[c];
  const int32_t* rows = arrival + row_start[c];
  int32_t* out = order + row_start[c];
  if (n <= 32 || m > 64) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) out[i] = rows[i];
    return;
  }
  place_rows_cta(n, m, K, window, place_smem,
                 [&](int a, int p) { return (int)codes[(size_t)rows[a] * m + p]; },
                 [&](int slot, int a) { out[slot] = rows[a]; });
}

// The same placement for a per-call target subset (`WHERE id IN (...)`): the selected rows, in table order, are cut into
// groups of `group` rows (a multiple of 32), one CTA places each group.  order[s] = index into sel_rows of the row that
// goes to slot s of the compact table (slots beyond the selection keep the identity).
__global__ void __launch_bounds__(1024)
subset_place_kernel(const uint2* __restrict__ src_units, int U, int m, int K, const int32_t* __restrict__ sel_rows,
                    const int32_t* __restrict__ n_sel, int window, int group, size_t place_bytes, int32_t* __restrict__ order) {
  extern __shared__ __align__(16) unsigned char place_smem[];
  const int g0 = blockIdx.x * group;
  const int n = min(group, *n_sel - g0);
  if (n <= 0) return;
  if (n <= 32) {
    for (int i = threadIdx.x; i < n; i += blockDim.x) order[g0 + i] = g0 + i;
    return;
  }
  // the group's codes first go to shared memory (one scattered 8-byte load per row and unit, all threads): the
  // placement loop is a chain of n dependent steps and must not wait for global memory in any of them
  int16_t* gcode = reinterpret_cast<int16_t*>(place_smem + place_bytes);     // [group][4 * U]
  const int mp = 4 * U;
  for (int i = threadIdx.x; i < n * U; i += blockDim.x) {
    const int a = i / U, u = i - a * U;
    const int row = sel_rows[g0 + a];
    const uint2 v = src_units[((size_t)(row >> 5) * U + u) * 32 + (row & 31)];
    int16_t* dst = gcode + (size_t)a * mp + 4 * u;
    dst[0] = (int16_t)((v.x & 0xFFFFu) >> 2); dst[1] = (int16_t)(v.x >> 18);
    dst[2] = (int16_t)((v.y & 0xFFFFu) >> 2); dst[3] = (int16_t)(v.y >> 18);
  }
  __syncthreads();
  place_rows_cta(n, m, K, window, place_smem,
                 [&](int a, int p) { return (int)gcode[(size_t)a * mp + p]; },
                 [&](int slot, int a) { order[g0 + slot] = g0 + a; });
}

inline size_t place_rows_smem(int m, int K, int window) {
  const size_t kw = (size_t)(K + 31) / 32;
  return ((size_t)m * kw + (((size_t)m * kw) & 1)) * 4 + 32 * 8 + (size_t)window * m * 2 + (size_t)m * 33 + 16;
}

// position i of the list-ordered sequence -> its slot of the blocked table
__global__ void pack_rows_kernel(const int16_t* __restrict__ codes, int m, int U, const int32_t* __restrict__ order,
                                 const int32_t* __restrict__ list_of_row, int64_t rows_per_pseudo, const int32_t* __restrict__ row_start,
                                 const int32_t* __restrict__ blk, int64_t N, uint2* __restrict__ units, uint4* __restrict__ units8,
                                 int32_t* __restrict__ rowno) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  const int64_t r = order != nullptr ? order[i] : i;
  const int c = list_of_row != nullptr ? list_of_row[r] : (int)(r / rows_per_pseudo);
  const int64_t slot = i - row_start[c];
  const int64_t b = blk[c] + (slot >> 5);
  const int l = (int)(slot & 31);
  const int16_t* cr = codes + (size_t)r * m;
  for (int u = 0; u < U; u++) {
    uint32_t f[4] = {0, 0, 0, 0};
    for (int t = 0; t < 4; t++)
      if (4 * u + t < m) f[t] = (uint32_t)cr[4 * u + t] * 4u;          // pre-scaled: byte offset into a K-float LUT row
    units[((size_t)b * U + u) * 32 + l] = make_uint2(f[0] | (f[1] << 16), f[2] | (f[3] << 16));
  }
  if (units8 != nullptr) {
    uint32_t w[4] = {0, 0, 0, 0};
    for (int p = 0; p < m && p < 16; p++) w[p >> 2] |= (uint32_t)(uint8_t)cr[p] << (8 * (p & 3));
    units8[(size_t)b * 32 + l] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  rowno[(size_t)b * 32 + l] = (int32_t)r;
}

// order-sensitive checksums of a pinned table: [0] over (slot, rowno), [1] over the code units (tests: two ways of
// building a table give the same layout)
__device__ __forceinline__ unsigned long long mix64(unsigned long long x) {
  x ^= x >> 33; x *= 0xff51afd7ed558ccdull; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ull; x ^= x >> 33;
  return x;
}
__global__ void table_checksum_kernel(const int32_t* __restrict__ rowno, const uint2* __restrict__ units, const uint4* __restrict__ units8,
                                      int64_t n_slots, int U, unsigned long long* __restrict__ out) {
  unsigned long long a = 0, b = 0;
  for (int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; s < n_slots; s += (int64_t)gridDim.x * blockDim.x) {
    a += mix64(((unsigned long long)s << 32) ^ (uint32_t)rowno[s]);
    const int64_t blkno = s >> 5;
    const int l = (int)(s & 31);
    for (int u = 0; u < U; u++) {
      const uint2 v = units[((size_t)blkno * U + u) * 32 + l];
      b += mix64(((unsigned long long)(s * U + u) << 1) ^ ((unsigned long long)v.x << 32 | v.y) * 0x9e3779b97f4a7c15ull);
    }
    if (units8 != nullptr) {
      const uint4 v = units8[s];
      b += mix64(((unsigned long long)v.x << 32 | v.y) ^ mix64((unsigned long long)v.z << 32 | v.w) ^ (unsigned long long)s);
    }
  }
  for (int sft = 16; sft >= 1; sft >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, sft);
    b += __shfl_xor_sync(0xffffffffu, b, sft);
  }
  if ((threadIdx.x & 31) == 0) { atomicAdd(out + 0, a); atomicAdd(out + 1, b); }
}

// multi-index cell frequencies of the rows whose id is listed (create_statistics, freddy--0.0.1.sql:150-171: a row
// counts once per listed occurrence of its id — the reference counts rows of the JOIN)
__global__ void cell_count_listed_kernel(const int32_t* __restrict__ sorted_ids, const int32_t* __restrict__ sorted_rows, int n_table,
                                         const int32_t* __restrict__ wanted, int n_wanted, const int32_t* __restrict__ cell_of_row,
                                         unsigned long long* __restrict__ counts, int cells) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_wanted) return;
  const int32_t id = wanted[i];
  int lo = 0, hi = n_table;
  while (lo < hi) { const int mid = (lo + hi) >> 1; if (sorted_ids[mid] < id) lo = mid + 1; else hi = mid; }
  for (int pos = lo; pos < n_table && sorted_ids[pos] == id; pos++) {
    atomicAdd(counts + cell_of_row[sorted_rows[pos]], 1ull);
    atomicAdd(counts + cells, 1ull);
  }
}
__global__ void cell_count_all_kernel(const int32_t* __restrict__ cell_of_row, int64_t N, unsigned long long* __restrict__ counts, int cells) {
  const int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= N) return;
  atomicAdd(counts + cell_of_row[r], 1ull);
  if (r == 0) counts[cells] = (unsigned long long)N;
}
// coarse_freq = (count::float8 / total)::float4; the last entry is the total itself (float4)
__global__ void cell_freq_kernel(const unsigned long long* __restrict__ counts, int cells, float* __restrict__ stats) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > cells) return;
  const double total = (double)counts[cells];
  stats[c] = c == cells ? (float)total : (float)((double)counts[c] / total);
}

}  // namespace fb
