// subset_kernels.cuh — `WHERE id IN (...)` on the device.
//
// pq_search_in / pq_search_in_batch (freddy.c:544-562, :1099-1113), grouping_pq (:1286-1300) and ivpq_search_in
// (ivpq_search_in.c:352-401) select the rows of a pinned code table whose id is in an int[] argument; the
// reference gets them from SPI in table order, every matching row once however often the id is listed.
// Here: ids -> rows through a sorted (id, row) image of the table (binary search per listed id), a row
// bitmap (duplicates in the list collapse, table order is restored for free), an exclusive scan of the
// bitmap's popcounts, and a gather of the selected rows' code units into a compact blocked table that the
// scan kernels read like any inverted list.  No host round trip: the row count stays on the device.
#pragma once
#include "common.cuh"

namespace fb {

__device__ __forceinline__ int lower_bound_i32(const int32_t* __restrict__ a, int n, int32_t key) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (a[mid] < key) lo = mid + 1; else hi = mid;
  }
  return lo;
}

// one thread per listed id: every table row carrying that id is marked
__global__ void subset_mark_kernel(const int32_t* __restrict__ sorted_ids, const int32_t* __restrict__ sorted_rows, int n_table,
                                   const int32_t* __restrict__ wanted, int n_wanted, uint32_t* __restrict__ bitmap) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_wanted) return;
  const int32_t id = wanted[i];
  for (int pos = lower_bound_i32(sorted_ids, n_table, id); pos < n_table && sorted_ids[pos] == id; pos++) {
    const int row = sorted_rows[pos];
    atomicOr(bitmap + (row >> 5), 1u << (row & 31));
  }
}

// exclusive scan of the popcounts of the bitmap words (one CTA of 1024 threads; 3M rows = 94k words);
// word_base[w] = selected rows before word w; *total = all selected rows (also written to total2 if non-null)
__global__ void __launch_bounds__(1024)
subset_scan_kernel(const uint32_t* __restrict__ bitmap, int n_words, int32_t* __restrict__ word_base, int32_t* __restrict__ total,
                   int32_t* __restrict__ total2) {
  __shared__ int s_part[1024];
  const int tid = threadIdx.x;
  const int per = (n_words + 1023) / 1024;
  const int w0 = tid * per, w1 = min(n_words, w0 + per);
  int sum = 0;
  for (int w = w0; w < w1; w++) sum += __popc(bitmap[w]);
  s_part[tid] = sum;
  __syncthreads();
  for (int off = 1; off < 1024; off <<= 1) {      // Hillis-Steele inclusive scan
    const int v = (tid >= off) ? s_part[tid - off] : 0;
    __syncthreads();
    s_part[tid] += v;
    __syncthreads();
  }
  int base = s_part[tid] - sum;
  for (int w = w0; w < w1; w++) { word_base[w] = base; base += __popc(bitmap[w]); }
  if (tid == 1023) { *total = s_part[1023]; if (total2) *total2 = s_part[1023]; }
}

// sel_rows[slot] = table row of the slot-th selected row (table order)
__global__ void subset_compact_kernel(const uint32_t* __restrict__ bitmap, int n_words, const int32_t* __restrict__ word_base,
                                      int32_t* __restrict__ sel_rows) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n_words) return;
  uint32_t bits = bitmap[w];
  int slot = word_base[w];
  while (bits) {
    const int b = __ffs(bits) - 1;
    sel_rows[slot++] = w * 32 + b;
    bits &= bits - 1;
  }
}

// gather the selected rows of a blocked code table (blocks in table order) into a compact blocked table;
// *n_sel rows are live, the remaining slots of the last block are padding (rowno -1)
__global__ void subset_gather_kernel(const uint2* __restrict__ src_units, int U, const int32_t* __restrict__ sel_rows,
                                     const int32_t* __restrict__ n_sel, uint2* __restrict__ dst_units,
                                     int32_t* __restrict__ dst_rowno, int n_dst_slots) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = *n_sel;
  if (s >= n_dst_slots || s >= ((n + 31) & ~31)) return;     // slots beyond the last live block are never read
  const int db = s >> 5, dl = s & 31;
  if (s < n) {
    const int r = sel_rows[s];
    const int sb = r >> 5, sl = r & 31;
    for (int u = 0; u < U; u++) dst_units[((size_t)db * U + u) * 32 + dl] = src_units[((size_t)sb * U + u) * 32 + sl];
    dst_rowno[s] = r;
  } else {
    for (int u = 0; u < U; u++) dst_units[((size_t)db * U + u) * 32 + dl] = make_uint2(0, 0);
    dst_rowno[s] = -1;
  }
}

}  // namespace fb
