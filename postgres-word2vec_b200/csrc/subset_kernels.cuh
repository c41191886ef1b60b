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
// `order` (may be null): slot s takes the order[s]-th selected row (conflict-aware placement inside groups of rows,
// subset_place_kernel); the rows' table order travels in dst_rowno
__global__ void subset_gather_kernel(const uint2* __restrict__ src_units, int U, const int32_t* __restrict__ sel_rows,
                                     const int32_t* __restrict__ n_sel, const int32_t* __restrict__ order,
                                     uint2* __restrict__ dst_units, int32_t* __restrict__ dst_rowno, int n_dst_slots) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  const int n = *n_sel;
  if (s >= n_dst_slots || s >= ((n + 31) & ~31)) return;     // slots beyond the last live block are never read
  const int db = s >> 5, dl = s & 31;
  if (s < n) {
    const int r = sel_rows[order != nullptr ? order[s] : s];
    const int sb = r >> 5, sl = r & 31;
    for (int u = 0; u < U; u++) dst_units[((size_t)db * U + u) * 32 + dl] = src_units[((size_t)sb * U + u) * 32 + sl];
    dst_rowno[s] = r;
  } else {
    for (int u = 0; u < U; u++) dst_units[((size_t)db * U + u) * 32 + dl] = make_uint2(0, 0);
    dst_rowno[s] = -1;
  }
}

}  // namespace fb

// ---------------------------------------------------------------------------------------------------
// In-place append to a pinned code table (insert_batch: freddy.c:1611-1625 adds rows to pq_quantization,
// fine_quantization and fine_quantization_ivpq).  Lists are packed back to back, so a list that grows moves
// every list behind it: the table is re-packed ON THE DEVICE into fresh buffers (one copy of ~28 B per row at
// HBM speed), the new rows land behind the old rows of their list (arrival order = old rows, then new rows in
// the order given), and nothing but the new rows crosses PCIe.
// ---------------------------------------------------------------------------------------------------
namespace fb {

// one CTA per list: move the list's old blocks to their new place, clear the blocks it gained
__global__ void append_repack_kernel(const uint2* __restrict__ old_units, const int32_t* __restrict__ old_rowno,
                                     const uint4* __restrict__ old_units8, int U,
                                     const int32_t* __restrict__ old_blk, const int32_t* __restrict__ old_len,
                                     const int32_t* __restrict__ new_blk, const int32_t* __restrict__ new_len,
                                     uint2* __restrict__ units, int32_t* __restrict__ rowno, uint4* __restrict__ units8) {
  const int c = blockIdx.x;
  const int ob = old_blk[c], nb = new_blk[c];
  const int o_blocks = (old_len[c] + 31) >> 5, n_blocks = (new_len[c] + 31) >> 5;
  for (int i = threadIdx.x; i < n_blocks * 32; i += blockDim.x) {
    const int b = i >> 5, l = i & 31;
    const bool live = b < o_blocks;
    for (int u = 0; u < U; u++)
      units[((size_t)(nb + b) * U + u) * 32 + l] = live ? old_units[((size_t)(ob + b) * U + u) * 32 + l] : make_uint2(0, 0);
    rowno[(size_t)(nb + b) * 32 + l] = live ? old_rowno[(size_t)(ob + b) * 32 + l] : -1;
    if (units8 != nullptr) units8[(size_t)(nb + b) * 32 + l] = live ? old_units8[(size_t)(ob + b) * 32 + l] : make_uint4(0, 0, 0, 0);
  }
}

// one thread per appended row: dst_slot = (new first block of its list) * 32 + position inside the list
__global__ void append_rows_kernel(const int16_t* __restrict__ codes, const int32_t* __restrict__ new_ids, const int64_t* __restrict__ dst_slot,
                                   int n, int m, int U, int64_t first_row, uint2* __restrict__ units, int32_t* __restrict__ rowno,
                                   uint4* __restrict__ units8, int32_t* __restrict__ ids) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t slot = dst_slot[i];
  const int64_t b = slot >> 5;
  const int l = (int)(slot & 31);
  const int16_t* cr = codes + (size_t)i * m;
  for (int u = 0; u < U; u++) {
    uint32_t f[4] = {0, 0, 0, 0};
    for (int t = 0; t < 4; t++)
      if (4 * u + t < m) f[t] = (uint32_t)cr[4 * u + t] * 4u;
    units[((size_t)b * U + u) * 32 + l] = make_uint2(f[0] | (f[1] << 16), f[2] | (f[3] << 16));
  }
  if (units8 != nullptr) {
    uint32_t w[4] = {0, 0, 0, 0};
    for (int p = 0; p < m && p < 16; p++) w[p >> 2] |= (uint32_t)(uint8_t)cr[p] << (8 * (p & 3));
    units8[(size_t)b * 32 + l] = make_uint4(w[0], w[1], w[2], w[3]);
  }
  rowno[slot] = (int32_t)(first_row + i);
  ids[first_row + i] = new_ids[i];
}

}  // namespace fb
