// rerank_kernels.cuh — exact cosine k-NN and post-verification.
//
// The reference does these in SQL around two C pieces that are on the hot path already:
//   k_nearest_neighbour(v, k)            freddy--0.0.1.sql:426-439   ORDER BY cosine_similarity_bytea(v, vector) DESC
//   knn_in_exact(v, k, ids)              freddy--0.0.1.sql:1026-1038 same, WHERE id = ANY(ids)
//   k_nearest_neighbour_ivfadc_pv(v, k)  freddy--0.0.1.sql:574-591   ivfadc_search(v, pvf*k) JOIN vectors ON idx = id,
//                                                                     ORDER BY cosine_similarity_bytea DESC FETCH FIRST k
// cosine_similarity_bytea (core_functions.c:67-81) is `scalar += v1[i] * v2[i]` in float4: product and sum
// rounded separately, left to right — the chain reproduced here bit for bit.
// ORDER BY ... FETCH FIRST k leaves the order of EQUAL similarities to the executor; this engine orders
// them by table row (the order a sequential scan delivers them in), stated in DESIGN.md.
#pragma once
#include "common.cuh"
#include "vector_kernels.cuh"

namespace fb {

constexpr int kKnnMaxK = 32;      // exact scan keeps k <= 32 keys per (warp, query)

// Exact scan with top-k.  Same tiling as analogy_scan_kernel: one CTA = 32 queries x one slab of rows, a
// lane owns one row of a 32-row block, the 32x32 score tile is transposed through shared memory so that
// lane j sees the 32 new scores of query j and inserts the ones that beat its k-th best (a descending
// key list per (warp, query) in shared memory; keys = (score desc, row asc), unique).
// Output: partial[slab][query][k] descending, 0 = empty.
__global__ void __launch_bounds__(kAnaThreads, 2)
exact_knn_scan_kernel(const float* __restrict__ vT, int64_t N, int d, int blocks_per_slab,
                      const float* __restrict__ qvecs, int nq, int k,
                      u64* __restrict__ partial,                   // [n_slabs][nq_pad][k]
                      int nq_pad, float one) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* qs = reinterpret_cast<float*>(smem_raw);                               // [d][32]
  float* tile = qs + (size_t)d * kAnaQT;                                        // [warps][32][33]
  u64* top = reinterpret_cast<u64*>(tile + (size_t)kAnaWarps * 32 * 33);        // [warps][32][k]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q0 = blockIdx.x * kAnaQT;
  const int64_t blk_begin = (int64_t)blockIdx.y * blocks_per_slab;
  const int64_t n_blocks = (N + 31) >> 5;
  const int64_t blk_end = min(blk_begin + blocks_per_slab, n_blocks);

  for (int idx = tid; idx < d * kAnaQT; idx += kAnaThreads) {
    const int i = idx >> 5, j = idx & 31;
    qs[idx] = (q0 + j < nq) ? qvecs[(size_t)(q0 + j) * d + i] : 0.0f;
  }
  u64* mylist = top + ((size_t)warp * 32 + lane) * k;
  for (int p = 0; p < k; p++) mylist[p] = 0ull;
  __syncthreads();

  const u64 one2 = pack2(one, one);
  u64 thr = 0ull;   // k-th best key of this (warp, query) so far
  float* mytile = tile + (size_t)warp * 32 * 33;
  for (int64_t blk = blk_begin + warp; blk < blk_end; blk += kAnaWarps) {
    const float* vp = vT + (size_t)blk * d * 32 + lane;
    u64 acc2[kAnaQT / 2];
#pragma unroll
    for (int j = 0; j < kAnaQT / 2; j++) acc2[j] = 0ull;
#pragma unroll 2
    for (int i = 0; i < d; i++) {
      const float v = __ldg(vp + (size_t)i * 32);
      const u64 v2 = pack2(v, v);
      const ulonglong2* qrow = reinterpret_cast<const ulonglong2*>(qs + i * kAnaQT);
#pragma unroll
      for (int t = 0; t < kAnaQT / 4; t++) {
        const ulonglong2 q4 = qrow[t];
        acc2[2 * t] = xacc2(xmul2(q4.x, v2), one2, acc2[2 * t]);
        acc2[2 * t + 1] = xacc2(xmul2(q4.y, v2), one2, acc2[2 * t + 1]);
      }
    }
#pragma unroll
    for (int j = 0; j < kAnaQT / 2; j++) {
      float lo, hi;
      unpack2(acc2[j], lo, hi);
      mytile[lane * 33 + 2 * j] = lo;
      mytile[lane * 33 + 2 * j + 1] = hi;
    }
    __syncwarp();
    const int64_t row0 = blk * 32;
    for (int r = 0; r < 32; r++) {
      const int64_t row = row0 + r;
      if (row >= N) break;
      const u64 key = score_key(mytile[r * 33 + lane], (uint32_t)row);
      if (key > thr) {
        int p = k - 1;
        while (p > 0 && mylist[p - 1] < key) { mylist[p] = mylist[p - 1]; p--; }
        mylist[p] = key;
        thr = mylist[k - 1];
      }
    }
    __syncwarp();
  }
  __syncthreads();
  if (warp == 0) {   // lane j merges the warps' lists of query j into its own
    for (int wv = 1; wv < kAnaWarps; wv++) {
      const u64* other = top + ((size_t)wv * 32 + lane) * k;
      for (int s = 0; s < k; s++) {
        const u64 key = other[s];
        if (key <= thr) break;                    // descending: nothing further can enter
        int p = k - 1;
        while (p > 0 && mylist[p - 1] < key) { mylist[p] = mylist[p - 1]; p--; }
        mylist[p] = key;
        thr = mylist[k - 1];
      }
    }
    u64* out = partial + ((size_t)blockIdx.y * nq_pad + q0 + lane) * k;
    for (int p = 0; p < k; p++) out[p] = mylist[p];
  }
}

// merge of the slabs' lists: one thread per query, k <= kKnnMaxK
__global__ void exact_knn_reduce_kernel(const u64* __restrict__ partial, int n_slabs, int nq, int nq_pad, int k,
                                        const int32_t* __restrict__ ids,     // id by row of the word-vector table
                                        const int32_t* __restrict__ row_map, // scanned (gathered) row -> word-vector row, or nullptr
                                        int32_t* __restrict__ out_ids, float* __restrict__ out_sims) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  u64 best[kKnnMaxK];
  for (int p = 0; p < k; p++) best[p] = 0ull;
  u64 thr = 0ull;
  for (int s = 0; s < n_slabs; s++) {
    const u64* in = partial + ((size_t)s * nq_pad + q) * k;
    for (int t = 0; t < k; t++) {
      const u64 key = in[t];
      if (key <= thr) break;
      int p = k - 1;
      while (p > 0 && best[p - 1] < key) { best[p] = best[p - 1]; p--; }
      best[p] = key;
      thr = best[k - 1];
    }
  }
  for (int p = 0; p < k; p++) {
    if (best[p] == 0ull) { out_ids[(size_t)q * k + p] = -1; out_sims[(size_t)q * k + p] = 0.0f; continue; }
    uint32_t row = key_row(best[p]);
    if (row_map) row = (uint32_t)row_map[row];
    out_ids[(size_t)q * k + p] = ids[row];
    out_sims[(size_t)q * k + p] = key_score(best[p]);
  }
}

// gather rows of the row-major image into a compact dimension-major blocked table (knn_in_exact subset)
__global__ void gather_vec_blocks_kernel(const float* __restrict__ vR, int d, const int32_t* __restrict__ rows, int n,
                                         float* __restrict__ out) {
  const int s = blockIdx.x * 32 + (threadIdx.x & 31);      // destination slot
  const int i0 = threadIdx.x >> 5, step = blockDim.x >> 5;
  const bool ok = s < n;
  const int r = ok ? rows[s] : 0;
  for (int i = i0; i < d; i += step)
    out[((size_t)blockIdx.x * d + i) * 32 + (s & 31)] = ok ? vR[(size_t)r * d + i] : 0.0f;
}

// Post-verification: one CTA per query.  cand_ids[q][kp] are the ids ivfadc_search / pq_search returned (rank
// order, -1 = unfilled).  INNER JOIN ... ON idx = id: unknown ids and -1 drop out.  Each thread computes the
// exact similarity of some candidates (one sequential fp32 chain each), the CTA sorts the keys
// (similarity desc, table row asc) and writes the first k.
constexpr int kPvThreads = 128;
constexpr int kPvMaxCand = 2048;

__device__ __forceinline__ int find_row_sorted(const int32_t* __restrict__ sorted_ids, const int32_t* __restrict__ sorted_rows,
                                               int n, int id) {
  int lo = 0, hi = n;
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (sorted_ids[mid] < id) lo = mid + 1; else hi = mid;
  }
  return (lo < n && sorted_ids[lo] == id) ? sorted_rows[lo] : -1;
}

__global__ void __launch_bounds__(kPvThreads)
pv_rerank_kernel(const float* __restrict__ queries, int d, const int32_t* __restrict__ cand_ids, int kp, int k,
                 const float* __restrict__ vR, const int32_t* __restrict__ vec_ids,
                 const int32_t* __restrict__ sorted_ids, const int32_t* __restrict__ sorted_rows, int n_vec,
                 int32_t* __restrict__ out_ids, float* __restrict__ out_sims) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* qs = reinterpret_cast<float*>(smem_raw);                    // [d]
  u64* keys = reinterpret_cast<u64*>(qs + ((d + 3) & ~3));           // [n_pad]
  const int q = blockIdx.x, tid = threadIdx.x;
  int n_pad = 32;
  while (n_pad < kp) n_pad <<= 1;
  for (int i = tid; i < d; i += kPvThreads) qs[i] = queries[(size_t)q * d + i];
  __syncthreads();
  for (int c = tid; c < n_pad; c += kPvThreads) {
    u64 key = 0ull;
    if (c < kp) {
      const int id = cand_ids[(size_t)q * kp + c];
      const int row = (id >= 0) ? find_row_sorted(sorted_ids, sorted_rows, n_vec, id) : -1;
      if (row >= 0) {
        const float* vp = vR + (size_t)row * d;                  // row-major fp32 image
        float acc = 0.0f;
        for (int i = 0; i < d; i++) acc = xadd(acc, xmul(qs[i], __ldg(vp + i)));
        key = score_key(acc, (uint32_t)row);
      }
    }
    keys[c] = key;
  }
  __syncthreads();
  // bitonic sort, descending
  for (int size = 2; size <= n_pad; size <<= 1) {
    for (int s = size >> 1; s >= 1; s >>= 1) {
      for (int i = tid; i < n_pad; i += kPvThreads) {
        const int j = i ^ s;
        if (j > i) {
          const bool desc = (i & size) == 0;
          const u64 a = keys[i], b = keys[j];
          if (desc ? (a < b) : (a > b)) { keys[i] = b; keys[j] = a; }
        }
      }
      __syncthreads();
    }
  }
  for (int p = tid; p < k; p += kPvThreads) {
    const u64 key = (p < n_pad) ? keys[p] : 0ull;
    // the same row can arrive twice only if the candidate list repeats an id; SQL would return it twice too
    out_ids[(size_t)q * k + p] = key ? vec_ids[key_row(key)] : -1;
    out_sims[(size_t)q * k + p] = key ? key_score(key) : 0.0f;
  }
}

}  // namespace fb
