// vector_kernels.cuh — the dense word-vector UDFs and the exact analogy scan.
//
//   cosine_similarity / _norm (double accumulation)   cosine_similarity.c:12-45, core_functions.c:23-63
//   cosine_similarity_bytea (fp32 sequential dot)      core_functions.c:65-81
//   vec_minus_bytea / vec_plus_bytea / vec_normalize_bytea   core_functions.c:120-139, :179-196, :243-269
//   analogy_3cosadd: ORDER BY cosine_similarity_bytea(v3 - v1 + v2, v4.vector) DESC FETCH FIRST 1
//                    over the whole table, the three input words excluded   freddy--0.0.1.sql:1270-1288
//
// Same exactness rule as the rest of the engine: every product and sum is rounded
// separately, in the reference's order, so scores (and therefore the arg-max and its
// ties) are bit-identical to the CPU extension.
#pragma once
#include "common.cuh"

namespace fb {

// ---- element-wise batch UDFs: one thread per pair / vector (tiny work, d ~ 300) ----
__global__ void cosine_pairs_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, int d, int variant,
                                    double* __restrict__ out) {
  const int p = blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const float* x = a + (size_t)p * d;
  const float* y = b + (size_t)p * d;
  if (variant == 2) {  // cosine_similarity_bytea: fp32, product and sum rounded separately
    float s = 0.0f;
    for (int i = 0; i < d; i++) s = xadd(s, xmul(x[i], y[i]));
    out[p] = (double)s;
    return;
  }
  double scalar = 0.0, sq1 = 0.0, sq2 = 0.0;
  for (int i = 0; i < d; i++) {
    const double xi = (double)x[i], yi = (double)y[i];
    scalar = __dadd_rn(scalar, __dmul_rn(xi, yi));
    if (variant == 0) {
      sq2 = __dadd_rn(sq2, __dmul_rn(yi, yi));
      sq1 = __dadd_rn(sq1, __dmul_rn(xi, xi));
    }
  }
  if (variant == 1) { out[p] = scalar; return; }                                   // cosine_similarity_norm
  out[p] = (sq1 > 0.0 && sq2 > 0.0) ? __ddiv_rn(scalar, __dmul_rn(__dsqrt_rn(sq1), __dsqrt_rn(sq2))) : 0.0;
}

// op 0: a - b   1: a + b   2: a / sqrt(sum a^2)  (fp32 sum, sqrt through double, fp32 division)
__global__ void vec_ops_kernel(const float* __restrict__ a, const float* __restrict__ b, int n, int d, int op,
                               float* __restrict__ out) {
  const int v = blockIdx.x;
  if (v >= n) return;
  const float* x = a + (size_t)v * d;
  float* o = out + (size_t)v * d;
  if (op == 2) {
    __shared__ float s_len;
    if (threadIdx.x == 0) {
      float sq = 0.0f;
      for (int i = 0; i < d; i++) sq = xadd(sq, xmul(x[i], x[i]));                // sequential, as the reference
      s_len = (float)__dsqrt_rn((double)sq);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < d; i += blockDim.x) o[i] = __fdiv_rn(x[i], s_len);
    return;
  }
  const float* y = b + (size_t)v * d;
  for (int i = threadIdx.x; i < d; i += blockDim.x) o[i] = (op == 0) ? xsub(x[i], y[i]) : xadd(x[i], y[i]);
}

// ---- word-vector table in HBM: blocks of 32 rows, dimension-major inside a block:
//      vT[(blk * d + i) * 32 + lane]  is dimension i of row blk*32+lane (zero padding) ----
__global__ void transpose_rows_kernel(const float* __restrict__ rows, int64_t N, int d, float* __restrict__ vT) {
  const int64_t blk = blockIdx.x;
  for (int idx = threadIdx.x; idx < d * 32; idx += blockDim.x) {
    const int lane = idx & 31, i = idx >> 5;
    const int64_t r = blk * 32 + lane;
    vT[(blk * d + i) * 32 + lane] = (r < N) ? rows[r * d + i] : 0.0f;
  }
}

// gather rows a, b, c of each query and form (v_c - v_a) + v_b  (vec_minus_bytea then vec_plus_bytea)
__global__ void analogy_query_kernel(const float* __restrict__ vT, int d, const int32_t* __restrict__ rows_abc, int nq,
                                     float* __restrict__ qvecs) {
  const int q = blockIdx.x;
  if (q >= nq) return;
  const int ra = rows_abc[3 * q], rb = rows_abc[3 * q + 1], rc = rows_abc[3 * q + 2];
  for (int i = threadIdx.x; i < d; i += blockDim.x) {
    const float va = vT[((size_t)(ra >> 5) * d + i) * 32 + (ra & 31)];
    const float vb = vT[((size_t)(rb >> 5) * d + i) * 32 + (rb & 31)];
    const float vc = vT[((size_t)(rc >> 5) * d + i) * 32 + (rc & 31)];
    qvecs[(size_t)q * d + i] = xadd(xsub(vc, va), vb);
  }
}

// order scores descending, rows ascending, in one unsigned key (max wins)
__device__ __forceinline__ u64 score_key(float s, uint32_t row) {
  uint32_t b = __float_as_uint(s);
  b ^= (b & 0x80000000u) ? 0xFFFFFFFFu : 0x80000000u;
  return ((u64)b << 32) | (u64)(0xFFFFFFFFu - row);
}
__device__ __forceinline__ float key_score(u64 key) {
  uint32_t b = (uint32_t)(key >> 32);
  b ^= (b & 0x80000000u) ? 0x80000000u : 0xFFFFFFFFu;
  return __uint_as_float(b);
}
__device__ __forceinline__ uint32_t key_row(u64 key) { return 0xFFFFFFFFu - (uint32_t)key; }

constexpr int kAnaQT = 32;        // queries per CTA tile
constexpr int kAnaThreads = 256;
constexpr int kAnaWarps = kAnaThreads / kWarp;

// The exact scan: score[q][r] = sum_i q[i] * v_r[i] (sequential fp32, product and sum
// rounded separately; two queries per FMUL2/FFMA2, see common.cuh).  One CTA = 32 queries x
// one slab of rows; a lane owns one row of a 32-row block (coalesced dimension-major loads),
// the 32 query values of a dimension are broadcast from shared memory.  After each block the
// 32x32 score tile is transposed through shared memory so that lane j tracks the best
// (score desc, row asc) of query j.  Output: one key per (slab, query).
__global__ void __launch_bounds__(kAnaThreads, 2)
analogy_scan_kernel(const float* __restrict__ vT, int64_t N, int d, int blocks_per_slab,
                    const float* __restrict__ qvecs, int nq,
                    const int32_t* __restrict__ exclude_rows,   // [nq][3] table rows, -1 = none
                    u64* __restrict__ partial,                   // [n_slabs][nq_pad]
                    int nq_pad, float one) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* qs = reinterpret_cast<float*>(smem_raw);                               // [d][32]
  float* tile = qs + (size_t)d * kAnaQT;                                        // [warps][32][33]
  __shared__ u64 s_best[kAnaWarps][kAnaQT];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int q0 = blockIdx.x * kAnaQT;
  const int64_t blk_begin = (int64_t)blockIdx.y * blocks_per_slab;
  const int64_t n_blocks = (N + 31) >> 5;
  const int64_t blk_end = min(blk_begin + blocks_per_slab, n_blocks);

  for (int idx = tid; idx < d * kAnaQT; idx += kAnaThreads) {
    const int i = idx >> 5, j = idx & 31;
    qs[idx] = (q0 + j < nq) ? qvecs[(size_t)(q0 + j) * d + i] : 0.0f;
  }
  int ex0 = -1, ex1 = -1, ex2 = -1;   // lane j: exclusions of query q0 + j
  if (q0 + lane < nq) {
    ex0 = exclude_rows[3 * (q0 + lane)];
    ex1 = exclude_rows[3 * (q0 + lane) + 1];
    ex2 = exclude_rows[3 * (q0 + lane) + 2];
  }
  __syncthreads();

  const u64 one2 = pack2(one, one);
  u64 best = 0ull;   // smaller than any real key
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
    // transpose: lane (= row) writes its 32 scores, then lane j (= query) scans the 32 rows
#pragma unroll
    for (int j = 0; j < kAnaQT / 2; j++) {
      float lo, hi;
      unpack2(acc2[j], lo, hi);
      mytile[lane * 33 + 2 * j] = lo;
      mytile[lane * 33 + 2 * j + 1] = hi;
    }
    __syncwarp();
    const int64_t row0 = blk * 32;
#pragma unroll 4
    for (int r = 0; r < 32; r++) {
      const int64_t row = row0 + r;
      if (row >= N) break;
      if ((int)row == ex0 || (int)row == ex1 || (int)row == ex2) continue;
      const u64 key = score_key(mytile[r * 33 + lane], (uint32_t)row);
      best = key > best ? key : best;
    }
    __syncwarp();
  }
  s_best[warp][lane] = best;
  __syncthreads();
  if (warp == 0) {
    u64 b = s_best[0][lane];
#pragma unroll
    for (int wv = 1; wv < kAnaWarps; wv++) b = s_best[wv][lane] > b ? s_best[wv][lane] : b;
    partial[(size_t)blockIdx.y * nq_pad + q0 + lane] = b;
  }
}

// arg-max over slabs: one thread per query
__global__ void analogy_reduce_kernel(const u64* __restrict__ partial, int n_slabs, int nq, int nq_pad,
                                      const int32_t* __restrict__ ids, int64_t row_base,
                                      int32_t* __restrict__ out_ids, float* __restrict__ out_scores,
                                      int32_t* __restrict__ out_rows) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  u64 best = 0ull;
  for (int s = 0; s < n_slabs; s++) {
    const u64 kx = partial[(size_t)s * nq_pad + q];
    best = kx > best ? kx : best;
  }
  if (best == 0ull) {   // no eligible row
    out_ids[q] = -1; out_scores[q] = 0.0f;
    if (out_rows) out_rows[q] = -1;
    return;
  }
  const uint32_t row = key_row(best);
  out_ids[q] = ids[row];
  out_scores[q] = key_score(best);
  if (out_rows) out_rows[q] = (int32_t)(row_base + row);
}

}  // namespace fb
