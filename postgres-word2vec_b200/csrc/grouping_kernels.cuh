// grouping_kernels.cuh — grouping_pq: for every selected row of the flat pq table, the
// nearest of G group vectors by ADC distance (freddy.c:1178-1401).
//   LUT[g] = getPrecomputedDistances(group vector g, pq codebook)          freddy.c:1291-1299
//   per row, groups in ascending group-id order: distance = sum_j LUT[g][j][code_j] (sequential fp32),
//   `if (distance < minDist)` from minDist = 100 -> first minimum wins      freddy.c:1340-1352
// One CTA = 8 blocks of 32 rows (a lane keeps its row's codes in registers); the G LUTs stream through two
// shared-memory buffers (1-D bulk async copies + mbarriers) and every thread walks them in order.
#pragma once
#include "common.cuh"
#include "ivfadc_kernels.cuh"

namespace fb {

constexpr int kGroupThreads = 256;

template <int M>
__global__ void __launch_bounds__(kGroupThreads, 2)
grouping_argmin_kernel(CodeTableDev tab, int n_blocks, int n_rows, const float* __restrict__ luts, int G, int K,
                       int32_t* __restrict__ nearest,      // [n_blocks * 32] group index per slot, -1 = padding / none
                       int32_t* __restrict__ err_flag) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar[2];
  const int m = (M > 0) ? M : tab.m;
  const int U = (M > 0) ? (M + 3) / 4 : tab.U;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const size_t lut_floats = (size_t)m * K;
  const uint32_t lut_bytes = (uint32_t)(lut_floats * sizeof(float));
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar[0], lut_bytes);
    bulk_g2s(smem_raw, luts, lut_bytes, &bar[0]);
    if (G > 1) {
      mbar_expect_tx(&bar[1], lut_bytes);
      bulk_g2s(smem_raw + lut_bytes, luts + lut_floats, lut_bytes, &bar[1]);
    }
  }
  const int blk = blockIdx.x * (kGroupThreads / 32) + warp;
  const bool have = blk < n_blocks;
  const int slot = blk * 32 + lane;
  const bool valid = have && tab.rowno[have ? slot : 0] >= 0 && slot < n_blocks * 32;
  constexpr int UU = (M > 0) ? (M + 3) / 4 : 1;
  uint2 codes[UU];
  if (M > 0 && have) {
    const uint2* up = tab.units + ((size_t)blk * U) * 32 + lane;
#pragma unroll
    for (int u = 0; u < UU; u++) codes[u] = __ldg(up + u * 32);
  }
  float min_dist = 100.0f;   // "sufficient high value", freddy.c:1326
  int best = -1;
  const uint32_t row_stride = (uint32_t)K * 4u;
  for (int g = 0; g < G; g++) {
    mbar_wait(&bar[g & 1], (uint32_t)((g >> 1) & 1));
    const char* lut_base = reinterpret_cast<const char*>(smem_raw) + (size_t)(g & 1) * lut_bytes;
    if (have) {
      float dist;
      if (M > 0) dist = adc_units<M, 0>(codes, lut_base, row_stride);
      else dist = adc_block_row<0, 0>(tab.units + ((size_t)blk * U) * 32 + lane, lut_base, m, U, row_stride);
      if (dist < min_dist) { min_dist = dist; best = g; }
    }
    __syncthreads();   // every warp is done with buffer g & 1
    if (tid == 0 && g + 2 < G) {
      mbar_expect_tx(&bar[g & 1], lut_bytes);
      bulk_g2s(smem_raw + (size_t)(g & 1) * lut_bytes, luts + (size_t)(g + 2) * lut_floats, lut_bytes, &bar[g & 1]);
    }
  }
  if (have) {
    nearest[slot] = valid ? best : -1;
    if (valid && best < 0) atomicExch(err_flag, 1);   // every distance >= 100: the reference reads an uninitialised slot
  }
  (void)n_rows;
}

}  // namespace fb
