// common.cuh — device helpers shared by the sm_100a kernels.
//
// Exact-arithmetic rule of this engine: every distance the
// reference computes is a left-to-right fp32 chain of individually rounded
// sub / mul / add (index_utils.c:500-508, :1126-1133).  All device arithmetic
// that feeds a returned distance or a rank therefore goes through the
// __f*_rn intrinsics below, which ptxas never contracts into FMA.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace fb {

typedef unsigned long long u64;

constexpr int kWarp = 32;
constexpr u64 kKeyInf = 0xFFFFFFFFFFFFFFFFull;

__device__ __forceinline__ float xsub(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ float xmul(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float xadd(float a, float b) { return __fadd_rn(a, b); }

// Packed fp32x2 forms of the same three individually rounded operations (sm_100a
// FADD2 / FMUL2 / FFMA2: two IEEE fp32 lanes per instruction, one issue slot).
// ptxas 12.9 contracts `mul.rn.f32x2` + `add.rn.f32x2` into one FFMA2 (single
// rounding) even under -fmad=false, so the accumulate step is written as
// fma(m, ONE, acc) with ONE = (1.0f, 1.0f) supplied at RUN time: m * 1.0 is exact,
// so the result is round(m + acc) — bit-identical to add.rn — and there is nothing
// left for ptxas to contract.  tests/ check the bits on the GPU.
__device__ __forceinline__ u64 pack2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ void unpack2(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 xsub2(u64 a, u64 b) {
  u64 r;
  asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 xmul2(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 xacc2(u64 m, u64 one2, u64 acc) {  // acc + m, rounded once per lane
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(m), "l"(one2), "l"(acc));
  return r;
}

// Selection key: distances are sums of squares (>= +0), so their bit patterns
// order like unsigned integers.  Low word = arrival order t (table row number):
// ascending key == (distance asc, arrival asc).
__device__ __forceinline__ u64 make_key(float d, uint32_t t) {
  return ((u64)__float_as_uint(d) << 32) | (u64)t;
}
__device__ __forceinline__ float key_dist(u64 key) { return __uint_as_float((uint32_t)(key >> 32)); }
__device__ __forceinline__ uint32_t key_dbits(u64 key) { return (uint32_t)(key >> 32); }
__device__ __forceinline__ uint32_t key_t(u64 key) { return (uint32_t)key; }

__device__ __forceinline__ u64 shfl_u64(u64 v, int src) {
  return __shfl_sync(0xffffffffu, v, src);
}
__device__ __forceinline__ u64 shfl_up_u64(u64 v, int delta) {
  return __shfl_up_sync(0xffffffffu, v, delta);
}
__device__ __forceinline__ u64 shfl_xor_u64(u64 v, int mask) {
  return __shfl_xor_sync(0xffffffffu, v, mask);
}

// A warp holds an ascending list of 32 keys, lane i = i-th smallest.
// Insert `nk` (warp-uniform) keeping the 32 smallest.
__device__ __forceinline__ void warp_list_insert(u64& mine, u64 nk, int lane) {
  u64 up = shfl_up_u64(mine, 1);
  if (lane == 0) up = 0;
  if (mine > nk) mine = (up > nk) ? up : nk;
}

// Merge an ascending 32-key list `other` (lane i = i-th) into `mine`.
__device__ __forceinline__ void warp_list_merge(u64& mine, u64 other, int lane) {
  u64 rev = shfl_u64(other, 31 - lane);
  mine = mine < rev ? mine : rev;  // 32 smallest of the union, bitonic
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    u64 partner = shfl_xor_u64(mine, s);
    bool upper = (lane & s) != 0;
    bool take = upper ? (partner > mine) : (partner < mine);
    if (take) mine = partner;
  }
}

// Full ascending sort of 32 keys, one per lane (bitonic network, 15 exchange steps).
__device__ __forceinline__ u64 warp_sort_u64(u64 v, int lane) {
#pragma unroll
  for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
    for (int s = size >> 1; s >= 1; s >>= 1) {
      const u64 partner = shfl_xor_u64(v, s);
      const bool up = ((lane & size) == 0) || size == 32;      // direction of this lane's subsequence
      const bool lower = (lane & s) == 0;
      const bool take_min = (lower == up);
      v = take_min ? (partner < v ? partner : v) : (partner > v ? partner : v);
    }
  }
  return v;
}

// ---- mbarrier + 1-D bulk async copy (TMA engine; SASS: UBLKCP / SYNCS) ----
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
          smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  }
}

}  // namespace fb
