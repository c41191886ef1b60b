// Measures issue throughput of scalar vs packed (f32x2) fp32 ops on sm_100a, in
// warp-instructions per clock per SM, to size the exact (unfused) LUT/coarse kernels.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o microbench_fp32 microbench_fp32.cu
#include <cstdio>
#include <cuda_runtime.h>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float lo(u64 v) { float a, b; asm("mov.b64 {%0,%1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); return a + b; }

template <int MODE>
__global__ void k(float* out, float x, float one, int iters) {
  float a[8]; u64 p[8];
  for (int i = 0; i < 8; i++) { a[i] = x + i + threadIdx.x; p[i] = pk(a[i], a[i] + 1.f); }
  u64 xx = pk(x, x * 1.5f), one2 = pk(one, one);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) a[i] = __fadd_rn(a[i], x);
      if (MODE == 1) a[i] = __fmul_rn(a[i], x);
      if (MODE == 2) a[i] = __fmaf_rn(a[i], x, one);
      if (MODE == 3) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(xx));
      if (MODE == 4) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p[i]) : "l"(xx));
      if (MODE == 5) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(p[i]) : "l"(xx), "l"(one2));
      if (MODE == 6) {  // the exact triple, scalar: t = r - c; acc += t*t
        float t = __fsub_rn(x, a[i]); a[i] = __fadd_rn(a[i], __fmul_rn(t, t));
      }
      if (MODE == 7) {  // the exact triple, packed: sub, mul, fma(m, 1.0(runtime), acc)
        u64 t, m2;
        asm volatile("sub.rn.f32x2 %0, %1, %2;" : "=l"(t) : "l"(xx), "l"(p[i]));
        asm volatile("mul.rn.f32x2 %0, %1, %1;" : "=l"(m2) : "l"(t));
        asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(p[i]) : "l"(m2), "l"(one2));
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 8; i++) s += a[i] + lo(p[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int flops_per_op, int ops_per_iter_per_chain) {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  float* out; cudaMalloc(&out, sms * 8 * 1024 * sizeof(float));
  const int iters = 20000;
  cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
  for (int blocks_per_sm : {1, 2}) {
    k<MODE><<<sms * blocks_per_sm, 1024>>>(out, 1.0001f, 1.0f, 100);
    cudaEventRecord(a);
    k<MODE><<<sms * blocks_per_sm, 1024>>>(out, 1.0001f, 1.0f, iters);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    double warp_instr = (double)sms * blocks_per_sm * 32 * iters * 8 * ops_per_iter_per_chain;
    double elems = warp_instr * 32 * flops_per_op / ops_per_iter_per_chain;
    printf("%-28s blocks/SM=%d  %.3f ms  %.2f warp-instr/clk/SM (at %d MHz nominal)  %.2f T elem-ops/s\n", name, blocks_per_sm, ms,
           warp_instr / (ms * 1e-3) / sms / (clk * 1e3), clk / 1000, elems / (ms * 1e-3) / 1e12);
  }
  cudaFree(out);
}

int main() {
  run<0>("FADD", 1, 1); run<1>("FMUL", 1, 1); run<2>("FFMA", 1, 1);
  run<3>("FADD2 (add.f32x2)", 2, 1); run<4>("FMUL2 (mul.f32x2)", 2, 1); run<5>("FFMA2 (fma.f32x2)", 2, 1);
  run<6>("exact triple scalar", 3, 3); run<7>("exact triple packed", 6, 3);
  return 0;
}
