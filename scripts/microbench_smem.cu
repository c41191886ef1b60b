// microbench_smem.cu — data-pipe cost of one warp-wide 32-bit shared-memory gather on sm_100a for
// controlled address patterns (which addresses conflict, and what do random gathers cost?).
// Addresses come from a table in shared memory itself (8 independent pointer chains per thread, 16
// warps per SM), so the inner loop is one LDS per load: no ALU work hides or adds to the LSU cost.
// Output: SM cycles per warp-level LDS (1.0 = conflict-free).
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
constexpr int WORDS = 8192;   // 32 KB table
__global__ void k(const unsigned* __restrict__ table, int iters, unsigned* out, long long* cyc) {
  extern __shared__ unsigned s[];
  for (int i = threadIdx.x; i < WORDS; i += blockDim.x) s[i] = table[i];
  __syncthreads();
  unsigned a[8];
#pragma unroll
  for (int u = 0; u < 8; u++) a[u] = s[(threadIdx.x & 31) + 32 * (((threadIdx.x >> 5) * 8 + u) % (WORDS / 32))];   // byte offsets; chain of lane L starts at a word = L mod 32
  long long t0 = clock64();
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int u = 0; u < 8; u++) a[u] = *reinterpret_cast<const unsigned*>(reinterpret_cast<const char*>(s) + a[u]);
  }
  long long t1 = clock64();
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  unsigned x = 0;
#pragma unroll
  for (int u = 0; u < 8; u++) x ^= a[u];
  out[blockIdx.x * blockDim.x + threadIdx.x] = x;
}
// table[i] = byte offset of the next word.  The lane identity of a chain is kept in (word % 32) for the
// structured patterns: start word = (tid*8+u) % WORDS, so lane = start % 32 only if we preserve it.
int main() {
  unsigned *d_table, *out; long long* cyc;
  cudaMalloc(&d_table, WORDS * 4); cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, WORDS * 4);
  const char* names[] = {"lane L -> bank L: conflict-free", "uniformly random next word", "lanes 2j,2j+1 -> bank 2j (2-way)",
                         "lanes 8j..8j+7 -> bank j (8-way)", "lanes 3j..3j+2 -> bank 3j (3-way)", "lanes 4j..4j+3 -> bank 4j (4-way)",
                         "lane L -> bank L or L^16 at random", "lane L -> random bank in its 16-bank half"};
  srand(1);
  for (int p = 0; p < 8; p++) {
    std::vector<unsigned> t(WORDS);
    for (int i = 0; i < WORDS; i++) {
      unsigned r = (unsigned)rand();
      unsigned w;
      switch (p) {
        case 0: w = (i % 32) + 32 * (r % (WORDS / 32)); break;
        case 1: w = r % WORDS; break;
        case 2: w = ((i % 32) / 2) * 2 + 32 * (r % (WORDS / 32)); break;          // lanes 2j,2j+1 -> bank 2j, different rows
        case 3: w = ((i % 32) / 8) * 8 + 32 * (r % (WORDS / 32)); break;             // 8 lanes -> one bank
        case 4: w = ((i % 32) / 3) * 3 + 32 * (r % (WORDS / 32)); break;
        case 5: w = ((i % 32) / 4) * 4 + 32 * (r % (WORDS / 32)); break;
        case 6: w = ((i % 32) ^ ((r & 1) * 16)) + 32 * ((r >> 8) % (WORDS / 32)); break;
        default: w = ((i % 32) / 16) * 16 + (r % 16) + 32 * ((r >> 8) % (WORDS / 32)); break;
      }
      t[i] = w * 4;
    }
    cudaMemcpy(d_table, t.data(), WORDS * 4, cudaMemcpyHostToDevice);
    const int iters = 4000;
    for (int rep = 0; rep < 2; rep++) { k<<<148, 512, WORDS * 4>>>(d_table, iters, out, cyc); cudaDeviceSynchronize(); }
    long long h[148]; cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
    double avg = 0; for (int i = 0; i < 148; i++) avg += h[i]; avg /= 148;
    printf("pattern %d %-50s %.2f SM-cycles per warp LDS\n", p, names[p], avg / (16.0 * iters * 8));
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
