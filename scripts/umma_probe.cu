// umma_probe.cu — stand-alone check of the tcgen05 pre-filter (csrc/prefilter_kernels.cuh) on a B200:
//   correctness (small N): every approximate score against a CPU fp32-chain score and the proven bound,
//     candidate sets against the true top-k, the re-scored top-k against the CPU's;
//   performance (large N): the GEMM launch alone and GEMM + re-score, CUDA events.
// Build:  nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo -o scripts/umma_probe scripts/umma_probe.cu
// Run:    scripts/umma_probe <nq> <N> <d> <k> [perf]
#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../postgres-word2vec_b200/csrc/prefilter_kernels.cuh"

using namespace fb;

#define CK(x)                                                                            \
  do {                                                                                   \
    cudaError_t e__ = (x);                                                               \
    if (e__ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e__), __FILE__, __LINE__); exit(2); } \
  } while (0)

__global__ void fill_rows_kernel(float* v, long long n, int d, unsigned seed, int n_clusters, float sigma) {
  // cluster centre (hash of cluster id, dim) + noise (hash of row, dim), then normalised: like the bench generator
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  auto h = [](unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
  };
  auto gauss = [&](unsigned long long key) {
    const unsigned long long a = h(key), b = h(key ^ 0x9e3779b97f4a7c15ULL);
    const float u1 = ((a >> 40) + 1.0f) / 16777217.0f, u2 = (b >> 40) / 16777216.0f;
    return sqrtf(-2.0f * logf(u1)) * cosf(6.2831853f * u2);
  };
  const unsigned cl = (unsigned)(h(r * 7919ULL + seed) % (unsigned)n_clusters);
  float s = 0.0f;
  for (int i = lane; i < d; i += 32) {
    const float x = gauss(((unsigned long long)cl << 20) + i + ((unsigned long long)seed << 44)) +
                    sigma * gauss(((unsigned long long)r << 10) + i + 0x5555ULL + ((unsigned long long)seed << 50));
    v[r * d + i] = x;
    s += x * x;
  }
  for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float inv = 1.0f / sqrtf(s);
  for (int i = lane; i < d; i += 32) v[r * d + i] *= inv;
}

static float chain_dot(const float* q, const float* v, int d) {
  volatile float acc = 0.0f;
  for (int i = 0; i < d; i++) {
    volatile float p = q[i] * v[i];
    acc = acc + p;
  }
  return acc;
}

static bool wait_stream(cudaStream_t s, double limit_s, const char* what) {
  const auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    cudaError_t e = cudaStreamQuery(s);
    if (e == cudaSuccess) return true;
    if (e != cudaErrorNotReady) { printf("%s: CUDA error %s\n", what, cudaGetErrorString(e)); exit(2); }
    if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > limit_s) {
      printf("%s: HUNG after %.0f s — giving up\n", what, limit_s);
      fflush(stdout);
      _Exit(3);
    }
  }
}

int main(int argc, char** argv) {
  const int nq = argc > 1 ? atoi(argv[1]) : 200;
  const long long N = argc > 2 ? atoll(argv[2]) : 5000;
  const int d = argc > 3 ? atoi(argv[3]) : 300;
  const int k = argc > 4 ? atoi(argv[4]) : 5;
  const bool perf = argc > 5 && !strcmp(argv[5], "perf");
  const int kch = (d + kPfBK - 1) / kPfBK, kpa = kch * kPfBK, ksteps = (d + 15) / 16;
  if (kch > kPfMaxKch) { printf("d too large\n"); return 1; }
  const int nq_pad = (nq + kPfBM - 1) / kPfBM * kPfBM, QT = nq_pad / kPfBM;
  const long long n_vt = (N + kPfBN - 1) / kPfBN, N_pad = n_vt * kPfBN;
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, 0));
  const int sms = prop.multiProcessorCount;
  printf("device %s, %d SMs; nq=%d N=%lld d=%d k=%d kch=%d ksteps=%d QT=%d vtiles=%lld\n", prop.name, sms, nq, N, d, k, kch, ksteps, QT, n_vt);
  cudaStream_t st;
  CK(cudaStreamCreate(&st));

  float *d_rows, *d_q, *d_vT, *d_eps2, *d_dump = nullptr;
  __nv_bfloat16 *d_vb, *d_qb;
  uint32_t *d_gthr, *d_norm;
  int32_t *d_cnt, *d_ids, *d_out_ids, *d_ovf_list, *d_ovf_cnt;
  float* d_out_sims;
  int2* d_cand;
  PfUnit* d_units;
  CK(cudaMalloc(&d_rows, (size_t)N * d * 4));
  CK(cudaMalloc(&d_q, (size_t)nq * d * 4));
  CK(cudaMalloc(&d_vT, (size_t)((N + 31) / 32) * 32 * d * 4));
  CK(cudaMalloc(&d_vb, (size_t)N_pad * kpa * 2));
  CK(cudaMalloc(&d_qb, (size_t)nq_pad * kpa * 2));
  CK(cudaMalloc(&d_eps2, nq_pad * 4));
  CK(cudaMalloc(&d_gthr, (size_t)nq_pad * kPfMaxK * 4));
  CK(cudaMalloc(&d_cnt, nq_pad * 4));
  CK(cudaMalloc(&d_norm, 4));
  CK(cudaMalloc(&d_cand, (size_t)nq_pad * kPfCandCap * sizeof(int2)));
  CK(cudaMalloc(&d_ids, (size_t)N * 4));
  CK(cudaMalloc(&d_out_ids, (size_t)nq * k * 4));
  CK(cudaMalloc(&d_out_sims, (size_t)nq * k * 4));
  CK(cudaMalloc(&d_ovf_list, nq_pad * 4));
  CK(cudaMalloc(&d_ovf_cnt, 4));
  CK(cudaMemsetAsync(d_vb, 0, (size_t)N_pad * kpa * 2, st));
  CK(cudaMemsetAsync(d_norm, 0, 4, st));
  CK(cudaMemsetAsync(d_ovf_cnt, 0, 4, st));
  fill_rows_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(d_rows, N, d, 1u, 1000, 1.0f);
  // queries: c - a + b of random rows would need a gather; use fresh vectors of the same mixture (norm 1)
  fill_rows_kernel<<<(nq + 7) / 8, 256, 0, st>>>(d_q, nq, d, 2u, 1000, 1.0f);
  {
    std::vector<int32_t> ids(N);
    for (long long i = 0; i < N; i++) ids[i] = (int32_t)(i + 1);
    CK(cudaMemcpyAsync(d_ids, ids.data(), (size_t)N * 4, cudaMemcpyHostToDevice, st));
    CK(cudaStreamSynchronize(st));
  }
  pf_rows_to_bf16_kernel<<<(unsigned)((N + 7) / 8), 256, 0, st>>>(d_rows, N, d, kpa, d_vb, d_norm);
  transpose_rows_kernel<<<(unsigned)((N + 31) / 32), 256, 0, st>>>(d_rows, N, d, d_vT);
  CK(cudaGetLastError());
  uint32_t nb = 0;
  CK(cudaMemcpyAsync(&nb, d_norm, 4, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  float n2;
  memcpy(&n2, &nb, 4);
  const float vmax = sqrtf(n2) * 1.000001f;
  printf("max row norm %.7f\n", vmax);

  CUtensorMap tm_q, tm_v;
  if (!pf_make_tensor_map(&tm_q, d_qb, nq_pad, kpa, kPfBM) || !pf_make_tensor_map(&tm_v, d_vb, N_pad, kpa, kPfBN)) {
    printf("cuTensorMapEncodeTiled failed\n");
    return 2;
  }
  std::vector<PfUnit> units(pf_max_units(QT, sms));
  int n_units = 0;
  pf_make_units(QT, (int)n_vt, sms, units.data(), &n_units);
  CK(cudaMalloc(&d_units, n_units * sizeof(PfUnit)));
  CK(cudaMemcpy(d_units, units.data(), n_units * sizeof(PfUnit), cudaMemcpyHostToDevice));
  printf("%d units; first: qt %d [%d,%d)\n", n_units, units[0].qt, units[0].v_begin, units[0].v_end);
  if (!perf) CK(cudaMalloc(&d_dump, (size_t)nq_pad * N_pad * 4));

  const int kk = k + 3;
  PfArgs a;
  memset(&a, 0, sizeof a);
  a.units = d_units; a.n_units = n_units; a.kch = kch; a.ksteps = ksteps; a.N = N; a.nq = nq; a.kk = kk;
  a.eps2 = d_eps2; a.gbest = d_gthr; a.cand_cnt = d_cnt; a.cand = d_cand; a.cap = kPfCandCap;
  a.dump = d_dump; a.dump_ld = N_pad;
  int* d_progress;
  CK(cudaMalloc(&d_progress, n_units * sizeof(int)));
  const int lockstep = getenv("PF_LOCKSTEP") ? atoi(getenv("PF_LOCKSTEP")) : 8;
  a.progress = d_progress; a.n_qt = QT; a.lockstep = (n_units <= sms && QT > 1 && lockstep > 0) ? (lockstep < 8 ? 8 : lockstep) : 0;
  a.dbg = getenv("PF_DBG") ? atoi(getenv("PF_DBG")) : 0;
  printf("lockstep window %d tiles, dbg %d\n", a.lockstep, a.dbg);
  CK(cudaFuncSetAttribute(prefilter_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PfSmem::total));
  const size_t rs_smem = (size_t)kPfCandCap * 8 + (size_t)d * 4;
  const int grid = std::min(sms, n_units);

  cudaEvent_t e0, e1, e2;
  CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&e2));
  const int reps = perf ? 5 : 1;
  for (int rep = 0; rep < reps; rep++) {
    pf_queries_prepare_kernel<<<(nq_pad + 7) / 8, 256, 0, st>>>(d_q, nq, nq_pad, d, kpa, vmax, d_qb, d_eps2, d_gthr, d_cnt);
    CK(cudaMemsetAsync(d_ovf_cnt, 0, 4, st));
    CK(cudaMemsetAsync(d_progress, 0, n_units * sizeof(int), st));
    CK(cudaEventRecord(e0, st));
    prefilter_gemm_kernel<<<grid, kPfThreads, PfSmem::total, st>>>(tm_q, tm_v, a);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e1, st));
    pf_rescore_kernel<<<nq, kPfRescoreThreads, rs_smem, st>>>(d_q, d, d_vT, d_cnt, d_cand, kPfCandCap, kk, d_eps2, nullptr, k, d_ids,
                                                             d_out_ids, d_out_sims, nullptr, d_ovf_list, d_ovf_cnt, nullptr);
    CK(cudaGetLastError());
    CK(cudaEventRecord(e2, st));
    wait_stream(st, 30.0, "prefilter + rescore");
    float ms_g = 0, ms_r = 0;
    CK(cudaEventElapsedTime(&ms_g, e0, e1));
    CK(cudaEventElapsedTime(&ms_r, e1, e2));
    const double flops = 2.0 * nq_pad * (double)N_pad * ksteps * 16;
    printf("rep %d: gemm %.3f ms (%.1f TFLOP/s bf16, %.2f TB/s of bf16 table per query tile), rescore %.3f ms\n", rep, ms_g,
           flops / ms_g / 1e9, (double)N_pad * kpa * 2 * QT / ms_g / 1e9, ms_r);
  }
  std::vector<int32_t> cnt(nq_pad);
  int ovf = 0;
  CK(cudaMemcpy(cnt.data(), d_cnt, nq_pad * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(&ovf, d_ovf_cnt, 4, cudaMemcpyDeviceToHost));
  long long tot = 0;
  int mx = 0;
  for (int q = 0; q < nq; q++) { tot += cnt[q]; mx = std::max(mx, cnt[q]); }
  printf("candidates per query: mean %.1f max %d; overflowed queries %d\n", (double)tot / nq, mx, ovf);
  std::vector<int32_t> ovf_list(nq_pad);
  CK(cudaMemcpy(ovf_list.data(), d_ovf_list, nq_pad * 4, cudaMemcpyDeviceToHost));
  std::vector<char> is_ovf(nq, 0);
  for (int i = 0; i < ovf; i++) is_ovf[ovf_list[i]] = 1;   // the engine re-does these with the fp32 scan

  // ---- checks against the CPU
  const int nq_chk = perf ? std::min(nq, 8) : nq;
  std::vector<float> hq((size_t)nq * d), hrows;
  CK(cudaMemcpy(hq.data(), d_q, (size_t)nq * d * 4, cudaMemcpyDeviceToHost));
  std::vector<int32_t> out_ids((size_t)nq * k);
  std::vector<float> out_sims((size_t)nq * k);
  CK(cudaMemcpy(out_ids.data(), d_out_ids, out_ids.size() * 4, cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(out_sims.data(), d_out_sims, out_sims.size() * 4, cudaMemcpyDeviceToHost));
  hrows.resize((size_t)N * d);
  CK(cudaMemcpy(hrows.data(), d_rows, (size_t)N * d * 4, cudaMemcpyDeviceToHost));
  std::vector<float> dump;
  if (!perf) {
    dump.resize((size_t)nq_pad * N_pad);
    CK(cudaMemcpy(dump.data(), d_dump, dump.size() * 4, cudaMemcpyDeviceToHost));
  }
  double worst_ratio = 0.0;
  int bad_topk = 0, bad_bound = 0;
  std::vector<std::pair<float, long long>> sc(N);
  for (int q = 0; q < nq_chk; q++) {
    double qn = 0;
    for (int i = 0; i < d; i++) qn += (double)hq[(size_t)q * d + i] * hq[(size_t)q * d + i];
    qn = sqrt(qn);
    for (long long r = 0; r < N; r++) {
      const float s = chain_dot(&hq[(size_t)q * d], &hrows[(size_t)r * d], d);
      sc[r] = {s, r};
      if (!perf) {
        const float ap = dump[(size_t)q * N_pad + r];
        const double ratio = fabs((double)ap - (double)s) / (qn * vmax);
        worst_ratio = std::max(worst_ratio, ratio);
        if (ratio > kPfEpsC) bad_bound++;
      }
    }
    std::partial_sort(sc.begin(), sc.begin() + k, sc.end(), [](const auto& x, const auto& y) {
      return x.first > y.first || (x.first == y.first && x.second < y.second);
    });
    for (int p = 0; p < k && !is_ovf[q]; p++) {
      const bool ok = out_ids[(size_t)q * k + p] == (int32_t)(sc[p].second + 1) &&
                      memcmp(&out_sims[(size_t)q * k + p], &sc[p].first, 4) == 0;
      if (!ok) {
        if (bad_topk < 10)
          printf("MISMATCH q=%d p=%d: gpu (%d, %.9g) cpu (%lld, %.9g)\n", q, p, out_ids[(size_t)q * k + p], out_sims[(size_t)q * k + p],
                 sc[p].second + 1, sc[p].first);
        bad_topk++;
      }
    }
  }
  if (!perf) printf("worst |approx - chain| / (|q| |v|max) = %.3e   (bound c = %.3e), violations %d\n", worst_ratio, (double)kPfEpsC, bad_bound);
  printf("top-%d of %d queries vs CPU fp32 chain: %d mismatching entries\n", k, nq_chk, bad_topk);
  printf(bad_topk == 0 && bad_bound == 0 ? "PROBE OK\n" : "PROBE FAILED\n");
  return (bad_topk == 0 && bad_bound == 0) ? 0 : 1;
}
