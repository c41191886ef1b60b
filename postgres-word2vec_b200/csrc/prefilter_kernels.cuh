// prefilter_kernels.cuh — tensor-core pre-filter + exact re-score for the one dense contraction of the
// path: the exact cosine scan behind k_nearest_neighbour (freddy--0.0.1.sql:426-439) and analogy_3cosadd
// (:1270-1288), score[q][r] = cosine_similarity_bytea(q, v_r) = sequential fp32 dot (core_functions.c:67-81).
//
// The reference's value is a chain of 2*d individually rounded fp32 operations, which tensor cores cannot
// reproduce.  They do not have to: what is returned is the top-k of the scores, so
//
//   1. prefilter_gemm_kernel  computes APPROXIMATE scores a[q][r] with tcgen05.mma (bf16 operands staged by
//      TMA into 128-byte-swizzled shared memory, fp32 accumulators in TMEM) and emits every (q, r) whose
//      approximate score is within 2*eps_q of a running lower bound of the k-th best approximate score;
//   2. pf_rescore_kernel      re-scores the emitted candidates with the reference's own fp32 chain and
//      selects the top-k by (score desc, table row asc) — the same keys the fp32 scan kernels use.
//
// Error bound (why the result is identical).  Let s = the reference's fp32 chain, a = the tensor-core value,
// x = the real-number dot product.  bf16 round-to-nearest has unit roundoff 2^-8, so each product carries a
// relative error <= 2^-7 + 2^-16; bf16 x bf16 products are exact in fp32; the accumulation inside the tensor
// core and the fp32 chain's own 2*d roundings are each bounded by a few d * 2^-24 (relative to sum|q_i v_i|).
// With c = 2^-7 + 2^-11 (the 2^-11 covers those second-order terms with a factor > 10 to spare) and
// Cauchy-Schwarz:  |a - s| <= eps_q = c * ||q||_2 * max_r ||v_r||_2.
// If a_(k) is the k-th largest approximate score, k rows have s >= a_(k) - eps, hence s_(k) >= a_(k) - eps,
// and every row with s >= s_(k) (all winners and all their ties) has a >= a_(k) - 2 eps: it is emitted.
// The running bound only rises towards a_(k), so emitting against it only adds candidates.  Queries whose
// candidate buffer overflows (heavy duplication) are re-done by the fp32 scan kernels: the result never
// depends on the pre-filter.  The measured |a - s| is checked against eps in tests/ and scripts/umma_probe.cu.
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "common.cuh"
#include "vector_kernels.cuh"

namespace fb {

constexpr int kPfBM = 128;          // queries per tile  = TMEM lanes   (UMMA M)
constexpr int kPfBN = 256;          // table rows per tile = TMEM columns of one accumulator (UMMA N)
constexpr int kPfBK = 64;           // bf16 elements per K chunk = 128 bytes = one swizzle-atom row
constexpr int kPfStages = 4;        // table-tile ring depth
constexpr int kPfMaxKch = 5;        // K chunks held for the query tile (d <= 320)
constexpr int kPfEpiWarps = 8;      // two warps per TMEM lane quarter, each draining half of an accumulator's columns
constexpr int kPfThreads = 64 + 32 * kPfEpiWarps;   // warp 0: TMA producer, warp 1: MMA issuer + TMEM owner, then the epilogue warps
constexpr int kPfCandCap = 4096;    // candidates buffered per query while the bound is still rising (S slabs start cold:
                                    // about S * k * ln(rows per slab / k) emissions); more = overflow -> fp32 scan
constexpr int kPfMaxK = 40;         // k' (k + excluded rows) the epilogue tracks
constexpr int kPfSlotBatch = 4;     // candidate slots a thread reserves per atomic
constexpr uint32_t kPfQChunkBytes = kPfBM * 128;        // 16 KB
constexpr uint32_t kPfVStageBytes = kPfBN * 128;        // 32 KB
constexpr float kPfEpsC = 0.0078125f + 0.00048828125f;  // 2^-7 + 2^-11

struct PfUnit { int qt, v_begin, v_end, pad; };   // one CTA-unit: query tile x range of table tiles

struct PfArgs {
  const PfUnit* units;
  int n_units;
  int kch;                 // K chunks of 64
  int ksteps;              // UMMA K steps of 16 covering d (ceil(d / 16))
  long long N;             // table rows
  int nq;                  // queries
  int kk;                  // k' <= kPfMaxK
  const float* eps2;       // [nq_pad] 2 * eps_q
  uint32_t* gbest;         // [nq_pad][kPfMaxK] per query: kk approximate scores of kk distinct rows seen so far by ANY
                           // CTA (ordered-float encoding, 0 = empty); their minimum is a lower bound of a_(kk)
  int32_t* cand_cnt;       // [nq_pad]
  int2* cand;              // [nq_pad][cap] (table row, approximate score bits)
  int cap;
  float* dump;             // debugging / tests: [nq_pad][dump_ld] every approximate score, or nullptr
  long long dump_ld;
  // lock-step of the CTAs that stream the same slab (one per query tile): progress[u] = table tiles unit u has
  // requested; a CTA does not run more than `lockstep` tiles ahead of the slowest sibling, so a tile fetched from
  // HBM by the first CTA is still in L2 when the others ask for it.  0 = off (units are not all co-resident).
  int* progress;
  int n_qt;
  int lockstep;
  int dbg;                 // timing experiments only (results invalid): 1 = epilogue drains nothing, 2 = no MMAs issued, 4 = no TMA loads
};

// shared-memory plan of prefilter_gemm_kernel (dynamic, base aligned to 1024 bytes by the kernel)
struct PfSmem {
  static constexpr uint32_t off_q = 0;
  static constexpr uint32_t off_v = kPfMaxKch * kPfQChunkBytes;                  // 80 KB
  static constexpr uint32_t off_bar = off_v + kPfStages * kPfVStageBytes;        // + 128 KB
  static constexpr uint32_t total = off_bar + 256 + 1024;                        // + barriers + alignment slack
};

__device__ __forceinline__ uint32_t ordered_u32(float f) {
  uint32_t b = __float_as_uint(f);
  return b ^ ((b & 0x80000000u) ? 0xFFFFFFFFu : 0x80000000u);
}
__device__ __forceinline__ float unordered_f32(uint32_t b) {
  return __uint_as_float(b ^ ((b & 0x80000000u) ? 0x80000000u : 0xFFFFFFFFu));
}

// ---- tcgen05 / TMA wrappers (inline PTX; SASS: UTCHMMA, LDTM, UTMALDG, UTCBAR) ----
__device__ __forceinline__ void tma_load_2d(uint32_t smem_dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(smem_dst),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* tm) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tm)) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// Warp-converged issue: the whole warp runs the role's loops with warp-uniform values (the compiler keeps descriptors,
// barrier addresses and counters in uniform registers), and the one-thread instructions are predicated on elect.sync
// INSIDE the asm, so there is no divergent branch around them.  (With `if (lane == 0)` around the loops the issue of
// one UTCHMMA took ~250 cycles of dependent R2UR / uniform-ALU work against the 128 cycles the MMA runs: ncu r2.)
__device__ __forceinline__ void tma_load_2d_elect(uint32_t smem_dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar, uint32_t tx_bytes) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%4], %5;\n\t"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n\t}"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar), "r"(tx_bytes)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_elect_noexpect(uint32_t smem_dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];\n\t}"
      ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void mbar_expect_tx_elect(uint32_t bar, uint32_t bytes) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n\t}" ::"r"(bar), "r"(bytes)
      : "memory");
}
__device__ __forceinline__ void tc_commit_elect(uint32_t bar) {
  asm volatile(
      "{\n\t.reg .pred q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "@q tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n\t}" ::"r"(bar)
      : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_elect(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p, q;\n\t"
      "elect.sync _|q, 0xffffffff;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "@q tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, bf16 x bf16 -> fp32, both operands K-major
__device__ __forceinline__ void tc_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// K-major operand tile in 128-byte-swizzled shared memory: rows of 128 bytes, 8-row atoms of 1024 bytes
// (stride byte offset), start address advanced by 32 bytes per UMMA K step inside the atom row.
__device__ __forceinline__ uint64_t umma_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);          // start address, bits [0,14)
  d |= (uint64_t)0 << 16;                                // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024u >> 4) << 32;                     // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                                // descriptor version (sm_100)
  d |= (uint64_t)2 << 61;                                // layout type: SWIZZLE_128B
  return d;
}
// instruction descriptor: D fp32 (bit 4), A and B bf16 (bits 7, 10), both K-major (bits 15, 16 = 0), N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t kPfIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(kPfBN >> 3) << 17) | ((uint32_t)(kPfBM >> 4) << 24);
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int j = 0; j < 32; j++) v[j] = __uint_as_float(r[j]);
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, uint32_t parity) {
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  }
}
__device__ __forceinline__ void mbar_init_u32(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx_u32(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive_u32(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

// ---------------------------------------------------------------------------------------------------
// The pre-filter.  Persistent: one CTA per SM walks its units; a unit = one tile of 128 queries (A operand,
// resident in shared memory) x a contiguous range of 256-row table tiles (B operand, streamed through a
// 4-stage TMA ring, one 64-wide K chunk per stage).  Per table tile the MMA thread issues `ksteps`
// tcgen05.mma (M=128, N=256, K=16) into one of two 256-column TMEM accumulators; the four epilogue warps
// drain the other one: thread = one query (TMEM lane), 32 scores per tcgen05.ld, one max + compare per 32
// scores on the fast path.
// ---------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kPfThreads, 1)
prefilter_gemm_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_v, const PfArgs a) {
  extern __shared__ unsigned char pf_smem_raw[];
  const uint32_t base = (smem_u32(pf_smem_raw) + 1023u) & ~1023u;
  const uint32_t q_smem = base + PfSmem::off_q, v_smem = base + PfSmem::off_v, bars = base + PfSmem::off_bar;
  // barrier map (8 bytes each): full[4] empty[4] q_full q_empty tmem_full[2] tmem_empty[2]; then the TMEM base slot
  auto bar_full = [&](int s) { return bars + 8u * s; };
  auto bar_empty = [&](int s) { return bars + 8u * (kPfStages + s); };
  const uint32_t bar_q_full = bars + 8u * (2 * kPfStages), bar_q_empty = bar_q_full + 8u;
  auto bar_t_full = [&](int s) { return bars + 8u * (2 * kPfStages + 2 + s); };
  auto bar_t_empty = [&](int s) { return bars + 8u * (2 * kPfStages + 4 + s); };
  const uint32_t tmem_slot = bars + 8u * (2 * kPfStages + 6);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  if (tid == 0) {
    for (int s = 0; s < kPfStages; s++) { mbar_init_u32(bar_full(s), 1); mbar_init_u32(bar_empty(s), 1); }
    mbar_init_u32(bar_q_full, 1);
    mbar_init_u32(bar_q_empty, 1);
    for (int s = 0; s < 2; s++) { mbar_init_u32(bar_t_full(s), 1); mbar_init_u32(bar_t_empty(s), kPfEpiWarps); }
    mbar_fence_init();
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_v);
  }
  if (warp == 1) {   // the whole TMEM (2 accumulators x 256 columns); one CTA per SM, so nobody else wants it
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(tmem_slot) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    // =========================================================== TMA producer (whole warp, one elected lane issues)
    int stage = 0;
    uint32_t phase = 0;
    int iter = 0;
    for (int u = blockIdx.x; u < a.n_units; u += gridDim.x, iter++) {
      const PfUnit un = a.units[u];
      if (iter > 0) mbar_wait_u32(bar_q_empty, (uint32_t)((iter - 1) & 1));   // MMAs of the previous unit are done with Q
      mbar_expect_tx_elect(bar_q_full, (uint32_t)a.kch * kPfQChunkBytes);
      for (int kc = 0; kc < a.kch; kc++)
        tma_load_2d_elect_noexpect(q_smem + kc * kPfQChunkBytes, &tm_q, kc * kPfBK, un.qt * kPfBM, bar_q_full);
      const int sib0 = (u / a.n_qt) * a.n_qt;                      // units sib0 .. sib0 + n_qt - 1 stream this slab
      for (int vt = un.v_begin; vt < un.v_end; vt++) {
        if (a.lockstep > 0) {
          const int t = vt - un.v_begin;
          if (t >= a.lockstep && (t & 3) == 0) {                  // a rate limiter, not a protocol: relaxed loads, every 4th tile
            for (;;) {
              int slowest = 0x7fffffff;
              const volatile int* pr = a.progress + sib0;
              for (int j = 0; j < a.n_qt; j++) { const int pj = pr[j]; slowest = pj < slowest ? pj : slowest; }
              if (slowest >= t - a.lockstep) break;
              __nanosleep(100);
            }
          }
          if ((t & 3) == 3 && lane == 0) *reinterpret_cast<volatile int*>(a.progress + u) = t + 1;
        }
        for (int kc = 0; kc < a.kch; kc++) {
          mbar_wait_u32(bar_empty(stage), phase ^ 1u);
          if (a.dbg & 4) { if (lane == 0) mbar_arrive_u32(bar_full(stage)); }
          else tma_load_2d_elect(v_smem + stage * kPfVStageBytes, &tm_v, kc * kPfBK, vt * kPfBN, bar_full(stage), kPfVStageBytes);
          if (++stage == kPfStages) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // =========================================================== MMA issuer (whole warp, one elected lane issues)
    constexpr uint32_t idesc = kPfIdesc;
    const uint64_t a_desc0 = umma_desc_sw128(q_smem), b_desc0 = umma_desc_sw128(v_smem);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    int iter = 0;
    for (int u = blockIdx.x; u < a.n_units; u += gridDim.x, iter++) {
      const PfUnit un = a.units[u];
      mbar_wait_u32(bar_q_full, (uint32_t)(iter & 1));
      tc_fence_after();
      for (int vt = un.v_begin; vt < un.v_end; vt++) {
        mbar_wait_u32(bar_t_empty(acc), acc_phase ^ 1u);   // the epilogue drained this accumulator
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + (uint32_t)acc * kPfBN;
        int ks_left = a.ksteps;
        for (int kc = 0; kc < a.kch; kc++) {
          mbar_wait_u32(bar_full(stage), phase);
          tc_fence_after();
          // descriptors differ from the base ones only in the start-address field (bytes >> 4, no carry out of it)
          const uint64_t ad = a_desc0 + (uint64_t)((kc * kPfQChunkBytes) >> 4);
          const uint64_t bd = b_desc0 + (uint64_t)((stage * kPfVStageBytes) >> 4);
#pragma unroll
          for (int k = 0; k < 4; k++)
            if (k < ks_left && !(a.dbg & 2)) tc_mma_bf16_elect(d_tmem, ad + 2u * k, bd + 2u * k, idesc, (kc | k) != 0 ? 1u : 0u);
          ks_left -= 4;
          tc_commit_elect(bar_empty(stage));          // the ring slot is free once these MMAs have read it
          if (++stage == kPfStages) { stage = 0; phase ^= 1u; }
        }
        tc_commit_elect(bar_t_full(acc));             // accumulator complete
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
      tc_commit_elect(bar_q_empty);
    }
  } else {
    // =========================================================== epilogue warps (thread = query)
    // The bound a thread emits against is min(gbest[q][0..kk)): kk scores of distinct rows, shared by every CTA
    // that works on query q (all slabs).  A row better than that minimum replaces it (compare-and-swap on the
    // slot), so the bound follows the kk-th best of ALL rows seen so far and the number of emissions per query
    // stays near kk * ln(N / kk) however many slabs there are.
    // (ncu, 4 epilogue warps: the MMA thread waited for a drained accumulator 20 % of the time and the tensor pipe
    // was busy 37 %: the drain, not TMA, set the pace.  Hence 8 warps, a max tree, and a bound refreshed every 4th tile.)
    const int quarter = warp & 3;                                  // TMEM lanes this warp may read
    const int chalf = (warp - 2) >> 2;                             // which half of the accumulator's columns it drains
    constexpr int kChunks = kPfBN / 32 / (kPfEpiWarps / 4);
    const int qlane = quarter * 32 + lane;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quarter * 32) << 16);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int u = blockIdx.x; u < a.n_units; u += gridDim.x) {
      const PfUnit un = a.units[u];
      const int q = un.qt * kPfBM + qlane;
      const bool live = q < a.nq;
      const float eps2 = live ? a.eps2[q] : 0.0f;
      uint32_t* gb = a.gbest + (size_t)(live ? q : 0) * kPfMaxK;
      // (min value, its slot) of the query's kk shared scores.  All loads are issued before the first compare: one L2
      // round trip per call instead of kk dependent ones (measured: the kk-proportional cost of this lookup was
      // 0.25 ms per unit of kk on the bench table).
      auto bound = [&]() {
        uint4 t[kPfMaxK / 4];
        const int nv = (a.kk + 3) >> 2;
#pragma unroll
        for (int i = 0; i < kPfMaxK / 4; i++)
          if (i < nv) t[i] = __ldcg(reinterpret_cast<const uint4*>(gb) + i);
        uint32_t mv = 0xFFFFFFFFu;
        int mi = 0;
#pragma unroll
        for (int i = 0; i < kPfMaxK / 4; i++) {
          if (i < nv) {
            const uint32_t x[4] = {t[i].x, t[i].y, t[i].z, t[i].w};
#pragma unroll
            for (int j = 0; j < 4; j++)
              if (4 * i + j < a.kk && x[j] < mv) { mv = x[j]; mi = 4 * i + j; }
          }
        }
        return make_uint2(mv, (uint32_t)mi);
      };
      int slot_base = 0, slot_left = 0;                            // candidate slots are reserved kPfSlotBatch at a time
      uint32_t gmin = 0u;
      for (int vt = un.v_begin; vt < un.v_end; vt++) {
        // the shared bound only rises: a stale copy just emits a few more candidates, so it is re-read every 4th tile
        if (((vt - un.v_begin) & 3) == 0) gmin = live ? bound().x : 0xFFFFFFFFu;
        mbar_wait_u32(bar_t_full(acc), acc_phase);
        tc_fence_after();
        float thr_emit = !live ? INFINITY : (gmin == 0u ? -INFINITY : unordered_f32(gmin) - eps2);
        const long long row0 = (long long)vt * kPfBN;
#pragma unroll 1
        for (int c = chalf * kChunks; c < (chalf + 1) * kChunks && !(a.dbg & 1); c++) {
          float v[32];
          tmem_ld32(t_lane + (uint32_t)acc * kPfBN + (uint32_t)(c * 32), v);
          float m8[8];                                             // max tree: 4 levels instead of a 31-deep chain
#pragma unroll
          for (int j = 0; j < 8; j++) m8[j] = fmaxf(fmaxf(v[4 * j], v[4 * j + 1]), fmaxf(v[4 * j + 2], v[4 * j + 3]));
          const float mx = fmaxf(fmaxf(fmaxf(m8[0], m8[1]), fmaxf(m8[2], m8[3])), fmaxf(fmaxf(m8[4], m8[5]), fmaxf(m8[6], m8[7])));
          if (mx >= thr_emit || a.dump != nullptr) {
            // Rare path, kept SMALL on purpose: a 32-times unrolled body (13 k SASS lines) thrashed the instruction cache
            // and cost as much as the whole GEMM.  The 32 scores go to a local array, a bit mask names the ones to look at.
            float vv[32];
            uint32_t todo = 0u;
#pragma unroll
            for (int j = 0; j < 32; j++) { vv[j] = v[j]; todo |= (v[j] >= thr_emit ? 1u : 0u) << j; }
            if (a.dump != nullptr && live) {
#pragma unroll 1
              for (int j = 0; j < 32; j++) {
                const long long row = row0 + c * 32 + j;
                if (row < a.N) a.dump[(long long)q * a.dump_ld + row] = vv[j];
              }
            }
#pragma unroll 1
            while (todo) {
              const int j = __ffs(todo) - 1;
              todo &= todo - 1;
              const float s = vv[j];
              const long long row = row0 + c * 32 + j;
              if (s >= thr_emit && row < a.N) {
                if (slot_left == 0) {                              // one atomic per kPfSlotBatch emissions; unused slots stay (-1, 0)
                  slot_base = atomicAdd(a.cand_cnt + q, kPfSlotBatch);
                  slot_left = kPfSlotBatch;
#pragma unroll
                  for (int z = 0; z < kPfSlotBatch; z++)
                    if (slot_base + z < a.cap) a.cand[(size_t)q * a.cap + slot_base + z] = make_int2(-1, 0);
                }
                const int slot = slot_base + (kPfSlotBatch - slot_left);
                slot_left--;
                if (slot < a.cap) a.cand[(size_t)q * a.cap + slot] = make_int2((int)row, __float_as_int(s));
                const uint32_t os = ordered_u32(s);
                if (os > gmin) {
                  for (;;) {                                       // replace the current minimum with this row's score
                    const uint2 mb = bound();
                    if (os <= mb.x) break;
                    if (atomicCAS(gb + mb.y, mb.x, os) == mb.x) break;
                  }
                  gmin = bound().x;
                  thr_emit = gmin == 0u ? -INFINITY : unordered_f32(gmin) - eps2;
                }
              }
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_u32(bar_t_empty(acc));
        if (++acc == 2) { acc = 0; acc_phase ^= 1u; }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---------------------------------------------------------------------------------------------------
// operand preparation
// ---------------------------------------------------------------------------------------------------
// rows [n][d] fp32 (row-major staging slice) -> bf16 [n][kpa] (zero padded), and the largest squared row norm
__global__ void pf_rows_to_bf16_kernel(const float* __restrict__ rows, long long n, int d, int kpa,
                                       __nv_bfloat16* __restrict__ out, uint32_t* __restrict__ max_norm2_bits) {
  const long long r = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n) return;
  double s = 0.0;
  for (int i = lane; i < kpa; i += 32) {
    const float v = (i < d) ? rows[r * d + i] : 0.0f;
    out[r * kpa + i] = __float2bfloat16_rn(v);
    s += (double)v * (double)v;
  }
  for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    float f = (float)s;
    if (!(f == f)) f = INFINITY;                      // NaN rows disable the pre-filter (host checks for inf)
    f = __uint_as_float(__float_as_uint(f) + 1u);     // round up (f >= 0)
    atomicMax(max_norm2_bits, __float_as_uint(f));
  }
}

// queries [nq][d] fp32 -> bf16 [nq_pad][kpa]; eps2[q] = 2 * c * ||q|| * vmax (rounded up); resets the per-query state
__global__ void pf_queries_prepare_kernel(const float* __restrict__ q, int nq, int nq_pad, int d, int kpa, float vmax,
                                          __nv_bfloat16* __restrict__ out, float* __restrict__ eps2,
                                          uint32_t* __restrict__ gbest, int32_t* __restrict__ cand_cnt) {
  const int r = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= nq_pad) return;
  double s = 0.0;
  for (int i = lane; i < kpa; i += 32) {
    const float v = (r < nq && i < d) ? q[(size_t)r * d + i] : 0.0f;
    out[(size_t)r * kpa + i] = __float2bfloat16_rn(v);
    s += (double)v * (double)v;
  }
  for (int o = 16; o >= 1; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) {
    const double e = 2.0 * (double)kPfEpsC * sqrt(s) * (double)vmax * 1.001 + 1e-30;
    eps2[r] = (float)e * 1.0000002f;                  // inf / NaN queries: nothing is emitted, the host re-does them
    cand_cnt[r] = 0;
  }
  for (int i = lane; i < kPfMaxK; i += 32) gbest[(size_t)r * kPfMaxK + i] = 0u;
}

// ---------------------------------------------------------------------------------------------------
// exact re-score of the candidates + selection: one CTA per query.
//   candidates sorted by approximate score; only those within 2 eps of the kk-th best approximate score can
//   hold a winner (see the header); their exact scores are the reference's fp32 chain over the
//   dimension-major table; keys (score desc, row asc); first k written.  exclude: up to three table rows
//   per query that never win (analogy), or nullptr.
// ---------------------------------------------------------------------------------------------------
constexpr int kPfRescoreThreads = 256;

__device__ __forceinline__ void pf_block_sort_desc(u64* s, int n_pad) {
  for (int size = 2; size <= n_pad; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int i = threadIdx.x; i < n_pad / 2; i += kPfRescoreThreads) {
        const int lo = 2 * i - (i & (stride - 1)), hi = lo + stride;
        const bool desc = (lo & size) == 0;
        const u64 x = s[lo], y = s[hi];
        if ((x < y) == desc) { s[lo] = y; s[hi] = x; }
      }
      __syncthreads();
    }
  }
}

__global__ void __launch_bounds__(kPfRescoreThreads)
pf_rescore_kernel(const float* __restrict__ queries, int d, const float* __restrict__ vR,
                  const int32_t* __restrict__ cand_cnt, const int2* __restrict__ cand, int cap, int kk,
                  const float* __restrict__ eps2, const int32_t* __restrict__ exclude_rows, int k,
                  const int32_t* __restrict__ ids, int32_t* __restrict__ out_ids, float* __restrict__ out_sims,
                  int32_t* __restrict__ out_rows,             // optional: winner's table row (k == 1 users), or nullptr
                  int32_t* __restrict__ ovf_list, int32_t* __restrict__ ovf_count, const float* __restrict__ q_ok_norm) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  u64* keys = reinterpret_cast<u64*>(smem_raw);                    // [cap]
  float* qs = reinterpret_cast<float*>(keys + cap);                // [d]
  __shared__ int s_ns;
  const int q = blockIdx.x, tid = threadIdx.x;
  const int n = cand_cnt[q];
  const float e2 = eps2[q];
  if (n > cap || !(e2 < INFINITY)) {                               // overflow, or a non-finite query: the fp32 scan re-does it
    if (tid == 0) ovf_list[atomicAdd(ovf_count, 1)] = q;
    return;
  }
  (void)q_ok_norm;
  for (int i = tid; i < d; i += kPfRescoreThreads) qs[i] = queries[(size_t)q * d + i];
  int n_pad = 32;
  while (n_pad < n) n_pad <<= 1;
  for (int i = tid; i < n_pad; i += kPfRescoreThreads) {
    u64 key = 0ull;                                                // reserved but unused slots (row -1) sort to the end
    if (i < n) {
      const int2 c = cand[(size_t)q * cap + i];
      if (c.x >= 0) key = ((u64)ordered_u32(__int_as_float(c.y)) << 32) | (u64)(0xFFFFFFFFu - (uint32_t)c.x);
    }
    keys[i] = key;
  }
  if (tid == 0) s_ns = 0;
  __syncthreads();
  pf_block_sort_desc(keys, n_pad);
  // survivors: a >= a_(kk) - 2 eps  (a prefix of the sorted list)
  float cut = -INFINITY;
  if (n >= kk && keys[kk - 1] != 0ull) cut = unordered_f32((uint32_t)(keys[kk - 1] >> 32)) - e2;
  for (int i = tid; i < n; i += kPfRescoreThreads) {
    const bool in = keys[i] != 0ull && unordered_f32((uint32_t)(keys[i] >> 32)) >= cut;
    const bool next_in = (i + 1 < n) && keys[i + 1] != 0ull && (unordered_f32((uint32_t)(keys[i + 1] >> 32)) >= cut);
    if (in && !next_in) s_ns = i + 1;
  }
  __syncthreads();
  const int ns = s_ns;
  int ex0 = -1, ex1 = -1, ex2 = -1;
  if (exclude_rows != nullptr) { ex0 = exclude_rows[3 * q]; ex1 = exclude_rows[3 * q + 1]; ex2 = exclude_rows[3 * q + 2]; }
  int ns_pad = 32;
  while (ns_pad < ns) ns_pad <<= 1;
  for (int i = tid; i < ns_pad; i += kPfRescoreThreads) {
    u64 key = 0ull;
    if (i < ns) {
      const int row = (int)(0xFFFFFFFFu - (uint32_t)keys[i]);
      if (row != ex0 && row != ex1 && row != ex2) {
        const float* vp = vR + (size_t)row * d;                  // row-major fp32 image
        float acc = 0.0f;
        for (int j = 0; j < d; j++) acc = xadd(acc, xmul(qs[j], __ldg(vp + j)));
        key = score_key(acc, (uint32_t)row);
      }
    }
    keys[i] = key;             // slot i is read and rewritten by this thread only
  }
  __syncthreads();
  pf_block_sort_desc(keys, ns_pad);
  for (int p = tid; p < k; p += kPfRescoreThreads) {
    const u64 key = (p < ns_pad) ? keys[p] : 0ull;
    out_ids[(size_t)q * k + p] = key ? ids[key_row(key)] : -1;
    out_sims[(size_t)q * k + p] = key ? key_score(key) : 0.0f;
    if (out_rows != nullptr) out_rows[(size_t)q * k + p] = key ? (int32_t)key_row(key) : -1;
  }
}

// ---------------------------------------------------------------------------------------------------
// host helpers
// ---------------------------------------------------------------------------------------------------
typedef CUresult (*PfEncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

inline PfEncodeTiledFn pf_encode_fn() {
  static PfEncodeTiledFn fn = nullptr;
  if (fn == nullptr) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PfEncodeTiledFn>(p);
  }
  return fn;
}

// bf16 [rows][kpa] row-major, box = 64 columns x box_rows, 128-byte swizzle
inline bool pf_make_tensor_map(CUtensorMap* tm, const void* base, long long rows, int kpa, int box_rows) {
  PfEncodeTiledFn fn = pf_encode_fn();
  if (fn == nullptr) return false;
  const cuuint64_t dims[2] = {(cuuint64_t)kpa, (cuuint64_t)rows};
  const cuuint64_t strides[1] = {(cuuint64_t)kpa * sizeof(__nv_bfloat16)};
  const cuuint32_t box[2] = {(cuuint32_t)kPfBK, (cuuint32_t)box_rows};
  const cuuint32_t estr[2] = {1, 1};
  return fn(tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// Split QT query tiles x n_vtiles table tiles into CTA units for `ctas` persistent CTAs.  Every query tile gets
// the SAME slab boundaries and the units of one slab are adjacent (u = slab * QT + qt): the QT CTAs that stream
// the same table tiles run side by side, so a tile is fetched from HBM once and served from L2 to the others.
// A slab keeps at least kPfMinSlabTiles tiles: every slab starts with a cold threshold and emits about
// k * ln(rows / k) candidates before it has warmed up.
constexpr int kPfMinSlabTiles = 32;
inline void pf_make_units(int QT, int n_vtiles, int ctas, PfUnit* out, int* n_out) {
  int n = 0;
  int slabs = QT >= ctas ? 1 : ctas / QT;
  const int max_slabs = (n_vtiles + kPfMinSlabTiles - 1) / kPfMinSlabTiles;
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  for (int s = 0; s < slabs; s++) {
    const int b = (int)((long long)n_vtiles * s / slabs), e = (int)((long long)n_vtiles * (s + 1) / slabs);
    if (e <= b) continue;
    for (int qt = 0; qt < QT; qt++) out[n++] = PfUnit{qt, b, e, 0};
  }
  *n_out = n;
}
inline int pf_max_units(int QT, int ctas) { return QT >= ctas ? QT : ctas; }

}  // namespace fb
