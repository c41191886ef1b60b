// ivfadc_kernels.cuh — the IVFADC / PQ hot path as sm_100a kernels.
//
// Pipeline for a chunk of queries (reference: freddy.c:247-378, ivfadc_search):
//   coarse_select_kernel   HOT(1) coarse L2 distances + top-w lists   freddy.c:264-283
//   lut_build_kernel       HOT(2) residuals + per-(query,probe) LUTs  freddy.c:295-314, index_utils.c:445-455
//   adc_scan_kernel        HOT(3)+(4) ADC over the probed lists' codes + per-warp top-(k+1)   freddy.c:347-372
//   finalize_kernel        merge per-warp lists, reference tie order, flag rare cases
//   (exact_kernels.cuh)    general path for flagged queries (boundary ties, re-probe loop)
//
// Exactness: every distance is the reference's own fp32 chain (common.cuh).
// Selection uses keys (distance bits, arrival order); the reference's
// order-dependent insertion (index_utils.c:19-33) is reproduced exactly when no
// tie straddles the k-th place, which finalize_kernel checks; the rest goes to
// the general kernel.
#pragma once
#include "common.cuh"

namespace fb {

// ---------------------------------------------------------------------------
// Device image of one code table (fine_quantization or pq_quantization).
// Rows are grouped by inverted list, each list padded to a multiple of 32 rows
// and stored in blocks of 32 rows x U units; a unit is 4 codes = 8 bytes, codes
// are pre-scaled by 4 (byte offset into one K-float LUT row):
//   unit u of lane L of block b  ->  units[(b*U + u)*32 + L]
// so a warp reads one block with U fully coalesced 256-byte loads.
// rowno[b*32 + L] is the row's position in the uploaded table (arrival order of
// the reference's loop), -1 for padding.  Algorithmic bytes per row: 2*m + 4.
// ---------------------------------------------------------------------------
struct CodeTableDev {
  const uint4* units8;      // [n_blocks][32] one byte per code, 16 bytes per row (tables with K <= 256, m <= 16), or nullptr:
                            // a warp reads a 32-row block with ONE coalesced 512-byte load; algorithmic bytes per row m + 4
  const uint2* units;       // [n_blocks][U][32]
  const int32_t* rowno;     // [n_blocks][32]
  const int32_t* list_blk;  // [n_lists] first block of each list
  const int32_t* list_len;  // [n_lists] rows in each list
  const int32_t* ids;       // [N] id by table row number
  int m, U, n_lists;
};

constexpr uint32_t kFlagExact = 1u;     // needs the general kernel
// why (statistics only)
constexpr uint32_t kWhyCoarseTie = 2u, kWhyCoarseFar = 4u, kWhyFewRows = 8u, kWhyScanTie = 16u, kWhyForced = 32u;
constexpr int kCoarseThreads = 512;
constexpr int kScanThreads = 256;
constexpr int kScanWarps = kScanThreads / kWarp;
constexpr int kLutMaxJobs = 16;

// Rare path of coarse_select_warp, kept out of line so that it costs the selection loop no registers: replays the
// reference's loop over S (see below); true = S may be incomplete, the general kernel has to decide.
__device__ __noinline__ bool coarse_tie_replay(u64& mine, u64 kw1, int w, int lane) {
  const bool in_s = mine != kKeyInf && key_dbits(mine) <= key_dbits(kw1);
  const unsigned smask = __ballot_sync(0xffffffffu, in_s);
  if (((smask >> 31) & 1u) || key_dist(kw1) >= 100.0f) return true;    // (or the sentinel quirk applies)
  const uint32_t my_id = key_t(mine);
  int rank = 0;                                             // arrival position of my entry inside S
  for (int j = 0; j < 32; j++) {
    const uint32_t oid = __shfl_sync(0xffffffffu, my_id, j);
    rank += (int)(((smask >> j) & 1u) && oid < my_id);
  }
  const int n_s = __popc(smask);
  float td = 100.0f;                                        // cqSelection[lane] (freddy.c:266-269)
  int tid = -1;
  for (int r = 0; r < n_s; r++) {
    const int src = __ffs(__ballot_sync(0xffffffffu, in_s && rank == r)) - 1;
    const float dd = key_dist(shfl_u64(mine, src));
    const int id = (int)__shfl_sync(0xffffffffu, my_id, src);
    const float worst = __shfl_sync(0xffffffffu, td, w - 1);
    if (dd < worst) {                                       // gate of freddy.c:278 (minDist = the w-th entry)
      const int slot = __popc(__ballot_sync(0xffffffffu, lane < w && td < dd));   // updateTopK, index_utils.c:19-33
      const float ud = __shfl_up_sync(0xffffffffu, td, 1);
      const int ui = __shfl_up_sync(0xffffffffu, tid, 1);
      if (lane > slot) { td = ud; tid = ui; }
      else if (lane == slot) { td = dd; tid = id; }
    }
  }
  if (lane < w) mine = make_key(td, (uint32_t)tid);
  return false;
}

// The w nearest lists of one query from its C coarse distances (one warp): an ascending list of (distance, centroid id)
// keys holding the w+1 smallest and everything tied with them; lanes 0..w-1 hold the selection, lane w the runner-up;
// flags the rare cases for the general kernel.
// A distance tie across the w-th place makes the kept SET depend on arrival order (updateTopK inserts before equal
// entries, so the earliest of the tied entries sits last and is evicted first, freddy.c:272-283): the literal loop is
// then replayed over S = {d <= v}, v = the w-th smallest distance, in centroid order — entries beyond v can neither
// stay nor rearrange the entries up to v — with lane i holding entry i of cqSelection.  S is complete whenever the
// 32nd key lies beyond v; only otherwise (more than 31 - w tied centroids) the query goes to the general kernel.
__device__ __forceinline__ void coarse_select_warp(const float* __restrict__ dist_row, int C, int Cs,
                                                   const int32_t* __restrict__ list_len, int w, int k, int q,
                                                   int32_t* __restrict__ probes, uint32_t* __restrict__ qflags,
                                                   int force_exact, int lane) {
  u64 mine = kKeyInf;
  uint32_t thr = 0xffffffffu;                        // distance bits of the (w+1)-th smallest key so far
  for (int c0 = 0; c0 < Cs; c0 += 128) {            // four independent loads in flight per lane
    float dv[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int c = c0 + 32 * u + lane;
      dv[u] = (c < C) ? dist_row[c] : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int c = c0 + 32 * u + lane;
      u64 key = (c < C) ? make_key(dv[u], (uint32_t)c) : kKeyInf;
      // `<=` on the distance alone: every centroid tied with the runner-up stays in the list (the replay below needs
      // all entries up to the w-th distance; ties are rare, so this admits what `key < (w+1)-th key` would)
      unsigned mask = __ballot_sync(0xffffffffu, c < C && key_dbits(key) <= thr);
      while (mask) {
        int src = __ffs(mask) - 1;
        u64 nk = shfl_u64(key, src);
        warp_list_insert(mine, nk, lane);
        thr = key_dbits(shfl_u64(mine, w));
        mask &= mask - 1;
        mask &= __ballot_sync(0xffffffffu, c < C && key_dbits(key) <= thr);
      }
    }
  }
  uint32_t flags = force_exact ? (kFlagExact | kWhyForced) : 0u;
  u64 kw = shfl_u64(mine, w), kw1 = shfl_u64(mine, w - 1);
  // a tie across the w-th place makes the kept set order dependent
  if (kw != kKeyInf && key_dbits(kw) == key_dbits(kw1)) {
    if (coarse_tie_replay(mine, kw1, w, lane)) flags |= kFlagExact | kWhyCoarseTie;   // S may be incomplete
  }
  // sentinel quirk of the reference: a selected distance >= 100 is undefined
  if (key_dist(kw1) >= 100.0f) flags |= kFlagExact | kWhyCoarseFar;
  int len = 0;
  if (lane < w) {
    int cid = (int)key_t(mine);
    probes[(size_t)q * w + lane] = cid;
    len = (list_len != nullptr) ? list_len[cid] : k;   // nullptr: quantisation only (fb_encode_*), no lists involved
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) len += __shfl_xor_sync(0xffffffffu, len, s);
  if (len < k) flags |= kFlagExact | kWhyFewRows;  // re-probe loop (freddy.c:262) needed
  if (lane == 0) qflags[q] = flags;
}

// ---------------------------------------------------------------------------
// HOT(1) coarse quantizer: squared L2 of each query against all C centroids
// (sequential 3-op chain per dimension, index_utils.c:500-508), then the w
// nearest lists per query (freddy.c:264-283).  One CTA = QT (<=16) queries x all
// centroids; one thread = one centroid x QT independent chains, centroid
// table transposed ([d][Cs]) so the per-dimension load is coalesced and shared
// by the QT queries.  Selection: one warp per query keeps an ascending key list.
// ---------------------------------------------------------------------------
// PACKED: a thread owns two adjacent centroids as one f32x2 pair (FADD2/FMUL2/FFMA2, common.cuh); the
// query values enter as scalar-broadcast operands, so one 8-byte coalesced load and QT/4 broadcast
// 16-byte reads feed 3*QT packed operations per dimension (half the shared-memory traffic per
// operation of the scalar form).  `one` must be 1.0f supplied at run time.
template <int QT, bool PACKED>
__global__ void __launch_bounds__(kCoarseThreads, 2)
coarse_select_kernel_t(const float* __restrict__ queries, int nq, int d,
                     const float* __restrict__ coarseT, int C, int Cs,
                     const int32_t* __restrict__ list_len, int w, int k,
                     int32_t* __restrict__ probes,      // [nq][w]
                     uint32_t* __restrict__ qflags,      // [nq]
                     int force_exact, float one, float* __restrict__ q_copy) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* qs = reinterpret_cast<float*>(smem_raw);     // [d][QT]
  float* dist = qs + (size_t)d * QT;           // [QT][Cs]
  const int q0 = blockIdx.x * QT;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

  // stage the tile's queries (coalesced rows -> [d][QT]).  `queries` may be pinned HOST memory mapped into the
  // device address space: the upload then happens here, overlapped across CTAs with the distance work, and
  // `q_copy` receives the device copy the later stages read (fb_ivfadc_search, large batches).
#pragma unroll 1
  for (int idx = tid; idx < d * QT; idx += kCoarseThreads) {
    const int qq = idx / d, i = idx - qq * d;
    const int q = q0 + qq;
    float v = 0.0f;
    if (q < nq) {
      v = queries[(size_t)q * d + i];
      if (q_copy != nullptr) q_copy[(size_t)q * d + i] = v;
    }
    qs[i * QT + qq] = v;
  }
  __syncthreads();

  if (PACKED && QT % 4 == 0) {
    const u64 one2 = pack2(one, one);
    const int Ch = Cs >> 1;                                   // Cs is a multiple of 32
    for (int c2 = tid; c2 < Ch; c2 += kCoarseThreads) {
      u64 acc[QT];
#pragma unroll
      for (int qq = 0; qq < QT; qq++) acc[qq] = 0ull;
      const u64* col = reinterpret_cast<const u64*>(coarseT) + c2;
      constexpr int PF = 2;
      u64 cur[PF], nxt[PF];
#pragma unroll
      for (int u = 0; u < PF; u++) cur[u] = (u < d) ? __ldg(col + (size_t)u * Ch) : 0ull;
      for (int i0 = 0; i0 < d; i0 += PF) {
#pragma unroll
        for (int u = 0; u < PF; u++) nxt[u] = (i0 + PF + u < d) ? __ldg(col + (size_t)(i0 + PF + u) * Ch) : 0ull;
#pragma unroll
        for (int u = 0; u < PF; u++) {
          const int i = i0 + u;
          if (i >= d) break;
          const u64 cv2 = cur[u];
          const float4* qrow = reinterpret_cast<const float4*>(qs + i * QT);
#pragma unroll
          for (int v = 0; v < QT / 4; v++) {
            const float4 qv = qrow[v];
            u64 t;
            t = xsub2(pack2(qv.x, qv.x), cv2); acc[4 * v + 0] = xacc2(xmul2(t, t), one2, acc[4 * v + 0]);
            t = xsub2(pack2(qv.y, qv.y), cv2); acc[4 * v + 1] = xacc2(xmul2(t, t), one2, acc[4 * v + 1]);
            t = xsub2(pack2(qv.z, qv.z), cv2); acc[4 * v + 2] = xacc2(xmul2(t, t), one2, acc[4 * v + 2]);
            t = xsub2(pack2(qv.w, qv.w), cv2); acc[4 * v + 3] = xacc2(xmul2(t, t), one2, acc[4 * v + 3]);
          }
        }
#pragma unroll
        for (int u = 0; u < PF; u++) cur[u] = nxt[u];
      }
#pragma unroll
      for (int qq = 0; qq < QT; qq++) reinterpret_cast<u64*>(dist + (size_t)qq * Cs)[c2] = acc[qq];
    }
  } else
  for (int c = tid; c < Cs; c += kCoarseThreads) {
    float acc[QT];
#pragma unroll
    for (int qq = 0; qq < QT; qq++) acc[qq] = 0.0f;
    const float* col = coarseT + c;
    constexpr int PF = 4;  // dimensions per prefetch group; the next group loads while this one computes
    float cur[PF], nxt[PF];
#pragma unroll
    for (int u = 0; u < PF; u++) cur[u] = (u < d) ? __ldg(col + (size_t)u * Cs) : 0.0f;
    for (int i0 = 0; i0 < d; i0 += PF) {
#pragma unroll
      for (int u = 0; u < PF; u++) nxt[u] = (i0 + PF + u < d) ? __ldg(col + (size_t)(i0 + PF + u) * Cs) : 0.0f;
#pragma unroll
      for (int u = 0; u < PF; u++) {
        const int i = i0 + u;
        if (i >= d) break;
        const float cv = cur[u];
        if (QT % 4 == 0) {
          const float4* qrow = reinterpret_cast<const float4*>(qs + i * QT);
#pragma unroll
          for (int v = 0; v < QT / 4; v++) {
            float4 qv = qrow[v];
            float t0 = xsub(qv.x, cv), t1 = xsub(qv.y, cv), t2 = xsub(qv.z, cv), t3 = xsub(qv.w, cv);
            acc[4 * v + 0] = xadd(acc[4 * v + 0], xmul(t0, t0));
            acc[4 * v + 1] = xadd(acc[4 * v + 1], xmul(t1, t1));
            acc[4 * v + 2] = xadd(acc[4 * v + 2], xmul(t2, t2));
            acc[4 * v + 3] = xadd(acc[4 * v + 3], xmul(t3, t3));
          }
        } else {
#pragma unroll
          for (int qq = 0; qq < QT; qq++) {
            float t = xsub(qs[i * QT + qq], cv);
            acc[qq] = xadd(acc[qq], xmul(t, t));
          }
        }
      }
#pragma unroll
      for (int u = 0; u < PF; u++) cur[u] = nxt[u];
    }
#pragma unroll
    for (int qq = 0; qq < QT; qq++) dist[qq * Cs + c] = acc[qq];
  }
  __syncthreads();

  // selection: top-(w+1) by (distance, centroid id); w + 1 <= 32 (host-checked)
  for (int qq = warp; qq < QT; qq += kCoarseThreads / kWarp) {
    const int q = q0 + qq;
    if (q >= nq) continue;
    coarse_select_warp(dist + (size_t)qq * Cs, C, Cs, list_len, w, k, q, probes, qflags, force_exact, lane);
  }
}

// Small batches (a single query is the reference's everyday call): one thread per (query, centroid) so that
// the 300-step chain of one query spreads over 8 CTAs instead of one; distances go through global memory to a
// second, one-warp-per-query selection kernel.
constexpr int kCoarseSmallThreads = 128;
__global__ void __launch_bounds__(kCoarseSmallThreads)
coarse_dist_small_kernel(const float* __restrict__ queries, int d, const float* __restrict__ coarseT, int Cs,
                         float* __restrict__ dist_out) {            // [nq][Cs]
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* qs = reinterpret_cast<float*>(smem_raw);                   // [d]
  const int q = blockIdx.y, c = blockIdx.x * kCoarseSmallThreads + threadIdx.x;
  for (int i = threadIdx.x; i < d; i += kCoarseSmallThreads) qs[i] = queries[(size_t)q * d + i];
  __syncthreads();
  if (c >= Cs) return;
  const float* col = coarseT + c;
  float acc = 0.0f;
  constexpr int B = 10;                              // dimensions per batch; the next batch loads while this one computes
  float cur[B], nxt[B];
#pragma unroll
  for (int u = 0; u < B; u++) cur[u] = (u < d) ? __ldg(col + (size_t)u * Cs) : 0.0f;
  for (int i = 0; i < d; i += B) {
#pragma unroll
    for (int u = 0; u < B; u++) nxt[u] = (i + B + u < d) ? __ldg(col + (size_t)(i + B + u) * Cs) : 0.0f;
#pragma unroll
    for (int u = 0; u < B; u++) {
      if (i + u < d) {
        const float t = xsub(qs[i + u], cur[u]);
        acc = xadd(acc, xmul(t, t));
      }
    }
#pragma unroll
    for (int u = 0; u < B; u++) cur[u] = nxt[u];
  }
  dist_out[(size_t)q * Cs + c] = acc;
}

__global__ void __launch_bounds__(128)
coarse_select_small_kernel(const float* __restrict__ dist, int nq, int C, int Cs, const int32_t* __restrict__ list_len,
                           int w, int k, int32_t* __restrict__ probes, uint32_t* __restrict__ qflags, int force_exact) {
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= nq) return;
  coarse_select_warp(dist + (size_t)q * Cs, C, Cs, list_len, w, k, q, probes, qflags, force_exact, threadIdx.x & 31);
}

// ---------------------------------------------------------------------------
// HOT(2) LUT build.  job = (query, probed centroid | none).  LUT[job][pos][code]
// = squareDistance(residual + pos*sub, codeword(pos,code), sub) with residual =
// query - centroid (freddy.c:296-314, index_utils.c:445-455); for flat PQ the
// raw query is used (freddy.c:519-525).
// One CTA owns one (pos, tile of <=1024 codes): its codebook slice [sub][TK] is
// staged ONCE into shared memory by the TMA engine (1-D bulk copies) and reused
// for every job the CTA processes; one thread = one code, W jobs at a time = W
// independent accumulation chains fed by broadcast reads of the residuals.
// ---------------------------------------------------------------------------
// TKS > 0: one CTA owns TKS codes, the shared-memory row stride of its codebook slice is
// the compile-time constant TKS (immediate LDS offsets in the unrolled loop) and
// 1024/TKS CTAs share an SM so that one CTA's per-group barrier/prologue bubbles are
// filled by the others; TKS == 0: generic (stride = TK, one CTA per SM).
template <int TKS> struct LutCfg {
  static constexpr int kThreads = (TKS > 0) ? TKS : 1024;
  static constexpr int kPerSm = (TKS > 0) ? (1024 / TKS) : 1;
};
// PACKED: the W chains run two per instruction (FADD2/FMUL2/FFMA2, see common.cuh);
// `one` must be 1.0f and must reach the kernel as a run-time value.
template <int W, int TKS, bool PACKED>
__global__ void __launch_bounds__((LutCfg<TKS>::kThreads), (LutCfg<TKS>::kPerSm))
lut_build_kernel(const float* __restrict__ queries, int d,
                 const float* __restrict__ coarse,        // [C][d] row-major or nullptr
                 const int32_t* __restrict__ probes,      // [njobs] centroid per job or nullptr
                 int jobs_per_query, int njobs,
                 const float* __restrict__ cbT,           // [m][sub][K]
                 int m, int K, int sub, int TK,
                 float* __restrict__ lut,                 // [njobs][m][K]
                 float one) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  float* cbs = reinterpret_cast<float*>(smem_raw);        // [sub][stride]
  constexpr int WS = (W + 3) & ~3;                        // residual row stride (16-byte rows)
  const int stride = (TKS > 0) ? TKS : TK;
  float* rs = cbs + (size_t)sub * stride;                 // 2 x [sub][WS]
  const int tiles = (K + TK - 1) / TK;
  const int pos = blockIdx.x / tiles, tile = blockIdx.x % tiles;
  const int code0 = tile * TK;
  const int ncodes = min(TK, K - code0);
  const int tid = threadIdx.x;

  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, (uint32_t)(sub * ncodes * sizeof(float)));
    for (int i = 0; i < sub; i++)
      bulk_g2s(cbs + (size_t)i * stride, cbT + ((size_t)pos * sub + i) * K + code0,
               (uint32_t)(ncodes * sizeof(float)), &bar);
  }
  mbar_wait(&bar, 0);

  // Residuals of the next job group are fetched (two dependent global loads) while the
  // current group is being computed; rs is double-buffered so one barrier per group suffices.
  // Ownership of the residual elements is fixed per thread, so all index arithmetic is
  // hoisted out of the group loop; only warps that own elements take part.
  constexpr int kPre = 4;                                  // residual elements a thread may own
  const int n_res = sub * W;
  const bool pipelined = n_res <= kPre * (int)blockDim.x;
  const bool aligned = (W % jobs_per_query) == 0;          // then job0 is a multiple of jobs_per_query
  float pre_q[kPre], pre_c[kPre];
  int e_dst[kPre], e_jj[kPre], e_qoff[kPre], e_src[kPre];
#pragma unroll
  for (int e = 0; e < kPre; e++) {
    const int idx = tid + e * (int)blockDim.x;
    const int i = idx / W, jj = idx % W;
    e_jj[e] = (idx < n_res) ? jj : -1;
    e_dst[e] = i * WS + jj;
    e_qoff[e] = jj / jobs_per_query;
    e_src[e] = pos * sub + i;
    pre_q[e] = 0.0f;
    pre_c[e] = 0.0f;
  }
  const bool owner_warp = (tid & ~31) < n_res;
  auto prefetch = [&](int job0) {
    if (!owner_warp) return;
    const int q0 = job0 / jobs_per_query;
#pragma unroll
    for (int e = 0; e < kPre; e++) {
      if (e_jj[e] >= 0) {
        const int job = min(job0 + e_jj[e], njobs - 1);
        const int q = aligned ? min(q0 + e_qoff[e], (njobs - 1) / jobs_per_query) : job / jobs_per_query;
        pre_q[e] = queries[(size_t)q * d + e_src[e]];
        if (probes != nullptr) pre_c[e] = coarse[(size_t)probes[job] * d + e_src[e]];
      }
    }
  };
  int cur = 0;
  if (pipelined) prefetch(blockIdx.y * W);
  for (int job0 = blockIdx.y * W; job0 < njobs; job0 += gridDim.y * W) {
    float* rsc = cur ? rs + (size_t)sub * WS : rs;
    if (pipelined) {
      if (owner_warp) {
#pragma unroll
        for (int e = 0; e < kPre; e++)
          if (e_jj[e] >= 0) rsc[e_dst[e]] = (probes != nullptr) ? xsub(pre_q[e], pre_c[e]) : pre_q[e];
      }
      __syncthreads();
      const int next = job0 + gridDim.y * W;
      if (next < njobs) prefetch(next);
    } else {
      __syncthreads();  // previous group's readers are done
      for (int idx = tid; idx < n_res; idx += blockDim.x) {
        const int i = idx / W, jj = idx % W;
        const int job = min(job0 + jj, njobs - 1);
        const int q = job / jobs_per_query;
        const float qv = queries[(size_t)q * d + pos * sub + i];
        rsc[i * WS + jj] = (probes != nullptr) ? xsub(qv, coarse[(size_t)probes[job] * d + pos * sub + i]) : qv;
      }
      __syncthreads();
    }
    float* const out = lut + ((size_t)job0 * m + pos) * K + code0 + tid;   // + j * m * K for job j of the group
    const size_t job_stride = (size_t)m * K;
    const bool full = job0 + W <= njobs;
    if (tid < ncodes) {
      const float* pc = cbs + tid;
      const float* pr = rsc;
      int i = 0;
      if (PACKED) {
        constexpr int NP = (W + 1) / 2;                     // chain pairs (a padding chain, if any, is never stored)
        const u64 one2 = pack2(one, one);
        u64 acc2[NP];
#pragma unroll
        for (int p2 = 0; p2 < NP; p2++) acc2[p2] = 0ull;    // (+0.0f, +0.0f)
        auto step2 = [&](const float* pcc, const float* prr) {
          const float cv = *pcc;
          const u64 cv2 = pack2(cv, cv);
          const ulonglong2* rrow = reinterpret_cast<const ulonglong2*>(prr);   // broadcast 16-byte reads
#pragma unroll
          for (int v = 0; v < WS / 4; v++) {
            const ulonglong2 r4 = rrow[v];
            if (2 * v < NP) {
              const u64 t0 = xsub2(r4.x, cv2);
              acc2[2 * v] = xacc2(xmul2(t0, t0), one2, acc2[2 * v]);
            }
            if (2 * v + 1 < NP) {
              const u64 t1 = xsub2(r4.y, cv2);
              acc2[2 * v + 1] = xacc2(xmul2(t1, t1), one2, acc2[2 * v + 1]);
            }
          }
        };
        for (; i + 5 <= sub; i += 5, pc += 5 * stride, pr += 5 * WS) {
#pragma unroll
          for (int u = 0; u < 5; u++) step2(pc + u * stride, pr + u * WS);
        }
        for (; i < sub; i++, pc += stride, pr += WS) step2(pc, pr);
        float* o = out;   // one pointer bump per job instead of a 64-bit multiply per store
        if (full) {
#pragma unroll
          for (int p2 = 0; p2 < NP; p2++) {
            float lo, hi;
            unpack2(acc2[p2], lo, hi);
            if (2 * p2 < W) { *o = lo; o += job_stride; }
            if (2 * p2 + 1 < W) { *o = hi; o += job_stride; }
          }
        } else {
#pragma unroll
          for (int p2 = 0; p2 < NP; p2++) {
            float lo, hi;
            unpack2(acc2[p2], lo, hi);
            if (2 * p2 < W && job0 + 2 * p2 < njobs) o[(size_t)(2 * p2) * job_stride] = lo;
            if (2 * p2 + 1 < W && job0 + 2 * p2 + 1 < njobs) o[(size_t)(2 * p2 + 1) * job_stride] = hi;
          }
        }
      } else {
        float acc[W];
#pragma unroll
        for (int jj = 0; jj < W; jj++) acc[jj] = 0.0f;
        // one dimension of the chain for all W jobs: r - c, squared, accumulated (3 rounded ops)
        auto step = [&](const float* pcc, const float* prr) {
          const float cv = *pcc;
          const float4* rrow4 = reinterpret_cast<const float4*>(prr);   // broadcast 16-byte reads
          float rv[WS];
#pragma unroll
          for (int v = 0; v < WS / 4; v++) {
            float4 t4 = rrow4[v];
            rv[4 * v + 0] = t4.x; rv[4 * v + 1] = t4.y; rv[4 * v + 2] = t4.z; rv[4 * v + 3] = t4.w;
          }
#pragma unroll
          for (int jj = 0; jj < W; jj++) {
            float t = xsub(rv[jj], cv);
            acc[jj] = xadd(acc[jj], xmul(t, t));
          }
        };
        for (; i + 5 <= sub; i += 5, pc += 5 * stride, pr += 5 * WS) {
#pragma unroll
          for (int u = 0; u < 5; u++) step(pc + u * stride, pr + u * WS);
        }
        for (; i < sub; i++, pc += stride, pr += WS) step(pc, pr);
#pragma unroll
        for (int jj = 0; jj < W; jj++)
          if (full || job0 + jj < njobs) out[(size_t)jj * job_stride] = acc[jj];
      }
    }
    if (pipelined) cur ^= 1;
  }
}

// ---------------------------------------------------------------------------
// ADC distance of the row owned by this lane in one 32-row block, LUT in shared
// memory.  M > 0 / KC > 0: compile-time m / K (unrolled, immediate LUT offsets).
// ---------------------------------------------------------------------------
template <int M, int KC>
__device__ __forceinline__ float adc_block_row(const uint2* __restrict__ up, const char* lut_base,
                                               int m_rt, int U_rt, uint32_t row_stride_rt) {
  float acc = 0.0f;
  if (M > 0) {
    constexpr int UU = (M + 3) / 4;
    const uint32_t rs = (KC > 0) ? (uint32_t)KC * 4u : row_stride_rt;
    uint2 v[UU > 0 ? UU : 1];
#pragma unroll
    for (int u = 0; u < UU; u++) v[u] = __ldg(up + u * 32);
#pragma unroll
    for (int u = 0; u < UU; u++) {
      const uint32_t wlo = v[u].x, whi = v[u].y;
      const char* base = lut_base + (size_t)(4 * u) * rs;
      if (4 * u + 0 < M) acc = xadd(acc, *reinterpret_cast<const float*>(base + (wlo & 0xFFFFu)));
      if (4 * u + 1 < M) acc = xadd(acc, *reinterpret_cast<const float*>(base + rs + (wlo >> 16)));
      if (4 * u + 2 < M) acc = xadd(acc, *reinterpret_cast<const float*>(base + 2 * rs + (whi & 0xFFFFu)));
      if (4 * u + 3 < M) acc = xadd(acc, *reinterpret_cast<const float*>(base + 3 * rs + (whi >> 16)));
    }
  } else {
    for (int u = 0; u < U_rt; u++) {
      const uint2 vv = __ldg(up + u * 32);
      const uint32_t wlo = vv.x, whi = vv.y;
      const char* base = lut_base + (size_t)(4 * u) * row_stride_rt;
      const int p = 4 * u;
      if (p + 0 < m_rt) acc = xadd(acc, *reinterpret_cast<const float*>(base + (wlo & 0xFFFFu)));
      if (p + 1 < m_rt) acc = xadd(acc, *reinterpret_cast<const float*>(base + row_stride_rt + (wlo >> 16)));
      if (p + 2 < m_rt) acc = xadd(acc, *reinterpret_cast<const float*>(base + 2 * row_stride_rt + (whi & 0xFFFFu)));
      if (p + 3 < m_rt) acc = xadd(acc, *reinterpret_cast<const float*>(base + 3 * row_stride_rt + (whi >> 16)));
    }
  }
  return acc;
}

// ---------------------------------------------------------------------------
// HOT(3)+(4) ADC scan + top-(k+1).  One CTA = one task = (LUT, list).  The
// task's LUT (m*K fp32, 48 KB at m=12,K=1024) is pulled into shared memory with
// one bulk async copy; each lane then owns one row per 32-row block: U coalesced
// 8-byte loads of pre-scaled codes, m shared-memory gathers, m sequential adds
// (freddy.c:364-368 / index_utils.c:1126-1133).  Each warp keeps the KK = k+1
// smallest (distance, arrival) keys it has seen in registers (lane i = i-th);
// a row is looked at again only if it beats the warp's current KK-th key.
// Output: per-warp ascending key lists, merged by finalize_kernel.
// ---------------------------------------------------------------------------
template <int M>  // M > 0: compile-time m (unrolled); M == 0: runtime m
__global__ void __launch_bounds__(kScanThreads, 4)
adc_scan_kernel(CodeTableDev tab,
                const int32_t* __restrict__ task_list,   // [ntasks] list per task, or nullptr (list = task % n_lists... see host)
                int tasks_per_lut,                       // LUT index = task / tasks_per_lut
                int lists_per_task_mod,                  // if task_list == nullptr: list = task % lists_per_task_mod
                int segs,                                // CTAs per task: each scans 1/segs of the list's blocks
                const float* __restrict__ lut, int K, int KK,
                u64* __restrict__ partial,               // [ntasks * segs][KK]: the CTA's KK smallest keys, ascending
                float sentinel) {                        // rows with distance >= sentinel are never admitted (freddy.c:369)
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar;
  float* slut = reinterpret_cast<float*>(smem_raw);
  const int m = (M > 0) ? M : tab.m;
  const int U = (M > 0) ? (M + 3) / 4 : tab.U;
  const int task = blockIdx.x / segs, seg = blockIdx.x % segs;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t lut_bytes = (uint32_t)((size_t)m * K * sizeof(float));

  if (tid == 0) {
    mbar_init(&bar, 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar, lut_bytes);
    bulk_g2s(slut, lut + (size_t)(task / tasks_per_lut) * m * K, lut_bytes, &bar);
  }
  const int list = (task_list != nullptr) ? task_list[task] : (task % lists_per_task_mod);
  const int blk0 = tab.list_blk[list];
  const int len = tab.list_len[list];
  const int nblk = (len + 31) >> 5;
  const int b_lo = (int)(((int64_t)nblk * seg) / segs), b_hi = (int)(((int64_t)nblk * (seg + 1)) / segs);
  mbar_wait(&bar, 0);

  u64 mine = kKeyInf;
  // the reference admits with `distance < maxDist`, maxDist starting at the sentinel: the largest admissible
  // distance is the float just below it (distances are >= +0, so bit patterns order like the values)
  const uint32_t thr0 = __float_as_uint(sentinel) - 1u;
  uint32_t thr_bits = thr0;
  const char* lut_bytes_base = reinterpret_cast<const char*>(slut);
  const uint32_t row_stride = (uint32_t)K * 4u;

  for (int b = b_lo + warp; b < b_hi; b += kScanWarps) {
    const uint2* up = tab.units + ((size_t)(blk0 + b) * U) * 32 + lane;
    const float acc = adc_block_row<M, 0>(up, lut_bytes_base, m, U, row_stride);
    const bool valid = (b * 32 + lane) < len;
    const uint32_t dbits = __float_as_uint(acc);
    bool cand = valid && (dbits <= thr_bits);
    unsigned mask = __ballot_sync(0xffffffffu, cand);
    if (mask) {
      u64 key = kKeyInf;
      if (cand) key = make_key(acc, (uint32_t)tab.rowno[(size_t)(blk0 + b) * 32 + lane]);
      while (mask) {
        int src = __ffs(mask) - 1;
        u64 nk = shfl_u64(key, src);
        warp_list_insert(mine, nk, lane);
        mask &= mask - 1;
      }
      thr_bits = min(thr0, key_dbits(shfl_u64(mine, KK - 1)));
    }
  }
  // merge the warps' lists inside the CTA (the LUT is no longer needed: reuse its shared memory)
  __syncthreads();
  u64* stage = reinterpret_cast<u64*>(smem_raw);
  stage[warp * 32 + lane] = mine;
  __syncthreads();
  if (warp == 0) {
    for (int l = 1; l < kScanWarps; l++) {
      const u64 other = stage[l * 32 + lane];
      if (__ballot_sync(0xffffffffu, other <= (shfl_u64(mine, KK - 1) | 0xFFFFFFFFull)) == 0) continue;   // keeps distance ties
      warp_list_merge(mine, other, lane);
    }
    if (lane < KK) partial[(size_t)blockIdx.x * KK + lane] = mine;
  }
}

// same chain from units already in registers (software-pipelined scan, M > 0)
template <int M, int KC>
__device__ __forceinline__ float adc_units(const uint2 (&v)[(M + 3) / 4 > 0 ? (M + 3) / 4 : 1], const char* lut_base,
                                           uint32_t row_stride_rt) {
  constexpr int UU = (M + 3) / 4;
  const uint32_t rs = (KC > 0) ? (uint32_t)KC * 4u : row_stride_rt;
  float acc = 0.0f;
#pragma unroll
  for (int u = 0; u < UU; u++) {
    const uint32_t wlo = v[u].x, whi = v[u].y;
    const char* base = lut_base + (size_t)(4 * u) * rs;
    if (4 * u + 0 < M) acc = xadd(acc, *reinterpret_cast<const float*>(base + (wlo & 0xFFFFu)));
    if (4 * u + 1 < M) acc = xadd(acc, *reinterpret_cast<const float*>(base + rs + (wlo >> 16)));
    if (4 * u + 2 < M) acc = xadd(acc, *reinterpret_cast<const float*>(base + 2 * rs + (whi & 0xFFFFu)));
    if (4 * u + 3 < M) acc = xadd(acc, *reinterpret_cast<const float*>(base + 3 * rs + (whi >> 16)));
  }
  return acc;
}

// Write the k results of one query in the reference's order (ascending distance,
// later arrival first among equal distances, index_utils.c:19-33) from an ascending
// key list held one key per lane, or flag the query for the general kernel when a
// distance tie straddles the k-th place.  Called by one full warp.
// q indexes the (chunk-offset) per-query arrays; q + q_base is what goes into exact_list.
__device__ __forceinline__ void warp_emit_topk(u64 mine, int lane, int q, int q_base, int k, uint32_t flags,
                                               const int32_t* __restrict__ ids, float sentinel,
                                               uint32_t* __restrict__ qflags,
                                               int32_t* __restrict__ out_ids, float* __restrict__ out_dists,
                                               int32_t* __restrict__ exact_list, int32_t* __restrict__ exact_count,
                                               u64* __restrict__ exact_total, u64* __restrict__ kth_key,
                                               int n_valid) {
  // lanes 0..n_valid-1 hold the n_valid smallest keys (lane i = i-th; n_valid = 32 when the warp lists were
  // merged untruncated, k+2 when they went through the partial-list buffer; every row whose distance is <=
  // the (k+2)-th smallest distance was admitted by the scan).  A distance tie across the k-th place makes the reference's result
  // depend on arrival order (strict `<` admission + insert-before-equal, index_utils.c:19-33).  With v = the
  // k-th smallest distance, rows with d > v can neither enter nor rearrange the entries <= v (fact B,
  // tests/test_topk_semantics.py), so the literal loop over S = {rows with d <= v} in arrival order gives
  // the reference's result.  S is complete in this list iff some lane < n_valid holds a key with d > v (or none):
  // then the warp replays it here (fact D).  S need not even be complete (fact E): the first k rows of S in
  // arrival order are all admitted (the array is not yet full of entries <= v), after that every row tied at v
  // is rejected (strict `<`) and every later row with d < v evicts the EARLIEST surviving tie (equal entries are
  // inserted in front of each other, so the earliest sits last).  Hence only the rows with d < v and the k
  // earliest rows tied at v matter; further ties replay as rejections.  Keys order ties by arrival, so the list
  // holds the earliest ones: the replay below is exact as soon as the list shows k ties (or all of S).
  // What remains for the general kernel: a list saturated with rows <= v that shows fewer than k ties
  // (needs more than 32 - k rows below v, or the k+2-key lists of the segmented scan).
  const u64 kk = shfl_u64(mine, k), kk1 = shfl_u64(mine, k - 1);
  bool replayed = false;
  if (kk != kKeyInf && key_dbits(kk) == key_dbits(kk1)) {
    const uint32_t v = key_dbits(kk1);
    const bool member = lane < n_valid && (mine != kKeyInf) && key_dbits(mine) <= v;
    const unsigned members = __ballot_sync(0xffffffffu, member);
    const int n_ties = __popc(__ballot_sync(0xffffffffu, member && key_dbits(mine) == v));
    if (members == (n_valid >= 32 ? 0xffffffffu : ((1u << n_valid) - 1u)) && n_ties < k) {
      flags |= kFlagExact | kWhyScanTie;
    } else {
      // S by arrival: (arrival << 32 | distance bits), non-members last
      u64 byt = member ? (((u64)key_t(mine) << 32) | (u64)key_dbits(mine)) : kKeyInf;
      byt = warp_sort_u64(byt, lane);
      const int n_s = __popc(members);
      // literal updateTopK on lanes 0..k-1 (lane i = tk[i]), starting from k sentinel entries
      float tk_d = sentinel;
      uint32_t tk_t = 0xFFFFFFFFu;
      for (int s_i = 0; s_i < n_s; s_i++) {
        const u64 e = shfl_u64(byt, s_i);
        const float dist = __uint_as_float((uint32_t)e);
        const uint32_t t = (uint32_t)(e >> 32);
        const float max_dist = __shfl_sync(0xffffffffu, tk_d, k - 1);
        const float up_d = __shfl_up_sync(0xffffffffu, tk_d, 1);
        const uint32_t up_t = __shfl_up_sync(0xffffffffu, tk_t, 1);
        const int pos = __popc(__ballot_sync(0xffffffffu, lane < k && tk_d < dist));   // entries strictly smaller
        if (dist < max_dist) {
          if (lane == pos) { tk_d = dist; tk_t = t; }
          else if (lane > pos && lane < k) { tk_d = up_d; tk_t = up_t; }
        }
      }
      if (lane < k) {
        out_ids[(size_t)q * k + lane] = (tk_t == 0xFFFFFFFFu) ? -1 : ids[tk_t];
        out_dists[(size_t)q * k + lane] = tk_d;
      }
      replayed = true;
    }
  }
  if (lane == 0) { kth_key[q] = kk1; qflags[q] = flags; }   // kk1: the k-th smallest key = the general kernel's bound
  if (flags & kFlagExact) {
    if (lane == 0) {
      exact_list[atomicAdd(exact_count, 1)] = q + q_base;
      atomicAdd(exact_total, 1ull);
      for (int b = 1; b <= 5; b++)
        if (flags & (1u << b)) atomicAdd(exact_total + b, 1ull);
    }
    return;
  }
  if (replayed) return;
  const uint32_t dbits = key_dbits(mine);
  const uint32_t prev = __shfl_up_sync(0xffffffffu, dbits, 1);
  const bool in_range = lane < k;
  const bool start = in_range && (lane == 0 || dbits != prev);
  const unsigned starts = __ballot_sync(0xffffffffu, start);
  if (in_range) {
    const unsigned below = starts & (0xffffffffu >> (31 - lane));
    const int s = 31 - __clz(below);
    const unsigned above = starts & ~(0xffffffffu >> (31 - lane));
    const int e = above ? (__ffs(above) - 2) : (k - 1);
    const int outpos = s + (e - lane);
    const bool filled = mine != kKeyInf;
    out_ids[(size_t)q * k + outpos] = filled ? ids[key_t(mine)] : -1;
    out_dists[(size_t)q * k + outpos] = filled ? key_dist(mine) : sentinel;
  }
}

// ---------------------------------------------------------------------------
// HOT(3)+(4), throughput form: one CTA = one query.  The CTA walks the query's w
// probed lists; LUT j+1 streams into the second shared-memory buffer (bulk async
// copy + mbarrier) while list j is scanned, so a warp's top-(k+1) key list and the
// CTA-wide admission threshold persist across all w lists (far fewer insertions
// than one list at a time), and the merge + reference ordering of finalize_kernel
// is fused at the end.
// ---------------------------------------------------------------------------
constexpr int kQScanThreads = 512;
constexpr int kQScanWarps = kQScanThreads / kWarp;

// ADC distance from a row of byte codes (uint4 = 16 codes), LUT in shared memory, M and KC compile-time
template <int M, int KC>
__device__ __forceinline__ float adc_bytes(const uint4& v, const char* lut_base) {
  const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
  float acc = 0.0f;
#pragma unroll
  for (int p = 0; p < M; p++) {
    const uint32_t code = (wds[p >> 2] >> (8 * (p & 3))) & 0xFFu;
    acc = xadd(acc, *reinterpret_cast<const float*>(lut_base + (size_t)p * KC * 4 + code * 4u));
  }
  return acc;
}

template <int M, int KC, bool C8 = false>
__global__ void __launch_bounds__(kQScanThreads, 2)
adc_scan_query_kernel(CodeTableDev tab, const int32_t* __restrict__ probes, int w,
                      const float* __restrict__ lut, int K, int KK, int k, float sentinel,
                      uint32_t* __restrict__ qflags,
                      int32_t* __restrict__ out_ids, float* __restrict__ out_dists,
                      int32_t* __restrict__ exact_list, int32_t* __restrict__ exact_count,
                      u64* __restrict__ exact_total, u64* __restrict__ kth_key, int q_base) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar[2];
  __shared__ uint32_t s_thr;
  const int m = (M > 0) ? M : tab.m;
  const int U = (M > 0) ? (M + 3) / 4 : tab.U;
  const int Kc = (KC > 0) ? KC : K;
  const size_t lut_floats = (size_t)m * Kc;
  const uint32_t lut_bytes = (uint32_t)(lut_floats * sizeof(float));
  const int q = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* qlut = lut + (size_t)q * w * lut_floats;

  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
    s_thr = __float_as_uint(sentinel) - 1u;   // strict admission below the sentinel (freddy.c:369)
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar[0], lut_bytes);
    bulk_g2s(smem_raw, qlut, lut_bytes, &bar[0]);
    if (w > 1) {
      mbar_expect_tx(&bar[1], lut_bytes);
      bulk_g2s(smem_raw + lut_bytes, qlut + lut_floats, lut_bytes, &bar[1]);
    }
  }

  u64 mine = kKeyInf;
  const uint32_t thr0 = __float_as_uint(sentinel) - 1u;
  uint32_t my_thr = thr0;
  const uint32_t row_stride = (uint32_t)Kc * 4u;

  for (int j = 0; j < w; j++) {
    const int list = probes[(size_t)q * w + j];
    const int blk0 = tab.list_blk[list];
    const int len = tab.list_len[list];
    const int nblk = (len + 31) >> 5;
    mbar_wait(&bar[j & 1], (uint32_t)((j >> 1) & 1));
    const char* lut_base = reinterpret_cast<const char*>(smem_raw) + (size_t)(j & 1) * lut_bytes;
    constexpr int UU = (M > 0) ? (M + 3) / 4 : 1;
    uint2 cur[UU], nxt[UU];
    uint4 cur8 = make_uint4(0, 0, 0, 0), nxt8 = make_uint4(0, 0, 0, 0);
    if (C8) {
      if (warp < nblk) cur8 = __ldg(tab.units8 + (size_t)(blk0 + warp) * 32 + lane);
    } else if (M > 0 && warp < nblk) {
      const uint2* up0 = tab.units + ((size_t)(blk0 + warp) * U) * 32 + lane;
#pragma unroll
      for (int u = 0; u < UU; u++) cur[u] = __ldg(up0 + u * 32);
    }
    for (int b = warp; b < nblk; b += kQScanWarps) {
      float acc;
      if (C8) {
        const int bn = b + kQScanWarps;
        if (bn < nblk) nxt8 = __ldg(tab.units8 + (size_t)(blk0 + bn) * 32 + lane);
        acc = adc_bytes<(M > 0 ? M : 1), (KC > 0 ? KC : 1)>(cur8, lut_base);
        cur8 = nxt8;
      } else if (M > 0) {
        // the next block's codes are requested before this block's gathers: their latency hides behind them
        const int bn = b + kQScanWarps;
        if (bn < nblk) {
          const uint2* upn = tab.units + ((size_t)(blk0 + bn) * U) * 32 + lane;
#pragma unroll
          for (int u = 0; u < UU; u++) nxt[u] = __ldg(upn + u * 32);
        }
        acc = adc_units<M, KC>(cur, lut_base, row_stride);
#pragma unroll
        for (int u = 0; u < UU; u++) cur[u] = nxt[u];
      } else {
        const uint2* up = tab.units + ((size_t)(blk0 + b) * U) * 32 + lane;
        acc = adc_block_row<M, KC>(up, lut_base, m, U, row_stride);
      }
      const uint32_t thr = min(my_thr, *reinterpret_cast<volatile uint32_t*>(&s_thr));
      const uint32_t dbits = __float_as_uint(acc);
      const bool cand = ((b * 32 + lane) < len) && (dbits <= thr);
      unsigned mask = __ballot_sync(0xffffffffu, cand);
      if (mask) {
        u64 key = kKeyInf;
        if (cand) key = make_key(acc, (uint32_t)tab.rowno[(size_t)(blk0 + b) * 32 + lane]);
        while (mask) {
          const int src = __ffs(mask) - 1;
          warp_list_insert(mine, shfl_u64(key, src), lane);
          mask &= mask - 1;
        }
        my_thr = min(thr0, key_dbits(shfl_u64(mine, KK - 1)));
        if (lane == 0 && my_thr < thr) atomicMin(&s_thr, my_thr);
      }
    }
    __syncthreads();  // every warp is done with buf[j & 1]
    if (tid == 0 && j + 2 < w) {
      mbar_expect_tx(&bar[j & 1], lut_bytes);
      bulk_g2s(smem_raw + (size_t)(j & 1) * lut_bytes, qlut + (size_t)(j + 2) * lut_floats, lut_bytes, &bar[j & 1]);
    }
  }
  // merge the warps' lists (all LUT loads have landed and been consumed: reuse buffer 0)
  u64* stage = reinterpret_cast<u64*>(smem_raw);
  stage[warp * 32 + lane] = mine;
  __syncthreads();
  if (warp == 0) {
    for (int l = 1; l < kQScanWarps; l++) {
      const u64 other = stage[l * 32 + lane];
      if (__ballot_sync(0xffffffffu, other <= (shfl_u64(mine, KK - 1) | 0xFFFFFFFFull)) == 0) continue;   // keeps distance ties
      warp_list_merge(mine, other, lane);
    }
    warp_emit_topk(mine, lane, q, q_base, k, qflags[q], tab.ids, sentinel, qflags, out_ids, out_dists,
                   exact_list, exact_count, exact_total, kth_key, 32);
  }
}

// ---------------------------------------------------------------------------
// Large-k form (k > 30, e.g. ivfadc_search(v, pvf*k) of the post-verification wrappers,
// freddy--0.0.1.sql:556-591): the same walk over the query's w lists, but every row's key
// (ADC distance bits, table row) is written out; topk_from_keys_kernel (knn_join_kernels.cuh)
// then applies the reference's top-k to the materialised stream.
// ---------------------------------------------------------------------------
template <int M, int KC>
__global__ void __launch_bounds__(kQScanThreads, 2)
adc_scan_keys_kernel(CodeTableDev tab, const int32_t* __restrict__ probes, int w,
                     const float* __restrict__ lut, int K,
                     u64* __restrict__ key_base, size_t stride, int32_t* __restrict__ n_keys) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  __shared__ __align__(8) uint64_t bar[2];
  const int m = (M > 0) ? M : tab.m;
  const int U = (M > 0) ? (M + 3) / 4 : tab.U;
  const int Kc = (KC > 0) ? KC : K;
  const size_t lut_floats = (size_t)m * Kc;
  const uint32_t lut_bytes = (uint32_t)(lut_floats * sizeof(float));
  const int q = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const float* qlut = lut + (size_t)q * w * lut_floats;
  u64* keys = key_base + (size_t)q * stride;
  if (tid == 0) {
    mbar_init(&bar[0], 1);
    mbar_init(&bar[1], 1);
    mbar_fence_init();
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&bar[0], lut_bytes);
    bulk_g2s(smem_raw, qlut, lut_bytes, &bar[0]);
    if (w > 1) {
      mbar_expect_tx(&bar[1], lut_bytes);
      bulk_g2s(smem_raw + lut_bytes, qlut + lut_floats, lut_bytes, &bar[1]);
    }
  }
  const uint32_t row_stride = (uint32_t)Kc * 4u;
  int base = 0;
  for (int j = 0; j < w; j++) {
    const int list = probes[(size_t)q * w + j];
    const int blk0 = tab.list_blk[list];
    const int len = tab.list_len[list];
    const int nblk = (len + 31) >> 5;
    mbar_wait(&bar[j & 1], (uint32_t)((j >> 1) & 1));
    const char* lut_base = reinterpret_cast<const char*>(smem_raw) + (size_t)(j & 1) * lut_bytes;
    for (int b = warp; b < nblk; b += kQScanWarps) {
      const uint2* up = tab.units + ((size_t)(blk0 + b) * U) * 32 + lane;
      const float acc = adc_block_row<M, KC>(up, lut_base, m, U, row_stride);
      const int r = b * 32 + lane;
      if (r < len) keys[base + r] = make_key(acc, (uint32_t)tab.rowno[(size_t)(blk0 + b) * 32 + lane]);
    }
    base += len;
    __syncthreads();
    if (tid == 0 && j + 2 < w) {
      mbar_expect_tx(&bar[j & 1], lut_bytes);
      bulk_g2s(smem_raw + (size_t)(j & 1) * lut_bytes, qlut + (size_t)(j + 2) * lut_floats, lut_bytes, &bar[j & 1]);
    }
  }
  if (tid == 0) n_keys[q] = base;
}

// ---------------------------------------------------------------------------
// finalize: one warp per query merges its n_lists per-warp key lists, checks the
// tie condition, and writes the k results in the reference's order: ascending
// distance, later arrival first among equal distances (index_utils.c:19-33).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
finalize_kernel(const u64* __restrict__ partial, int lists_per_query, int KK, int k,
                int nq, const int32_t* __restrict__ ids, float sentinel,
                uint32_t* __restrict__ qflags, int has_input_flags,   // in (from the coarse kernel) / out
                int32_t* __restrict__ out_ids, float* __restrict__ out_dists,   // [nq][k]
                int32_t* __restrict__ exact_list, int32_t* __restrict__ exact_count,
                u64* __restrict__ exact_total,            // cumulative statistic
                u64* __restrict__ kth_key,                // [nq] k-th smallest key (bound for the general kernel)
                int q_base) {
  const int lane = threadIdx.x & 31;
  const int q = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (q >= nq) return;
  // the query's lists_per_query * KK keys as one flat stream, 4 x 32 independent loads in flight; a key enters
  // the warp's 32-key list if it is not larger than the current (k+2)-th distance (keeps distance ties)
  u64 mine = kKeyInf;
  const u64* base = partial + (size_t)q * lists_per_query * KK;
  const int n_keys = lists_per_query * KK;
  u64 thr = kKeyInf;
  for (int i0 = 0; i0 < n_keys; i0 += 128) {
    u64 kv[4];
#pragma unroll
    for (int u = 0; u < 4; u++) {
      const int i = i0 + 32 * u + lane;
      kv[u] = (i < n_keys) ? base[i] : kKeyInf;
    }
#pragma unroll
    for (int u = 0; u < 4; u++) {
      unsigned mask = __ballot_sync(0xffffffffu, kv[u] != kKeyInf && kv[u] <= thr);
      while (mask) {
        const int src = __ffs(mask) - 1;
        const u64 nk = shfl_u64(kv[u], src);
        if (nk <= thr) {                                   // thr may have dropped since the ballot
          warp_list_insert(mine, nk, lane);
          const u64 kth = shfl_u64(mine, KK - 1);
          thr = (kth == kKeyInf) ? kKeyInf : (kth | 0xFFFFFFFFull);
        }
        mask &= mask - 1;
      }
    }
  }
  // lanes >= KK may hold keys beyond the (k+2)-th distance group that other lists truncated: not part of the contract
  if (lane >= KK) mine = kKeyInf;
  const uint32_t flags = has_input_flags ? qflags[q] : 0u;
  warp_emit_topk(mine, lane, q, q_base, k, flags, ids, sentinel, qflags, out_ids, out_dists, exact_list, exact_count,
                 exact_total, kth_key, KK);
}

// ---------------------------------------------------------------------------
// Quantisation of new rows (insert_batch): code[job][pos] = first minimum over the K entries of LUT row
// (job, pos) with the reference's strict `<` from 100 (index_utils.c:926-939): key = (distance bits, code),
// smallest key wins.  One warp per (job, pos).  A row whose distances are all >= 100 leaves the
// reference's assignment uninitialised: error flag.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
lut_argmin_kernel(const float* __restrict__ lut, int njobs, int m, int K, int16_t* __restrict__ codes, int32_t* __restrict__ err_flag) {
  const int lane = threadIdx.x & 31;
  const int64_t wid = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (wid >= (int64_t)njobs * m) return;
  const float* row = lut + (size_t)wid * K;
  u64 best = kKeyInf;
  for (int c = lane; c < K; c += 32) {
    const u64 key = make_key(row[c], (uint32_t)c);
    best = key < best ? key : best;
  }
#pragma unroll
  for (int s = 16; s >= 1; s >>= 1) {
    const u64 o = shfl_xor_u64(best, s);
    best = o < best ? o : best;
  }
  if (lane == 0) {
    if (!(key_dist(best) < 100.0f)) atomicExch(err_flag, 1);
    codes[wid] = (int16_t)key_t(best);
  }
}

}  // namespace fb
