// pipeline_kernels.cuh — the throughput form of the IVFADC hot path as ONE
// warp-specialised kernel per pipeline beat (sm_100a).
//
// The two heavy stages of ivfadc_search (freddy.c:247-378) stress different parts
// of an SM: the residual LUT build (freddy.c:295-314, index_utils.c:445-455) is
// bound by the fp32 pipe, the ADC scan (freddy.c:347-372) by shared-memory gathers
// (LSU).  Launch c of this kernel therefore does both at once on every SM:
//
//   producer warps (the last warps of the CTA) build the LUTs of query chunk c+1 into global scratch;
//   scan warps walk the probed lists of query chunk c, LUTs streamed through a ring
//       of shared-memory buffers by a loader thread (1-D bulk async copies +
//       mbarriers), a merger warp writes each query's k results;
//
// and the only dependency between the roles is launch order on the stream (chunk
// c+1's LUTs are complete when launch c ends), so there is no cross-CTA waiting.
//
// Exactness is unchanged: every distance is the reference's own chain of
// individually rounded fp32 operations (common.cuh), selection and tie handling are
// warp_emit_topk's (ivfadc_kernels.cuh).
#pragma once
#include "common.cuh"
#include "ivfadc_kernels.cuh"

namespace fb {

constexpr int kPipeTile = 512;                       // codes per producer CTA slice: 128 code threads x 4 codes
constexpr int kPipeBufs = 3;                         // LUT ring depth in shared memory

// Role split of the CTA.  PW producer warps (a multiple of 4: one job split of 8 jobs per 4 warps),
// SW scan warps, one loader warp, one merger warp; roles are whole warpgroups so that setmaxnreg can
// move registers from the scan side to the producers.
template <int PW, int SW, int JT = 8>
struct PipeCfg {
  static constexpr int kJobsPerThread = JT;          // jobs per producer thread (8 or 16): 4 codes x JT jobs = 2*JT packed accumulators
  static constexpr int kProdWarps = PW, kScanWarps = SW;
  static constexpr int kWarps = PW + SW + 2;
  static constexpr int kThreads = kWarps * 32;
  static constexpr int kProdThreads = PW * 32;
  static constexpr int kSplits = PW / 4;             // job splits: 4 warps = 128 threads x 4 codes = one 512-code slice
  static constexpr int kGroup = JT * kSplits;        // jobs per producer group
  static constexpr int kLaunchRegs = ((65536 / kThreads) & ~7) > 255 ? 248 : ((65536 / kThreads) & ~7);
  static constexpr int kScanRegs = 64;               // 56 left a loop bound of the scan warps in local memory (LDL in the hot loop)
  // setmaxnreg moves registers inside the CTA's own allocation (kLaunchRegs * kThreads), not the SM's 64 K: what the
  // scan side gives back is all the producers can take (asking for more blocks forever)
  static constexpr int kProdRegsRaw = ((kLaunchRegs * kThreads - kScanRegs * (kThreads - kProdThreads)) / kProdThreads) & ~7;
  static constexpr int kProdRegs = kProdRegsRaw > 232 ? 232 : kProdRegsRaw;
  static_assert(PW % 4 == 0 && (SW + 2) % 4 == 0, "roles must be whole warpgroups");
  static_assert(kThreads <= 1024, "CTA too large");
  static_assert(kProdRegs >= 72, "producers need about 80 registers");
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// 32-bit shared-window loads (constant offsets fold into LDS [R + imm]; no generic-address arithmetic)
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds_u32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

struct PipeArgs {
  // producer half: LUTs of `lut_njobs` jobs (job = query * jobs_per_query + probe)
  const float* lut_queries;     // [lut_nq][d]
  const int32_t* lut_probes;    // [lut_njobs] coarse centroid per job
  float* lut_out;               // [lut_njobs][m][K]
  int lut_njobs;
  int jobs_per_query;
  const float* coarse;          // [C][d]
  const float* cbT;             // [m][sub][K]
  int d, K;
  int n_slices, n_groups, tiles;
  float one;
  // scan half: queries [0, scan_nq) of the previous chunk
  CodeTableDev tab;
  const int32_t* scan_probes;   // [scan_nq][w]
  const float* scan_lut;        // [scan_nq][w][m][K]
  int scan_nq, w, KK, k;
  float sentinel;
  uint32_t* qflags;
  int32_t* out_ids;
  float* out_dists;
  int32_t* exact_list;
  int32_t* exact_count;
  u64* exact_total;
  u64* kth_key;
  int q_base;
  int32_t* work_counter;        // zeroed before the launch
};

// shared-memory plan (dynamic): [kPipeBufs][M*KC] LUT ring | [SUB][kPipeTile] codebook slice |
// 2 x [SUB][jobs per group] residuals | 2 x [scan warps][32] key staging | control block
template <int M, int KC, int SUB, class Cfg>
struct PipeSmem {
  static constexpr size_t lut_bytes = (size_t)M * KC * sizeof(float);
  static constexpr size_t off_cb = kPipeBufs * lut_bytes;
  static constexpr size_t cb_bytes = (size_t)SUB * kPipeTile * sizeof(float);
  static constexpr size_t off_rs = off_cb + cb_bytes;
  static constexpr size_t rs_bytes = 2 * (size_t)SUB * Cfg::kGroup * sizeof(float);
  static constexpr size_t off_stage = off_rs + rs_bytes;
  static constexpr size_t stage_bytes = 2 * (size_t)Cfg::kScanWarps * 32 * sizeof(u64);
  static constexpr size_t off_ctl = off_stage + stage_bytes;
  static constexpr size_t total = off_ctl + 256;
};

struct PipeCtl {
  uint64_t full[kPipeBufs], empty[kPipeBufs], stg_full[2], stg_empty[2], cb_bar;   // mbarriers
  int4 desc[kPipeBufs];        // {first block, rows, query, probe index} of the task in each ring slot
  int stage_q[2];
  uint32_t thr[2];
};
static_assert(sizeof(PipeCtl) <= 256, "control block");

// C8: the scan warps read the byte-code image of the table (CodeTableDev::units8, K <= 256): one 16-byte load per row
template <int M, int KC, int SUB, class Cfg, bool C8 = false>
__global__ void __launch_bounds__(Cfg::kThreads, 1)
ivfadc_pipe_kernel(const PipeArgs a) {
  using L = PipeSmem<M, KC, SUB, Cfg>;
  constexpr int kPipeWarps = Cfg::kWarps, kPipeProdWarps = Cfg::kProdWarps, kPipeProdThreads = Cfg::kProdThreads;
  constexpr int kPipeScanWarps = Cfg::kScanWarps, kPipeGroup = Cfg::kGroup;
  constexpr int kPipeProdRegs = Cfg::kProdRegs, kPipeScanRegs = Cfg::kScanRegs;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  PipeCtl* ctl = reinterpret_cast<PipeCtl*>(smem_raw + L::off_ctl);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr uint32_t lut_bytes = (uint32_t)L::lut_bytes;
  constexpr size_t lut_floats = (size_t)M * KC;

  if (tid == 0) {
    for (int b = 0; b < kPipeBufs; b++) {
      mbar_init(&ctl->full[b], 1);
      mbar_init(&ctl->empty[b], kPipeScanWarps);
    }
    for (int p = 0; p < 2; p++) {
      mbar_init(&ctl->stg_full[p], kPipeScanWarps);
      mbar_init(&ctl->stg_empty[p], 1);
      ctl->thr[p] = __float_as_uint(a.sentinel) - 1u;   // strict admission below the sentinel (freddy.c:369)
      ctl->stage_q[p] = -1;
    }
    mbar_init(&ctl->cb_bar, 1);
    mbar_fence_init();
  }
  __syncthreads();

  // Register re-balancing between the roles (warpgroup-wide setmaxnreg): the CTA is launched with
  // Cfg::kLaunchRegs registers per thread; the scan-side warpgroups give theirs back down to kScanRegs,
  // the producer warpgroups take them (12 + 14 warps: 512 * 64 + 384 * 80 <= 896 * 72), so the LUT build can keep its
  // operands for several dimensions in flight while the scan warps only need a few dozen registers.
  if (warp >= kPipeWarps - kPipeProdWarps) {
    if constexpr (kPipeProdRegs > Cfg::kLaunchRegs) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(kPipeProdRegs));
  } else {
    if constexpr (kPipeScanRegs < Cfg::kLaunchRegs) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(kPipeScanRegs));
  }

  // ============================================================ producers
  // PW warps = PW/4 job splits x 4 warps (128 code threads); a thread owns 4 adjacent codes (two packed pairs)
  // x JT jobs of the group: per dimension one 16-byte read of its codes and JT/4 broadcast 16-byte reads of
  // residuals feed 6*JT packed operations (JT = 8: shared-memory wavefronts per operation 0.17).
  if (warp >= kPipeWarps - kPipeProdWarps) {
    const int pt = (warp - (kPipeWarps - kPipeProdWarps)) * kWarp + lane;   // 0..255
    const int half = pt >> 7, ct = pt & 127;            // job split index (JT jobs each), code thread
    const int slice = blockIdx.x % a.n_slices, group = blockIdx.x / a.n_slices;
    if (a.lut_njobs <= 0 || group >= a.n_groups) return;
    const int pos = slice / a.tiles, tile = slice % a.tiles;
    const int code0 = tile * kPipeTile;
    const int ncodes = min(kPipeTile, a.K - code0);
    float* cbs = reinterpret_cast<float*>(smem_raw + L::off_cb);       // [SUB][kPipeTile]
    float* rs2 = reinterpret_cast<float*>(smem_raw + L::off_rs);       // 2 x [SUB][kPipeGroup] residuals
    if (pt == 0) {
      mbar_expect_tx(&ctl->cb_bar, (uint32_t)(SUB * ncodes * sizeof(float)));
      for (int i = 0; i < SUB; i++)
        bulk_g2s(cbs + (size_t)i * kPipeTile, a.cbT + ((size_t)pos * SUB + i) * a.K + code0,
                 (uint32_t)(ncodes * sizeof(float)), &ctl->cb_bar);
    }
    // residual element ownership: element e = i * kPipeGroup + jj of a group, kPre per thread
    constexpr int n_res = SUB * kPipeGroup;
    constexpr int kPre = (n_res + kPipeProdThreads - 1) / kPipeProdThreads;
    const int jpq = a.jobs_per_query;
    const int last_job = a.lut_njobs - 1;
    float pre_q[kPre], pre_c[kPre];
    auto prefetch = [&](int job0) {
#pragma unroll
      for (int e = 0; e < kPre; e++) {
        const int idx = pt + e * kPipeProdThreads;
        if (idx < n_res) {
          const int i = idx / kPipeGroup, jj = idx % kPipeGroup;
          const int job = min(job0 + jj, last_job);
          const int q = job / jpq;
          pre_q[e] = a.lut_queries[(size_t)q * a.d + pos * SUB + i];
          pre_c[e] = a.coarse[(size_t)a.lut_probes[job] * a.d + pos * SUB + i];
        }
      }
    };
    const int job_step = a.n_groups * kPipeGroup;
    int cur = 0;
    prefetch(group * kPipeGroup);
    mbar_wait(&ctl->cb_bar, 0);
    const u64 one2 = pack2(a.one, a.one);
    const bool active = 4 * ct < ncodes;
    const float* pc = cbs + 4 * ct;
    constexpr int WJ = Cfg::kJobsPerThread;                              // jobs per thread
    for (int job0 = group * kPipeGroup; job0 < a.lut_njobs; job0 += job_step) {
      float* rsc = rs2 + (size_t)cur * n_res;
#pragma unroll
      for (int e = 0; e < kPre; e++) {
        const int idx = pt + e * kPipeProdThreads;
        if (idx < n_res) {
          rsc[idx] = xsub(pre_q[e], pre_c[e]);
        }
      }
      named_bar_sync(1, kPipeProdThreads);   // rs2[cur] complete; rs2[cur^1] readers (previous group) are done
      if (job0 + job_step < a.lut_njobs) prefetch(job0 + job_step);
      if (active) {
        u64 acc[WJ][2];
#pragma unroll
        for (int j = 0; j < WJ; j++) acc[j][0] = acc[j][1] = 0ull;
        const float* rh = rsc + half * WJ;
#pragma unroll
        for (int i = 0; i < SUB; i++) {
          const ulonglong2 cv = *reinterpret_cast<const ulonglong2*>(pc + i * kPipeTile);   // codes 4ct..4ct+3, dimension i
          const float4* rr = reinterpret_cast<const float4*>(rh + i * kPipeGroup);            // broadcast reads, 4 jobs each
#pragma unroll
          for (int j4 = 0; j4 < WJ / 4; j4++) {
            const float4 r4 = rr[j4];
            const float rj[4] = {r4.x, r4.y, r4.z, r4.w};
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
              // pack2(r, r) costs nothing: FADD2 takes the scalar as a broadcast operand (R.F32)
              const u64 r2 = pack2(rj[jj], rj[jj]);
              u64 t;
              t = xsub2(r2, cv.x); acc[4 * j4 + jj][0] = xacc2(xmul2(t, t), one2, acc[4 * j4 + jj][0]);
              t = xsub2(r2, cv.y); acc[4 * j4 + jj][1] = xacc2(xmul2(t, t), one2, acc[4 * j4 + jj][1]);
            }
          }
        }
        const int jb = job0 + half * WJ;
        ulonglong2* o = reinterpret_cast<ulonglong2*>(a.lut_out + ((size_t)jb * M + pos) * a.K + code0) + ct;
        constexpr size_t job_stride = (size_t)M * KC / 4;          // in 16-byte units
        if (jb + WJ <= a.lut_njobs) {
#pragma unroll
          for (int j = 0; j < WJ; j++) o[(size_t)j * job_stride] = make_ulonglong2(acc[j][0], acc[j][1]);
        } else {
#pragma unroll
          for (int j = 0; j < WJ; j++)
            if (jb + j < a.lut_njobs) o[(size_t)j * job_stride] = make_ulonglong2(acc[j][0], acc[j][1]);
        }
      }
      cur ^= 1;
    }
    return;
  }

  // ============================================================ loader (one thread)
  if (warp == kPipeScanWarps) {
    if (lane != 0) return;
    int t = 0;
    for (;;) {
      const int q = (a.scan_nq > 0) ? atomicAdd(a.work_counter, 1) : 0x7fffffff;
      const bool end = q >= a.scan_nq;
      const int nj = end ? 1 : a.w;
      for (int j = 0; j < nj; j++, t++) {
        const int b = t % kPipeBufs, use = t / kPipeBufs;
        if (use > 0) mbar_wait(&ctl->empty[b], (uint32_t)((use - 1) & 1));
        if (end) {
          ctl->desc[b] = make_int4(0, 0, -1, 0);
          mbar_arrive(&ctl->full[b]);
        } else {
          const int list = a.scan_probes[(size_t)q * a.w + j];
          ctl->desc[b] = make_int4(a.tab.list_blk[list], a.tab.list_len[list], q, j);
          mbar_expect_tx(&ctl->full[b], lut_bytes);
          bulk_g2s(smem_raw + (size_t)b * lut_bytes, a.scan_lut + ((size_t)q * a.w + j) * lut_floats, lut_bytes,
                   &ctl->full[b]);
        }
      }
      if (end) return;
    }
  }

  // ============================================================ merger (one warp)
  if (warp == kPipeScanWarps + 1) {
    const u64* stage = reinterpret_cast<const u64*>(smem_raw + L::off_stage);
    for (int n = 0;; n++) {
      const int p = n & 1;
      mbar_wait(&ctl->stg_full[p], (uint32_t)((n >> 1) & 1));
      const int q = ctl->stage_q[p];
      if (q < 0) return;
      const u64* sp = stage + (size_t)p * kPipeScanWarps * 32;
      u64 mine = sp[lane];
      for (int l = 1; l < kPipeScanWarps; l++) {
        const u64 other = sp[l * 32 + lane];
        if (__ballot_sync(0xffffffffu, other <= (shfl_u64(mine, a.KK - 1) | 0xFFFFFFFFull)) == 0) continue;   // keeps distance ties
        warp_list_merge(mine, other, lane);
      }
      warp_emit_topk(mine, lane, q, a.q_base, a.k, a.qflags[q], a.tab.ids, a.sentinel, a.qflags, a.out_ids, a.out_dists,
                     a.exact_list, a.exact_count, a.exact_total, a.kth_key, 32);
      if (lane == 0) ctl->thr[p] = __float_as_uint(a.sentinel) - 1u;
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->stg_empty[p]);
    }
  }

  // ============================================================ scan warps
  {
    u64* stage = reinterpret_cast<u64*>(smem_raw + L::off_stage);
    constexpr int UU = (M + 3) / 4;
    const uint32_t smem_base = smem_u32(smem_raw);
    u64 mine = kKeyInf;
    const uint32_t thr0 = __float_as_uint(a.sentinel) - 1u;
    uint32_t my_thr = thr0;
    int nq_seen = 0, p = 0;
    for (int t = 0;; t++) {
      const int b = t % kPipeBufs;
      mbar_wait(&ctl->full[b], (uint32_t)((t / kPipeBufs) & 1));
      const int4 ds = ctl->desc[b];
      if (ds.w == 0) {   // first list of a query (or the end marker): staging slot and threshold of parity p must be free
        p = nq_seen & 1;
        if (nq_seen >= 2) mbar_wait(&ctl->stg_empty[p], (uint32_t)(((nq_seen >> 1) - 1) & 1));
        mine = kKeyInf;
        my_thr = thr0;
      }
      if (ds.z < 0) {
        if (warp == 0 && lane == 0) ctl->stage_q[p] = -1;
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->stg_full[p]);
        return;
      }
      const int blk0 = ds.x, len = ds.y;
      const int nblk = (len + 31) >> 5;
      const uint32_t lut_s = smem_base + (uint32_t)b * lut_bytes;        // shared-window address of this task's LUT
      const uint32_t thr_s = smem_u32(&ctl->thr[p]);
      // this warp's blocks: warp, warp + S, ...; their codes travel through three register stages
      // (A, B, C) so that two blocks are always in flight behind the one being gathered
      const int n_mine = (nblk > warp) ? (nblk - warp + kPipeScanWarps - 1) / kPipeScanWarps : 0;
      const uint2* ubase = a.tab.units + ((size_t)(blk0 + warp) * UU) * 32 + lane;
      constexpr size_t ustep = (size_t)kPipeScanWarps * UU * 32;           // uint2 elements between a warp's blocks
      uint2 sa[UU], sb[UU], sc[UU];
      auto load = [&](uint2 (&v)[UU], int i) {
        const uint2* up = ubase + (size_t)i * ustep;
#pragma unroll
        for (int u = 0; u < UU; u++) v[u] = __ldg(up + u * 32);
      };
      // a row's ADC distance is done: admission against the CTA-wide threshold, insertion into the warp's key list
      auto consider = [&](float acc, int i) {
        const uint32_t thr = min(my_thr, lds_u32(thr_s));
        const uint32_t dbits = __float_as_uint(acc);
        const int blk = warp + i * kPipeScanWarps;
        const bool cand = ((blk * 32 + lane) < len) && (dbits <= thr);
        unsigned mask = __ballot_sync(0xffffffffu, cand);
        if (mask) {
          u64 key = kKeyInf;
          if (cand) key = make_key(acc, (uint32_t)a.tab.rowno[(size_t)(blk0 + blk) * 32 + lane]);
          while (mask) {
            const int src = __ffs(mask) - 1;
            warp_list_insert(mine, shfl_u64(key, src), lane);
            mask &= mask - 1;
          }
          my_thr = min(thr0, key_dbits(shfl_u64(mine, a.KK - 1)));
          if (lane == 0 && my_thr < thr) atomicMin(&ctl->thr[p], my_thr);
        }
      };
      auto process = [&](const uint2 (&v)[UU], int i) {
        float acc = 0.0f;
#pragma unroll
        for (int u = 0; u < UU; u++) {
          const uint32_t wlo = v[u].x, whi = v[u].y;
          if (4 * u + 0 < M) acc = xadd(acc, lds_f32((uint32_t)((4 * u + 0) * KC * 4) + lut_s + (wlo & 0xFFFFu)));
          if (4 * u + 1 < M) acc = xadd(acc, lds_f32((uint32_t)((4 * u + 1) * KC * 4) + lut_s + (wlo >> 16)));
          if (4 * u + 2 < M) acc = xadd(acc, lds_f32((uint32_t)((4 * u + 2) * KC * 4) + lut_s + (whi & 0xFFFFu)));
          if (4 * u + 3 < M) acc = xadd(acc, lds_f32((uint32_t)((4 * u + 3) * KC * 4) + lut_s + (whi >> 16)));
        }
        consider(acc, i);
      };
      if constexpr (C8) {
        // byte codes: one 16-byte load per row, three blocks in flight like the 16-bit path
        const uint4* ubase8 = a.tab.units8 + (size_t)(blk0 + warp) * 32 + lane;
        constexpr size_t ustep8 = (size_t)kPipeScanWarps * 32;
        uint4 ta = make_uint4(0, 0, 0, 0), tb = ta, tc = ta;
        auto load8 = [&](uint4& v, int i) { v = __ldg(ubase8 + (size_t)i * ustep8); };
        auto process8 = [&](const uint4& v, int i) {
          const uint32_t wds[4] = {v.x, v.y, v.z, v.w};
          float acc = 0.0f;
#pragma unroll
          for (int pp = 0; pp < M; pp++) {
            const uint32_t code4 = ((wds[pp >> 2] >> (8 * (pp & 3))) & 0xFFu) << 2;
            acc = xadd(acc, lds_f32((uint32_t)(pp * KC * 4) + lut_s + code4));
          }
          consider(acc, i);
        };
        if (n_mine > 0) load8(ta, 0);
        if (n_mine > 1) load8(tb, 1);
        for (int i = 0; i < n_mine; i += 3) {
          if (i + 2 < n_mine) load8(tc, i + 2);
          process8(ta, i);
          if (i + 1 >= n_mine) break;
          if (i + 3 < n_mine) load8(ta, i + 3);
          process8(tb, i + 1);
          if (i + 2 >= n_mine) break;
          if (i + 4 < n_mine) load8(tb, i + 4);
          process8(tc, i + 2);
        }
      } else {
      if (n_mine > 0) load(sa, 0);
      if (n_mine > 1) load(sb, 1);
      for (int i = 0; i < n_mine; i += 3) {
        if (i + 2 < n_mine) load(sc, i + 2);
        process(sa, i);
        if (i + 1 >= n_mine) break;
        if (i + 3 < n_mine) load(sa, i + 3);
        process(sb, i + 1);
        if (i + 2 >= n_mine) break;
        if (i + 4 < n_mine) load(sb, i + 4);
        process(sc, i + 2);
      }
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(&ctl->empty[b]);   // this warp is done with ring slot b
      if (ds.w == a.w - 1) {                        // last list of the query: hand the warp's keys to the merger
        stage[((size_t)p * kPipeScanWarps + warp) * 32 + lane] = mine;
        if (warp == 0 && lane == 0) ctl->stage_q[p] = ds.z;
        __syncwarp();
        if (lane == 0) mbar_arrive(&ctl->stg_full[p]);
        nq_seen++;
      }
    }
  }
}

}  // namespace fb
