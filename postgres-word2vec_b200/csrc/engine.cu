// engine.cu — host side of libfreddy_b200.so: the C-ABI (include/freddy_b200.h),
// index upload ("pin once per session"), scratch management and the kernel
// pipeline.  No Postgres, no torch: plain CUDA runtime.  Every entry point
// returns a status code; C++ exceptions never cross the boundary.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <new>
#include <atomic>
#include <string>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/freddy_b200.h"
#include "../../include/freddy_sidecar.h"
#include "exact_kernels.cuh"
#include "ivfadc_kernels.cuh"
#include "vector_kernels.cuh"
#include "knn_join_kernels.cuh"
#include "pipeline_kernels.cuh"
#include "rerank_kernels.cuh"
#include "grouping_kernels.cuh"
#include "subset_kernels.cuh"
#include "prefilter_kernels.cuh"
#include "build_kernels.cuh"

using namespace fb;

namespace {

std::string g_create_error;

// bumped whenever a device buffer moves: captured CUDA graphs hold raw pointers and are re-captured after that
uint64_t g_alloc_generation = 0;

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  ~DevBuf() { release(); }          // fb_destroy selects the device before the engine (and its buffers) goes away
  cudaError_t ensure(size_t count) {
    if (count <= n && p) return cudaSuccess;
    g_alloc_generation++;
    if (p) cudaFree(p);
    p = nullptr;
    n = 0;
    if (count == 0) count = 1;
    cudaError_t err = cudaMalloc(&p, count * sizeof(T));
    if (err == cudaSuccess) n = count;
    return err;
  }
  // like ensure(), but what is there survives (append paths); grows by at least a quarter to amortise repeated appends
  cudaError_t grow(size_t count, size_t live) {
    if (count <= n && p) return cudaSuccess;
    const size_t cap = std::max(count, n + n / 4 + 1024);
    T* q = nullptr;
    cudaError_t err = cudaMalloc(&q, cap * sizeof(T));
    if (err != cudaSuccess) return err;
    if (p && live) err = cudaMemcpy(q, p, std::min(live, n) * sizeof(T), cudaMemcpyDeviceToDevice);
    if (p) cudaFree(p);
    g_alloc_generation++;
    p = q;
    n = cap;
    return err;
  }
  void release() {
    if (p) { cudaFree(p); g_alloc_generation++; }
    p = nullptr;
    n = 0;
  }
};

struct CodeTable {
  DevBuf<uint2> units;
  DevBuf<uint4> units8;                      // byte-code image (K <= 256, m <= 16): 16 bytes per row, same slots as `units`
  bool has8 = false;
  DevBuf<int32_t> rowno, list_blk, list_len, ids;
  DevBuf<int32_t> sorted_ids, sorted_rows;   // (id, row) pairs ordered by id, rows ascending inside an id: `WHERE id IN (...)` on the device
  int m = 0, U = 0, n_lists = 0;
  int64_t N = 0, n_blocks = 0;
  bool loaded = false;
  std::vector<int32_t> h_list_len, h_list_blk;
  int32_t max_id = 0;                        // largest id in the table (appended rows with larger ids keep the id index sorted)
  int K = 0;
  CodeTableDev dev() const {
    CodeTableDev t;
    t.units8 = has8 ? units8.p : nullptr;
    t.units = units.p; t.rowno = rowno.p; t.list_blk = list_blk.p; t.list_len = list_len.p;
    t.ids = ids.p; t.m = m; t.U = U; t.n_lists = n_lists;
    return t;
  }
  void release() {
    units.release(); units8.release(); rowno.release(); list_blk.release(); list_len.release(); ids.release();
    sorted_ids.release(); sorted_rows.release(); loaded = false; has8 = false;
  }
};

struct Codebook {
  DevBuf<float> cbT;  // [m][sub][K]
  int m = 0, K = 0, sub = 0;
  bool loaded = false;
};

enum Stage { ST_COARSE = 0, ST_LUT, ST_SCAN, ST_FINALIZE, ST_EXACT, ST_PIPE, ST_COUNT };

}  // namespace

struct fb_engine {
  int device = 0;
  int num_sms = 0;
  size_t smem_optin = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t own_stream = nullptr;
  std::string err;

  // index
  DevBuf<float> coarse, coarseT;
  int C = 0, Cs = 0, d = 0;
  bool coarse_loaded = false;
  Codebook cb[FB_CB_KINDS];
  CodeTable fine, pq, tmp;  // tmp: per-call target subset of pq (units / rowno only; compact, one list)
  DevBuf<uint32_t> sel_bitmap;          // `WHERE id IN (...)` on the device (subset_kernels.cuh)
  DevBuf<int32_t> sel_word_base, sel_total, sel_wanted, zero_i32;
  std::vector<int32_t> pq_ids_host;
  DevBuf<int32_t> iota_lists;
  // IVPQ index (kNN-join)
  CodeTable ivpq, jtmp;
  DevBuf<float> coarse_multi, ivpq_stats;
  DevBuf<int32_t> ivpq_cells;
  std::vector<float> ivpq_stats_host;
  int ivpq_Kc = 0, ivpq_d = 0;
  bool ivpq_loaded = false;
  DevBuf<int32_t> j_cell, j_vrow, j_id, j_active, j_active2, j_ncells, j_filled, j_tcounts;
  DevBuf<int32_t> j_counts, j_cell_start, j_perm, j_state;
  DevBuf<uint16_t> j_sel_cells;
  DevBuf<uint32_t> j_bitmaps;
  DevBuf<u64> j_keys;
  int64_t join_rounds = 0, join_pairs = 0;
  // word-vector table (analogy / exact rerank)
  DevBuf<float> vecT;                        // dimension-major 32-row blocks: the whole-table fp32 scans
  DevBuf<float> vecR;                        // row-major [N][d]: everything that gathers rows (post-verification, re-score, re-rank)
  DevBuf<int32_t> vec_ids, vec_sorted_ids, vec_sorted_rows;   // sorted (id, row) pairs: id -> row on the device
  DevBuf<float> sub_vT;                                        // gathered subset (knn_in_exact)
  DevBuf<int32_t> sub_rows, pv_cand;
  DevBuf<u64> knn_partial;
  std::vector<int32_t> vec_ids_host;
  bool vec_ids_sorted = true;
  std::unordered_map<int32_t, int32_t> vec_id_to_row;
  int64_t vec_N = 0;
  int vec_d = 0;
  bool vec_loaded = false;
  DevBuf<float> va, vb, vo;
  DevBuf<double> vdo;
  // tensor-core pre-filter of the exact scans (prefilter_kernels.cuh): bf16 image of the table [N_pad][kpa], its TMA map,
  // the largest row norm, and per-call scratch
  DevBuf<__nv_bfloat16> vec_bf16, pf_qb;
  CUtensorMap pf_tm_v;
  bool pf_ready = false;
  int pf_kpa = 0;
  int64_t pf_N_pad = 0;
  float pf_vmax = 0.0f;
  int prefilter = 1;          // FB_OPT_PREFILTER
  int byte_codes = 1;         // FB_OPT_BYTE_CODES
  bool used8 = false;         // the last scan read the byte image (accounting of algorithmic bytes)
  DevBuf<float> pf_eps2;
  DevBuf<uint32_t> pf_gbest, pf_norm;
  DevBuf<int32_t> pf_cnt, pf_ovf, pf_rows_out, pf_ex;
  DevBuf<int2> pf_cand;
  DevBuf<PfUnit> pf_units;
  DevBuf<int32_t> pf_progress;
  int pf_lockstep = 0;        // FB_OPT_PREFILTER_LOCKSTEP (off: once the epilogue stopped being the limiter the CTAs of a slab stay
                              // together by themselves — 1.94 GB of DRAM reads for the 1.92 GB table, L2 hit rate 80 %)
  int64_t pf_queries = 0, pf_overflow_queries = 0, pf_candidates = 0;
  DevBuf<int32_t> ana_rows;
  DevBuf<u64> ana_partial;

  // scratch
  DevBuf<float> lut, exact_lut, q_stage, dist_stage, coarse_dist;
  DevBuf<int32_t> probes, exact_list, id_stage, sel_rows, sel_order;
  DevBuf<uint32_t> qflags;
  DevBuf<u64> partial, kth;
  DevBuf<int32_t> small;   // [0]=exact_count [1]=work_counter [2]=error_flag
  DevBuf<u64> counters64;  // [0]=rows scanned [1]=exact queries

  // options
  bool force_exact = false;
  bool profile = false;
  int64_t query_chunk = 2048;
  int qscan_min_queries = 64;
  bool packed_fp32 = true;
  int lut_tile = 512;
  int lut_ctas_per_sm = 0;   // 0: as many as fit; overlap mode leaves room for scan CTAs
  int overlap = 0;           // 1: LUT build of chunk c+1 runs concurrently with the scan of chunk c (two streams)
  cudaStream_t s_lut = nullptr, s_scan = nullptr;
  cudaEvent_t ev_main = nullptr, ev_all = nullptr, ev_lut_done[2] = {nullptr, nullptr}, ev_scan_done[2] = {nullptr, nullptr};
  DevBuf<float> lut2;
  int pipeline = 1;          // 1: warp-specialised pipeline kernel for large batches of the headline shapes
  int64_t pipe_chunk = 2048; // queries per pipeline beat
  DevBuf<int32_t> pipe_counters;
  int pipe_debug = 0;
  // small host-buffer calls (<= kGraphMaxQueries queries) replay a captured CUDA graph of the whole call
  struct GraphEntry { int nq, k, w; uint64_t gen, epoch; cudaGraphExec_t exec; int launches; };
  std::vector<GraphEntry> graphs;
  uint64_t graph_epoch = 0;        // bumped by loads / option changes: kernel arguments are baked into a graph
  int use_graphs = 1;
  // set by fb_ivfadc_search for large batches in pinned host memory: the coarse kernel reads the queries through
  // this host-mapped pointer and writes the device copy (d_q) itself, so the upload overlaps the coarse step
  const float* coarse_src = nullptr;
  int zero_copy = 1;
  float* pin_q = nullptr; int32_t* pin_ids = nullptr; float* pin_d = nullptr; int32_t* pin_flag = nullptr;
  size_t pin_q_floats = 0;
  int pipe_shape = 0;
  int pipe_ramp = 0;
  int subset_placement = 1;       // FB_OPT_SUBSET_PLACEMENT: 0 off, 1 when it pays, 2 always
  bool device_build = true;   // FB_OPT_DEVICE_BUILD: CSR, placement and packing of a pinned table run as kernels
  int placement_window = 256; // rows considered per slot by the conflict-aware placement of the fine table (<= 1: off)
  volatile float one = 1.0f;

  // profiling
  struct Ev { cudaEvent_t a, b; int stage; };
  std::vector<Ev> events;
  size_t events_used = 0;
  double ms[ST_COUNT] = {0, 0, 0, 0, 0, 0};
  int64_t n_scan_launches = 0;
  int64_t n_pipe_launches = 0;
  int64_t launches = 0;
  int64_t queries_done = 0;
  int bytes_per_row = 0;
  int64_t host_rows = 0;
};

static int knn_prefilter_dev(fb_engine* e, const float* d_q, int nq, int k, int kk, const int32_t* d_exclude,
                             int32_t* d_out_ids, float* d_out_sims, int32_t* d_out_rows, std::vector<int32_t>& overflow);
static bool pf_usable(const fb_engine* e, int kk);

namespace {

int fail(fb_engine* e, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  if (e) e->err = buf; else g_create_error = buf;
  return code;
}

#define FB_CUDA(e, call)                                                                       \
  do {                                                                                         \
    cudaError_t err__ = (call);                                                                \
    if (err__ != cudaSuccess)                                                                  \
      return fail((e), FB_ERR_CUDA, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(err__), \
                  __FILE__, __LINE__);                                                         \
  } while (0)

struct StageTimer {
  fb_engine* e;
  int idx = -1;
  StageTimer(fb_engine* eng, int stage) : e(eng) {
    if (!e->profile) return;
    if (e->events_used == e->events.size()) {
      fb_engine::Ev ev;
      cudaEventCreate(&ev.a);
      cudaEventCreate(&ev.b);
      e->events.push_back(ev);
    }
    idx = (int)e->events_used++;
    e->events[idx].stage = stage;
    cudaEventRecord(e->events[idx].a, e->stream);
  }
  ~StageTimer() {
    if (idx >= 0) cudaEventRecord(e->events[idx].b, e->stream);
  }
};

void drain_events(fb_engine* e) {
  for (size_t i = 0; i < e->events_used; i++) {
    float t = 0;
    if (cudaEventElapsedTime(&t, e->events[i].a, e->events[i].b) == cudaSuccess) e->ms[e->events[i].stage] += t;
  }
  e->events_used = 0;
}

// ---- host-side layout transform of a code table ---------------------------
// Bank-conflict-aware row placement.  The ADC scan gathers LUT[pos][code] from shared memory with one
// row per lane; lanes whose codes differ but share a bank (code % 32: every LUT row starts on a 128-byte
// boundary) serialise, and a gather costs max-over-banks(distinct codes) data-pipe cycles (measured:
// scripts/microbench_smem.cu).  Rows of a list may sit in any slot — arrival order travels in rowno —
// so each 32-row block is filled greedily from a window of the list's remaining rows with the row that
// raises the fewest per-position bank maxima.  `order` receives, per list, the table rows in slot order.
void place_rows_of_list(const int16_t* codes, int m, int K, const std::vector<int32_t>& rows, int window,
                        std::vector<int32_t>& order) {
  const int n = (int)rows.size();
  order.clear();
  order.reserve(n);
  if (n <= 32 || m > 64) { order = rows; return; }
  std::vector<int32_t> rem(rows);
  size_t head = 0;                                  // rem[head..) are unplaced, in arrival order
  std::vector<uint8_t> cnt((size_t)m * 32);         // distinct codes per (pos, bank) in the open block
  std::vector<uint8_t> mx(m);
  std::vector<uint32_t> seen((size_t)m * ((K + 31) / 32));
  const int kw = (K + 31) / 32;
  while (head < rem.size()) {
    std::fill(cnt.begin(), cnt.end(), 0);
    std::fill(mx.begin(), mx.end(), 0);
    std::fill(seen.begin(), seen.end(), 0u);
    for (int slot = 0; slot < 32 && head < rem.size(); slot++) {
      const size_t avail = std::min<size_t>(window, rem.size() - head);
      size_t best = 0;
      if (slot > 0) {
        int best_cost = 1 << 30;
        for (size_t c = 0; c < avail; c++) {
          const int16_t* cr = codes + (size_t)rem[head + c] * m;
          int raises = 0, load = 0;
          for (int p = 0; p < m; p++) {
            const int code = cr[p];
            if (seen[(size_t)p * kw + (code >> 5)] >> (code & 31) & 1u) continue;   // same address: merged
            const int l = cnt[(size_t)p * 32 + (code & 31)];
            raises += (l + 1 > mx[p]);
            load += l;
          }
          const int cost = raises * 4096 + load;
          if (cost < best_cost) { best_cost = cost; best = c; if (cost == 0) break; }
        }
      }
      const int32_t r = rem[head + best];
      // keep the window in arrival order: shift the skipped rows up by one
      for (size_t c = best; c > 0; c--) rem[head + c] = rem[head + c - 1];
      head++;
      order.push_back(r);
      const int16_t* cr = codes + (size_t)r * m;
      for (int p = 0; p < m; p++) {
        const int code = cr[p];
        uint32_t& wd = seen[(size_t)p * kw + (code >> 5)];
        if (wd >> (code & 31) & 1u) continue;
        wd |= 1u << (code & 31);
        uint8_t& l = cnt[(size_t)p * 32 + (code & 31)];
        l++;
        if (l > mx[p]) mx[p] = l;
      }
    }
  }
}

int build_table_device(fb_engine* e, CodeTable& tab, const int32_t* ids, const int32_t* list_of_row,
                       int n_lists, int rows_per_pseudo_list, const int16_t* codes, int64_t N, int m, int K, int placement_window);
int build_table_host(fb_engine* e, CodeTable& tab, const int32_t* ids, const int32_t* list_of_row,
                     int n_lists, int rows_per_pseudo_list, const int16_t* codes, int64_t N, int m, int K, int placement_window);

int build_table(fb_engine* e, CodeTable& tab, const int32_t* ids, const int32_t* list_of_row,
                int n_lists, int rows_per_pseudo_list, const int16_t* codes, int64_t N, int m, int K,
                int placement_window = 0) {
  if (N < 0 || m <= 0) return fail(e, FB_ERR_INVALID, "bad table shape N=%lld m=%d", (long long)N, m);
  if (N >= (1ll << 31) - 64) return fail(e, FB_ERR_UNSUPPORTED, "table too large");
  if (e->device_build && N > 0 && placement_window <= 1024)
    return build_table_device(e, tab, ids, list_of_row, n_lists, rows_per_pseudo_list, codes, N, m, K, placement_window);
  return build_table_host(e, tab, ids, list_of_row, n_lists, rows_per_pseudo_list, codes, N, m, K, placement_window);
}

// FB_TRACE_BUILD=1: phase times of the device builder on stderr (diagnostics)
struct BuildTrace {
  bool on; cudaStream_t st; double t0; const char* what;
  static double now() { timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec + 1e-9 * t.tv_nsec; }
  BuildTrace(cudaStream_t s, const char* w) : on(getenv("FB_TRACE_BUILD") != nullptr), st(s), t0(0), what(w) { if (on) { cudaStreamSynchronize(st); t0 = now(); } }
  void lap(const char* phase) { if (on) { cudaStreamSynchronize(st); const double t = now(); fprintf(stderr, "[fb build] %s: %s %.2f ms\n", what, phase, (t - t0) * 1e3); t0 = t; } }
};

// CSR + placement + packing on the device (build_kernels.cuh); same layout as build_table_host
int build_table_device(fb_engine* e, CodeTable& tab, const int32_t* ids, const int32_t* list_of_row,
                       int n_lists, int rows_per_pseudo_list, const int16_t* codes, int64_t N, int m, int K, int placement_window) {
  const int U = (m + 3) / 4;
  const bool pseudo = list_of_row == nullptr;
  if (pseudo) n_lists = (int)std::max<int64_t>(1, (N + rows_per_pseudo_list - 1) / rows_per_pseudo_list);
  cudaStream_t st = e->stream;
  tab.loaded = false;          // the id column is overwritten before the rows are validated: a failed load leaves no table
  BuildTrace tr(st, "table");
  DevBuf<int32_t> d_list, d_len, d_diag, d_arrival, d_order, d_keys, d_iota, d_row_start;
  DevBuf<int16_t> d_codes;
  FB_CUDA(e, d_codes.ensure((size_t)N * m));
  FB_CUDA(e, d_len.ensure((size_t)n_lists));
  FB_CUDA(e, d_diag.ensure(4));
  FB_CUDA(e, tab.ids.ensure((size_t)N));
  FB_CUDA(e, cudaMemcpyAsync(d_codes.p, codes, (size_t)N * m * sizeof(int16_t), cudaMemcpyHostToDevice, st));
  FB_CUDA(e, cudaMemcpyAsync(tab.ids.p, ids, (size_t)N * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  if (!pseudo) {
    FB_CUDA(e, d_list.ensure((size_t)N));
    FB_CUDA(e, cudaMemcpyAsync(d_list.p, list_of_row, (size_t)N * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  }
  tr.lap("alloc + H2D of the raw columns");
  const int32_t diag0[4] = {0x7fffffff, 0x7fffffff, 0, 0};
  FB_CUDA(e, cudaMemcpyAsync(d_diag.p, diag0, sizeof diag0, cudaMemcpyHostToDevice, st));
  FB_CUDA(e, cudaMemsetAsync(d_len.p, 0, (size_t)n_lists * sizeof(int32_t), st));
  const int grid_rows = (int)std::min<int64_t>((N + 255) / 256, (int64_t)e->num_sms * 16);
  table_count_kernel<<<grid_rows, 256, 0, st>>>(pseudo ? nullptr : d_list.p, N, n_lists, d_codes.p, m, K, tab.ids.p, d_len.p, d_diag.p);
  e->launches++;
  std::vector<int32_t> len(n_lists, 0), blk(n_lists, 0), row_start(n_lists, 0);
  int32_t diag[4];
  FB_CUDA(e, cudaMemcpyAsync(diag, d_diag.p, sizeof diag, cudaMemcpyDeviceToHost, st));
  if (!pseudo) FB_CUDA(e, cudaMemcpyAsync(len.data(), d_len.p, (size_t)n_lists * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
  FB_CUDA(e, cudaStreamSynchronize(st));
  tr.lap("count + validate");
  if (diag[0] != 0x7fffffff)
    return fail(e, FB_ERR_INVALID, "row %lld: coarse_id %d out of range [0,%d)", (long long)diag[0], list_of_row[diag[0]], n_lists);
  if (diag[1] != 0x7fffffff) {
    const int64_t r = diag[1];
    for (int p = 0; p < m; p++) {
      const int code = codes[(size_t)r * m + p];
      if (code < 0 || code >= K) return fail(e, FB_ERR_INVALID, "row %lld pos %d: code %d out of range [0,%d)", (long long)r, p, code, K);
    }
  }
  if (pseudo)
    for (int c = 0; c < n_lists; c++) len[c] = (int32_t)std::min<int64_t>(rows_per_pseudo_list, N - (int64_t)c * rows_per_pseudo_list);
  int64_t n_blocks = 0, rows_before = 0;
  for (int c = 0; c < n_lists; c++) {
    blk[c] = (int32_t)n_blocks; row_start[c] = (int32_t)rows_before;
    n_blocks += (len[c] + 31) / 32; rows_before += len[c];
  }
  const bool want8 = K <= 256 && m <= 16;
  const size_t n_slots = (size_t)std::max<int64_t>(1, n_blocks) * 32;
  FB_CUDA(e, tab.units.ensure(n_slots * U));
  FB_CUDA(e, tab.rowno.ensure(n_slots));
  FB_CUDA(e, tab.list_blk.ensure(n_lists));
  FB_CUDA(e, tab.list_len.ensure(n_lists));
  FB_CUDA(e, d_row_start.ensure(n_lists));
  if (want8) FB_CUDA(e, tab.units8.ensure(n_slots));
  FB_CUDA(e, cudaMemsetAsync(tab.units.p, 0, n_slots * U * sizeof(uint2), st));
  FB_CUDA(e, cudaMemsetAsync(tab.rowno.p, 0xff, n_slots * sizeof(int32_t), st));
  if (want8) FB_CUDA(e, cudaMemsetAsync(tab.units8.p, 0, n_slots * sizeof(uint4), st));
  FB_CUDA(e, cudaMemcpyAsync(tab.list_blk.p, blk.data(), (size_t)n_lists * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  FB_CUDA(e, cudaMemcpyAsync(tab.list_len.p, len.data(), (size_t)n_lists * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  FB_CUDA(e, cudaMemcpyAsync(d_row_start.p, row_start.data(), (size_t)n_lists * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  tr.lap("table alloc + clear");
  const int32_t* d_seq = nullptr;          // rows in list order (nullptr: table order is list order)
  if (!pseudo) {
    // stable sort of the rows by list = arrival order inside every list
    int bits = 1;
    while ((1ll << bits) < n_lists) bits++;
    FB_CUDA(e, d_iota.ensure((size_t)N));
    FB_CUDA(e, d_keys.ensure((size_t)N));
    FB_CUDA(e, d_arrival.ensure((size_t)N));
    iota_i32_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(d_iota.p, N);
    size_t tmp_bytes = 0;
    FB_CUDA(e, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_list.p, d_keys.p, d_iota.p, d_arrival.p, (int)N, 0, bits, st));
    DevBuf<unsigned char> d_tmp;
    FB_CUDA(e, d_tmp.ensure(tmp_bytes));
    FB_CUDA(e, cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, d_list.p, d_keys.p, d_iota.p, d_arrival.p, (int)N, 0, bits, st));
    e->launches += 2;
    d_seq = d_arrival.p;
    tr.lap("stable sort by list");
    const size_t smem = place_rows_smem(m, K, std::max(32, placement_window));
    if (placement_window > 1 && K <= 65536 && smem <= std::min<size_t>(e->smem_optin, 160 * 1024)) {
      const int threads = std::min(1024, (std::max(32, std::max(placement_window, m)) + 31) / 32 * 32);
      FB_CUDA(e, d_order.ensure((size_t)N));
      FB_CUDA(e, cudaFuncSetAttribute(place_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
      place_rows_kernel<<<n_lists, threads, smem, st>>>(d_codes.p, m, K, d_arrival.p, d_row_start.p, tab.list_len.p, placement_window, d_order.p);
      e->launches++;
      d_seq = d_order.p;
    }
    FB_CUDA(e, cudaStreamSynchronize(st));   // d_tmp goes out of scope
    tr.lap("placement");
  }
  pack_rows_kernel<<<(unsigned)((N + 255) / 256), 256, 0, st>>>(d_codes.p, m, U, d_seq, pseudo ? nullptr : d_list.p, rows_per_pseudo_list,
                                                               d_row_start.p, tab.list_blk.p, N, tab.units.p, want8 ? tab.units8.p : nullptr,
                                                               tab.rowno.p);
  e->launches++;
  FB_CUDA(e, cudaGetLastError());
  FB_CUDA(e, cudaStreamSynchronize(st));     // the staging buffers are freed on return
  tr.lap("pack");
  tab.has8 = want8;
  tab.m = m; tab.U = U; tab.n_lists = n_lists; tab.N = N; tab.n_blocks = n_blocks;
  tab.h_list_len = len;
  tab.h_list_blk = blk;
  tab.K = K;
  tab.max_id = diag[2];
  tab.loaded = true;
  return FB_OK;
}

int build_table_host(fb_engine* e, CodeTable& tab, const int32_t* ids, const int32_t* list_of_row,
                     int n_lists, int rows_per_pseudo_list, const int16_t* codes, int64_t N, int m, int K,
                     int placement_window) {
  if (N < 0 || m <= 0) return fail(e, FB_ERR_INVALID, "bad table shape N=%lld m=%d", (long long)N, m);
  if (N >= (1ll << 31) - 64) return fail(e, FB_ERR_UNSUPPORTED, "table too large");
  const int U = (m + 3) / 4;
  if (list_of_row == nullptr) n_lists = (int)std::max<int64_t>(1, (N + rows_per_pseudo_list - 1) / rows_per_pseudo_list);
  std::vector<int32_t> len(n_lists, 0), blk(n_lists, 0);
  for (int64_t r = 0; r < N; r++) {
    int c = list_of_row ? list_of_row[r] : (int)(r / rows_per_pseudo_list);
    if (c < 0 || c >= n_lists) return fail(e, FB_ERR_INVALID, "row %lld: coarse_id %d out of range [0,%d)", (long long)r, c, n_lists);
    len[c]++;
  }
  int64_t n_blocks = 0;
  for (int c = 0; c < n_lists; c++) { blk[c] = (int32_t)n_blocks; n_blocks += (len[c] + 31) / 32; }
  std::vector<uint2> units((size_t)std::max<int64_t>(1, n_blocks) * U * 32, make_uint2(0, 0));
  const bool want8 = K <= 256 && m <= 16;   // true uint8 code table: what index_creation/config/*_config.json (k = 256) produce
  std::vector<uint4> units8(want8 ? (size_t)std::max<int64_t>(1, n_blocks) * 32 : 0, make_uint4(0, 0, 0, 0));
  std::vector<int32_t> rowno((size_t)std::max<int64_t>(1, n_blocks) * 32, -1);
  for (int64_t r = 0; r < N; r++)
    for (int p = 0; p < m; p++) {
      const int code = codes[(size_t)r * m + p];
      if (code < 0 || code >= K) return fail(e, FB_ERR_INVALID, "row %lld pos %d: code %d out of range [0,%d)", (long long)r, p, code, K);
    }
  // slot of every row inside its list: arrival order, or the conflict-aware placement
  std::vector<int32_t> slot_of((size_t)std::max<int64_t>(1, N));
  {
    std::vector<int32_t> cursor(n_lists, 0);
    for (int64_t r = 0; r < N; r++) {
      int c = list_of_row ? list_of_row[r] : (int)(r / rows_per_pseudo_list);
      slot_of[r] = cursor[c]++;
    }
  }
  if (placement_window > 1 && list_of_row != nullptr && K <= 65536) {
    std::vector<std::vector<int32_t>> rows_of(n_lists);
    for (int c = 0; c < n_lists; c++) rows_of[c].reserve(len[c]);
    for (int64_t r = 0; r < N; r++) rows_of[list_of_row[r]].push_back((int32_t)r);
    const int n_threads = (int)std::max(1u, std::min(32u, std::thread::hardware_concurrency()));
    std::atomic<int> next(0);
    auto worker = [&]() {
      std::vector<int32_t> order;
      for (int c = next.fetch_add(1); c < n_lists; c = next.fetch_add(1)) {
        place_rows_of_list(codes, m, K, rows_of[c], placement_window, order);
        for (size_t s = 0; s < order.size(); s++) slot_of[order[s]] = (int32_t)s;
      }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; t++) pool.emplace_back(worker);
    worker();
    for (auto& t : pool) t.join();
  }
  for (int64_t r = 0; r < N; r++) {
    int c = list_of_row ? list_of_row[r] : (int)(r / rows_per_pseudo_list);
    int slot = slot_of[r];
    int64_t b = blk[c] + slot / 32;
    int L = slot % 32;
    const int16_t* cr = codes + (size_t)r * m;
    for (int u = 0; u < U; u++) {
      uint32_t f[4] = {0, 0, 0, 0};
      for (int t = 0; t < 4; t++) {
        int p = 4 * u + t;
        if (p < m) {
          int code = cr[p];
          f[t] = (uint32_t)code * 4u;  // pre-scaled: byte offset into a K-float LUT row
        }
      }
      units[((size_t)b * U + u) * 32 + L] = make_uint2(f[0] | (f[1] << 16), f[2] | (f[3] << 16));
    }
    if (want8) {
      uint32_t wds[4] = {0, 0, 0, 0};
      for (int p2 = 0; p2 < m; p2++) wds[p2 >> 2] |= (uint32_t)(uint8_t)cr[p2] << (8 * (p2 & 3));
      units8[(size_t)b * 32 + L] = make_uint4(wds[0], wds[1], wds[2], wds[3]);
    }
    rowno[(size_t)b * 32 + L] = (int32_t)r;
  }
  FB_CUDA(e, tab.units.ensure(units.size()));
  FB_CUDA(e, tab.rowno.ensure(rowno.size()));
  FB_CUDA(e, tab.list_blk.ensure(n_lists));
  FB_CUDA(e, tab.list_len.ensure(n_lists));
  FB_CUDA(e, tab.ids.ensure((size_t)std::max<int64_t>(1, N)));
  FB_CUDA(e, cudaMemcpy(tab.units.p, units.data(), units.size() * sizeof(uint2), cudaMemcpyHostToDevice));
  tab.has8 = want8;
  if (want8) {
    FB_CUDA(e, tab.units8.ensure(units8.size()));
    FB_CUDA(e, cudaMemcpy(tab.units8.p, units8.data(), units8.size() * sizeof(uint4), cudaMemcpyHostToDevice));
  }
  FB_CUDA(e, cudaMemcpy(tab.rowno.p, rowno.data(), rowno.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  FB_CUDA(e, cudaMemcpy(tab.list_blk.p, blk.data(), n_lists * sizeof(int32_t), cudaMemcpyHostToDevice));
  FB_CUDA(e, cudaMemcpy(tab.list_len.p, len.data(), n_lists * sizeof(int32_t), cudaMemcpyHostToDevice));
  if (N > 0) FB_CUDA(e, cudaMemcpy(tab.ids.p, ids, (size_t)N * sizeof(int32_t), cudaMemcpyHostToDevice));
  tab.m = m; tab.U = U; tab.n_lists = n_lists; tab.N = N; tab.n_blocks = n_blocks;
  tab.h_list_len = len;
  tab.h_list_blk = blk;
  tab.K = K;
  tab.max_id = 0;
  for (int64_t r = 0; r < N; r++) tab.max_id = std::max(tab.max_id, ids[r]);
  tab.loaded = true;
  return FB_OK;
}

__global__ void iota_kernel(int32_t* out, int n) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = i;
}

__global__ void count_rows_kernel(const int32_t* __restrict__ probes, int n, const int32_t* __restrict__ list_len,
                                  u64* __restrict__ counter) {
  u64 local = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) local += (u64)list_len[probes[i]];
  for (int s = 16; s >= 1; s >>= 1) local += __shfl_xor_sync(0xffffffffu, local, s);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(counter, local);
}

// ---- kernel launch helpers -------------------------------------------------
template <int QT>
int launch_coarse_t(fb_engine* e, const float* d_q, int nq, int w, int k, int64_t q0) {
  size_t smem = ((size_t)e->d * QT + (size_t)QT * e->Cs) * sizeof(float);
  auto kern = e->packed_fp32 ? coarse_select_kernel_t<QT, true> : coarse_select_kernel_t<QT, false>;
  FB_CUDA(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<(nq + QT - 1) / QT, kCoarseThreads, smem, e->stream>>>(
      (e->coarse_src ? e->coarse_src : d_q) + (size_t)q0 * e->d, nq, e->d, e->coarseT.p, e->C, e->Cs, e->fine.list_len.p, w, k,
      e->probes.p + (size_t)q0 * w, e->qflags.p + q0, e->force_exact ? 1 : 0, e->one,
      e->coarse_src ? const_cast<float*>(d_q) + (size_t)q0 * e->d : nullptr);
  e->launches++;
  FB_CUDA(e, cudaGetLastError());
  return FB_OK;
}

int launch_coarse_range(fb_engine* e, const float* d_q, int nq, int w, int k, int64_t q0) {
  StageTimer t(e, ST_COARSE);
  if (nq <= 32 && (size_t)e->d * sizeof(float) <= 48 * 1024) {
    // small batch: spread each query's C chains over C/128 CTAs, then one warp per query selects
    if (e->coarse_dist.ensure((size_t)nq * e->Cs) != cudaSuccess) return fail(e, FB_ERR_CUDA, "out of device memory");
    dim3 grid((e->Cs + kCoarseSmallThreads - 1) / kCoarseSmallThreads, nq);
    coarse_dist_small_kernel<<<grid, kCoarseSmallThreads, (size_t)e->d * sizeof(float), e->stream>>>(
        d_q + (size_t)q0 * e->d, e->d, e->coarseT.p, e->Cs, e->coarse_dist.p);
    coarse_select_small_kernel<<<(nq + 3) / 4, 128, 0, e->stream>>>(e->coarse_dist.p, nq, e->C, e->Cs, e->fine.list_len.p, w, k,
                                                                  e->probes.p + (size_t)q0 * w, e->qflags.p + q0,
                                                                  e->force_exact ? 1 : 0);
    e->launches += 2;
    FB_CUDA(e, cudaGetLastError());
    return FB_OK;
  }
  auto need = [&](int QT) { return ((size_t)e->d * QT + (size_t)QT * e->Cs) * sizeof(float); };
  const size_t two_per_sm = 100 * 1024;
  if (need(16) <= two_per_sm) {
    // all tiles cost the same, so the launch takes ceil(tiles / resident CTAs) rounds of one tile time
    // (about QT + 1.5 units): pick the tile height that wastes the least of the last round
    const int slots = 2 * e->num_sms;
    auto cost = [&](int QT) { const int tiles = (nq + QT - 1) / QT; return (double)((tiles + slots - 1) / slots) * (QT + 1.5); };
    const double c16 = cost(16), c12 = cost(12), c8 = cost(8);
    if (c12 < c16 && c12 <= c8) return launch_coarse_t<12>(e, d_q, nq, w, k, q0);
    if (c8 < c16) return launch_coarse_t<8>(e, d_q, nq, w, k, q0);
    return launch_coarse_t<16>(e, d_q, nq, w, k, q0);
  }
  if (need(8) <= two_per_sm) return launch_coarse_t<8>(e, d_q, nq, w, k, q0);
  if (need(4) <= e->smem_optin) return launch_coarse_t<4>(e, d_q, nq, w, k, q0);
  if (need(1) <= e->smem_optin) return launch_coarse_t<1>(e, d_q, nq, w, k, q0);
  return fail(e, FB_ERR_UNSUPPORTED, "coarse table too large for shared memory (C=%d d=%d)", e->C, e->d);
}

int launch_coarse(fb_engine* e, const float* d_q, int nq, int w, int k) { return launch_coarse_range(e, d_q, nq, w, k, 0); }

template <int W, int TKS, bool PACKED>
int launch_lut_cfg(fb_engine* e, const Codebook& cb, const float* d_q, const float* d_coarse, const int32_t* d_probes,
                   int jobs_per_query, int njobs, float* d_lut, int TK, size_t smem) {
  const int tiles = (cb.K + TK - 1) / TK;
  const int per_sm = TKS > 0 ? 1024 / TKS : 1;
  const int ctas_per_sm = (e->lut_ctas_per_sm > 0) ? std::min(e->lut_ctas_per_sm, per_sm) : per_sm;
  int groups = std::max(1, e->num_sms * ctas_per_sm / std::max(1, cb.m * tiles));
  groups = std::min(groups, (njobs + W - 1) / W);
  dim3 grid(cb.m * tiles, groups);
  auto kern = lut_build_kernel<W, TKS, PACKED>;
  FB_CUDA(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // e->one: run-time 1.0f (common.cuh: keeps ptxas from contracting the packed chain)
  kern<<<grid, TK, smem, e->stream>>>(d_q, cb.m * cb.sub, d_coarse, d_probes, jobs_per_query, njobs, cb.cbT.p, cb.m, cb.K, cb.sub,
                                      TK, d_lut, e->one);
  e->launches++;
  FB_CUDA(e, cudaGetLastError());
  return FB_OK;
}

template <int W>
int launch_lut_w(fb_engine* e, const Codebook& cb, const float* d_q, const float* d_coarse, const int32_t* d_probes,
                 int jobs_per_query, int njobs, float* d_lut) {
  const int K = cb.K, sub = cb.sub;
  constexpr int WS = (W + 3) & ~3;
  const size_t rs_bytes = 2 * (size_t)sub * WS * sizeof(float);
  // preferred: TKS codes per CTA, 1024/TKS CTAs per SM, constant row stride
  const int tks = e->lut_tile;
  if ((tks == 256 || tks == 512 || tks == 1024) && K % 4 == 0) {
    const int TK = std::min(tks, (K + 31) / 32 * 32);
    const size_t smem = (size_t)sub * tks * sizeof(float) + rs_bytes;
    const size_t per_sm_budget = (e->smem_optin - 2048) / (1024 / tks);
    if (smem + 1024 <= per_sm_budget) {
#define FB_LUT_GO(T_) \
  return e->packed_fp32 ? launch_lut_cfg<W, T_, true>(e, cb, d_q, d_coarse, d_probes, jobs_per_query, njobs, d_lut, TK, smem) \
                        : launch_lut_cfg<W, T_, false>(e, cb, d_q, d_coarse, d_probes, jobs_per_query, njobs, d_lut, TK, smem)
      if (tks == 256) FB_LUT_GO(256);
      if (tks == 512) FB_LUT_GO(512);
      FB_LUT_GO(1024);
#undef FB_LUT_GO
    }
  }
  // generic: shrink the code tile until the slice fits
  const size_t budget = std::min<size_t>(e->smem_optin, 200 * 1024) - 4096;
  int TK = std::min(1024, (K + 31) / 32 * 32);
  while ((size_t)sub * TK * sizeof(float) + rs_bytes > budget && TK > 32) TK -= 32;
  if ((size_t)sub * TK * sizeof(float) + rs_bytes > budget)
    return fail(e, FB_ERR_UNSUPPORTED, "sub-vector too long for shared memory (sub=%d)", sub);
  return launch_lut_cfg<W, 0, false>(e, cb, d_q, d_coarse, d_probes, jobs_per_query, njobs, d_lut, TK,
                                     (size_t)sub * TK * sizeof(float) + rs_bytes);
}

int launch_lut(fb_engine* e, const Codebook& cb, const float* d_q, const float* d_coarse, const int32_t* d_probes,
               int jobs_per_query, int njobs, float* d_lut) {
  StageTimer t(e, ST_LUT);
  int W;
  if (jobs_per_query >= kLutMaxJobs) W = kLutMaxJobs;
  else W = jobs_per_query * (kLutMaxJobs / jobs_per_query);
  if (njobs < W) W = njobs;
  switch (W) {
#define FB_LUT_CASE(n) case n: return launch_lut_w<n>(e, cb, d_q, d_coarse, d_probes, jobs_per_query, njobs, d_lut);
    FB_LUT_CASE(1) FB_LUT_CASE(2) FB_LUT_CASE(3) FB_LUT_CASE(4) FB_LUT_CASE(5) FB_LUT_CASE(6) FB_LUT_CASE(7)
    FB_LUT_CASE(8) FB_LUT_CASE(9) FB_LUT_CASE(10) FB_LUT_CASE(11) FB_LUT_CASE(12) FB_LUT_CASE(13)
    FB_LUT_CASE(14) FB_LUT_CASE(15) FB_LUT_CASE(16)
#undef FB_LUT_CASE
  }
  return fail(e, FB_ERR_INVALID, "bad LUT job width %d", W);
}

template <int M>
int launch_scan_m(fb_engine* e, const CodeTableDev& tab, const int32_t* d_task_list, int ntasks, int tasks_per_lut,
                  int list_mod, int segs, const float* d_lut, int K, int KK, u64* d_partial, float sentinel) {
  size_t smem = std::max<size_t>((size_t)tab.m * K * sizeof(float), (size_t)kScanWarps * 32 * sizeof(u64));
  if (smem > e->smem_optin - 1024) return fail(e, FB_ERR_UNSUPPORTED, "LUT (%zu bytes) exceeds shared memory", smem);
  auto kern = adc_scan_kernel<M>;
  FB_CUDA(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<ntasks * segs, kScanThreads, smem, e->stream>>>(tab, d_task_list, tasks_per_lut, list_mod, segs, d_lut, K, KK, d_partial, sentinel);
  e->launches++;
  e->n_scan_launches++;
  FB_CUDA(e, cudaGetLastError());
  return FB_OK;
}

int launch_scan(fb_engine* e, const CodeTableDev& tab, const int32_t* d_task_list, int ntasks, int tasks_per_lut,
                int list_mod, int segs, const float* d_lut, int K, int KK, u64* d_partial, float sentinel) {
  StageTimer t(e, ST_SCAN);
  switch (tab.m) {
    case 8: return launch_scan_m<8>(e, tab, d_task_list, ntasks, tasks_per_lut, list_mod, segs, d_lut, K, KK, d_partial, sentinel);
    case 12: return launch_scan_m<12>(e, tab, d_task_list, ntasks, tasks_per_lut, list_mod, segs, d_lut, K, KK, d_partial, sentinel);
    case 16: return launch_scan_m<16>(e, tab, d_task_list, ntasks, tasks_per_lut, list_mod, segs, d_lut, K, KK, d_partial, sentinel);
    default: return launch_scan_m<0>(e, tab, d_task_list, ntasks, tasks_per_lut, list_mod, segs, d_lut, K, KK, d_partial, sentinel);
  }
}

template <int M, int KC, bool C8 = false>
int launch_qscan_mk(fb_engine* e, const CodeTableDev& tab, int q0, int nq, int w, const float* d_lut, int K, int KK, int k,
                    float sentinel, int32_t* d_out_ids, float* d_out_dists) {
  size_t smem = std::max<size_t>(2 * (size_t)tab.m * K * sizeof(float), kQScanWarps * 32 * sizeof(u64));
  if (smem > e->smem_optin - 1024) return FB_ERR_UNSUPPORTED;  // caller falls back to one list per CTA
  auto kern = adc_scan_query_kernel<M, KC, C8>;
  FB_CUDA(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<nq, kQScanThreads, smem, e->stream>>>(tab, e->probes.p + (size_t)q0 * w, w, d_lut, K, KK, k, sentinel,
                                               e->qflags.p + q0, d_out_ids, d_out_dists, e->exact_list.p,
                                               e->small.p + 0, e->counters64.p + 1, e->kth.p + q0, q0);
  e->launches++;
  e->n_scan_launches++;
  FB_CUDA(e, cudaGetLastError());
  return FB_OK;
}

// throughput form: one CTA per query, finalize fused
int launch_qscan(fb_engine* e, const CodeTableDev& tab, int q0, int nq, int w, const float* d_lut, int K, int KK, int k,
                 float sentinel, int32_t* oi, float* od) {
  StageTimer t(e, ST_SCAN);
#define FB_QS(M_, K_) return launch_qscan_mk<M_, K_>(e, tab, q0, nq, w, d_lut, K, KK, k, sentinel, oi, od)
  if (K == 1024) {
    if (tab.m == 12) FB_QS(12, 1024);
    if (tab.m == 8) FB_QS(8, 1024);
    if (tab.m == 16) FB_QS(16, 1024);
  } else if (K == 256) {
    if (tab.m == 12 && tab.units8 != nullptr && e->byte_codes) {
      e->used8 = true;
      return launch_qscan_mk<12, 256, true>(e, tab, q0, nq, w, d_lut, K, KK, k, sentinel, oi, od);
    }
    if (tab.m == 12) FB_QS(12, 256);
    if (tab.m == 8) FB_QS(8, 256);
    if (tab.m == 16) FB_QS(16, 256);
  }
  FB_QS(0, 0);
#undef FB_QS
}

// ivfadc_batch_search: queries whose k-th slot is still empty after the first round and that are not flagged yet
__global__ void batch_unfilled_kernel(const int32_t* __restrict__ out_ids, int nq, int k, uint32_t* __restrict__ qflags,
                                      int32_t* __restrict__ exact_list, int32_t* __restrict__ exact_count, u64* __restrict__ exact_total) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= nq) return;
  const uint32_t f = qflags[q];
  if ((f & kFlagExact) || out_ids[(size_t)q * k + k - 1] != -1) return;
  qflags[q] = f | kFlagExact | kWhyFewRows;
  exact_list[atomicAdd(exact_count, 1)] = q;
  atomicAdd(exact_total, 1ull);
  atomicAdd(exact_total + 3, 1ull);
}

__global__ void collect_flagged_kernel(const uint32_t* __restrict__ qflags, int n, int q_base, int32_t* __restrict__ exact_list,
                                       int32_t* __restrict__ exact_count, u64* __restrict__ exact_total) {
  const int q = blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= n) return;
  const uint32_t f = qflags[q];
  if (f & kFlagExact) {
    exact_list[atomicAdd(exact_count, 1)] = q + q_base;
    atomicAdd(exact_total, 1ull);
    for (int b = 1; b <= 5; b++)
      if (f & (1u << b)) atomicAdd(exact_total + b, 1ull);
  }
}

template <int M, int KC>
int launch_scan_keys_mk(fb_engine* e, const CodeTable& tab, const int32_t* d_probes, int nq, int w, const float* d_lut, int K,
                        u64* d_keys, size_t stride, int32_t* d_nkeys) {
  size_t smem = 2 * (size_t)tab.m * K * sizeof(float);
  if (smem > e->smem_optin - 1024) return fail(e, FB_ERR_UNSUPPORTED, "LUT too large for shared memory");
  auto kern = adc_scan_keys_kernel<M, KC>;
  FB_CUDA(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  kern<<<nq, kQScanThreads, smem, e->stream>>>(tab.dev(), d_probes, w, d_lut, K, d_keys, stride, d_nkeys);
  e->launches++;
  e->n_scan_launches++;
  FB_CUDA(e, cudaGetLastError());
  return FB_OK;
}

int launch_scan_keys(fb_engine* e, const CodeTable& tab, const int32_t* d_probes, int nq, int w, const float* d_lut, int K,
                     u64* d_keys, size_t stride, int32_t* d_nkeys) {
  StageTimer t(e, ST_SCAN);
  if (K == 1024 && tab.m == 12) return launch_scan_keys_mk<12, 1024>(e, tab, d_probes, nq, w, d_lut, K, d_keys, stride, d_nkeys);
  if (K == 256 && tab.m == 12) return launch_scan_keys_mk<12, 256>(e, tab, d_probes, nq, w, d_lut, K, d_keys, stride, d_nkeys);
  return launch_scan_keys_mk<0, 0>(e, tab, d_probes, nq, w, d_lut, K, d_keys, stride, d_nkeys);
}

int launch_finalize(fb_engine* e, const CodeTableDev& tab, int q0, int lists_per_query, int KK, int k, int nq, float sentinel,
                    bool has_input_flags, int32_t* d_out_ids, float* d_out_dists) {
  StageTimer t(e, ST_FINALIZE);
  const int warps = 8;
  finalize_kernel<<<(nq + warps - 1) / warps, warps * 32, 0, e->stream>>>(
      e->partial.p, lists_per_query, KK, k, nq, tab.ids, sentinel, e->qflags.p + q0, has_input_flags ? 1 : 0, d_out_ids,
      d_out_dists, e->exact_list.p, e->small.p + 0, e->counters64.p + 1, e->kth.p + q0, q0);
  e->launches++;
  FB_CUDA(e, cudaGetLastError());
  return FB_OK;
}

int check_common(fb_engine* e, int nq, int k) {
  if (!e) return FB_ERR_INVALID;
  if (nq < 0) return fail(e, FB_ERR_INVALID, "nq < 0");
  if (k < 1 || k > kExactMaxK) return fail(e, FB_ERR_UNSUPPORTED, "k=%d outside [1,%d]", k, kExactMaxK);
  return FB_OK;
}

size_t exact_smem_bytes(const fb_engine* e, int w) {
  return kExactFixedSmem + sizeof(float) * e->Cs + sizeof(float) * ((e->d + 3) & ~3) +
         (sizeof(float) + sizeof(int)) * ((w + 3) & ~3) + (size_t)e->C + 16;
}

// ---- throughput form: warp-specialised pipeline (pipeline_kernels.cuh) ------
// Beat c (= one launch) builds the LUTs of chunk c+1 and scans chunk c; two LUT
// buffers alternate.  nchunks + 1 launches on the engine stream.
template <int M, int KC, int SUB, class Cfg, bool C8 = false>
int run_pipeline_t(fb_engine* e, const Codebook& cb, const float* d_q, int nq, int k, int w, int KK, float sentinel,
                   int32_t* d_out_ids, float* d_out_dists, int64_t chunk) {
  using L = PipeSmem<M, KC, SUB, Cfg>;
  auto kern = ivfadc_pipe_kernel<M, KC, SUB, Cfg, C8>;
  FB_CUDA(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)L::total));
  // Chunk boundaries.  The first launch only builds LUTs and the last one only scans, so the chunks ramp
  // up at the start and down at the end (quarter, half, full ... half, quarter of `chunk`): the two
  // un-overlapped launches become short.
  std::vector<int64_t> bounds{0};
  {
    const int64_t ramp_up[2] = {std::max<int64_t>(64, chunk / 4), std::max<int64_t>(64, chunk / 2)};
    const int64_t tail = ramp_up[0] + ramp_up[1];
    int64_t pos = 0;
    if (e->pipe_ramp && nq >= 2 * tail + chunk) {
      for (int i = 0; i < 2; i++) { pos += ramp_up[i]; bounds.push_back(pos); }
      while (nq - pos > tail + chunk) { pos += chunk; bounds.push_back(pos); }
      const int64_t mid = nq - pos - tail;            // in (0, chunk]
      if (mid > 0) { pos += mid; bounds.push_back(pos); }
      pos += ramp_up[1]; bounds.push_back(pos);
      bounds.push_back(nq);
    } else {
      while (pos < nq) { pos = std::min<int64_t>(nq, pos + chunk); bounds.push_back(pos); }
    }
  }
  const int nchunks = (int)bounds.size() - 1;
  const size_t lut_per_query = (size_t)w * M * KC;
  FB_CUDA(e, e->lut.ensure((size_t)chunk * lut_per_query));
  FB_CUDA(e, e->lut2.ensure((size_t)chunk * lut_per_query));
  FB_CUDA(e, e->pipe_counters.ensure((size_t)nchunks + 1));
  FB_CUDA(e, cudaMemsetAsync(e->pipe_counters.p, 0, ((size_t)nchunks + 1) * sizeof(int32_t), e->stream));
  const int tiles = (KC + kPipeTile - 1) / kPipeTile;
  PipeArgs a;
  memset(&a, 0, sizeof a);
  a.jobs_per_query = w;
  a.coarse = e->coarse.p;
  a.cbT = cb.cbT.p;
  a.d = e->d;
  a.K = KC;
  a.tiles = tiles;
  a.n_slices = M * tiles;
  a.n_groups = std::max(1, e->num_sms / a.n_slices);
  a.one = e->one;
  a.tab = e->fine.dev();
  a.w = w; a.KK = KK; a.k = k;
  a.sentinel = sentinel;
  a.exact_list = e->exact_list.p;
  a.exact_count = e->small.p + 0;
  a.exact_total = e->counters64.p + 1;
  const int grid = std::max(e->num_sms, a.n_slices);   // one CTA per SM; producers beyond n_groups * n_slices idle
  if (grid > e->num_sms) a.n_groups = 1;
  for (int c = -1; c < nchunks; c++) {
    // producer half: chunk c + 1
    const int64_t p0 = (c + 1 < nchunks) ? bounds[c + 1] : nq;
    const int pn = (c + 1 < nchunks) ? (int)(bounds[c + 2] - bounds[c + 1]) : 0;
    a.lut_queries = d_q + (size_t)p0 * e->d;
    a.lut_probes = e->probes.p + (size_t)p0 * w;
    a.lut_out = ((c + 1) & 1) ? e->lut2.p : e->lut.p;
    a.lut_njobs = pn * w;
    // scan half: chunk c
    const int64_t s0 = (c >= 0) ? bounds[c] : 0;
    const int sn = (c >= 0) ? (int)(bounds[c + 1] - bounds[c]) : 0;
    a.scan_probes = e->probes.p + (size_t)s0 * w;
    a.scan_lut = (c & 1) ? e->lut2.p : e->lut.p;
    a.scan_nq = sn;
    a.qflags = e->qflags.p + s0;
    a.out_ids = d_out_ids + (size_t)s0 * k;
    a.out_dists = d_out_dists + (size_t)s0 * k;
    a.kth_key = e->kth.p + s0;
    a.q_base = (int)s0;
    a.work_counter = e->pipe_counters.p + (c + 1);
    if (e->pipe_debug == 1) a.scan_nq = 0;
    if (e->pipe_debug == 2) a.lut_njobs = 0;
    StageTimer t(e, ST_PIPE);
    kern<<<grid, Cfg::kThreads, L::total, e->stream>>>(a);
    e->launches++;
    e->n_pipe_launches++;
    FB_CUDA(e, cudaGetLastError());
  }
  return FB_OK;
}

// FB_ERR_UNSUPPORTED: shape not covered by the pipeline kernel (caller uses the separate kernels)
int run_pipeline(fb_engine* e, const Codebook& cb, const float* d_q, int nq, int k, int w, int KK, float sentinel,
                 int32_t* d_out_ids, float* d_out_dists) {
  if (!e->pipeline || nq < 512 || cb.m != 12 || cb.sub != 25) return FB_ERR_UNSUPPORTED;
  int64_t chunk = std::max<int64_t>(128, std::min<int64_t>(e->pipe_chunk, (nq + 3) / 4));
#define FB_PIPE_GO(K_, ...)                                                                              \
  do {                                                                                                   \
    if (PipeSmem<12, K_, 25, PipeCfg<__VA_ARGS__>>::total > e->smem_optin) return FB_ERR_UNSUPPORTED;    \
    return run_pipeline_t<12, K_, 25, PipeCfg<__VA_ARGS__>>(e, cb, d_q, nq, k, w, KK, sentinel, d_out_ids, \
                                                            d_out_dists, chunk);                         \
  } while (0)
  if (cb.K == 1024) {
    switch (e->pipe_shape) {   // (producer warps, scan warps[, jobs per producer thread]): tuning knob, FB_OPT_PIPE_SHAPE
      case 1: FB_PIPE_GO(1024, 8, 18);
      case 2: FB_PIPE_GO(1024, 8, 14, 16);
      case 3: FB_PIPE_GO(1024, 16, 10);      // LUT-bound indexes (probed lists near the nominal N*w/C rows): more producers
      default: FB_PIPE_GO(1024, 12, 14);
    }
  }
  if (cb.K == 256) {
    if (e->fine.has8 && e->byte_codes) {
      if (PipeSmem<12, 256, 25, PipeCfg<12, 14>>::total > e->smem_optin) return FB_ERR_UNSUPPORTED;
      e->used8 = true;
      return run_pipeline_t<12, 256, 25, PipeCfg<12, 14>, true>(e, cb, d_q, nq, k, w, KK, sentinel, d_out_ids, d_out_dists, chunk);
    }
    FB_PIPE_GO(256, 12, 14);
  }
#undef FB_PIPE_GO
  return FB_ERR_UNSUPPORTED;
}

// ---- the IVFADC pipeline on device pointers --------------------------------
// batch_mode: ivfadc_batch_search's loop (one list per round, sentinel 100.0, rounds until k rows were ADMITTED)
int ivfadc_dev(fb_engine* e, const float* d_q, int nq, int k, int w, int32_t* d_out_ids, float* d_out_dists,
               float sentinel = 1000.0f, bool batch_mode = false) {
  int rc = check_common(e, nq, k);
  if (rc) return rc;
  if (!e->coarse_loaded || !e->cb[FB_CB_RESIDUAL].loaded || !e->fine.loaded)
    return fail(e, FB_ERR_INVALID, "IVFADC index not loaded (coarse / residual codebook / fine table)");
  const Codebook& cb = e->cb[FB_CB_RESIDUAL];
  if (cb.m != e->fine.m) return fail(e, FB_ERR_INVALID, "codebook m=%d but fine table m=%d", cb.m, e->fine.m);
  if (cb.m * cb.sub != e->d) return fail(e, FB_ERR_INVALID, "m*sub=%d != d=%d", cb.m * cb.sub, e->d);
  if (w < 1) return fail(e, FB_ERR_INVALID, "w=%d", w);
  if (w > e->C) return fail(e, FB_ERR_REFERENCE_UB, "w=%d > %d coarse centroids: the reference indexes cq[-1] (freddy.c:296-302)", w, e->C);
  if (nq == 0) return FB_OK;
  const int m = cb.m, K = cb.K;
  e->used8 = false;
  const bool fast = (k <= 30 && w <= 31);
  const bool large_k = (!fast && w <= 31);   // k in 31..1024: materialised-key path (scan -> keys -> reference top-k)
  const int KK = k + 2;   // k + 2 keys: enough to settle a boundary tie in the merge (warp_emit_topk)
  int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(e->query_chunk, nq));
  const size_t lut_per_query = (size_t)w * m * K;
  size_t key_cap = 0;   // keys one query can produce: the w longest lists
  if (large_k) {
    std::vector<int32_t> lens = e->fine.h_list_len;
    std::sort(lens.begin(), lens.end(), std::greater<int32_t>());
    for (int j = 0; j < w && j < (int)lens.size(); j++) key_cap += (size_t)lens[j];
    key_cap = std::max<size_t>(key_cap, 1);
    const size_t budget = (size_t)1 << 30;
    chunk = std::max<int64_t>(1, std::min<int64_t>(chunk, (int64_t)(budget / (key_cap * sizeof(u64)))));
  }

  // per-query products of the streaming pass live for the whole call (the general
  // kernel runs once at the end); only the LUT scratch is per chunk
  FB_CUDA(e, e->probes.ensure((size_t)nq * w));
  FB_CUDA(e, e->qflags.ensure((size_t)nq));
  FB_CUDA(e, e->exact_list.ensure((size_t)nq));
  FB_CUDA(e, e->kth.ensure((size_t)nq));
  if (fast || large_k) FB_CUDA(e, e->lut.ensure((size_t)chunk * lut_per_query));
  if (fast && chunk < e->qscan_min_queries) FB_CUDA(e, e->partial.ensure((size_t)chunk * w * 16 * KK));
  if (large_k) {
    FB_CUDA(e, e->j_keys.ensure((size_t)chunk * key_cap));
    FB_CUDA(e, e->j_ncells.ensure((size_t)chunk));
  }
  // general-kernel scratch: one LUT set per resident CTA
  int exact_ctas = 2 * e->num_sms;
  const size_t exact_budget = (size_t)1 << 30;
  while (exact_ctas > 1 && (size_t)exact_ctas * lut_per_query * sizeof(float) > exact_budget) exact_ctas /= 2;
  FB_CUDA(e, e->exact_lut.ensure((size_t)exact_ctas * lut_per_query));
  size_t ex_smem = exact_smem_bytes(e, w);
  if (ex_smem > e->smem_optin) return fail(e, FB_ERR_UNSUPPORTED, "general kernel needs %zu bytes of shared memory", ex_smem);
  int ex_stage = 0;  // LUT staging buffer when it fits next to the rest
  if (ex_smem + (size_t)m * K * sizeof(float) <= e->smem_optin) { ex_stage = m * K; ex_smem += (size_t)m * K * sizeof(float); }
  FB_CUDA(e, cudaFuncSetAttribute(ivfadc_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ex_smem));

  FB_CUDA(e, cudaMemsetAsync(e->small.p, 0, 2 * sizeof(int32_t), e->stream));
  if (fast) {
    if ((rc = launch_coarse(e, d_q, nq, w, k))) return rc;   // HOT(1) for the whole batch
    count_rows_kernel<<<64, 256, 0, e->stream>>>(e->probes.p, nq * w, e->fine.list_len.p, e->counters64.p + 0);
    e->launches++;
    // With overlap on, the LUT build of chunk c+1 (fp32-pipe bound) runs on its own stream while the
    // scan of chunk c (shared-memory / LSU bound) runs on another; two LUT buffers alternate.
    const bool overlap = e->overlap && nq > chunk && chunk >= e->qscan_min_queries;
    cudaStream_t main_stream = e->stream;
    rc = run_pipeline(e, cb, d_q, nq, k, w, KK, sentinel, d_out_ids, d_out_dists);
    if (rc != FB_OK && rc != FB_ERR_UNSUPPORTED) return rc;
    const bool piped = (rc == FB_OK);
    if (overlap && !piped) {
      FB_CUDA(e, e->lut2.ensure((size_t)chunk * lut_per_query));
      FB_CUDA(e, cudaEventRecord(e->ev_main, main_stream));
      FB_CUDA(e, cudaStreamWaitEvent(e->s_lut, e->ev_main, 0));
      FB_CUDA(e, cudaStreamWaitEvent(e->s_scan, e->ev_main, 0));
    }
    int c = 0;
    for (int64_t q0 = 0; q0 < nq && !piped; q0 += chunk, c++) {
      const int n = (int)std::min<int64_t>(chunk, nq - q0);
      const float* dq = d_q + (size_t)q0 * e->d;
      const int32_t* pr = e->probes.p + (size_t)q0 * w;
      int32_t* oi = d_out_ids + (size_t)q0 * k;
      float* od = d_out_dists + (size_t)q0 * k;
      float* lutbuf = (overlap && (c & 1)) ? e->lut2.p : e->lut.p;
      if (overlap) {
        if (c >= 2) FB_CUDA(e, cudaStreamWaitEvent(e->s_lut, e->ev_scan_done[c & 1], 0));   // buffer free again
        e->stream = e->s_lut;
      }
      rc = launch_lut(e, cb, dq, e->coarse.p, pr, w, n * w, lutbuf);                               // HOT(2)
      if (overlap) {
        cudaEventRecord(e->ev_lut_done[c & 1], e->s_lut);
        cudaStreamWaitEvent(e->s_scan, e->ev_lut_done[c & 1], 0);
        e->stream = e->s_scan;
      }
      if (rc) { e->stream = main_stream; return rc; }
      // HOT(3)+(4): one CTA per query when the chunk fills the GPU, else one CTA per (query, list)
      rc = (n >= e->qscan_min_queries) ? launch_qscan(e, e->fine.dev(), (int)q0, n, w, lutbuf, K, KK, k, sentinel, oi, od)
                                       : FB_ERR_UNSUPPORTED;
      if (rc == FB_ERR_UNSUPPORTED) {
        // few (query, list) tasks: split every list over `segs` CTAs so that the launch still fills the GPU
        const int segs = std::max(1, std::min(16, (2 * e->num_sms) / std::max(1, n * w)));
        rc = e->partial.ensure((size_t)chunk * w * 16 * KK) == cudaSuccess ? FB_OK : FB_ERR_CUDA;
        if (!rc) rc = launch_scan(e, e->fine.dev(), pr, n * w, 1, 1, segs, lutbuf, K, KK, e->partial.p, sentinel);
        if (!rc) rc = launch_finalize(e, e->fine.dev(), (int)q0, w * segs, KK, k, n, sentinel, true, oi, od);
      }
      if (overlap) cudaEventRecord(e->ev_scan_done[c & 1], e->s_scan);
      e->stream = main_stream;
      if (rc) return rc;
    }
    if (overlap && !piped) {
      FB_CUDA(e, cudaEventRecord(e->ev_all, e->s_scan));
      FB_CUDA(e, cudaStreamWaitEvent(main_stream, e->ev_all, 0));
    }
  } else if (large_k) {
    if ((rc = launch_coarse(e, d_q, nq, w, k))) return rc;
    count_rows_kernel<<<64, 256, 0, e->stream>>>(e->probes.p, nq * w, e->fine.list_len.p, e->counters64.p + 0);
    collect_flagged_kernel<<<(nq + 255) / 256, 256, 0, e->stream>>>(e->qflags.p, nq, 0, e->exact_list.p, e->small.p + 0,
                                                                    e->counters64.p + 1);
    e->launches += 2;
    for (int64_t q0 = 0; q0 < nq; q0 += chunk) {
      const int n = (int)std::min<int64_t>(chunk, nq - q0);
      const int32_t* pr = e->probes.p + (size_t)q0 * w;
      if ((rc = launch_lut(e, cb, d_q + (size_t)q0 * e->d, e->coarse.p, pr, w, n * w, e->lut.p))) return rc;
      if ((rc = launch_scan_keys(e, e->fine, pr, n, w, e->lut.p, K, e->j_keys.p, key_cap, e->j_ncells.p))) return rc;
      StageTimer t(e, ST_FINALIZE);
      topk_from_keys_kernel<<<n, kJoinThreads, 0, e->stream>>>(e->j_keys.p, key_cap, e->j_ncells.p, k, sentinel, e->fine.ids.p,
                                                              e->qflags.p + q0, d_out_ids + (size_t)q0 * k,
                                                              d_out_dists + (size_t)q0 * k);
      e->launches++;
      FB_CUDA(e, cudaGetLastError());
    }
  } else {
    iota_kernel<<<(nq + 255) / 256, 256, 0, e->stream>>>(e->exact_list.p, nq);
    int32_t cnt = nq;
    FB_CUDA(e, cudaMemcpyAsync(e->small.p, &cnt, sizeof cnt, cudaMemcpyHostToDevice, e->stream));
    e->launches++;
  }
  if (batch_mode && (fast || large_k)) {
    // ivfadc_batch_search goes on probing while a query's top-k has an empty slot (admissions, not rows, are counted)
    batch_unfilled_kernel<<<(nq + 255) / 256, 256, 0, e->stream>>>(d_out_ids, nq, k, e->qflags.p, e->exact_list.p, e->small.p + 0,
                                                                   e->counters64.p + 1);
    e->launches++;
  }
  {
    // flagged queries (boundary ties, re-probe loop, large k/w): the literal kernel, once per call
    StageTimer t(e, ST_EXACT);
    ivfadc_exact_kernel<<<exact_ctas, kExactThreads, ex_smem, e->stream>>>(
        d_q, e->d, e->coarse.p, e->coarseT.p, e->C, e->Cs, cb.cbT.p, K, cb.sub, e->fine.dev(), w, k,
        e->exact_list.p, e->small.p + 0, e->small.p + 1, e->exact_lut.p,
        fast ? e->qflags.p : nullptr, e->probes.p, e->kth.p, d_out_ids, d_out_dists, e->small.p + 2, ex_stage, sentinel,
        (fast || large_k) ? nullptr : e->counters64.p + 0, batch_mode ? 1 : 0);
    e->launches++;
    FB_CUDA(e, cudaGetLastError());
  }
  e->queries_done += nq;
  // algorithmic bytes per scanned row: int2 codes + int4 id, or one byte per code when the byte image is read
  e->bytes_per_row = e->used8 ? m + 4 : 2 * m + 4;
  return FB_OK;
}

int check_error_flag(fb_engine* e) {
  int32_t flag = 0;
  FB_CUDA(e, cudaMemcpy(&flag, e->small.p + 2, sizeof flag, cudaMemcpyDeviceToHost));
  if (flag) {
    cudaMemset(e->small.p + 2, 0, sizeof(int32_t));
    return fail(e, FB_ERR_REFERENCE_UB,
                "a query hit a state where the reference is undefined (fewer than w unprobed lists left, or a coarse distance >= 100)");
  }
  return FB_OK;
}

// ---- small host-buffer calls as one CUDA graph ------------------------------------------------
// A single-query call is ~9 kernel launches and 4 copies: launch-bound.  The second call with the same
// (nq, k, w) captures the whole sequence — upload from a pinned staging buffer, coarse, LUT, scan, finalize,
// general kernel, download of ids / distances / error flag — and later calls replay it with one
// cudaGraphLaunch.  A graph is dropped when a buffer moved (g_alloc_generation) or the index / options
// changed (graph_epoch).  *handled = false: the caller runs the ordinary path.
constexpr int kGraphMaxQueries = 16;

void drop_graphs(fb_engine* e) {
  for (auto& g : e->graphs)
    if (g.exec) cudaGraphExecDestroy(g.exec);
  e->graphs.clear();
}

int ivfadc_search_graph(fb_engine* e, const float* queries, int nq, int k, int w, int32_t* out_ids, float* out_dists,
                        bool* handled) {
  *handled = false;
  fb_engine::GraphEntry* ent = nullptr;
  for (auto& g : e->graphs)
    if (g.nq == nq && g.k == k && g.w == w) ent = &g;
  if (ent == nullptr) {   // first sighting: the ordinary path sizes every buffer
    e->graphs.push_back({nq, k, w, 0, 0, nullptr, 0});
    return FB_OK;
  }
  const size_t q_floats = (size_t)kGraphMaxQueries * e->d;
  if (e->pin_q == nullptr || e->pin_q_floats < q_floats) {
    if (e->pin_q) cudaFreeHost(e->pin_q);
    e->pin_q = nullptr;
    FB_CUDA(e, cudaMallocHost(&e->pin_q, q_floats * sizeof(float)));
    e->pin_q_floats = q_floats;
    drop_graphs(e);
    e->graphs.push_back({nq, k, w, 0, 0, nullptr, 0});
    ent = &e->graphs.back();
  }
  if (e->pin_ids == nullptr) {
    FB_CUDA(e, cudaMallocHost(&e->pin_ids, (size_t)kGraphMaxQueries * 32 * sizeof(int32_t)));
    FB_CUDA(e, cudaMallocHost(&e->pin_d, (size_t)kGraphMaxQueries * 32 * sizeof(float)));
    FB_CUDA(e, cudaMallocHost(&e->pin_flag, sizeof(int32_t)));
  }
  if (ent->exec == nullptr || ent->gen != g_alloc_generation || ent->epoch != e->graph_epoch) {
    if (ent->exec) { cudaGraphExecDestroy(ent->exec); ent->exec = nullptr; }
    FB_CUDA(e, e->q_stage.ensure((size_t)nq * e->d));
    FB_CUDA(e, e->id_stage.ensure((size_t)nq * k));
    FB_CUDA(e, e->dist_stage.ensure((size_t)nq * k));
    const uint64_t gen0 = g_alloc_generation;
    const int64_t launches0 = e->launches, queries0 = e->queries_done;
    FB_CUDA(e, cudaStreamSynchronize(e->stream));
    FB_CUDA(e, cudaStreamBeginCapture(e->stream, cudaStreamCaptureModeRelaxed));
    cudaMemcpyAsync(e->q_stage.p, e->pin_q, (size_t)nq * e->d * sizeof(float), cudaMemcpyHostToDevice, e->stream);
    int rc = ivfadc_dev(e, e->q_stage.p, nq, k, w, e->id_stage.p, e->dist_stage.p);
    cudaMemcpyAsync(e->pin_ids, e->id_stage.p, (size_t)nq * k * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream);
    cudaMemcpyAsync(e->pin_d, e->dist_stage.p, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, e->stream);
    cudaMemcpyAsync(e->pin_flag, e->small.p + 2, sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream);
    cudaGraph_t graph = nullptr;
    cudaError_t cerr = cudaStreamEndCapture(e->stream, &graph);
    const int n_launches = (int)(e->launches - launches0);
    e->launches = launches0;             // nothing ran during the capture
    e->queries_done = queries0;
    if (rc != FB_OK || cerr != cudaSuccess || graph == nullptr || gen0 != g_alloc_generation) {
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      ent->gen = 0;                      // try again on a later call; this one takes the ordinary path
      return rc == FB_OK ? FB_OK : rc;
    }
    cerr = cudaGraphInstantiate(&ent->exec, graph, 0);
    cudaGraphDestroy(graph);
    if (cerr != cudaSuccess) { ent->exec = nullptr; cudaGetLastError(); return FB_OK; }
    ent->gen = g_alloc_generation;
    ent->epoch = e->graph_epoch;
    ent->launches = n_launches;
  }
  memcpy(e->pin_q, queries, (size_t)nq * e->d * sizeof(float));
  FB_CUDA(e, cudaGraphLaunch(ent->exec, e->stream));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  memcpy(out_ids, e->pin_ids, (size_t)nq * k * sizeof(int32_t));
  memcpy(out_dists, e->pin_d, (size_t)nq * k * sizeof(float));
  e->launches += ent->launches;
  e->queries_done += nq;
  *handled = true;
  if (*e->pin_flag) {
    cudaMemset(e->small.p + 2, 0, sizeof(int32_t));
    return fail(e, FB_ERR_REFERENCE_UB,
                "a query hit a state where the reference is undefined (fewer than w unprobed lists left, or a coarse distance >= 100)");
  }
  return FB_OK;
}

// ---- flat PQ pipeline over a one-list table (the pq table or a per-call subset) --------
// freddy.c:74-134 (pq_search), :514-631 (pq_search_in_batch), :1070-1143 (pq_search_in): one LUT per query on the
// raw query, every row of the table scanned by every query.  Throughput form: one CTA per query keeps its LUT
// resident in shared memory and walks the whole table (codes come from L2: 100k targets = 2.4 MB), top-k and
// the reference's tie order fused (adc_scan_query_kernel, w = 1).  Few queries: the table is cut into `segs`
// segments per query so that the launch still fills the GPU (adc_scan_kernel + finalize_kernel).
// rows_upper: host-side upper bound of the table's row count (the exact count of a subset lives on the device).
int pq_dev(fb_engine* e, const CodeTableDev& tab, int64_t rows_upper, const float* d_q, int nq, int k, float sentinel,
           int32_t* d_out_ids, float* d_out_dists) {
  const Codebook& cb = e->cb[FB_CB_PQ];
  const int m = cb.m, K = cb.K, d = cb.m * cb.sub;
  const bool fast = (k <= 30);
  const int KK = k + 2;
  int64_t chunk = std::max<int64_t>(1, std::min<int64_t>(e->query_chunk, nq));
  const int64_t blocks_upper = std::max<int64_t>(1, (rows_upper + 31) / 32);
  auto segs_for = [&](int n) {
    if (n >= 2 * e->num_sms) return 1;
    return (int)std::max<int64_t>(1, std::min<int64_t>({(int64_t)(4 * e->num_sms + n - 1) / n, blocks_upper / 8, (int64_t)1024}));
  };
  const int segs_max = segs_for((int)std::min<int64_t>(chunk, nq));
  FB_CUDA(e, e->lut.ensure((size_t)chunk * m * K));
  FB_CUDA(e, e->qflags.ensure((size_t)chunk));
  FB_CUDA(e, e->exact_list.ensure((size_t)chunk));
  FB_CUDA(e, e->kth.ensure((size_t)chunk));
  FB_CUDA(e, e->probes.ensure((size_t)chunk));
  FB_CUDA(e, e->zero_i32.ensure(4));
  FB_CUDA(e, cudaMemsetAsync(e->zero_i32.p, 0, 4 * sizeof(int32_t), e->stream));
  FB_CUDA(e, cudaMemsetAsync(e->probes.p, 0, (size_t)chunk * sizeof(int32_t), e->stream));   // every query "probes" list 0
  if (fast && segs_max > 1) FB_CUDA(e, e->partial.ensure((size_t)chunk * segs_max * KK));
  size_t ex_smem = kExactFixedSmem;
  int ex_stage = 0;
  if (ex_smem + (size_t)m * K * sizeof(float) <= e->smem_optin) { ex_stage = m * K; ex_smem += (size_t)m * K * sizeof(float); }
  FB_CUDA(e, cudaFuncSetAttribute(pq_exact_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ex_smem));
  int rc;
  for (int64_t q0 = 0; q0 < nq; q0 += chunk) {
    const int n = (int)std::min<int64_t>(chunk, nq - q0);
    const float* dq = d_q + (size_t)q0 * d;
    int32_t* oi = d_out_ids + (size_t)q0 * k;
    float* od = d_out_dists + (size_t)q0 * k;
    FB_CUDA(e, cudaMemsetAsync(e->small.p, 0, 2 * sizeof(int32_t), e->stream));
    FB_CUDA(e, cudaMemsetAsync(e->qflags.p, 0, (size_t)n * sizeof(uint32_t), e->stream));
    if ((rc = launch_lut(e, cb, dq, nullptr, nullptr, 1, n, e->lut.p))) return rc;   // freddy.c:519-525
    if (fast && !e->force_exact) {
      const int segs = segs_for(n);
      rc = (segs == 1) ? launch_qscan(e, tab, 0, n, 1, e->lut.p, K, KK, k, sentinel, oi, od) : FB_ERR_UNSUPPORTED;
      if (rc == FB_ERR_UNSUPPORTED) {
        const int sg = std::max(segs, 1);
        if (e->partial.ensure((size_t)chunk * sg * KK) != cudaSuccess) return fail(e, FB_ERR_CUDA, "out of device memory");
        if ((rc = launch_scan(e, tab, e->probes.p, n, 1, 1, sg, e->lut.p, K, KK, e->partial.p, sentinel))) return rc;
        if ((rc = launch_finalize(e, tab, 0, sg, KK, k, n, sentinel, false, oi, od))) return rc;
      } else if (rc) {
        return rc;
      }
    } else {
      iota_kernel<<<(n + 255) / 256, 256, 0, e->stream>>>(e->exact_list.p, n);
      int32_t cnt = n;
      FB_CUDA(e, cudaMemcpyAsync(e->small.p, &cnt, sizeof cnt, cudaMemcpyHostToDevice, e->stream));
      e->launches++;
    }
    {
      StageTimer t(e, ST_EXACT);
      pq_exact_kernel<<<2 * e->num_sms, kExactThreads, ex_smem, e->stream>>>(
          tab, e->zero_i32.p, e->lut.p, K, k, sentinel, e->exact_list.p, e->small.p + 0, e->small.p + 1,
          (fast && !e->force_exact) ? e->kth.p : nullptr, oi, od, ex_stage);
      e->launches++;
      FB_CUDA(e, cudaGetLastError());
    }
  }
  e->queries_done += nq;
  e->bytes_per_row = 2 * m + 4;
  return FB_OK;
}

int check_pq_ready(fb_engine* e) {
  if (!e->cb[FB_CB_PQ].loaded || !e->pq.loaded) return fail(e, FB_ERR_INVALID, "flat PQ index not loaded (pq codebook / pq table)");
  const Codebook& cb = e->cb[FB_CB_PQ];
  if (cb.m != e->pq.m) return fail(e, FB_ERR_INVALID, "pq codebook m=%d but pq table m=%d", cb.m, e->pq.m);
  return FB_OK;
}
int pq_dim(const fb_engine* e) { return e->cb[FB_CB_PQ].m * e->cb[FB_CB_PQ].sub; }

// sorted (id, row) image of a table's id column on the device: the `WHERE id IN (...)` index
// (id, row) pairs ordered by id, rows ascending inside an id: stable radix sort of an id column that is already on the device
int device_id_index(fb_engine* e, const int32_t* d_ids, int64_t N, int32_t* d_sorted_ids, int32_t* d_sorted_rows) {
  DevBuf<int32_t> d_iota;
  DevBuf<unsigned char> d_tmp;
  FB_CUDA(e, d_iota.ensure((size_t)N));
  iota_i32_kernel<<<(unsigned)((N + 255) / 256), 256, 0, e->stream>>>(d_iota.p, N);
  size_t tmp_bytes = 0;
  FB_CUDA(e, cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_ids, d_sorted_ids, d_iota.p, d_sorted_rows, (int)N, 0, 32, e->stream));
  FB_CUDA(e, d_tmp.ensure(tmp_bytes));
  FB_CUDA(e, cub::DeviceRadixSort::SortPairs(d_tmp.p, tmp_bytes, d_ids, d_sorted_ids, d_iota.p, d_sorted_rows, (int)N, 0, 32, e->stream));
  e->launches += 2;
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  return FB_OK;
}

int build_id_index(fb_engine* e, CodeTable& tab, const int32_t* ids, int64_t N) {
  FB_CUDA(e, tab.sorted_ids.ensure((size_t)std::max<int64_t>(1, N)));
  FB_CUDA(e, tab.sorted_rows.ensure((size_t)std::max<int64_t>(1, N)));
  if (N == 0) return FB_OK;
  if (e->device_build) return device_id_index(e, tab.ids.p, N, tab.sorted_ids.p, tab.sorted_rows.p);   // tab.ids = table order
  std::vector<int32_t> order((size_t)N), sid((size_t)N);
  for (int64_t r = 0; r < N; r++) order[r] = (int32_t)r;
  if (!std::is_sorted(ids, ids + N))
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return ids[a] < ids[b]; });
  for (int64_t r = 0; r < N; r++) sid[r] = ids[order[r]];
  FB_CUDA(e, cudaMemcpy(tab.sorted_ids.p, sid.data(), (size_t)N * sizeof(int32_t), cudaMemcpyHostToDevice));
  FB_CUDA(e, cudaMemcpy(tab.sorted_rows.p, order.data(), (size_t)N * sizeof(int32_t), cudaMemcpyHostToDevice));
  return FB_OK;
}

// Rows of `src` selected by `WHERE id IN (wanted)` in table order, every matching row once (freddy.c:544-562,
// :1286-1300), gathered on the device into the compact one-list table e->tmp.  Nothing comes back to the host:
// the row count stays in e->sel_total[0] (= the list length the scan kernels read).  `view` describes the
// subset for the kernels; rows_upper bounds its row count.
int build_subset(fb_engine* e, const CodeTable& src, const int32_t* wanted, int n_wanted, CodeTableDev& view, int64_t& rows_upper,
                 int64_t n_queries = 0) {
  const int64_t N = src.N;
  const int n_words = (int)std::max<int64_t>(1, (N + 31) / 32);
  rows_upper = std::min<int64_t>(N, n_wanted);
  const int n_slots = (int)std::max<int64_t>(32, (rows_upper + 31) / 32 * 32);
  CodeTable& tmp = e->tmp;
  FB_CUDA(e, tmp.units.ensure((size_t)n_slots * src.U));
  FB_CUDA(e, tmp.rowno.ensure((size_t)n_slots));
  FB_CUDA(e, e->sel_rows.ensure((size_t)n_slots));
  FB_CUDA(e, e->sel_bitmap.ensure((size_t)n_words));
  FB_CUDA(e, e->sel_word_base.ensure((size_t)n_words));
  FB_CUDA(e, e->sel_total.ensure(4));
  FB_CUDA(e, e->sel_wanted.ensure((size_t)std::max(1, n_wanted)));
  FB_CUDA(e, e->zero_i32.ensure(4));
  FB_CUDA(e, cudaMemsetAsync(e->zero_i32.p, 0, 4 * sizeof(int32_t), e->stream));
  FB_CUDA(e, cudaMemsetAsync(e->sel_bitmap.p, 0, (size_t)n_words * sizeof(uint32_t), e->stream));
  if (n_wanted > 0) {
    FB_CUDA(e, cudaMemcpyAsync(e->sel_wanted.p, wanted, (size_t)n_wanted * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    subset_mark_kernel<<<(n_wanted + 255) / 256, 256, 0, e->stream>>>(src.sorted_ids.p, src.sorted_rows.p, (int)N, e->sel_wanted.p,
                                                                      n_wanted, e->sel_bitmap.p);
    e->launches++;
  }
  subset_scan_kernel<<<1, 1024, 0, e->stream>>>(e->sel_bitmap.p, n_words, e->sel_word_base.p, e->sel_total.p, nullptr);
  subset_compact_kernel<<<(n_words + 255) / 256, 256, 0, e->stream>>>(e->sel_bitmap.p, n_words, e->sel_word_base.p, e->sel_rows.p);
  // Many queries over the subset: its rows get the conflict-aware placement of the pinned fine table, in groups of 512
  // rows (one CTA each, ~0.1 ms) — the shared-memory gather of the scan then replays far fewer bank conflicts
  // (profiles/r2_config3_config4_ncu.txt: two thirds of the wavefronts were replays).  Few queries: not worth the kernel.
  const int32_t* d_order = nullptr;
  constexpr int kGroup = 512, kWindow = 128;
  const size_t place_bytes = (place_rows_smem(src.m, src.K, kWindow) + 15) / 16 * 16;
  const size_t place_smem = place_bytes + (size_t)kGroup * 4 * src.U * sizeof(int16_t);
  if ((e->subset_placement == 2 || (e->subset_placement == 1 && n_queries * rows_upper >= ((int64_t)1 << 26))) && src.m <= 64 && src.K > 0 &&
      place_smem <= std::min<size_t>(e->smem_optin, 96 * 1024)) {
    FB_CUDA(e, e->sel_order.ensure((size_t)n_slots));
    FB_CUDA(e, cudaFuncSetAttribute(subset_place_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)place_smem));
    subset_place_kernel<<<(unsigned)((rows_upper + kGroup - 1) / kGroup), std::max(kWindow, (src.m + 31) / 32 * 32), place_smem, e->stream>>>(
        src.units.p, src.U, src.m, src.K, e->sel_rows.p, e->sel_total.p, kWindow, kGroup, place_bytes, e->sel_order.p);
    e->launches++;
    d_order = e->sel_order.p;
  }
  subset_gather_kernel<<<(n_slots + 255) / 256, 256, 0, e->stream>>>(src.units.p, src.U, e->sel_rows.p, e->sel_total.p, d_order,
                                                                     tmp.units.p, tmp.rowno.p, n_slots);
  e->launches += 3;
  FB_CUDA(e, cudaGetLastError());
  view.units8 = nullptr;
  view.units = tmp.units.p; view.rowno = tmp.rowno.p; view.list_blk = e->zero_i32.p; view.list_len = e->sel_total.p;
  view.ids = src.ids.p; view.m = src.m; view.U = src.U; view.n_lists = 1;
  return FB_OK;
}

}  // namespace

// ============================ C-ABI =========================================
extern "C" {

const char* fb_version(void) { return "freddy_b200 0.1 (sm_100a)"; }

const char* fb_last_error(const fb_engine* e) { return e ? e->err.c_str() : g_create_error.c_str(); }

int fb_create(int device, fb_engine** out) {
  if (!out) return FB_ERR_INVALID;
  *out = nullptr;
  int count = 0;
  cudaError_t err = cudaGetDeviceCount(&count);
  if (err != cudaSuccess || count == 0)
    return fail(nullptr, FB_ERR_CUDA, "no CUDA device: %s (this library has no CPU fallback)", cudaGetErrorString(err));
  if (device < 0 || device >= count) return fail(nullptr, FB_ERR_INVALID, "device %d out of range [0,%d)", device, count);
  fb_engine* e = new (std::nothrow) fb_engine();
  if (!e) return fail(nullptr, FB_ERR_INVALID, "out of host memory");
  e->device = device;
  cudaDeviceProp prop;
  if ((err = cudaSetDevice(device)) != cudaSuccess || (err = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) {
    delete e;
    return fail(nullptr, FB_ERR_CUDA, "cudaSetDevice(%d): %s", device, cudaGetErrorString(err));
  }
  if (prop.major < 10) {
    delete e;
    return fail(nullptr, FB_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor);
  }
  e->num_sms = prop.multiProcessorCount;
  e->smem_optin = prop.sharedMemPerBlockOptin;
  if ((err = cudaStreamCreateWithFlags(&e->own_stream, cudaStreamNonBlocking)) != cudaSuccess ||
      (err = e->small.ensure(4)) != cudaSuccess || (err = e->counters64.ensure(8)) != cudaSuccess) {
    delete e;
    return fail(nullptr, FB_ERR_CUDA, "engine setup: %s", cudaGetErrorString(err));
  }
  e->stream = e->own_stream;
  cudaStreamCreateWithFlags(&e->s_lut, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&e->s_scan, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&e->ev_main, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&e->ev_all, cudaEventDisableTiming);
  for (int i = 0; i < 2; i++) {
    cudaEventCreateWithFlags(&e->ev_lut_done[i], cudaEventDisableTiming);
    cudaEventCreateWithFlags(&e->ev_scan_done[i], cudaEventDisableTiming);
  }
  cudaMemset(e->small.p, 0, 4 * sizeof(int32_t));
  cudaMemset(e->counters64.p, 0, 8 * sizeof(u64));
  *out = e;
  return FB_OK;
}

void fb_destroy(fb_engine* e) {
  if (!e) return;
  cudaSetDevice(e->device);
  cudaStreamSynchronize(e->stream);
  drop_graphs(e);
  if (e->pin_q) cudaFreeHost(e->pin_q);
  if (e->pin_ids) cudaFreeHost(e->pin_ids);
  if (e->pin_d) cudaFreeHost(e->pin_d);
  if (e->pin_flag) cudaFreeHost(e->pin_flag);
  for (auto& ev : e->events) { cudaEventDestroy(ev.a); cudaEventDestroy(ev.b); }
  e->coarse.release(); e->coarseT.release();
  for (auto& c : e->cb) c.cbT.release();
  e->fine.release(); e->pq.release(); e->tmp.release();
  e->iota_lists.release();
  e->ivpq.release(); e->jtmp.release(); e->coarse_multi.release(); e->ivpq_stats.release(); e->ivpq_cells.release();
  e->j_cell.release(); e->j_vrow.release(); e->j_id.release(); e->j_active.release(); e->j_ncells.release();
  e->j_filled.release(); e->j_tcounts.release(); e->j_bitmaps.release(); e->j_keys.release();
  e->vecT.release(); e->vecR.release(); e->vec_ids.release(); e->va.release(); e->vb.release(); e->vo.release(); e->vdo.release();
  e->ana_rows.release(); e->ana_partial.release();
  e->lut.release(); e->exact_lut.release(); e->q_stage.release(); e->dist_stage.release();
  e->probes.release(); e->exact_list.release(); e->id_stage.release(); e->sel_rows.release();
  e->qflags.release(); e->partial.release(); e->kth.release(); e->small.release(); e->counters64.release();
  cudaStreamDestroy(e->own_stream);
  cudaStreamDestroy(e->s_lut); cudaStreamDestroy(e->s_scan);
  cudaEventDestroy(e->ev_main); cudaEventDestroy(e->ev_all);
  for (int i = 0; i < 2; i++) { cudaEventDestroy(e->ev_lut_done[i]); cudaEventDestroy(e->ev_scan_done[i]); }
  e->lut2.release();
  delete e;
}

int fb_load_coarse(fb_engine* e, const float* coarse, int C, int d) {
  if (e) e->graph_epoch++;
  if (!e || !coarse || C < 1 || d < 1) return fail(e, FB_ERR_INVALID, "fb_load_coarse: bad arguments");
  FB_CUDA(e, cudaSetDevice(e->device));
  const int Cs = (C + 31) / 32 * 32;
  std::vector<float> t((size_t)d * Cs, 0.0f);
  for (int c = 0; c < C; c++)
    for (int i = 0; i < d; i++) t[(size_t)i * Cs + c] = coarse[(size_t)c * d + i];
  FB_CUDA(e, e->coarse.ensure((size_t)C * d));
  FB_CUDA(e, e->coarseT.ensure(t.size()));
  FB_CUDA(e, cudaMemcpy(e->coarse.p, coarse, (size_t)C * d * sizeof(float), cudaMemcpyHostToDevice));
  FB_CUDA(e, cudaMemcpy(e->coarseT.p, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice));
  e->C = C; e->Cs = Cs; e->d = d; e->coarse_loaded = true;
  return FB_OK;
}

int fb_load_codebook(fb_engine* e, int kind, const float* codebook, int m, int K, int sub) {
  if (e) e->graph_epoch++;
  if (!e || !codebook || kind < 0 || kind >= FB_CB_KINDS || m < 1 || K < 1 || sub < 1)
    return fail(e, FB_ERR_INVALID, "fb_load_codebook: bad arguments");
  if (K % 4 != 0 || K > 16384) return fail(e, FB_ERR_UNSUPPORTED, "K=%d: need K %% 4 == 0 and K <= 16384", K);
  FB_CUDA(e, cudaSetDevice(e->device));
  std::vector<float> t((size_t)m * sub * K);
  for (int p = 0; p < m; p++)
    for (int c = 0; c < K; c++)
      for (int i = 0; i < sub; i++) t[((size_t)p * sub + i) * K + c] = codebook[((size_t)p * K + c) * sub + i];
  Codebook& cb = e->cb[kind];
  FB_CUDA(e, cb.cbT.ensure(t.size()));
  FB_CUDA(e, cudaMemcpy(cb.cbT.p, t.data(), t.size() * sizeof(float), cudaMemcpyHostToDevice));
  cb.m = m; cb.K = K; cb.sub = sub; cb.loaded = true;
  return FB_OK;
}

int fb_load_fine(fb_engine* e, const int32_t* ids, const int32_t* coarse_ids, const int16_t* codes, int64_t N, int m) {
  if (e) e->graph_epoch++;
  if (!e || (N > 0 && (!ids || !coarse_ids || !codes))) return fail(e, FB_ERR_INVALID, "fb_load_fine: bad arguments");
  if (!e->coarse_loaded || !e->cb[FB_CB_RESIDUAL].loaded)
    return fail(e, FB_ERR_INVALID, "fb_load_fine: load the coarse table and the residual codebook first");
  FB_CUDA(e, cudaSetDevice(e->device));
  return build_table(e, e->fine, ids, coarse_ids, e->C, 0, codes, N, m, e->cb[FB_CB_RESIDUAL].K, e->placement_window);
}

int fb_load_pq(fb_engine* e, const int32_t* ids, const int16_t* codes, int64_t N, int m) {
  if (!e || (N > 0 && (!ids || !codes))) return fail(e, FB_ERR_INVALID, "fb_load_pq: bad arguments");
  if (!e->cb[FB_CB_PQ].loaded) return fail(e, FB_ERR_INVALID, "fb_load_pq: load the pq codebook first");
  FB_CUDA(e, cudaSetDevice(e->device));
  // one list holding the whole table in table order
  int rc = build_table(e, e->pq, ids, nullptr, 0, 0x7fffffff, codes, N, m, e->cb[FB_CB_PQ].K);
  if (rc) return rc;
  e->pq_ids_host.assign(ids, ids + N);
  return build_id_index(e, e->pq, ids, N);
}

int fb_ivfadc_search_dev(fb_engine* e, const float* d_queries, int nq, int k, int w, int32_t* d_out_ids, float* d_out_dists) {
  if (!e) return FB_ERR_INVALID;
  FB_CUDA(e, cudaSetDevice(e->device));
  return ivfadc_dev(e, d_queries, nq, k, w, d_out_ids, d_out_dists);
}

int fb_ivfadc_search(fb_engine* e, const float* queries, int nq, int k, int w, int32_t* out_ids, float* out_dists) {
  if (!e) return FB_ERR_INVALID;
  int rc = check_common(e, nq, k);
  if (rc) return rc;
  if (nq == 0) return FB_OK;
  if (!queries || !out_ids || !out_dists) return fail(e, FB_ERR_INVALID, "null buffer");
  FB_CUDA(e, cudaSetDevice(e->device));
  if (e->use_graphs && !e->profile && nq <= kGraphMaxQueries && k <= 30 && w >= 1 && w <= 31 && w <= e->C &&
      e->coarse_loaded && e->cb[FB_CB_RESIDUAL].loaded && e->fine.loaded) {
    bool handled = false;
    rc = ivfadc_search_graph(e, queries, nq, k, w, out_ids, out_dists, &handled);
    if (rc || handled) return rc;
  }
  FB_CUDA(e, e->q_stage.ensure((size_t)nq * e->d));
  FB_CUDA(e, e->id_stage.ensure((size_t)nq * k));
  FB_CUDA(e, e->dist_stage.ensure((size_t)nq * k));
  const float* mapped = nullptr;
  if (e->zero_copy && nq > 32 && k <= 30 && w <= 31) {      // the batched coarse kernel will run (it does the staging)
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, queries) == cudaSuccess && attr.type == cudaMemoryTypeHost && attr.devicePointer != nullptr)
      mapped = static_cast<const float*>(attr.devicePointer);
    else
      cudaGetLastError();
  }
  if (mapped == nullptr)
    FB_CUDA(e, cudaMemcpyAsync(e->q_stage.p, queries, (size_t)nq * e->d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  e->coarse_src = mapped;
  rc = ivfadc_dev(e, e->q_stage.p, nq, k, w, e->id_stage.p, e->dist_stage.p);
  e->coarse_src = nullptr;
  if (rc) return rc;
  FB_CUDA(e, cudaMemcpyAsync(out_ids, e->id_stage.p, (size_t)nq * k * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaMemcpyAsync(out_dists, e->dist_stage.p, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  return check_error_flag(e);
}

int fb_pq_search(fb_engine* e, const float* queries, int nq, int k, int32_t* out_ids, float* out_dists) {
  if (!e) return FB_ERR_INVALID;
  int rc = check_common(e, nq, k);
  if (rc) return rc;
  if ((rc = check_pq_ready(e))) return rc;
  if (nq == 0) return FB_OK;
  if (!queries || !out_ids || !out_dists) return fail(e, FB_ERR_INVALID, "null buffer");
  FB_CUDA(e, cudaSetDevice(e->device));
  const int d = pq_dim(e);
  FB_CUDA(e, e->q_stage.ensure((size_t)nq * d));
  FB_CUDA(e, e->id_stage.ensure((size_t)nq * k));
  FB_CUDA(e, e->dist_stage.ensure((size_t)nq * k));
  FB_CUDA(e, cudaMemcpyAsync(e->q_stage.p, queries, (size_t)nq * d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  if ((rc = pq_dev(e, e->pq.dev(), e->pq.N, e->q_stage.p, nq, k, 100.0f, e->id_stage.p, e->dist_stage.p))) return rc;  // freddy.c:90-92
  e->host_rows += (int64_t)nq * e->pq.N;   // every query sees every row of the table
  FB_CUDA(e, cudaMemcpyAsync(out_ids, e->id_stage.p, (size_t)nq * k * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaMemcpyAsync(out_dists, e->dist_stage.p, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  return FB_OK;
}

int fb_pq_search_in_batch(fb_engine* e, const float* queries, int nq, int k, const int32_t* targets, int n_targets,
                          int use_target_lists, int32_t* out_ids, float* out_dists) {
  (void)use_target_lists;  // loop-nest choice of the reference (freddy.c:600-631); results are identical
  if (!e) return FB_ERR_INVALID;
  int rc = check_common(e, nq, k);
  if (rc) return rc;
  if ((rc = check_pq_ready(e))) return rc;
  if (n_targets < 0 || (n_targets > 0 && !targets)) return fail(e, FB_ERR_INVALID, "bad target array");
  if (nq == 0) return FB_OK;
  if (!queries || !out_ids || !out_dists) return fail(e, FB_ERR_INVALID, "null buffer");
  FB_CUDA(e, cudaSetDevice(e->device));
  const int d = pq_dim(e);
  FB_CUDA(e, e->q_stage.ensure((size_t)nq * d));
  FB_CUDA(e, e->id_stage.ensure((size_t)nq * k));
  FB_CUDA(e, e->dist_stage.ensure((size_t)nq * k));
  FB_CUDA(e, cudaMemcpyAsync(e->q_stage.p, queries, (size_t)nq * d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  CodeTableDev view;
  int64_t rows_upper = 0;
  if ((rc = build_subset(e, e->pq, targets, n_targets, view, rows_upper, nq))) return rc;
  rc = pq_dev(e, view, rows_upper, e->q_stage.p, nq, k, 1000.0f, e->id_stage.p, e->dist_stage.p);  // freddy.c:415
  if (rc) return rc;
  int32_t n_sel = 0;
  FB_CUDA(e, cudaMemcpyAsync(&n_sel, e->sel_total.p, sizeof n_sel, cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaMemcpyAsync(out_ids, e->id_stage.p, (size_t)nq * k * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaMemcpyAsync(out_dists, e->dist_stage.p, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  e->host_rows += (int64_t)nq * n_sel;     // every query sees every selected row
  return FB_OK;
}

}  // extern "C"

namespace {
// Append n rows (ids, list of each row or nullptr for one-list tables, codes [n][m]) to a pinned code table:
// device-side re-pack (subset_kernels.cuh), id index extended, host mirrors updated.
int append_rows(fb_engine* e, CodeTable& tab, const int32_t* ids, const int32_t* list_of_row, const int16_t* codes, int64_t n) {
  if (!tab.loaded) return fail(e, FB_ERR_INVALID, "append: the table is not loaded");
  if (n == 0) return FB_OK;
  const int m = tab.m, U = tab.U, nl = tab.n_lists;
  if (tab.N + n >= (1ll << 31) - 64) return fail(e, FB_ERR_UNSUPPORTED, "table too large");
  for (int64_t i = 0; i < n; i++) {
    const int c = list_of_row ? list_of_row[i] : 0;
    if (c < 0 || c >= nl) return fail(e, FB_ERR_INVALID, "appended row %lld: coarse_id %d out of range [0,%d)", (long long)i, c, nl);
    for (int p = 0; p < m; p++)
      if (codes[(size_t)i * m + p] < 0 || codes[(size_t)i * m + p] >= tab.K)
        return fail(e, FB_ERR_INVALID, "appended row %lld pos %d: code out of range", (long long)i, p);
  }
  // new geometry; position of every appended row inside its list
  std::vector<int32_t> new_len(tab.h_list_len), new_blk(nl);
  std::vector<int64_t> dst((size_t)n);
  for (int64_t i = 0; i < n; i++) dst[i] = new_len[list_of_row ? list_of_row[i] : 0]++;
  int64_t n_blocks = 0;
  for (int c = 0; c < nl; c++) { new_blk[c] = (int32_t)n_blocks; n_blocks += (new_len[c] + 31) / 32; }
  for (int64_t i = 0; i < n; i++) dst[i] += (int64_t)new_blk[list_of_row ? list_of_row[i] : 0] * 32;
  // fresh buffers, old lists moved on the device
  DevBuf<uint2> units;
  DevBuf<uint4> units8;
  DevBuf<int32_t> rowno, d_new_blk, d_new_len, d_ids;
  DevBuf<int16_t> d_codes;
  DevBuf<int64_t> d_dst;
  const size_t slots = (size_t)std::max<int64_t>(1, n_blocks) * 32;
  FB_CUDA(e, units.ensure(slots * U));
  FB_CUDA(e, rowno.ensure(slots));
  if (tab.has8) FB_CUDA(e, units8.ensure(slots));
  FB_CUDA(e, d_new_blk.ensure(nl));
  FB_CUDA(e, d_new_len.ensure(nl));
  FB_CUDA(e, d_ids.ensure((size_t)n));
  FB_CUDA(e, d_codes.ensure((size_t)n * m));
  FB_CUDA(e, d_dst.ensure((size_t)n));
  FB_CUDA(e, tab.ids.grow((size_t)(tab.N + n), (size_t)tab.N));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  FB_CUDA(e, cudaMemcpy(d_new_blk.p, new_blk.data(), nl * sizeof(int32_t), cudaMemcpyHostToDevice));
  FB_CUDA(e, cudaMemcpy(d_new_len.p, new_len.data(), nl * sizeof(int32_t), cudaMemcpyHostToDevice));
  FB_CUDA(e, cudaMemcpy(d_ids.p, ids, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice));
  FB_CUDA(e, cudaMemcpy(d_codes.p, codes, (size_t)n * m * sizeof(int16_t), cudaMemcpyHostToDevice));
  FB_CUDA(e, cudaMemcpy(d_dst.p, dst.data(), (size_t)n * sizeof(int64_t), cudaMemcpyHostToDevice));
  append_repack_kernel<<<nl, 256, 0, e->stream>>>(tab.units.p, tab.rowno.p, tab.has8 ? tab.units8.p : nullptr, U, tab.list_blk.p,
                                                  tab.list_len.p, d_new_blk.p, d_new_len.p, units.p, rowno.p,
                                                  tab.has8 ? units8.p : nullptr);
  append_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, e->stream>>>(d_codes.p, d_ids.p, d_dst.p, (int)n, m, U, tab.N, units.p, rowno.p,
                                                                       tab.has8 ? units8.p : nullptr, tab.ids.p);
  e->launches += 2;
  FB_CUDA(e, cudaGetLastError());
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  std::swap(tab.units.p, units.p); std::swap(tab.units.n, units.n);
  std::swap(tab.rowno.p, rowno.p); std::swap(tab.rowno.n, rowno.n);
  if (tab.has8) { std::swap(tab.units8.p, units8.p); std::swap(tab.units8.n, units8.n); }
  std::swap(tab.list_blk.p, d_new_blk.p); std::swap(tab.list_blk.n, d_new_blk.n);
  std::swap(tab.list_len.p, d_new_len.p); std::swap(tab.list_len.n, d_new_len.n);
  g_alloc_generation++;
  // id index: ids beyond the current maximum, ascending, keep it sorted by appending; anything else re-sorts
  if (tab.sorted_ids.p != nullptr) {
    bool tail = true;
    int32_t prev = tab.max_id;
    for (int64_t i = 0; i < n && tail; i++) { tail = ids[i] > prev || (i > 0 && ids[i] == prev); prev = ids[i]; }
    if (tail) {
      std::vector<int32_t> rows((size_t)n);
      for (int64_t i = 0; i < n; i++) rows[i] = (int32_t)(tab.N + i);
      FB_CUDA(e, tab.sorted_ids.grow((size_t)(tab.N + n), (size_t)tab.N));
      FB_CUDA(e, tab.sorted_rows.grow((size_t)(tab.N + n), (size_t)tab.N));
      FB_CUDA(e, cudaMemcpy(tab.sorted_ids.p + tab.N, ids, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice));
      FB_CUDA(e, cudaMemcpy(tab.sorted_rows.p + tab.N, rows.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice));
    } else {
      std::vector<int32_t> all((size_t)(tab.N + n));
      FB_CUDA(e, cudaMemcpy(all.data(), tab.ids.p, all.size() * sizeof(int32_t), cudaMemcpyDeviceToHost));
      tab.N += n;
      int rc = build_id_index(e, tab, all.data(), tab.N);
      tab.N -= n;
      if (rc) return rc;
    }
  }
  for (int64_t i = 0; i < n; i++) tab.max_id = std::max(tab.max_id, ids[i]);
  tab.h_list_len = new_len;
  tab.h_list_blk = new_blk;
  tab.N += n;
  tab.n_blocks = n_blocks;
  e->graph_epoch++;
  return FB_OK;
}
}  // namespace

extern "C" {

int fb_append_fine(fb_engine* e, const int32_t* ids, const int32_t* coarse_ids, const int16_t* codes, int64_t n) {
  if (!e || n < 0 || (n > 0 && (!ids || !coarse_ids || !codes))) return fail(e, FB_ERR_INVALID, "fb_append_fine: bad arguments");
  FB_CUDA(e, cudaSetDevice(e->device));
  return append_rows(e, e->fine, ids, coarse_ids, codes, n);
}

int fb_append_pq(fb_engine* e, int kind, const int32_t* ids, const int32_t* cells, const int16_t* codes, int64_t n) {
  if (!e || n < 0 || (n > 0 && (!ids || !codes))) return fail(e, FB_ERR_INVALID, "fb_append_pq: bad arguments");
  FB_CUDA(e, cudaSetDevice(e->device));
  if (kind == FB_CB_PQ) {
    int rc = append_rows(e, e->pq, ids, nullptr, codes, n);
    if (rc == FB_OK) e->pq_ids_host.insert(e->pq_ids_host.end(), ids, ids + n);
    return rc;
  }
  if (kind == FB_CB_IVPQ) {
    if (!e->ivpq_loaded) return fail(e, FB_ERR_INVALID, "fb_append_pq: the IVPQ index is not loaded");
    if (n > 0 && !cells) return fail(e, FB_ERR_INVALID, "fb_append_pq(FB_CB_IVPQ) needs the multi-index cell of every row");
    const int ncell = e->ivpq_Kc * e->ivpq_Kc;
    for (int64_t i = 0; i < n; i++)
      if (cells[i] < 0 || cells[i] >= ncell) return fail(e, FB_ERR_INVALID, "appended row %lld: cell %d out of range", (long long)i, cells[i]);
    const int64_t N0 = e->ivpq.N;
    int rc = append_rows(e, e->ivpq, ids, nullptr, codes, n);
    if (rc) return rc;
    FB_CUDA(e, e->ivpq_cells.grow((size_t)(N0 + n), (size_t)N0));
    if (n > 0) FB_CUDA(e, cudaMemcpy(e->ivpq_cells.p + N0, cells, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice));
    return FB_OK;
  }
  return fail(e, FB_ERR_INVALID, "fb_append_pq: kind must be FB_CB_PQ or FB_CB_IVPQ");
}

int fb_synchronize(fb_engine* e) {
  if (!e) return FB_ERR_INVALID;
  FB_CUDA(e, cudaSetDevice(e->device));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  return check_error_flag(e);
}

int fb_set_stream(fb_engine* e, void* cuda_stream) {
  if (e) e->graph_epoch++;
  if (!e) return FB_ERR_INVALID;
  FB_CUDA(e, cudaSetDevice(e->device));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  drain_events(e);
  e->stream = cuda_stream ? (cudaStream_t)cuda_stream : e->own_stream;
  return FB_OK;
}

int fb_set_option(fb_engine* e, int option, int64_t value) {
  if (e) e->graph_epoch++;
  if (!e) return FB_ERR_INVALID;
  switch (option) {
    case FB_OPT_FORCE_EXACT_PATH: e->force_exact = value != 0; return FB_OK;
    case FB_OPT_PROFILE: e->profile = value != 0; return FB_OK;
    case FB_OPT_PACKED_FP32: e->packed_fp32 = value != 0; return FB_OK;
    case FB_OPT_LUT_TILE: e->lut_tile = (int)value; return FB_OK;
    case FB_OPT_LUT_CTAS_PER_SM: e->lut_ctas_per_sm = (int)value; return FB_OK;
    case FB_OPT_OVERLAP: e->overlap = value != 0; return FB_OK;
    case FB_OPT_PIPELINE: e->pipeline = value != 0; return FB_OK;
    case FB_OPT_PIPE_DEBUG: e->pipe_debug = (int)value; return FB_OK;
    case FB_OPT_PIPE_SHAPE: e->pipe_shape = (int)value; return FB_OK;
    case FB_OPT_PIPE_RAMP: e->pipe_ramp = value != 0; return FB_OK;
    case FB_OPT_CUDA_GRAPHS: e->use_graphs = value != 0; return FB_OK;
    case FB_OPT_ZERO_COPY_UPLOAD: e->zero_copy = value != 0; return FB_OK;
    case FB_OPT_PREFILTER: e->prefilter = value != 0; return FB_OK;
    case FB_OPT_BYTE_CODES: e->byte_codes = value != 0; return FB_OK;
    case FB_OPT_PREFILTER_LOCKSTEP: e->pf_lockstep = (int)std::max<int64_t>(0, std::min<int64_t>(value, 1 << 20)); return FB_OK;
    case FB_OPT_DEVICE_BUILD: e->device_build = value != 0; return FB_OK;
    case FB_OPT_SUBSET_PLACEMENT: e->subset_placement = (int)std::max<int64_t>(0, std::min<int64_t>(2, value)); return FB_OK;
    case FB_OPT_PLACEMENT_WINDOW: e->placement_window = (int)std::max<int64_t>(0, std::min<int64_t>(value, 1 << 20)); return FB_OK;
    case FB_OPT_PIPE_CHUNK:
      if (value < 1) return fail(e, FB_ERR_INVALID, "pipeline chunk must be >= 1");
      e->pipe_chunk = value;
      return FB_OK;
    case FB_OPT_QSCAN_MIN_QUERIES:
      e->qscan_min_queries = (int)std::max<int64_t>(0, std::min<int64_t>(value, 1 << 30));
      return FB_OK;
    case FB_OPT_QUERY_CHUNK:
      if (value < 1) return fail(e, FB_ERR_INVALID, "query chunk must be >= 1");
      e->query_chunk = value;
      return FB_OK;
  }
  return fail(e, FB_ERR_INVALID, "unknown option %d", option);
}

int fb_get_counters(fb_engine* e, fb_counters* out) {
  if (!e || !out) return FB_ERR_INVALID;
  FB_CUDA(e, cudaSetDevice(e->device));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  drain_events(e);
  u64 c64[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  FB_CUDA(e, cudaMemcpy(c64, e->counters64.p, sizeof c64, cudaMemcpyDeviceToHost));
  memset(out, 0, sizeof *out);
  out->queries = e->queries_done;
  out->rows_scanned = (int64_t)c64[0] + e->host_rows;
  out->scan_bytes = out->rows_scanned * e->bytes_per_row;
  out->exact_path_queries = (int64_t)c64[1];
  out->exact_coarse_tie = (int64_t)c64[2]; out->exact_coarse_far = (int64_t)c64[3];
  out->exact_few_rows = (int64_t)c64[4]; out->exact_scan_tie = (int64_t)c64[5]; out->exact_forced = (int64_t)c64[6];
  out->kernel_launches = e->launches;
  out->ms_coarse = e->ms[ST_COARSE]; out->ms_lut = e->ms[ST_LUT]; out->ms_scan = e->ms[ST_SCAN];
  out->ms_finalize = e->ms[ST_FINALIZE]; out->ms_exact = e->ms[ST_EXACT];
  out->n_scan_launches = e->n_scan_launches;
  out->ms_pipe = e->ms[ST_PIPE];
  out->n_pipe_launches = e->n_pipe_launches;
  out->prefilter_queries = e->pf_queries;
  out->prefilter_overflow_queries = e->pf_overflow_queries;
  out->prefilter_candidates = e->pf_candidates;
  return FB_OK;
}

int fb_reset_counters(fb_engine* e) {
  if (!e) return FB_ERR_INVALID;
  FB_CUDA(e, cudaSetDevice(e->device));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  drain_events(e);
  FB_CUDA(e, cudaMemset(e->counters64.p, 0, 8 * sizeof(u64)));
  for (double& v : e->ms) v = 0;
  e->launches = 0; e->queries_done = 0; e->n_scan_launches = 0; e->n_pipe_launches = 0; e->host_rows = 0;
  e->pf_queries = 0; e->pf_overflow_queries = 0; e->pf_candidates = 0;
  return FB_OK;
}

int fb_placement_order(const int16_t* codes, int n, int m, int K, int window, int32_t* order_out) {
  if (!codes || !order_out || n < 0 || m <= 0 || K <= 0) return FB_ERR_INVALID;
  for (int64_t i = 0; i < (int64_t)n * m; i++)
    if (codes[i] < 0 || codes[i] >= K) return FB_ERR_INVALID;
  std::vector<int32_t> rows(n), order;
  for (int i = 0; i < n; i++) rows[i] = i;
  if (window > 1) place_rows_of_list(codes, m, K, rows, window, order);
  else order = rows;
  memcpy(order_out, order.data(), (size_t)n * sizeof(int32_t));
  return FB_OK;
}

float fb_round_through_text(float distance) {
  char buf[16];
  snprintf(buf, sizeof buf, "%f", distance);
  return strtof(buf, nullptr);
}

}  // extern "C"

namespace {
int vec_row_of(const fb_engine* e, int32_t id) {
  if (id < 0) return -1;
  if (e->vec_ids_sorted) {
    auto it = std::lower_bound(e->vec_ids_host.begin(), e->vec_ids_host.end(), id);
    return (it != e->vec_ids_host.end() && *it == id) ? (int)(it - e->vec_ids_host.begin()) : -1;
  }
  auto it = e->vec_id_to_row.find(id);
  return it == e->vec_id_to_row.end() ? -1 : it->second;
}

// d_q: device [nq][d] query vectors; h_rows: host [nq][3] excluded table rows
int analogy_scan_fp32(fb_engine* e, const float* d_q, const std::vector<int32_t>& h_rows, int nq, int32_t* d_out_ids,
                      float* d_out_scores) {
  const int d = e->vec_d;
  const int64_t N = e->vec_N;
  const int tiles = (nq + kAnaQT - 1) / kAnaQT, nq_pad = tiles * kAnaQT;
  const int64_t n_blocks = (N + 31) / 32;
  const int blocks_per_slab = kAnaWarps * 8;
  const int n_slabs = (int)((n_blocks + blocks_per_slab - 1) / blocks_per_slab);
  FB_CUDA(e, e->ana_rows.ensure((size_t)nq * 3));
  FB_CUDA(e, e->ana_partial.ensure((size_t)std::max(1, n_slabs) * nq_pad));
  FB_CUDA(e, cudaMemcpyAsync(e->ana_rows.p, h_rows.data(), (size_t)nq * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));  // h_rows is the caller's stack/heap
  const size_t smem = ((size_t)d * kAnaQT + (size_t)kAnaWarps * 32 * 33) * sizeof(float);
  if (smem > e->smem_optin - 2048) return fail(e, FB_ERR_UNSUPPORTED, "d=%d too large for the analogy scan", d);
  FB_CUDA(e, cudaFuncSetAttribute(analogy_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (n_slabs > 0) {
    StageTimer t(e, ST_SCAN);
    dim3 grid(tiles, n_slabs);
    analogy_scan_kernel<<<grid, kAnaThreads, smem, e->stream>>>(e->vecT.p, N, d, blocks_per_slab, d_q, nq, e->ana_rows.p,
                                                                e->ana_partial.p, nq_pad, e->one);
    e->launches++;
    FB_CUDA(e, cudaGetLastError());
  }
  analogy_reduce_kernel<<<(nq + 127) / 128, 128, 0, e->stream>>>(e->ana_partial.p, n_slabs, nq, nq_pad, e->vec_ids.p, 0,
                                                                d_out_ids, d_out_scores, nullptr);
  e->launches++;
  FB_CUDA(e, cudaGetLastError());
  e->host_rows += (int64_t)nq * N;
  e->bytes_per_row = d * 4;
  return FB_OK;
}

// arg-max with exclusions: the tensor-core pre-filter with k' = 1 + 3 (the three excluded rows may occupy the top
// places), the fp32 scan for whatever it hands back
int analogy_scan_dev(fb_engine* e, const float* d_q, const std::vector<int32_t>& h_rows, int nq, int32_t* d_out_ids,
                     float* d_out_scores) {
  if (!pf_usable(e, 4)) return analogy_scan_fp32(e, d_q, h_rows, nq, d_out_ids, d_out_scores);
  const int d = e->vec_d;
  FB_CUDA(e, e->pf_ex.ensure((size_t)nq * 3));
  FB_CUDA(e, cudaMemcpyAsync(e->pf_ex.p, h_rows.data(), (size_t)nq * 3 * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
  std::vector<int32_t> ovf;
  int rc = knn_prefilter_dev(e, d_q, nq, 1, 4, e->pf_ex.p, d_out_ids, d_out_scores, nullptr, ovf);   // synchronises: h_rows may go away
  if (rc || ovf.empty()) return rc;
  const int no = (int)ovf.size();
  std::vector<int32_t> sub_rows((size_t)no * 3);
  FB_CUDA(e, e->vb.ensure((size_t)no * d));
  FB_CUDA(e, e->pv_cand.ensure((size_t)no));
  FB_CUDA(e, e->sub_vT.ensure((size_t)no));
  for (int i = 0; i < no; i++) {
    for (int j = 0; j < 3; j++) sub_rows[(size_t)i * 3 + j] = h_rows[(size_t)ovf[i] * 3 + j];
    FB_CUDA(e, cudaMemcpyAsync(e->vb.p + (size_t)i * d, d_q + (size_t)ovf[i] * d, (size_t)d * sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
  }
  if ((rc = analogy_scan_fp32(e, e->vb.p, sub_rows, no, e->pv_cand.p, e->sub_vT.p))) return rc;
  for (int i = 0; i < no; i++) {
    FB_CUDA(e, cudaMemcpyAsync(d_out_ids + ovf[i], e->pv_cand.p + i, sizeof(int32_t), cudaMemcpyDeviceToDevice, e->stream));
    FB_CUDA(e, cudaMemcpyAsync(d_out_scores + ovf[i], e->sub_vT.p + i, sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
  }
  return FB_OK;
}
}  // namespace

extern "C" {

int fb_load_vectors(fb_engine* e, const int32_t* ids, const float* vectors, int64_t N, int d) {
  if (!e || N < 0 || d < 1 || (N > 0 && (!ids || !vectors))) return fail(e, FB_ERR_INVALID, "fb_load_vectors: bad arguments");
  if (N >= (1ll << 31) - 64) return fail(e, FB_ERR_UNSUPPORTED, "table too large");
  FB_CUDA(e, cudaSetDevice(e->device));
  const int64_t n_blocks = std::max<int64_t>(1, (N + 31) / 32);
  FB_CUDA(e, e->vecT.ensure((size_t)n_blocks * d * 32));
  FB_CUDA(e, e->vec_ids.ensure((size_t)std::max<int64_t>(1, N)));
  // the rows land in the row-major image (what the gathers read); the dimension-major blocks of the whole-table scans
  // and the bf16 image of the pre-filter are made from it on the device, slice by slice
  const int64_t slice_rows = 32 * 8192;
  FB_CUDA(e, e->vecR.ensure((size_t)std::max<int64_t>(N, 1) * d));
  // bf16 image for the tensor-core pre-filter of the exact scans (d <= 320): [N_pad][kpa], zero padded
  e->pf_ready = false;
  const int kch = (d + kPfBK - 1) / kPfBK;
  const bool want_pf = kch <= kPfMaxKch && N > 0 && pf_encode_fn() != nullptr;
  if (want_pf) {
    e->pf_kpa = kch * kPfBK;
    e->pf_N_pad = (N + kPfBN - 1) / kPfBN * kPfBN;
    FB_CUDA(e, e->vec_bf16.ensure((size_t)e->pf_N_pad * e->pf_kpa));
    FB_CUDA(e, e->pf_norm.ensure(1));
    FB_CUDA(e, cudaMemsetAsync(e->vec_bf16.p, 0, (size_t)e->pf_N_pad * e->pf_kpa * sizeof(__nv_bfloat16), e->stream));
    FB_CUDA(e, cudaMemsetAsync(e->pf_norm.p, 0, sizeof(uint32_t), e->stream));
  }
  for (int64_t r0 = 0; r0 < N; r0 += slice_rows) {
    const int64_t n = std::min(slice_rows, N - r0);
    float* slice = e->vecR.p + (size_t)r0 * d;
    FB_CUDA(e, cudaMemcpyAsync(slice, vectors + (size_t)r0 * d, (size_t)n * d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    transpose_rows_kernel<<<(unsigned)((n + 31) / 32), 256, 0, e->stream>>>(slice, n, d, e->vecT.p + (size_t)(r0 / 32) * d * 32);
    if (want_pf)
      pf_rows_to_bf16_kernel<<<(unsigned)((n + 7) / 8), 256, 0, e->stream>>>(slice, n, d, e->pf_kpa, e->vec_bf16.p + (size_t)r0 * e->pf_kpa,
                                                                            e->pf_norm.p);
    FB_CUDA(e, cudaGetLastError());
  }
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  if (want_pf) {
    uint32_t nb = 0;
    FB_CUDA(e, cudaMemcpy(&nb, e->pf_norm.p, sizeof nb, cudaMemcpyDeviceToHost));
    float n2;
    memcpy(&n2, &nb, sizeof n2);
    e->pf_vmax = sqrtf(n2) * 1.000001f;
    // a non-finite row (or an all-zero table) leaves the exact scans on the fp32 kernels
    e->pf_ready = std::isfinite(e->pf_vmax) && e->pf_vmax > 0.0f &&
                  pf_make_tensor_map(&e->pf_tm_v, e->vec_bf16.p, e->pf_N_pad, e->pf_kpa, kPfBN);
  }
  if (N > 0) FB_CUDA(e, cudaMemcpy(e->vec_ids.p, ids, (size_t)N * sizeof(int32_t), cudaMemcpyHostToDevice));
  e->vec_ids_host.assign(ids, ids + N);
  e->vec_ids_sorted = std::is_sorted(e->vec_ids_host.begin(), e->vec_ids_host.end());
  e->vec_id_to_row.clear();
  if (!e->vec_ids_sorted)
    for (int64_t r = 0; r < N; r++) e->vec_id_to_row.emplace(ids[r], (int32_t)r);
  FB_CUDA(e, e->vec_sorted_ids.ensure((size_t)std::max<int64_t>(1, N)));
  FB_CUDA(e, e->vec_sorted_rows.ensure((size_t)std::max<int64_t>(1, N)));
  if (N > 0 && e->device_build) {
    int rc = device_id_index(e, e->vec_ids.p, N, e->vec_sorted_ids.p, e->vec_sorted_rows.p);
    if (rc) return rc;
  } else if (N > 0) {
    std::vector<int32_t> order((size_t)N), sid((size_t)N);
    for (int64_t r = 0; r < N; r++) order[r] = (int32_t)r;
    std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return ids[a] < ids[b]; });
    for (int64_t r = 0; r < N; r++) sid[r] = ids[order[r]];
    FB_CUDA(e, cudaMemcpy(e->vec_sorted_ids.p, sid.data(), (size_t)N * sizeof(int32_t), cudaMemcpyHostToDevice));
    FB_CUDA(e, cudaMemcpy(e->vec_sorted_rows.p, order.data(), (size_t)N * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  e->vec_N = N; e->vec_d = d; e->vec_loaded = true;
  return FB_OK;
}

__global__ void append_vec_rows_kernel(const float* __restrict__ rows, int64_t n, int d, int64_t first_row, float* __restrict__ vT) {
  const int64_t i = blockIdx.x;
  if (i >= n) return;
  const int64_t r = first_row + i;
  for (int j = threadIdx.x; j < d; j += blockDim.x) vT[((size_t)(r >> 5) * d + j) * 32 + (r & 31)] = rows[i * d + j];
}

// rows insert_batch adds to the word-vector table (updateWordVectorsRelation, index_utils.c:1045-1074), appended to the
// pinned image: dimension-major blocks, bf16 image of the pre-filter, id index
int fb_append_vectors(fb_engine* e, const int32_t* ids, const float* vectors, int64_t n) {
  if (!e || n < 0 || (n > 0 && (!ids || !vectors))) return fail(e, FB_ERR_INVALID, "fb_append_vectors: bad arguments");
  if (!e->vec_loaded) return fail(e, FB_ERR_INVALID, "word-vector table not loaded (fb_load_vectors)");
  if (n == 0) return FB_OK;
  FB_CUDA(e, cudaSetDevice(e->device));
  const int d = e->vec_d;
  const int64_t N0 = e->vec_N, N1 = N0 + n;
  if (N1 >= (1ll << 31) - 64) return fail(e, FB_ERR_UNSUPPORTED, "table too large");
  const int64_t b0 = std::max<int64_t>(1, (N0 + 31) / 32), b1 = (N1 + 31) / 32;
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  FB_CUDA(e, e->vecT.grow((size_t)b1 * d * 32, (size_t)b0 * d * 32));
  if (b1 > b0) FB_CUDA(e, cudaMemset(e->vecT.p + (size_t)b0 * d * 32, 0, (size_t)(b1 - b0) * d * 32 * sizeof(float)));
  FB_CUDA(e, e->vecR.grow((size_t)N1 * d, (size_t)N0 * d));
  float* fresh = e->vecR.p + (size_t)N0 * d;                 // the new rows, row-major, behind the old ones
  FB_CUDA(e, cudaMemcpy(fresh, vectors, (size_t)n * d * sizeof(float), cudaMemcpyHostToDevice));
  append_vec_rows_kernel<<<(unsigned)n, 128, 0, e->stream>>>(fresh, n, d, N0, e->vecT.p);
  e->launches++;
  FB_CUDA(e, e->vec_ids.grow((size_t)N1, (size_t)N0));
  FB_CUDA(e, cudaMemcpy(e->vec_ids.p + N0, ids, (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice));
  if (e->pf_ready || e->pf_kpa > 0) {
    const int64_t pad0 = e->pf_N_pad, pad1 = (N1 + kPfBN - 1) / kPfBN * kPfBN;
    FB_CUDA(e, e->vec_bf16.grow((size_t)pad1 * e->pf_kpa, (size_t)pad0 * e->pf_kpa));
    if (pad1 > pad0) FB_CUDA(e, cudaMemset(e->vec_bf16.p + (size_t)pad0 * e->pf_kpa, 0, (size_t)(pad1 - pad0) * e->pf_kpa * sizeof(__nv_bfloat16)));
    pf_rows_to_bf16_kernel<<<(unsigned)((n + 7) / 8), 256, 0, e->stream>>>(fresh, n, d, e->pf_kpa, e->vec_bf16.p + (size_t)N0 * e->pf_kpa, e->pf_norm.p);
    e->launches++;
    FB_CUDA(e, cudaGetLastError());
    uint32_t nb = 0;
    FB_CUDA(e, cudaMemcpyAsync(&nb, e->pf_norm.p, sizeof nb, cudaMemcpyDeviceToHost, e->stream));
    FB_CUDA(e, cudaStreamSynchronize(e->stream));
    float n2;
    memcpy(&n2, &nb, sizeof n2);
    e->pf_vmax = sqrtf(n2) * 1.000001f;
    e->pf_N_pad = pad1;
    e->pf_ready = std::isfinite(e->pf_vmax) && e->pf_vmax > 0.0f &&
                  pf_make_tensor_map(&e->pf_tm_v, e->vec_bf16.p, e->pf_N_pad, e->pf_kpa, kPfBN);
  }
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  // id -> row: host mirror and the sorted device image
  const bool tail = e->vec_ids_sorted && std::is_sorted(ids, ids + n) && (N0 == 0 || ids[0] >= e->vec_ids_host.back());
  e->vec_ids_host.insert(e->vec_ids_host.end(), ids, ids + n);
  if (!tail) {
    if (e->vec_ids_sorted) {
      e->vec_id_to_row.clear();
      for (int64_t r = 0; r < N0; r++) e->vec_id_to_row.emplace(e->vec_ids_host[r], (int32_t)r);
      e->vec_ids_sorted = false;
    }
    for (int64_t r = N0; r < N1; r++) e->vec_id_to_row.emplace(e->vec_ids_host[r], (int32_t)r);
  }
  {
    std::vector<int32_t> order((size_t)N1), sid((size_t)N1);
    for (int64_t r = 0; r < N1; r++) order[r] = (int32_t)r;
    const int32_t* all = e->vec_ids_host.data();
    if (!tail || !e->vec_ids_sorted) std::stable_sort(order.begin(), order.end(), [&](int32_t a, int32_t b) { return all[a] < all[b]; });
    for (int64_t r = 0; r < N1; r++) sid[r] = all[order[r]];
    FB_CUDA(e, e->vec_sorted_ids.ensure((size_t)N1));
    FB_CUDA(e, e->vec_sorted_rows.ensure((size_t)N1));
    FB_CUDA(e, cudaMemcpy(e->vec_sorted_ids.p, sid.data(), (size_t)N1 * sizeof(int32_t), cudaMemcpyHostToDevice));
    FB_CUDA(e, cudaMemcpy(e->vec_sorted_rows.p, order.data(), (size_t)N1 * sizeof(int32_t), cudaMemcpyHostToDevice));
  }
  e->vec_N = N1;
  e->graph_epoch++;
  return FB_OK;
}

int fb_cosine_similarity(fb_engine* e, int variant, const float* a, const float* b, int n, int d, double* out) {
  if (!e || variant < 0 || variant > 2 || n < 0 || d < 1) return fail(e, FB_ERR_INVALID, "fb_cosine_similarity: bad arguments");
  if (n == 0) return FB_OK;
  if (!a || !b || !out) return fail(e, FB_ERR_INVALID, "null buffer");
  FB_CUDA(e, cudaSetDevice(e->device));
  FB_CUDA(e, e->va.ensure((size_t)n * d));
  FB_CUDA(e, e->vb.ensure((size_t)n * d));
  FB_CUDA(e, e->vdo.ensure((size_t)n));
  FB_CUDA(e, cudaMemcpyAsync(e->va.p, a, (size_t)n * d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  FB_CUDA(e, cudaMemcpyAsync(e->vb.p, b, (size_t)n * d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  cosine_pairs_kernel<<<(n + 127) / 128, 128, 0, e->stream>>>(e->va.p, e->vb.p, n, d, variant, e->vdo.p);
  e->launches++;
  FB_CUDA(e, cudaGetLastError());
  FB_CUDA(e, cudaMemcpyAsync(out, e->vdo.p, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  return FB_OK;
}

int fb_vec_op(fb_engine* e, int op, const float* a, const float* b, int n, int d, float* out) {
  if (!e || op < 0 || op > 2 || n < 0 || d < 1) return fail(e, FB_ERR_INVALID, "fb_vec_op: bad arguments");
  if (n == 0) return FB_OK;
  if (!a || !out || (op != 2 && !b)) return fail(e, FB_ERR_INVALID, "null buffer");
  FB_CUDA(e, cudaSetDevice(e->device));
  FB_CUDA(e, e->va.ensure((size_t)n * d));
  FB_CUDA(e, e->vb.ensure((size_t)n * d));
  FB_CUDA(e, e->vo.ensure((size_t)n * d));
  FB_CUDA(e, cudaMemcpyAsync(e->va.p, a, (size_t)n * d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  if (op != 2) FB_CUDA(e, cudaMemcpyAsync(e->vb.p, b, (size_t)n * d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  vec_ops_kernel<<<n, 128, 0, e->stream>>>(e->va.p, e->vb.p, n, d, op, e->vo.p);
  e->launches++;
  FB_CUDA(e, cudaGetLastError());
  FB_CUDA(e, cudaMemcpyAsync(out, e->vo.p, (size_t)n * d * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  return FB_OK;
}

int fb_analogy_3cosadd(fb_engine* e, const int32_t* ids_abc, int nq, int32_t* out_ids, float* out_scores) {
  if (!e || nq < 0) return fail(e, FB_ERR_INVALID, "fb_analogy_3cosadd: bad arguments");
  if (!e->vec_loaded) return fail(e, FB_ERR_INVALID, "word-vector table not loaded (fb_load_vectors)");
  if (nq == 0) return FB_OK;
  if (!ids_abc || !out_ids || !out_scores) return fail(e, FB_ERR_INVALID, "null buffer");
  FB_CUDA(e, cudaSetDevice(e->device));
  std::vector<int32_t> rows((size_t)nq * 3);
  for (int i = 0; i < nq * 3; i++) {
    rows[i] = vec_row_of(e, ids_abc[i]);
    // the SQL inner joins on the three words: an unknown word yields no row at all
    if (rows[i] < 0) return fail(e, FB_ERR_INVALID, "analogy: id %d is not in the word-vector table", ids_abc[i]);
  }
  const int d = e->vec_d;
  FB_CUDA(e, e->ana_rows.ensure((size_t)nq * 3));
  FB_CUDA(e, e->vo.ensure((size_t)nq * d));
  FB_CUDA(e, e->id_stage.ensure((size_t)nq));
  FB_CUDA(e, e->dist_stage.ensure((size_t)nq));
  FB_CUDA(e, cudaMemcpyAsync(e->ana_rows.p, rows.data(), rows.size() * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
  analogy_query_kernel<<<nq, 128, 0, e->stream>>>(e->vecT.p, d, e->ana_rows.p, nq, e->vo.p);
  e->launches++;
  FB_CUDA(e, cudaGetLastError());
  int rc = analogy_scan_dev(e, e->vo.p, rows, nq, e->id_stage.p, e->dist_stage.p);
  if (rc) return rc;
  FB_CUDA(e, cudaMemcpyAsync(out_ids, e->id_stage.p, (size_t)nq * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaMemcpyAsync(out_scores, e->dist_stage.p, (size_t)nq * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  e->queries_done += nq;
  return FB_OK;
}

int fb_analogy_scan(fb_engine* e, const float* qvecs, const int32_t* exclude_ids, int nq, int32_t* out_ids, float* out_scores) {
  if (!e || nq < 0) return fail(e, FB_ERR_INVALID, "fb_analogy_scan: bad arguments");
  if (!e->vec_loaded) return fail(e, FB_ERR_INVALID, "word-vector table not loaded (fb_load_vectors)");
  if (nq == 0) return FB_OK;
  if (!qvecs || !out_ids || !out_scores) return fail(e, FB_ERR_INVALID, "null buffer");
  FB_CUDA(e, cudaSetDevice(e->device));
  std::vector<int32_t> rows((size_t)nq * 3, -1);
  if (exclude_ids)
    for (int i = 0; i < nq * 3; i++) rows[i] = vec_row_of(e, exclude_ids[i]);  // ids living on another shard: no local row
  const int d = e->vec_d;
  FB_CUDA(e, e->va.ensure((size_t)nq * d));
  FB_CUDA(e, e->id_stage.ensure((size_t)nq));
  FB_CUDA(e, e->dist_stage.ensure((size_t)nq));
  FB_CUDA(e, cudaMemcpyAsync(e->va.p, qvecs, (size_t)nq * d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  int rc = analogy_scan_dev(e, e->va.p, rows, nq, e->id_stage.p, e->dist_stage.p);
  if (rc) return rc;
  FB_CUDA(e, cudaMemcpyAsync(out_ids, e->id_stage.p, (size_t)nq * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaMemcpyAsync(out_scores, e->dist_stage.p, (size_t)nq * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  e->queries_done += nq;
  return FB_OK;
}

}  // extern "C"

extern "C" {

// exact cosine top-k over a dimension-major blocked table (the whole word-vector table or a gathered subset)
static int knn_exact_dev(fb_engine* e, const float* vT, int64_t N, const int32_t* d_row_map, const float* d_q, int nq, int k,
                         int32_t* d_out_ids, float* d_out_sims) {
  const int d = e->vec_d;
  const int tiles = (nq + kAnaQT - 1) / kAnaQT, nq_pad = tiles * kAnaQT;
  const int64_t n_blocks = (N + 31) / 32;
  const int blocks_per_slab = kAnaWarps * 8;
  const int n_slabs = (int)((n_blocks + blocks_per_slab - 1) / blocks_per_slab);
  FB_CUDA(e, e->knn_partial.ensure((size_t)std::max(1, n_slabs) * nq_pad * k));
  const size_t smem = ((size_t)d * kAnaQT + (size_t)kAnaWarps * 32 * 33) * sizeof(float) + (size_t)kAnaWarps * 32 * k * sizeof(u64);
  if (smem > e->smem_optin - 2048) return fail(e, FB_ERR_UNSUPPORTED, "d=%d, k=%d too large for the exact scan", d, k);
  FB_CUDA(e, cudaFuncSetAttribute(exact_knn_scan_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (n_slabs > 0) {
    StageTimer t(e, ST_SCAN);
    dim3 grid(tiles, n_slabs);
    exact_knn_scan_kernel<<<grid, kAnaThreads, smem, e->stream>>>(vT, N, d, blocks_per_slab, d_q, nq, k, e->knn_partial.p,
                                                                  nq_pad, e->one);
    e->launches++;
    FB_CUDA(e, cudaGetLastError());
  }
  exact_knn_reduce_kernel<<<(nq + 63) / 64, 64, 0, e->stream>>>(e->knn_partial.p, n_slabs, nq, nq_pad, k, e->vec_ids.p,
                                                               d_row_map, d_out_ids, d_out_sims);
  e->launches++;
  FB_CUDA(e, cudaGetLastError());
  e->host_rows += (int64_t)nq * N;
  e->bytes_per_row = d * 4;
  return FB_OK;
}

}  // extern "C"

// Exact top-k over the WHOLE word-vector table through the tensor-core pre-filter (prefilter_kernels.cuh):
// bf16 tcgen05 scores select candidates, the reference's fp32 chain decides.  d_q: device [nq][d]; d_exclude:
// device [nq][3] table rows that never win (analogy) or nullptr; kk = k + excluded rows.  Queries whose candidate
// buffer overflowed come back in `overflow` (query indices): the caller re-does them with the fp32 scan.
static int knn_prefilter_dev(fb_engine* e, const float* d_q, int nq, int k, int kk, const int32_t* d_exclude,
                             int32_t* d_out_ids, float* d_out_sims, int32_t* d_out_rows, std::vector<int32_t>& overflow) {
  const int d = e->vec_d, kpa = e->pf_kpa, kch = kpa / kPfBK;
  const int64_t N = e->vec_N;
  const int n_vt = (int)(e->pf_N_pad / kPfBN);
  const int q_batch = 8192;                                   // bounds the candidate buffers (8192 x 4096 x 8 B = 256 MB)
  const int nb_max = std::min(nq, q_batch), nb_pad = (nb_max + kPfBM - 1) / kPfBM * kPfBM;
  FB_CUDA(e, e->pf_qb.ensure((size_t)nb_pad * kpa));
  FB_CUDA(e, e->pf_eps2.ensure((size_t)nb_pad));
  FB_CUDA(e, e->pf_gbest.ensure((size_t)nb_pad * kPfMaxK));
  FB_CUDA(e, e->pf_cnt.ensure((size_t)nb_pad));
  FB_CUDA(e, e->pf_cand.ensure((size_t)nb_pad * kPfCandCap));
  FB_CUDA(e, e->pf_ovf.ensure((size_t)nb_pad + 1));
  FB_CUDA(e, e->pf_units.ensure((size_t)pf_max_units(nb_pad / kPfBM, e->num_sms)));
  FB_CUDA(e, e->pf_progress.ensure((size_t)pf_max_units(nb_pad / kPfBM, e->num_sms)));
  FB_CUDA(e, cudaFuncSetAttribute(prefilter_gemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)PfSmem::total));
  std::vector<PfUnit> units((size_t)pf_max_units(nb_pad / kPfBM, e->num_sms));
  std::vector<int32_t> h_ovf((size_t)nb_pad + 1), h_cnt;
  for (int q0 = 0; q0 < nq; q0 += q_batch) {
    const int n = std::min(q_batch, nq - q0), n_pad = (n + kPfBM - 1) / kPfBM * kPfBM, QT = n_pad / kPfBM;
    int n_units = 0;
    pf_make_units(QT, n_vt, e->num_sms, units.data(), &n_units);
    CUtensorMap tm_q;
    if (!pf_make_tensor_map(&tm_q, e->pf_qb.p, n_pad, kpa, kPfBM)) return fail(e, FB_ERR_CUDA, "cuTensorMapEncodeTiled failed");
    FB_CUDA(e, cudaMemcpyAsync(e->pf_units.p, units.data(), (size_t)n_units * sizeof(PfUnit), cudaMemcpyHostToDevice, e->stream));
    FB_CUDA(e, cudaMemsetAsync(e->pf_ovf.p, 0, sizeof(int32_t), e->stream));
    FB_CUDA(e, cudaMemsetAsync(e->pf_progress.p, 0, (size_t)n_units * sizeof(int32_t), e->stream));
    pf_queries_prepare_kernel<<<(n_pad + 7) / 8, 256, 0, e->stream>>>(d_q + (size_t)q0 * d, n, n_pad, d, kpa, e->pf_vmax, e->pf_qb.p,
                                                                     e->pf_eps2.p, e->pf_gbest.p, e->pf_cnt.p);
    e->launches++;
    PfArgs a;
    memset(&a, 0, sizeof a);
    a.units = e->pf_units.p; a.n_units = n_units; a.kch = kch; a.ksteps = (d + 15) / 16; a.N = N; a.nq = n; a.kk = kk;
    a.eps2 = e->pf_eps2.p; a.gbest = e->pf_gbest.p; a.cand_cnt = e->pf_cnt.p; a.cand = e->pf_cand.p; a.cap = kPfCandCap;
    a.progress = e->pf_progress.p; a.n_qt = QT;
    // only when every unit has its own resident CTA; progress is published every 4th tile, so the window is >= 8
    a.lockstep = (n_units <= e->num_sms && QT > 1 && e->pf_lockstep > 0) ? std::max(8, e->pf_lockstep) : 0;
    {
      StageTimer t(e, ST_SCAN);
      prefilter_gemm_kernel<<<std::min(e->num_sms, n_units), kPfThreads, PfSmem::total, e->stream>>>(tm_q, e->pf_tm_v, a);
      e->launches++;
      e->n_scan_launches++;
      FB_CUDA(e, cudaGetLastError());
    }
    {
      StageTimer t(e, ST_FINALIZE);
      const size_t smem = (size_t)kPfCandCap * sizeof(u64) + (size_t)d * sizeof(float);
      pf_rescore_kernel<<<n, kPfRescoreThreads, smem, e->stream>>>(
          d_q + (size_t)q0 * d, d, e->vecR.p, e->pf_cnt.p, e->pf_cand.p, kPfCandCap, kk, e->pf_eps2.p,
          d_exclude ? d_exclude + (size_t)q0 * 3 : nullptr, k, e->vec_ids.p, d_out_ids + (size_t)q0 * k, d_out_sims + (size_t)q0 * k,
          d_out_rows ? d_out_rows + (size_t)q0 * k : nullptr, e->pf_ovf.p + 1, e->pf_ovf.p, nullptr);
      e->launches++;
      FB_CUDA(e, cudaGetLastError());
    }
    FB_CUDA(e, cudaMemcpyAsync(h_ovf.data(), e->pf_ovf.p, ((size_t)n + 1) * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    if (e->profile) {
      h_cnt.resize((size_t)n);
      FB_CUDA(e, cudaMemcpyAsync(h_cnt.data(), e->pf_cnt.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
    }
    FB_CUDA(e, cudaStreamSynchronize(e->stream));
    for (int i = 0; i < h_ovf[0]; i++) overflow.push_back(q0 + h_ovf[1 + i]);
    if (e->profile) for (int i = 0; i < n; i++) e->pf_candidates += h_cnt[i];
  }
  e->pf_queries += nq;
  e->pf_overflow_queries += (int64_t)overflow.size();
  e->host_rows += (int64_t)nq * N;
  e->bytes_per_row = d * 4;
  return FB_OK;
}

static bool pf_usable(const fb_engine* e, int kk) {
  return e->prefilter && e->pf_ready && kk <= kPfMaxK;
}

extern "C" {

int fb_knn_exact(fb_engine* e, const float* queries, int nq, int k, const int32_t* targets, int n_targets,
                 int32_t* out_ids, float* out_sims) {
  if (!e || nq < 0) return fail(e, FB_ERR_INVALID, "fb_knn_exact: bad arguments");
  if (k < 1 || k > kKnnMaxK) return fail(e, FB_ERR_UNSUPPORTED, "fb_knn_exact: k=%d outside [1,%d]", k, kKnnMaxK);
  if (!e->vec_loaded) return fail(e, FB_ERR_INVALID, "word-vector table not loaded (fb_load_vectors)");
  if (n_targets < 0 || (n_targets > 0 && !targets)) return fail(e, FB_ERR_INVALID, "bad target array");
  if (nq == 0) return FB_OK;
  if (!queries || !out_ids || !out_sims) return fail(e, FB_ERR_INVALID, "null buffer");
  FB_CUDA(e, cudaSetDevice(e->device));
  const int d = e->vec_d;
  FB_CUDA(e, e->va.ensure((size_t)nq * d));
  FB_CUDA(e, e->id_stage.ensure((size_t)nq * k));
  FB_CUDA(e, e->dist_stage.ensure((size_t)nq * k));
  FB_CUDA(e, cudaMemcpyAsync(e->va.p, queries, (size_t)nq * d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  int rc;
  if (targets == nullptr && pf_usable(e, k)) {
    std::vector<int32_t> ovf;
    rc = knn_prefilter_dev(e, e->va.p, nq, k, k, nullptr, e->id_stage.p, e->dist_stage.p, nullptr, ovf);
    if (rc == FB_OK && !ovf.empty()) {
      // overflowed queries (heavy duplication around the k-th place): the fp32 scan answers them
      const int no = (int)ovf.size();
      FB_CUDA(e, e->vb.ensure((size_t)no * d));
      FB_CUDA(e, e->pv_cand.ensure((size_t)no * k));
      FB_CUDA(e, e->vo.ensure((size_t)no * k));
      for (int i = 0; i < no; i++)
        FB_CUDA(e, cudaMemcpyAsync(e->vb.p + (size_t)i * d, e->va.p + (size_t)ovf[i] * d, (size_t)d * sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
      rc = knn_exact_dev(e, e->vecT.p, e->vec_N, nullptr, e->vb.p, no, k, e->pv_cand.p, e->vo.p);
      for (int i = 0; i < no && rc == FB_OK; i++) {
        FB_CUDA(e, cudaMemcpyAsync(e->id_stage.p + (size_t)ovf[i] * k, e->pv_cand.p + (size_t)i * k, (size_t)k * sizeof(int32_t), cudaMemcpyDeviceToDevice, e->stream));
        FB_CUDA(e, cudaMemcpyAsync(e->dist_stage.p + (size_t)ovf[i] * k, e->vo.p + (size_t)i * k, (size_t)k * sizeof(float), cudaMemcpyDeviceToDevice, e->stream));
      }
    }
  } else if (targets == nullptr) {
    rc = knn_exact_dev(e, e->vecT.p, e->vec_N, nullptr, e->va.p, nq, k, e->id_stage.p, e->dist_stage.p);
  } else {
    // WHERE id = ANY(targets): the matching rows in table order, each once (freddy--0.0.1.sql:1026-1038)
    std::vector<int32_t> rows;
    rows.reserve(n_targets);
    for (int i = 0; i < n_targets; i++) { const int r = vec_row_of(e, targets[i]); if (r >= 0) rows.push_back(r); }
    std::sort(rows.begin(), rows.end());
    rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
    const int n = (int)rows.size();
    const int nb = std::max(1, (n + 31) / 32);
    FB_CUDA(e, e->sub_rows.ensure((size_t)std::max(1, n)));
    FB_CUDA(e, e->sub_vT.ensure((size_t)nb * d * 32));
    if (n > 0) {
      FB_CUDA(e, cudaMemcpyAsync(e->sub_rows.p, rows.data(), (size_t)n * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
      gather_vec_blocks_kernel<<<nb, 256, 0, e->stream>>>(e->vecR.p, d, e->sub_rows.p, n, e->sub_vT.p);
      e->launches++;
      FB_CUDA(e, cudaGetLastError());
      FB_CUDA(e, cudaStreamSynchronize(e->stream));   // rows is a local
    }
    rc = knn_exact_dev(e, e->sub_vT.p, n, e->sub_rows.p, e->va.p, nq, k, e->id_stage.p, e->dist_stage.p);
  }
  if (rc) return rc;
  FB_CUDA(e, cudaMemcpyAsync(out_ids, e->id_stage.p, (size_t)nq * k * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaMemcpyAsync(out_sims, e->dist_stage.p, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  e->queries_done += nq;
  return FB_OK;
}

int fb_ivfadc_search_pv(fb_engine* e, const float* queries, int nq, int k, int pvf, int w, int32_t* out_ids, float* out_sims) {
  if (!e || nq < 0) return fail(e, FB_ERR_INVALID, "fb_ivfadc_search_pv: bad arguments");
  if (k < 1 || pvf < 1) return fail(e, FB_ERR_INVALID, "fb_ivfadc_search_pv: k=%d pvf=%d", k, pvf);
  const int64_t kp64 = (int64_t)k * pvf;
  if (kp64 > kExactMaxK || kp64 > kPvMaxCand) return fail(e, FB_ERR_UNSUPPORTED, "pvf*k=%lld outside [1,%d]", (long long)kp64, kExactMaxK);
  const int kp = (int)kp64;
  if (!e->vec_loaded) return fail(e, FB_ERR_INVALID, "post-verification joins the word-vector table: fb_load_vectors first");
  if (e->vec_d != e->d) return fail(e, FB_ERR_INVALID, "word vectors have d=%d, IVFADC index d=%d", e->vec_d, e->d);
  if (nq == 0) return FB_OK;
  if (!queries || !out_ids || !out_sims) return fail(e, FB_ERR_INVALID, "null buffer");
  FB_CUDA(e, cudaSetDevice(e->device));
  const int d = e->d;
  FB_CUDA(e, e->q_stage.ensure((size_t)nq * d));
  FB_CUDA(e, e->pv_cand.ensure((size_t)nq * kp));
  FB_CUDA(e, e->vo.ensure((size_t)nq * kp));
  FB_CUDA(e, e->id_stage.ensure((size_t)nq * k));
  FB_CUDA(e, e->dist_stage.ensure((size_t)nq * k));
  FB_CUDA(e, cudaMemcpyAsync(e->q_stage.p, queries, (size_t)nq * d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  // candidates: ivfadc_search(v, pvf * k)  (freddy--0.0.1.sql:583-584)
  int rc = ivfadc_dev(e, e->q_stage.p, nq, kp, w, e->pv_cand.p, e->vo.p);
  if (rc) return rc;
  int n_pad = 32;
  while (n_pad < kp) n_pad <<= 1;
  const size_t smem = (size_t)((d + 3) & ~3) * sizeof(float) + (size_t)n_pad * sizeof(u64);
  {
    StageTimer t(e, ST_FINALIZE);
    pv_rerank_kernel<<<nq, kPvThreads, smem, e->stream>>>(e->q_stage.p, d, e->pv_cand.p, kp, k, e->vecR.p, e->vec_ids.p,
                                                          e->vec_sorted_ids.p, e->vec_sorted_rows.p, (int)e->vec_N,
                                                          e->id_stage.p, e->dist_stage.p);
    e->launches++;
    FB_CUDA(e, cudaGetLastError());
  }
  FB_CUDA(e, cudaMemcpyAsync(out_ids, e->id_stage.p, (size_t)nq * k * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaMemcpyAsync(out_sims, e->dist_stage.p, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  return check_error_flag(e);
}

// k_nearest_neighbour_pq_pv (freddy--0.0.1.sql:624-662): candidates = pq_search(v, pvf * k), joined with the word
// vectors by id, ordered by the exact cosine_similarity_bytea, first k.  (As written, the reference's SQL hands the
// candidate's WORD to cosine_similarity_bytea and cannot run; this is the query its ivfadc twin at :574-591 spells out.)
int fb_pq_search_pv(fb_engine* e, const float* queries, int nq, int k, int pvf, int32_t* out_ids, float* out_sims) {
  if (!e || nq < 0) return fail(e, FB_ERR_INVALID, "fb_pq_search_pv: bad arguments");
  if (k < 1 || pvf < 1) return fail(e, FB_ERR_INVALID, "fb_pq_search_pv: k=%d pvf=%d", k, pvf);
  const int64_t kp64 = (int64_t)k * pvf;
  if (kp64 > kExactMaxK || kp64 > kPvMaxCand) return fail(e, FB_ERR_UNSUPPORTED, "pvf*k=%lld outside [1,%d]", (long long)kp64, kExactMaxK);
  const int kp = (int)kp64;
  int rc = check_pq_ready(e);
  if (rc) return rc;
  const int d = pq_dim(e);
  if (!e->vec_loaded) return fail(e, FB_ERR_INVALID, "post-verification joins the word-vector table: fb_load_vectors first");
  if (e->vec_d != d) return fail(e, FB_ERR_INVALID, "word vectors have d=%d, PQ index d=%d", e->vec_d, d);
  if (nq == 0) return FB_OK;
  if (!queries || !out_ids || !out_sims) return fail(e, FB_ERR_INVALID, "null buffer");
  FB_CUDA(e, cudaSetDevice(e->device));
  FB_CUDA(e, e->q_stage.ensure((size_t)nq * d));
  FB_CUDA(e, e->pv_cand.ensure((size_t)nq * kp));
  FB_CUDA(e, e->vo.ensure((size_t)nq * kp));
  FB_CUDA(e, e->id_stage.ensure((size_t)nq * k));
  FB_CUDA(e, e->dist_stage.ensure((size_t)nq * k));
  FB_CUDA(e, cudaMemcpyAsync(e->q_stage.p, queries, (size_t)nq * d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  if ((rc = pq_dev(e, e->pq.dev(), e->pq.N, e->q_stage.p, nq, kp, 100.0f, e->pv_cand.p, e->vo.p))) return rc;   // pq_search: sentinel 100.0 (freddy.c:90-92)
  e->host_rows += (int64_t)nq * e->pq.N;
  int n_pad = 32;
  while (n_pad < kp) n_pad <<= 1;
  const size_t smem = (size_t)((d + 3) & ~3) * sizeof(float) + (size_t)n_pad * sizeof(u64);
  {
    StageTimer t(e, ST_FINALIZE);
    pv_rerank_kernel<<<nq, kPvThreads, smem, e->stream>>>(e->q_stage.p, d, e->pv_cand.p, kp, k, e->vecR.p, e->vec_ids.p,
                                                          e->vec_sorted_ids.p, e->vec_sorted_rows.p, (int)e->vec_N,
                                                          e->id_stage.p, e->dist_stage.p);
    e->launches++;
    FB_CUDA(e, cudaGetLastError());
  }
  FB_CUDA(e, cudaMemcpyAsync(out_ids, e->id_stage.p, (size_t)nq * k * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaMemcpyAsync(out_sims, e->dist_stage.p, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  return check_error_flag(e);
}

// quantisation of new rows as insert_batch assigns them (freddy.c:1567-1582, index_utils.c:923-939):
// coarse argmin (optional) -> residual LUT rows -> first minimum per position
__global__ void any_coarse_far_kernel(const uint32_t* __restrict__ qflags, int n, int32_t* __restrict__ flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && (qflags[i] & kWhyCoarseFar)) atomicExch(flag, 1);
}

// on_device: vectors / out_coarse_ids / out_codes are DEVICE pointers (index build: the rows are already in HBM)
static int encode_impl(fb_engine* e, const Codebook& cb, bool with_coarse, const float* vectors, int64_t n,
                       int32_t* out_coarse_ids, int16_t* out_codes, bool on_device = false) {
  if (!cb.loaded) return fail(e, FB_ERR_INVALID, "codebook not loaded");
  if (with_coarse && !e->coarse_loaded) return fail(e, FB_ERR_INVALID, "coarse table not loaded");
  const int d = cb.m * cb.sub, m = cb.m, K = cb.K;
  if (with_coarse && d != e->d) return fail(e, FB_ERR_INVALID, "codebook d=%d, coarse table d=%d", d, e->d);
  if (n == 0) return FB_OK;
  if (!vectors || !out_codes || (with_coarse && !out_coarse_ids)) return fail(e, FB_ERR_INVALID, "null buffer");
  FB_CUDA(e, cudaSetDevice(e->device));
  const int64_t chunk = std::min<int64_t>(n, on_device ? 16384 : 4096);
  DevBuf<int16_t> d_codes;
  if (!on_device) FB_CUDA(e, e->q_stage.ensure((size_t)chunk * d));
  FB_CUDA(e, e->lut.ensure((size_t)chunk * m * K));
  if (!on_device) FB_CUDA(e, d_codes.ensure((size_t)chunk * m));
  FB_CUDA(e, e->probes.ensure((size_t)chunk));
  FB_CUDA(e, e->qflags.ensure((size_t)chunk));
  FB_CUDA(e, cudaMemsetAsync(e->small.p + 2, 0, 2 * sizeof(int32_t), e->stream));   // [2] sub-vector far, [3] coarse far
  std::vector<uint32_t> h_flags;
  int rc = FB_OK;
  for (int64_t r0 = 0; r0 < n && rc == FB_OK; r0 += chunk) {
    const int c = (int)std::min<int64_t>(chunk, n - r0);
    const float* dq = on_device ? vectors + (size_t)r0 * d : e->q_stage.p;
    int16_t* dc = on_device ? out_codes + (size_t)r0 * m : d_codes.p;
    if (!on_device)
      FB_CUDA(e, cudaMemcpyAsync(e->q_stage.p, vectors + (size_t)r0 * d, (size_t)c * d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
    if (with_coarse) {
      // the w = 1 selection of the coarse kernel is the reference's argmin: smallest (distance, id) key
      rc = launch_coarse(e, dq, c, 1, 1);   // (list lengths only feed the search flags; may be absent)
      if (rc) break;
      rc = launch_lut(e, cb, dq, e->coarse.p, e->probes.p, 1, c, e->lut.p);
    } else {
      rc = launch_lut(e, cb, dq, nullptr, nullptr, 1, c, e->lut.p);
    }
    if (rc) break;
    const int64_t warps = (int64_t)c * m;
    lut_argmin_kernel<<<(unsigned)((warps + 7) / 8), 256, 0, e->stream>>>(e->lut.p, c, m, K, dc, e->small.p + 2);
    e->launches++;
    FB_CUDA(e, cudaGetLastError());
    if (on_device) {
      if (with_coarse) {
        FB_CUDA(e, cudaMemcpyAsync(out_coarse_ids + r0, e->probes.p, (size_t)c * sizeof(int32_t), cudaMemcpyDeviceToDevice, e->stream));
        any_coarse_far_kernel<<<(c + 255) / 256, 256, 0, e->stream>>>(e->qflags.p, c, e->small.p + 3);
        e->launches++;
      }
      continue;
    }
    FB_CUDA(e, cudaMemcpyAsync(out_codes + (size_t)r0 * m, d_codes.p, (size_t)c * m * sizeof(int16_t), cudaMemcpyDeviceToHost, e->stream));
    if (with_coarse) {
      FB_CUDA(e, cudaMemcpyAsync(out_coarse_ids + r0, e->probes.p, (size_t)c * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
      h_flags.resize(c);
      FB_CUDA(e, cudaMemcpyAsync(h_flags.data(), e->qflags.p, (size_t)c * sizeof(uint32_t), cudaMemcpyDeviceToHost, e->stream));
    }
    FB_CUDA(e, cudaStreamSynchronize(e->stream));
    if (with_coarse)
      for (int i = 0; i < c; i++)
        if (h_flags[i] & kWhyCoarseFar)
          return fail(e, FB_ERR_REFERENCE_UB, "row %lld: every coarse distance >= 100, the reference's assignment is uninitialised (freddy.c:1570-1577)",
                      (long long)(r0 + i));
  }
  if (rc) return rc;
  int32_t flag[2] = {0, 0};
  FB_CUDA(e, cudaMemcpyAsync(flag, e->small.p + 2, sizeof flag, cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  if (flag[0] || flag[1]) {
    cudaMemset(e->small.p + 2, 0, 2 * sizeof(int32_t));
    if (flag[1]) return fail(e, FB_ERR_REFERENCE_UB, "a row is >= 100 away from every coarse centroid: the reference's assignment is uninitialised (freddy.c:1570-1577)");
    return fail(e, FB_ERR_REFERENCE_UB, "a sub-vector is >= 100 away from every codeword: the reference's assignment is uninitialised (index_utils.c:926-938)");
  }
  return FB_OK;
}

int fb_encode_ivfadc_dev(fb_engine* e, const float* d_vectors, int64_t n, int32_t* d_out_coarse_ids, int16_t* d_out_codes) {
  if (!e || n < 0) return fail(e, FB_ERR_INVALID, "fb_encode_ivfadc_dev: bad arguments");
  return encode_impl(e, e->cb[FB_CB_RESIDUAL], true, d_vectors, n, d_out_coarse_ids, d_out_codes, true);
}

int fb_encode_pq_dev(fb_engine* e, int kind, const float* d_vectors, int64_t n, int16_t* d_out_codes) {
  if (!e || n < 0 || kind < 0 || kind >= FB_CB_KINDS) return fail(e, FB_ERR_INVALID, "fb_encode_pq_dev: bad arguments");
  return encode_impl(e, e->cb[kind], false, d_vectors, n, nullptr, d_out_codes, true);
}

int fb_encode_ivfadc(fb_engine* e, const float* vectors, int64_t n, int32_t* out_coarse_ids, int16_t* out_codes) {
  if (!e || n < 0) return fail(e, FB_ERR_INVALID, "fb_encode_ivfadc: bad arguments");
  return encode_impl(e, e->cb[FB_CB_RESIDUAL], true, vectors, n, out_coarse_ids, out_codes);
}

int fb_encode_pq(fb_engine* e, int kind, const float* vectors, int64_t n, int16_t* out_codes) {
  if (!e || n < 0 || kind < 0 || kind >= FB_CB_KINDS) return fail(e, FB_ERR_INVALID, "fb_encode_pq: bad arguments");
  return encode_impl(e, e->cb[kind], false, vectors, n, nullptr, out_codes);
}

int fb_load_ivpq(fb_engine* e, const float* coarse_multi, int Kc, int d, const int32_t* ids, const int32_t* coarse_ids,
                 const int16_t* codes, int64_t N, int m, const float* stats) {
  if (!e || !coarse_multi || Kc < 1 || d < 2 || (N > 0 && (!ids || !coarse_ids || !codes)))
    return fail(e, FB_ERR_INVALID, "fb_load_ivpq: bad arguments");
  if (!stats && N == 0) return fail(e, FB_ERR_INVALID, "fb_load_ivpq: statistics of an empty table are undefined (division by zero in create_statistics)");
  if (Kc > 32) return fail(e, FB_ERR_UNSUPPORTED, "Kc=%d: the cell-selection kernel handles up to 32 x 32 cells", Kc);
  if (d % 2) return fail(e, FB_ERR_INVALID, "d must be even for the 2-way multi-index");
  if (!e->cb[FB_CB_IVPQ].loaded) return fail(e, FB_ERR_INVALID, "fb_load_ivpq: load the ivpq codebook first");
  FB_CUDA(e, cudaSetDevice(e->device));
  const int cells = Kc * Kc;
  for (int64_t r = 0; r < N; r++)
    if (coarse_ids[r] < 0 || coarse_ids[r] >= cells) return fail(e, FB_ERR_INVALID, "row %lld: coarse_id %d out of range", (long long)r, coarse_ids[r]);
  int rc = build_table(e, e->ivpq, ids, nullptr, 0, 0x7fffffff, codes, N, m, e->cb[FB_CB_IVPQ].K);   // one list, table order
  if (rc) return rc;
  if ((rc = build_id_index(e, e->ivpq, ids, N))) return rc;
  FB_CUDA(e, e->coarse_multi.ensure((size_t)2 * Kc * (d / 2)));
  FB_CUDA(e, e->ivpq_stats.ensure((size_t)cells + 1));
  FB_CUDA(e, e->ivpq_cells.ensure((size_t)std::max<int64_t>(1, N)));
  FB_CUDA(e, cudaMemcpy(e->coarse_multi.p, coarse_multi, (size_t)2 * Kc * (d / 2) * sizeof(float), cudaMemcpyHostToDevice));
  if (N > 0) FB_CUDA(e, cudaMemcpy(e->ivpq_cells.p, coarse_ids, (size_t)N * sizeof(int32_t), cudaMemcpyHostToDevice));
  e->ivpq_Kc = Kc; e->ivpq_d = d; e->ivpq_loaded = true;
  if (stats == nullptr) return fb_ivpq_statistics(e, nullptr, 0, nullptr, 1);   // over every row of the table
  FB_CUDA(e, cudaMemcpy(e->ivpq_stats.p, stats, ((size_t)cells + 1) * sizeof(float), cudaMemcpyHostToDevice));
  e->ivpq_stats_host.assign(stats, stats + cells + 1);
  return FB_OK;
}

// create_statistics (freddy--0.0.1.sql:150-171) on the device: cell frequencies of the table rows whose id is listed
// (ids = NULL: all rows), coarse_freq = (count::float8 / total)::float4, last entry = total
int fb_ivpq_statistics(fb_engine* e, const int32_t* ids, int64_t n_ids, float* out_stats, int install) {
  if (!e || n_ids < 0 || n_ids > 0x7fffffff) return fail(e, FB_ERR_INVALID, "fb_ivpq_statistics: bad arguments");
  if (!e->ivpq_loaded) return fail(e, FB_ERR_INVALID, "IVPQ index not loaded (fb_load_ivpq)");
  FB_CUDA(e, cudaSetDevice(e->device));
  const int cells = e->ivpq_Kc * e->ivpq_Kc;
  const int64_t N = e->ivpq.N;
  DevBuf<unsigned long long> d_counts;
  DevBuf<float> d_stats;
  DevBuf<int32_t> d_wanted;
  FB_CUDA(e, d_counts.ensure((size_t)cells + 1));
  FB_CUDA(e, d_stats.ensure((size_t)cells + 1));
  FB_CUDA(e, cudaMemsetAsync(d_counts.p, 0, ((size_t)cells + 1) * sizeof(unsigned long long), e->stream));
  if (ids == nullptr) {
    if (N > 0) cell_count_all_kernel<<<(unsigned)((N + 255) / 256), 256, 0, e->stream>>>(e->ivpq_cells.p, N, d_counts.p, cells);
  } else if (n_ids > 0) {
    FB_CUDA(e, d_wanted.ensure((size_t)n_ids));
    FB_CUDA(e, cudaMemcpyAsync(d_wanted.p, ids, (size_t)n_ids * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
    cell_count_listed_kernel<<<(unsigned)((n_ids + 255) / 256), 256, 0, e->stream>>>(e->ivpq.sorted_ids.p, e->ivpq.sorted_rows.p, (int)N, d_wanted.p,
                                                                                     (int)n_ids, e->ivpq_cells.p, d_counts.p, cells);
  }
  cell_freq_kernel<<<(cells + 256) / 256, 256, 0, e->stream>>>(d_counts.p, cells, d_stats.p);
  e->launches += 2;
  FB_CUDA(e, cudaGetLastError());
  std::vector<float> h((size_t)cells + 1);
  FB_CUDA(e, cudaMemcpyAsync(h.data(), d_stats.p, h.size() * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  if (!(h[cells] > 0.0f)) return fail(e, FB_ERR_INVALID, "fb_ivpq_statistics: no table row matches (create_statistics divides by the match count)");
  if (out_stats) memcpy(out_stats, h.data(), h.size() * sizeof(float));
  if (install) {
    FB_CUDA(e, e->ivpq_stats.ensure((size_t)cells + 1));
    FB_CUDA(e, cudaMemcpy(e->ivpq_stats.p, h.data(), h.size() * sizeof(float), cudaMemcpyHostToDevice));
    e->ivpq_stats_host = h;
    e->graph_epoch++;
  }
  return FB_OK;
}

// ---- sidecar: one engine answering the single-query calls of many backends (include/freddy_sidecar.h) ----
struct fb_sidecar {
  fb_engine* e = nullptr;
  fbsc_server* srv = nullptr;
  std::thread worker;
  void* pinned[3] = {nullptr, nullptr, nullptr};
  int run_rc = 0;
  std::atomic<int> running{1};
};

static int sidecar_batch(void* ctx, const float* queries, int nq, int k, int w, int32_t* out_ids, float* out_dists) {
  cudaSetDevice(((fb_engine*)ctx)->device);     // the server loop runs on a thread of its own
  return fb_ivfadc_search((fb_engine*)ctx, queries, nq, k, w, out_ids, out_dists);
}

int fb_sidecar_start(fb_engine* e, const char* name, int max_k, int slots, int max_batch, int linger_us, fb_sidecar** out) {
  if (!e || !name || !out || max_k < 1 || slots < 1 || max_batch < 1) return fail(e, FB_ERR_INVALID, "fb_sidecar_start: bad arguments");
  if (!e->fine.loaded) return fail(e, FB_ERR_INVALID, "fb_sidecar_start: IVFADC index not loaded");
  FB_CUDA(e, cudaSetDevice(e->device));
  fb_sidecar* sc = new (std::nothrow) fb_sidecar;
  if (!sc) return fail(e, FB_ERR_INVALID, "out of memory");
  sc->e = e;
  int rc = fbsc_server_create(name, e->d, max_k, slots, &sc->srv);
  if (rc) { delete sc; return fail(e, FB_ERR_INVALID, "fbsc_server_create(%s) failed: %d", name, rc); }
  float* bq; int32_t* bi; float* bd;
  max_batch = std::min(max_batch, slots);
  if ((rc = fbsc_server_buffers(sc->srv, max_batch, &bq, &bi, &bd))) { fbsc_server_destroy(sc->srv); delete sc; return fail(e, FB_ERR_INVALID, "sidecar buffers: %d", rc); }
  // page-locked batch buffers: the engine's small-batch path copies from / to them without a staging pass
  const size_t al = 4096;
  const size_t sz[3] = {((size_t)max_batch * e->d * 4 + al - 1) / al * al, ((size_t)max_batch * max_k * 4 + al - 1) / al * al,
                        ((size_t)max_batch * max_k * 4 + al - 1) / al * al};
  void* ptr[3] = {bq, bi, bd};
  for (int i = 0; i < 3; i++)
    if (cudaHostRegister(ptr[i], sz[i], cudaHostRegisterDefault) == cudaSuccess) sc->pinned[i] = ptr[i]; else cudaGetLastError();
  sc->worker = std::thread([sc, max_batch, linger_us]() { sc->run_rc = fbsc_server_run(sc->srv, sidecar_batch, sc->e, max_batch, linger_us); sc->running = 0; });
  *out = sc;
  return FB_OK;
}

int fb_sidecar_running(fb_sidecar* sc) { return sc ? sc->running.load() : 0; }

int fb_sidecar_stop(fb_sidecar* sc, int64_t* counters3) {
  if (!sc) return FB_ERR_INVALID;
  fbsc_server_stop(sc->srv);
  if (sc->worker.joinable()) sc->worker.join();
  if (counters3) fbsc_server_counters(sc->srv, counters3 + 0, counters3 + 1, counters3 + 2);
  cudaSetDevice(sc->e->device);
  for (int i = 0; i < 3; i++)
    if (sc->pinned[i]) cudaHostUnregister(sc->pinned[i]);
  fbsc_server_destroy(sc->srv);
  const int rc = sc->run_rc;
  delete sc;
  return rc;
}

// order-sensitive checksums of a pinned table's layout: out[0] over (slot, row), out[1] over the packed codes
int fb_table_checksum(fb_engine* e, int table, uint64_t* out) {
  if (!e || !out || table < 0 || table > 2) return fail(e, FB_ERR_INVALID, "fb_table_checksum: bad arguments");
  CodeTable& tab = table == 0 ? e->fine : table == 1 ? e->pq : e->ivpq;
  if (!tab.loaded) return fail(e, FB_ERR_INVALID, "fb_table_checksum: table not loaded");
  FB_CUDA(e, cudaSetDevice(e->device));
  DevBuf<unsigned long long> d_out;
  FB_CUDA(e, d_out.ensure(2));
  FB_CUDA(e, cudaMemsetAsync(d_out.p, 0, 2 * sizeof(unsigned long long), e->stream));
  const int64_t n_slots = tab.n_blocks * 32;
  if (n_slots > 0)
    table_checksum_kernel<<<(unsigned)std::min<int64_t>((n_slots + 255) / 256, 4096), 256, 0, e->stream>>>(
        tab.rowno.p, tab.units.p, tab.has8 ? tab.units8.p : nullptr, n_slots, tab.U, d_out.p);
  unsigned long long h[2];
  FB_CUDA(e, cudaMemcpyAsync(h, d_out.p, sizeof h, cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  out[0] = h[0]; out[1] = h[1];
  return FB_OK;
}

// ivpq_search_in (ivpq_search_in.c:61-721).  Per call: the target set becomes a compact CELL-MAJOR table on the
// device (`id IN (targets)` -> rows -> stable counting sort by multi-index cell); per round of the reference's
// retry loop: cell selection per active query, then a scan that touches only the rows of the selected cells.
// The active list is compacted on the device; the host reads back two integers per round.
int fb_ivpq_search_in(fb_engine* e, const float* queries, int nq, int k, const int32_t* targets, int n_targets, int alpha,
                      int pvf, int method, int use_target_lists, float confidence, int double_threshold,
                      int32_t* out_ids, float* out_dists) {
  if (!e) return FB_ERR_INVALID;
  int rc = check_common(e, nq, k);
  if (rc) return rc;
  if (!e->ivpq_loaded) return fail(e, FB_ERR_INVALID, "IVPQ index not loaded (fb_load_ivpq)");
  if (method < 0 || method > 2) return fail(e, FB_ERR_INVALID, "Unknown computation method!");   // ivpq_search_in.c:376
  if (n_targets < 0 || (n_targets > 0 && !targets)) return fail(e, FB_ERR_INVALID, "bad target array");
  if (alpha < 1) return fail(e, FB_ERR_INVALID, "alpha=%d", alpha);
  if (method != 0 && !e->vec_loaded) return fail(e, FB_ERR_INVALID, "methods 1/2 join the word-vector table: fb_load_vectors first");
  if (pvf < 1) pvf = 1;                                                                           // :206-208
  const Codebook& cb = e->cb[FB_CB_IVPQ];
  const int d = e->ivpq_d, m = cb.m, K = cb.K, Kc = e->ivpq_Kc, cells = Kc * Kc;
  if (cb.m != e->ivpq.m || cb.m * cb.sub != d) return fail(e, FB_ERR_INVALID, "ivpq codebook / table shape mismatch");
  if (method != 0 && e->vec_d != d) return fail(e, FB_ERR_INVALID, "word vectors have d=%d, index d=%d", e->vec_d, d);
  // alpha*k > double_threshold selects the pair-LUT variant (ivpq_search_in.c:261-275); decided once from the
  // alpha of the call, like the reference.  Its pair codes live in an int16 array (:417, :447-451).
  const bool pair_sums = method != 1 && (int64_t)alpha * k > double_threshold;
  if (pair_sums && (int64_t)K * K > 32768)
    return fail(e, FB_ERR_REFERENCE_UB, "pair-LUT variant with K=%d: the reference's int16 pair codes overflow (ivpq_search_in.c:447-451)", K);
  if ((int64_t)k * pvf > kJoinMaxP) return fail(e, FB_ERR_UNSUPPORTED, "k*pvf=%lld > %d", (long long)k * pvf, kJoinMaxP);
  if (nq == 0) return FB_OK;
  if (!queries || !out_ids || !out_dists) return fail(e, FB_ERR_INVALID, "null buffer");
  FB_CUDA(e, cudaSetDevice(e->device));

  // ---- the target set: rows selected by `fq.id IN (targets)` in table order (ivpq_search_in.c:352-401) ----
  FB_CUDA(e, e->q_stage.ensure((size_t)nq * d));
  FB_CUDA(e, cudaMemcpyAsync(e->q_stage.p, queries, (size_t)nq * d * sizeof(float), cudaMemcpyHostToDevice, e->stream));
  CodeTableDev sub;
  int64_t rows_upper = 0;
  if ((rc = build_subset(e, e->ivpq, targets, n_targets, sub, rows_upper))) return rc;   // only sel_rows / sel_total are used
  const int nt_up = (int)std::max<int64_t>(1, rows_upper);
  const int n_chunks = (nt_up + kJoinChunk - 1) / kJoinChunk;
  const int U = e->ivpq.U;
  const int n_slots = (nt_up + 31) / 32 * 32;
  FB_CUDA(e, e->j_cell.ensure((size_t)nt_up));
  FB_CUDA(e, e->j_vrow.ensure((size_t)nt_up));
  FB_CUDA(e, e->j_id.ensure((size_t)nt_up));
  FB_CUDA(e, e->j_perm.ensure((size_t)nt_up));
  FB_CUDA(e, e->j_counts.ensure((size_t)n_chunks * kJoinCells));
  FB_CUDA(e, e->j_cell_start.ensure((size_t)kJoinCells + 1));
  FB_CUDA(e, e->jtmp.units.ensure((size_t)n_slots * U));
  FB_CUDA(e, e->j_state.ensure(4));
  join_rows_kernel<<<(nt_up + 255) / 256, 256, 0, e->stream>>>(e->sel_rows.p, e->sel_total.p, e->ivpq_cells.p, e->ivpq.ids.p,
                                                               e->vec_sorted_ids.p, e->vec_sorted_rows.p, (int)e->vec_N,
                                                               method != 0 ? 1 : 0, e->j_cell.p, e->j_id.p, e->j_vrow.p);
  join_cell_count_kernel<<<n_chunks, 32, 0, e->stream>>>(e->j_cell.p, e->sel_total.p, e->j_counts.p);
  join_cell_scan_kernel<<<1, kJoinCells, 0, e->stream>>>(e->j_counts.p, n_chunks, e->j_cell_start.p);
  join_cell_scatter_kernel<<<n_chunks, 32, 0, e->stream>>>(e->j_cell.p, e->sel_total.p, e->j_counts.p, e->sel_rows.p, e->ivpq.units.p, U,
                                                           e->j_perm.p, e->jtmp.units.p);
  e->launches += 4;
  FB_CUDA(e, cudaGetLastError());
  CodeTableDev ctab;
  ctab.units8 = nullptr;
  ctab.units = e->jtmp.units.p; ctab.rowno = nullptr; ctab.list_blk = nullptr; ctab.list_len = nullptr;
  ctab.ids = nullptr; ctab.m = m; ctab.U = U; ctab.n_lists = 1;

  FB_CUDA(e, e->id_stage.ensure((size_t)nq * k));
  FB_CUDA(e, e->dist_stage.ensure((size_t)nq * k));
  FB_CUDA(e, e->j_active.ensure((size_t)nq));
  FB_CUDA(e, e->j_active2.ensure((size_t)nq));
  FB_CUDA(e, e->j_ncells.ensure((size_t)nq));
  FB_CUDA(e, e->j_filled.ensure((size_t)nq));
  FB_CUDA(e, e->j_tcounts.ensure((size_t)nq));
  FB_CUDA(e, e->j_sel_cells.ensure((size_t)nq * 1024));
  FB_CUDA(e, cudaMemsetAsync(e->j_tcounts.p, 0, (size_t)nq * sizeof(int32_t), e->stream));
  iota_kernel<<<(nq + 255) / 256, 256, 0, e->stream>>>(e->j_active.p, nq);
  e->launches++;
  if (method != 1) {   // LUT per query on the raw query (ivpq_search_in.c:279-290)
    FB_CUDA(e, e->lut.ensure((size_t)nq * m * K));
    if ((rc = launch_lut(e, cb, e->q_stage.p, nullptr, nullptr, 1, nq, e->lut.p))) return rc;
  }
  const size_t smem = sizeof(u64) * 2 * kJoinSortN + (sizeof(float) * 2 + sizeof(int32_t)) * kJoinMaxP +
                      (method != 1 ? (size_t)m * K * sizeof(float) : 0);
  if (smem > e->smem_optin - 4096) return fail(e, FB_ERR_UNSUPPORTED, "LUT too large for the join kernel");
  FB_CUDA(e, cudaFuncSetAttribute(ivpq_scan_cells_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // persistent scan CTAs; each owns a slice of global key scratch for queries with more than kJoinSortN candidate rows
  int scan_ctas = std::max(1, std::min(nq, 2 * e->num_sms));
  while (scan_ctas > 1 && (size_t)scan_ctas * nt_up * sizeof(u64) > ((size_t)256 << 20)) scan_ctas /= 2;
  FB_CUDA(e, e->j_keys.ensure((size_t)scan_ctas * nt_up));

  int n_active = nq;
  int64_t cur_alpha = alpha;
  int32_t* act = e->j_active.p;
  int32_t* act_next = e->j_active2.p;
  JoinParams prm;
  prm.d = d; prm.m = m; prm.K = K; prm.Kc = Kc; prm.k = k; prm.pvf = pvf; prm.method = method;
  prm.n_targets_sql = n_targets; prm.confidence = confidence; prm.stat_total = (int)e->ivpq_stats_host[cells];
  prm.skip_below = use_target_lists ? k * alpha : 0;
  prm.pair_sums = pair_sums ? 1 : 0;
  prm.last_iteration = 0;
  while (n_active > 0) {                                                                        // :299
    prm.min_target = (int)std::min<int64_t>((int64_t)k * cur_alpha, 0x7fffffff);
    const int32_t st0[4] = {1, 0, 0, 0};      // [0] all active queries exhausted every cell (AND), [1] next active count, [2] work counter
    FB_CUDA(e, cudaMemcpyAsync(e->j_state.p, st0, sizeof st0, cudaMemcpyHostToDevice, e->stream));
    {
      StageTimer t(e, ST_COARSE);
      ivpq_select_kernel<<<n_active, 1024, 0, e->stream>>>(e->q_stage.p, act, d, Kc, e->coarse_multi.p, e->ivpq_stats.p, prm,
                                                           e->j_sel_cells.p, e->j_ncells.p, e->j_state.p);
      e->launches++;
      FB_CUDA(e, cudaGetLastError());
    }
    {
      StageTimer t(e, ST_SCAN);
      ivpq_scan_cells_kernel<<<std::min(scan_ctas, n_active), kJoinThreads, smem, e->stream>>>(
          e->q_stage.p, act, n_active, prm, ctab, e->j_cell_start.p, e->j_perm.p, e->j_vrow.p, e->j_id.p, e->vecR.p, e->lut.p,
          e->j_sel_cells.p, e->j_ncells.p, e->j_state.p, e->j_tcounts.p, e->j_keys.p, (size_t)nt_up, e->id_stage.p, e->dist_stage.p,
          e->j_filled.p, e->j_state.p + 2, e->counters64.p + 7);
      e->launches++;
      e->n_scan_launches++;
      FB_CUDA(e, cudaGetLastError());
    }
    join_next_active_kernel<<<(n_active + 255) / 256, 256, 0, e->stream>>>(act, e->j_filled.p, n_active, act_next, e->j_state.p);
    e->launches++;
    int32_t st[2] = {0, 0};
    FB_CUDA(e, cudaMemcpyAsync(st, e->j_state.p, sizeof st, cudaMemcpyDeviceToHost, e->stream));
    FB_CUDA(e, cudaStreamSynchronize(e->stream));
    e->join_rounds++;
    n_active = st[0] ? 0 : st[1];                                                                // :639-666 (lastIteration ends the loop)
    std::swap(act, act_next);
    cur_alpha += cur_alpha;                                                                     // :680
  }
  FB_CUDA(e, cudaMemcpyAsync(out_ids, e->id_stage.p, (size_t)nq * k * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaMemcpyAsync(out_dists, e->dist_stage.p, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  e->queries_done += nq;
  e->bytes_per_row = 2 * m + 4;
  return FB_OK;
}

}  // extern "C"

extern "C" {

__global__ void gather_vec_rows_kernel(const float* __restrict__ vR, int d, const int32_t* __restrict__ rows, int n,
                                       float* __restrict__ out) {
  const int q = blockIdx.x;
  if (q >= n) return;
  const int r = rows[q];
  for (int i = threadIdx.x; i < d; i += blockDim.x) out[(size_t)q * d + i] = vR[(size_t)r * d + i];
}

int fb_ivfadc_batch_search(fb_engine* e, const int32_t* query_ids, int n_ids, int k, int32_t* out_query_ids,
                           int32_t* out_ids, float* out_dists, int* n_queries_out) {
  if (!e || n_ids < 0 || !n_queries_out) return fail(e, FB_ERR_INVALID, "fb_ivfadc_batch_search: bad arguments");
  int rc = check_common(e, n_ids, k);
  if (rc) return rc;
  if (!e->vec_loaded) return fail(e, FB_ERR_INVALID, "ivfadc_batch_search reads the query vectors from the word-vector table: fb_load_vectors first");
  *n_queries_out = 0;
  if (n_ids == 0) return FB_OK;
  if (!query_ids || !out_query_ids || !out_ids || !out_dists) return fail(e, FB_ERR_INVALID, "null buffer");
  FB_CUDA(e, cudaSetDevice(e->device));
  // `SELECT id, vector FROM <normalized> WHERE id IN (...)` (freddy.c:767-804): table order, each id once
  std::vector<int32_t> rows;
  for (int i = 0; i < n_ids; i++) { int r = vec_row_of(e, query_ids[i]); if (r >= 0) rows.push_back(r); }
  std::sort(rows.begin(), rows.end());
  rows.erase(std::unique(rows.begin(), rows.end()), rows.end());
  const int nq = (int)rows.size();
  *n_queries_out = nq;
  if (nq == 0) return FB_OK;
  if (e->vec_d != e->d) return fail(e, FB_ERR_INVALID, "word vectors have d=%d, IVFADC index d=%d", e->vec_d, e->d);
  const int d = e->vec_d;
  FB_CUDA(e, e->ana_rows.ensure((size_t)nq));
  FB_CUDA(e, e->q_stage.ensure((size_t)nq * d));
  FB_CUDA(e, e->id_stage.ensure((size_t)nq * k));
  FB_CUDA(e, e->dist_stage.ensure((size_t)nq * k));
  FB_CUDA(e, cudaMemcpyAsync(e->ana_rows.p, rows.data(), (size_t)nq * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
  gather_vec_rows_kernel<<<nq, 128, 0, e->stream>>>(e->vecR.p, d, e->ana_rows.p, nq, e->q_stage.p);
  e->launches++;
  FB_CUDA(e, cudaGetLastError());
  FB_CUDA(e, cudaStreamSynchronize(e->stream));   // rows is a local
  // one list per round, argmin with the first minimum winning (freddy.c:846-866) == the w = 1 case of
  // ivfadc_search's selection; rounds continue until k rows were seen; sentinel 100.0 (:823-827)
  if ((rc = ivfadc_dev(e, e->q_stage.p, nq, k, 1, e->id_stage.p, e->dist_stage.p, 100.0f, true))) return rc;
  FB_CUDA(e, cudaMemcpyAsync(out_ids, e->id_stage.p, (size_t)nq * k * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaMemcpyAsync(out_dists, e->dist_stage.p, (size_t)nq * k * sizeof(float), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  for (int i = 0; i < nq; i++) out_query_ids[i] = e->vec_ids_host[rows[i]];
  return check_error_flag(e);
}

}  // extern "C"

extern "C" {

// grouping_pq(int[] ids, int[] group_ids) (freddy.c:1178-1401)
int fb_grouping_pq(fb_engine* e, const int32_t* ids, int n_ids, const int32_t* group_ids, int n_groups,
                   int32_t* out_ids, int32_t* out_group_ids, int* n_out) {
  if (!e || n_ids < 0 || n_groups < 0 || !n_out) return fail(e, FB_ERR_INVALID, "fb_grouping_pq: bad arguments");
  int rc = check_pq_ready(e);
  if (rc) return rc;
  const int d = pq_dim(e);
  if (!e->vec_loaded) return fail(e, FB_ERR_INVALID, "grouping_pq reads the group vectors from the word-vector table: fb_load_vectors first");
  if (e->vec_d != d) return fail(e, FB_ERR_INVALID, "word vectors have d=%d, pq index d=%d", e->vec_d, d);
  *n_out = 0;
  if ((n_ids > 0 && (!ids || !out_ids || !out_group_ids)) || (n_groups > 0 && !group_ids)) return fail(e, FB_ERR_INVALID, "null buffer");
  FB_CUDA(e, cudaSetDevice(e->device));
  // group vectors: `WHERE id IN (group ids) ORDER BY id ASC`; every id must select its own row (freddy.c:1228-1246)
  std::vector<int32_t> groups(group_ids, group_ids + n_groups);
  std::sort(groups.begin(), groups.end());
  std::vector<int32_t> grows(n_groups);
  for (int g = 0; g < n_groups; g++) {
    grows[g] = vec_row_of(e, groups[g]);
    if (grows[g] < 0 || (g > 0 && groups[g] == groups[g - 1])) return fail(e, FB_ERR_INVALID, "Group ids do not exist");
  }
  if (n_ids == 0) return FB_OK;
  CodeTableDev view;
  int64_t rows_upper = 0;
  if ((rc = build_subset(e, e->pq, ids, n_ids, view, rows_upper))) return rc;
  int32_t n = 0;
  FB_CUDA(e, cudaMemcpyAsync(&n, e->sel_total.p, sizeof n, cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  *n_out = n;
  if (n == 0) return FB_OK;
  std::vector<int32_t> rows((size_t)n);
  FB_CUDA(e, cudaMemcpy(rows.data(), e->sel_rows.p, (size_t)n * sizeof(int32_t), cudaMemcpyDeviceToHost));
  for (int i = 0; i < n; i++) out_ids[i] = e->pq_ids_host[rows[i]];
  if (n_groups == 0) return fail(e, FB_ERR_REFERENCE_UB, "no groups: the reference reads an uninitialised assignment (freddy.c:1326-1352)");
  const Codebook& cb = e->cb[FB_CB_PQ];
  const int m = cb.m, K = cb.K;
  const size_t lut_bytes = (size_t)m * K * sizeof(float);
  if (2 * lut_bytes > e->smem_optin - 1024) return fail(e, FB_ERR_UNSUPPORTED, "LUT too large for shared memory");
  const int n_blocks = (n + 31) / 32, n_slots = n_blocks * 32;
  FB_CUDA(e, e->ana_rows.ensure((size_t)n_groups));
  FB_CUDA(e, e->q_stage.ensure((size_t)n_groups * d));
  FB_CUDA(e, e->lut.ensure((size_t)n_groups * m * K));
  FB_CUDA(e, e->id_stage.ensure((size_t)n_slots));
  FB_CUDA(e, cudaMemcpyAsync(e->ana_rows.p, grows.data(), (size_t)n_groups * sizeof(int32_t), cudaMemcpyHostToDevice, e->stream));
  gather_vec_rows_kernel<<<n_groups, 128, 0, e->stream>>>(e->vecR.p, d, e->ana_rows.p, n_groups, e->q_stage.p);
  e->launches++;
  FB_CUDA(e, cudaGetLastError());
  if ((rc = launch_lut(e, cb, e->q_stage.p, nullptr, nullptr, 1, n_groups, e->lut.p))) return rc;   // freddy.c:1291-1299
  {
    StageTimer t(e, ST_SCAN);
    const int ctas = (n_blocks + kGroupThreads / 32 - 1) / (kGroupThreads / 32);
    auto kern = (m == 12) ? grouping_argmin_kernel<12> : grouping_argmin_kernel<0>;
    FB_CUDA(e, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(2 * lut_bytes)));
    kern<<<ctas, kGroupThreads, 2 * lut_bytes, e->stream>>>(view, n_blocks, n, e->lut.p, n_groups, K, e->id_stage.p, e->small.p + 2);
    e->launches++;
    FB_CUDA(e, cudaGetLastError());
  }
  std::vector<int32_t> nearest((size_t)n_slots);
  FB_CUDA(e, cudaMemcpyAsync(nearest.data(), e->id_stage.p, (size_t)n_slots * sizeof(int32_t), cudaMemcpyDeviceToHost, e->stream));
  FB_CUDA(e, cudaStreamSynchronize(e->stream));
  int32_t flag = 0;
  FB_CUDA(e, cudaMemcpy(&flag, e->small.p + 2, sizeof flag, cudaMemcpyDeviceToHost));
  if (flag) {
    cudaMemset(e->small.p + 2, 0, sizeof(int32_t));
    return fail(e, FB_ERR_REFERENCE_UB, "a row is >= 100 away from every group: the reference's assignment is uninitialised (freddy.c:1326-1352)");
  }
  for (int i = 0; i < n; i++) {            // the subset is one compact list: slot i = i-th selected row
    if (nearest[i] < 0) return fail(e, FB_ERR_CUDA, "grouping_pq: internal slot accounting (slot %d of %d)", i, n);
    out_group_ids[i] = groups[nearest[i]];
  }
  e->host_rows += (int64_t)n * n_groups;
  e->bytes_per_row = 2 * m + 4;
  return FB_OK;
}

}  // extern "C"
