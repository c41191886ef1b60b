#!/usr/bin/env python
"""bench.py — IVFADC queries/s at k=5, nprobe(w)=10 on a synthetic 3M x 300, m=12,
K=1024, C=1000 index (BASELINE.json metric; config[1] in its throughput form).

  python bench.py --gpus N --steps K --warmup W            our arm (CUDA engine, C-ABI)
  python bench.py --impl reference ...                      CPU arm (the oracle on host threads)

A "step" = one pass of the hot path over one batch of --batch queries per GPU:
coarse quantizer -> residual LUTs -> ADC scan of the w probed lists -> top-k.
  value : whole-job queries/s, query batch already resident in HBM
  e2e   : same through fb_ivfadc_search with pinned HOST buffers (H2D queries +
          D2H results inside the timed region)
  roofline : ADC-scan kernel, algorithmic bytes (rows * (2m+4)) / CUDA-event time
          of the scan launches (events recorded by the engine on the launch stream)
Multi-GPU: index replicated, queries sharded (weak scaling: --batch per GPU),
one NCCL all-gather of the per-rank top-k per step.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "postgres-word2vec_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "ivfadc_queries_per_sec_k5_w10"
UNIT = "queries/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n", type=int, default=3_000_000)
    ap.add_argument("--d", type=int, default=300)
    ap.add_argument("--m", type=int, default=12)
    ap.add_argument("--K", type=int, default=1024)
    ap.add_argument("--C", type=int, default=1000)
    ap.add_argument("--k", type=int, default=5)
    ap.add_argument("--w", type=int, default=10)
    ap.add_argument("--batch", type=int, default=32768,
                    help="queries per GPU per step (16 pipeline chunks of 2048: the two un-overlapped launches of a call, "
                         "LUT-only first and scan-only last, are 2/17 of the step; 10000 gives 1.86 M q/s, 32768 1.98 M)")
    ap.add_argument("--sigma", type=float, default=1.0,
                    help="within-cluster noise of the synthetic vectors (per dimension, centres ~ N(0,I)); "
                         "1.0 keeps PQ codes diverse like real word embeddings, 0.3 collapses clusters onto "
                         "identical codes (25%% duplicate rows) and sends a quarter of the queries down the tie path")
    ap.add_argument("--zipf", type=float, default=0.35,
                    help="cluster-size skew of the synthetic vectors: size ~ 1/(10+rank)^zipf")
    ap.add_argument("--cpu-seconds", type=float, default=12.0, help="CPU-baseline sample budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--lut-tile", type=int, default=None, help="codes per LUT-build CTA (256/512/1024)")
    ap.add_argument("--overlap", type=int, default=None, help="1: overlap LUT build and scan of consecutive chunks")
    ap.add_argument("--lut-ctas", type=int, default=None, help="cap on LUT-build CTAs per SM")
    ap.add_argument("--chunk", type=int, default=None, help="queries per pipeline chunk")
    ap.add_argument("--scalar-lut", action="store_true", help="LUT build with scalar instead of packed f32x2 ops")
    ap.add_argument("--pipeline", type=int, default=None, help="0: separate LUT/scan kernels instead of the pipeline kernel")
    ap.add_argument("--pipe-chunk", type=int, default=None, help="queries per pipeline beat")
    ap.add_argument("--placement", type=int, default=None, help="window of the conflict-aware row placement (0 = arrival order)")
    ap.add_argument("--pipe-shape", type=int, default=None, help="role split of the pipeline CTA (FB_OPT_PIPE_SHAPE)")
    ap.add_argument("--pipe-ramp", type=int, default=None, help="0: uniform pipeline chunks")
    ap.add_argument("--pipe-debug", type=int, default=None, help="timing aid (invalid results): 1 producers only, 2 scan only")
    ap.add_argument("--secondary", default=None,
                    help="comma list of secondary workloads reported in config.secondary: 3,4,5 (BASELINE configs), sigma03, nominal; "
                         "'none' skips them.  Default: all five on one GPU, 4 and 5 (the sharded ones) on several")
    ap.add_argument("--gpu-rotate", type=int, default=int(os.environ.get("FB_GPU_ROTATE", "0")),
                    help="rank r runs on GPU (r + ROTATE) %% n_gpus: separates a slow GPU from a slow rank (VERDICT r1 item 5)")
    ap.add_argument("--secondary-sample", type=int, default=256, help="queries of each secondary workload checked against the reference")
    return ap.parse_args()


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def workload_name(a):
    return (f"k_nearest_neighbour_ivfadc throughput form: {a.batch} queries/GPU/step, k={a.k}, w={a.w}, "
            f"synthetic N={a.n} d={a.d} m={a.m} K={a.K} C={a.C} (1000 Zipf-sized Gaussian clusters, sigma={a.sigma}, zipf={a.zipf}, L2-normalised)")


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region"""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows = []
        self.gpu = gpu_index
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "20", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
            import atexit
            atexit.register(lambda: self.proc and self.proc.poll() is None and self.proc.kill())
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [x.strip() for x in line.split(",")]))

    def mark(self):
        """samples before this moment are dropped (nvidia-smi is started early: with 8 instances on an 8-GPU box its
        start-up alone outlasts a short timed region)"""
        self.t_from = time.time()

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        t_from = getattr(self, "t_from", 0.0)
        rows = [r for (t, r) in self.rows if t >= t_from]
        sm = [float(r[1]) for r in rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            if len(r) >= 9:
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def build_index(a, device):
    """coarse / codebook training in torch (index construction is outside the parity boundary); the assignment
    and encoding of all N rows by the engine (fb_encode_ivfadc_dev: the reference's rule, freddy.c:1567-1582)"""
    from freddy_b200.index_build import make_synthetic_index
    t0 = time.time()
    enc = None
    if str(device).startswith("cuda") and a.impl == "ours":      # the reference arm stays free of this repo's library
        from freddy_b200 import Engine
        import torch
        enc = Engine(torch.device(device).index or 0)
    ix = make_synthetic_index(a.n, d=a.d, m=a.m, K=a.K, C=a.C, n_train=min(100_000, a.n), n_clusters=1000,
                              sigma=a.sigma, zipf=a.zipf, kmeans_iters=10, seed=1234, device=device, keep_vectors=True, encoder=enc)
    if enc is not None:
        enc.close()
    return ix, time.time() - t0


def cpu_arm(a, ix, queries, seconds):
    """times the oracle (plain-C restatement of freddy.c:247-378) on all host threads"""
    return cpu_arm_threads(a, ix, queries, seconds, os.cpu_count() or 1)


def cpu_arm_threads(a, ix, queries, seconds, threads):
    from oracle import oracle
    oi = oracle.OracleIndex(ix)
    # calibrate on a small sample, then size one run to ~`seconds`
    n0 = min(len(queries), 4 * threads)
    t = time.time()
    oi.ivfadc_search(queries[:n0], a.k, a.w, threads=threads)
    per_q = max((time.time() - t) / n0, 1e-6)
    n = int(max(n0, min(len(queries), seconds / per_q)))
    t = time.time()
    _, _, rc, rows = oi.ivfadc_search(queries[:n], a.k, a.w, threads=threads)
    dt = time.time() - t
    return {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": f"{n} of the step's queries, one pass, {threads} host threads (oracle/freddy_oracle.c, SPI-free upper bound)",
            "seconds": dt, "rows_per_query": rows / max(n, 1), "rc": rc}


# ---- the reference's own SRF (oracle/_ref: freddy.c compiled unmodified, SPI/fmgr emulated) on all host cores:
# one single-threaded "backend" process per core, each with the index tables registered (as one Postgres
# connection per core would), each answering a slice of the sample.
_REF = {}


def _ref_worker_init(shm_dir, w):
    from oracle import oracle
    ix = {k: np.load(os.path.join(shm_dir, k + ".npy"), mmap_mode="r") for k in
          ("coarse", "residual_codebook", "ids", "coarse_ids", "codes")}
    meta = json.load(open(os.path.join(shm_dir, "meta.json")))
    ix.update(meta)
    rs = oracle.ReferenceSession()
    rs.load_ivfadc(ix, w)
    _REF["rs"] = rs


def _ref_worker_run(args):
    q, k = args
    ids, raw, _ = _REF["rs"].ivfadc_search(q, k)
    return ids, raw


class ReferencePool:
    def __init__(self, a, ix):
        import multiprocessing as mp
        import shutil
        import tempfile
        need = sum(np.asarray(ix[k]).nbytes for k in ("coarse", "residual_codebook", "ids", "coarse_ids", "codes"))
        shm = "/dev/shm"
        use_shm = os.path.isdir(shm) and shutil.disk_usage(shm).free > 2 * need + (64 << 20)
        self.dir = tempfile.mkdtemp(prefix="fb_ref_", dir=shm if use_shm else None)
        self._rm = shutil.rmtree
        for k in ("coarse", "residual_codebook", "ids", "coarse_ids", "codes"):
            np.save(os.path.join(self.dir, k + ".npy"), np.ascontiguousarray(ix[k]))
        json.dump({k: int(ix[k]) for k in ("d", "m", "K", "C", "N")}, open(os.path.join(self.dir, "meta.json"), "w"))
        self.procs = os.cpu_count() or 1
        self.k = a.k
        self.pool = mp.get_context("spawn").Pool(self.procs, initializer=_ref_worker_init, initargs=(self.dir, a.w))
        self.pool.map(_ref_worker_run, [(np.zeros((1, a.d), np.float32) + 0.01, a.k)] * self.procs)   # every backend is up

    def run(self, queries):
        parts = [p for p in np.array_split(np.ascontiguousarray(queries, np.float32), self.procs) if len(p)]
        t = time.time()
        res = self.pool.map(_ref_worker_run, [(p, self.k) for p in parts], chunksize=1)
        dt = time.time() - t
        return np.concatenate([r[0] for r in res]), np.concatenate([r[1] for r in res]), dt

    def close(self):
        self.pool.close()
        self.pool.join()
        self._rm(self.dir, ignore_errors=True)


def reference_arm(a, ix, queries, seconds, pool):
    """times the reference's own ivfadc_search SRF (oracle/_ref) on all host cores"""
    n0 = min(len(queries), 2 * pool.procs)
    _, _, dt0 = pool.run(queries[:n0])
    per_q = max(dt0 / n0, 1e-6)
    n = int(max(n0, min(len(queries), seconds / per_q)))
    ids, raw, dt = pool.run(queries[:n])
    return {"value": n / dt, "unit": UNIT, "cores": pool.procs, "kind": "reference",
            "sample": f"{n} of the step's queries, one pass, {pool.procs} single-threaded backends "
                      "(oracle/_ref: the reference's freddy.c SRF, SPI emulated in memory)",
            "seconds": dt, "ids": ids, "raw": raw}


def main():
    a = parse()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    import torch

    # ------------------------------------------------------------------ CPU arm
    if a.impl == "reference":
        if rank != 0:
            return
        dev = "cuda" if torch.cuda.is_available() else "cpu"
        ix, _ = build_index(a, dev)
        vec = ix.pop("vectors_t")
        g = torch.Generator(); g.manual_seed(4321)
        sel = torch.randperm(a.n, generator=g)[:a.batch]
        queries = vec[sel.to(vec.device)].cpu().numpy()
        del vec
        from oracle import oracle
        use_ref = os.path.exists(oracle.REF_SO)
        pool = None
        if use_ref:
            try:
                pool = ReferencePool(a, ix)
            except Exception as ex:
                print(f"reference pool unavailable ({ex}); using the oracle port", file=sys.stderr)
                use_ref = False
        vals = []
        budget = max(1.0, a.cpu_seconds / max(1, a.steps))
        for i in range(a.warmup + a.steps):
            r = reference_arm(a, ix, queries, budget, pool) if use_ref else cpu_arm(a, ix, queries, budget)
            if i >= a.warmup:
                vals.append(r)
        port = cpu_arm(a, ix, queries, 3.0)
        parity = None
        if use_ref:      # the port answers the same sample: the two CPU implementations must agree bit for bit
            n_chk = min(len(vals[-1]["ids"]), 64)
            eids, ed, _, _ = oracle.OracleIndex(ix).ivfadc_search(queries[:n_chk], a.k, a.w, threads=os.cpu_count() or 1)
            parity = bool((eids == vals[-1]["ids"][:n_chk]).all() and
                          (ed.view(np.uint32) == vals[-1]["raw"][:n_chk].view(np.uint32)).all())
            pool.close()
        v = float(np.mean([r["value"] for r in vals]))
        n_s = int(np.mean([r["value"] * r["seconds"] for r in vals]))
        kind = vals[0]["kind"]
        what = ("single-threaded backends running the reference's own freddy.c SRF (oracle/_ref)" if use_ref
                else "host threads (oracle port)")
        out = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
               "warmup": a.warmup, "ms_per_step": 1e3 * float(np.mean([r["seconds"] for r in vals])),
               "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
               "config": {"workload": workload_name(a), "note": "each step = a bounded sample of the step's query batch",
                          "oracle_port_all_threads_queries_per_s": port["value"], "port_equals_reference_on_sample": parity},
               "cpu_baseline": {"value": v, "unit": UNIT, "cores": vals[0]["cores"], "kind": kind,
                                "sample": f"~{n_s} queries per step on {vals[0]['cores']} {what}"},
               "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(out))
        return

    # ------------------------------------------------------------------ our arm
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    if a.gpu_rotate:
        local_rank = (local_rank + a.gpu_rotate) % max(1, torch.cuda.device_count())
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    from freddy_b200 import Engine, _lib

    clocks = ClockSampler(local_rank)   # started now, windowed by clocks.mark() below
    clocks.start()
    # index: built once on rank 0, replicated to every rank (broadcast over NCCL)
    names = ["coarse", "residual_codebook", "coarse_ids", "codes"]
    if rank == 0:
        ix, t_build = build_index(a, dev)
        vec = ix.pop("vectors_t")
        g = torch.Generator(); g.manual_seed(4321)
        sel = torch.randperm(a.n, generator=g)[:a.batch * world]
        all_q = vec[sel.to(dev)].contiguous()
        vec_t = vec                              # kept for the secondary workloads (3.6 GB of the 180)
        del vec
    else:
        ix, t_build, all_q, vec_t = {"d": a.d, "m": a.m, "K": a.K, "C": a.C, "N": a.n}, 0.0, None, None
    if world > 1:
        shapes = {"coarse": ((a.C, a.d), torch.float32), "residual_codebook": ((a.m, a.K, a.d // a.m), torch.float32),
                  "coarse_ids": ((a.n,), torch.int32), "codes": ((a.n, a.m), torch.int16)}
        for nm in names:
            shp, dt = shapes[nm]
            t = torch.from_numpy(ix[nm]).to(dev) if rank == 0 else torch.empty(shp, dtype=dt, device=dev)
            dist.broadcast(t.view(torch.uint8), 0)      # NCCL has no int16: ship the raw bytes
            ix[nm] = t.cpu().numpy()
        if rank != 0:
            all_q = torch.empty(a.batch * world, a.d, device=dev)
            ix["ids"] = np.arange(1, a.n + 1, dtype=np.int32)
        dist.broadcast(all_q, 0)
    torch.cuda.empty_cache()
    my_q = all_q[rank * a.batch:(rank + 1) * a.batch].contiguous()

    eng = Engine(local_rank)
    if a.placement is not None:
        eng.set_option(_lib.FB_OPT_PLACEMENT_WINDOW, a.placement)
    t_load = time.time()
    eng.load_ivfadc_index(ix)
    t_load = time.time() - t_load
    stream = torch.cuda.Stream(device=dev)          # a real (non-default) stream: events and kernels share it
    torch.cuda.set_stream(stream)
    assert stream.cuda_stream != 0
    eng.set_stream(stream.cuda_stream)
    eng.set_option(_lib.FB_OPT_PROFILE, 1)
    if a.lut_tile is not None:
        eng.set_option(_lib.FB_OPT_LUT_TILE, a.lut_tile)
    if a.overlap is not None:
        eng.set_option(_lib.FB_OPT_OVERLAP, a.overlap)
    if a.lut_ctas is not None:
        eng.set_option(_lib.FB_OPT_LUT_CTAS_PER_SM, a.lut_ctas)
    if a.chunk is not None:
        eng.set_option(_lib.FB_OPT_QUERY_CHUNK, a.chunk)
    if a.scalar_lut:
        eng.set_option(_lib.FB_OPT_PACKED_FP32, 0)
    if a.pipeline is not None:
        eng.set_option(_lib.FB_OPT_PIPELINE, a.pipeline)
    if a.pipe_chunk is not None:
        eng.set_option(_lib.FB_OPT_PIPE_CHUNK, a.pipe_chunk)
    if a.pipe_ramp is not None:
        eng.set_option(_lib.FB_OPT_PIPE_RAMP, a.pipe_ramp)
    if a.pipe_shape is not None:
        eng.set_option(_lib.FB_OPT_PIPE_SHAPE, a.pipe_shape)

    nq, k, w = a.batch, a.k, a.w
    d_ids = torch.empty(nq, k, dtype=torch.int32, device=dev)
    d_dist = torch.empty(nq, k, dtype=torch.float32, device=dev)
    g_ids = torch.empty(world * nq, k, dtype=torch.int32, device=dev) if world > 1 else None
    g_dist = torch.empty(world * nq, k, dtype=torch.float32, device=dev) if world > 1 else None
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2

    def step_dev():
        eng.ivfadc_search_dev(my_q.data_ptr(), nq, k, w, d_ids.data_ptr(), d_dist.data_ptr())
        if world > 1:   # the path's one exchange: all-gather of per-rank top-k
            dist.all_gather_into_tensor(g_ids, d_ids)
            dist.all_gather_into_tensor(g_dist, d_dist)

    h_q = my_q.cpu().pin_memory()
    h_ids = torch.empty(nq, k, dtype=torch.int32).pin_memory()
    h_dist = torch.empty(nq, k, dtype=torch.float32).pin_memory()

    def step_e2e():
        eng.ivfadc_search_ptr(h_q.data_ptr(), nq, k, w, h_ids.data_ptr(), h_dist.data_ptr())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(step_fn, steps, use_events):
        """K steps, L2 flushed between steps; device time by CUDA events on the launch stream"""
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        wall = []
        barrier()
        for i in range(steps):
            flush.fill_(i & 0xFF)
            if use_events:
                evs[i][0].record(stream)
                step_fn()
                evs[i][1].record(stream)
            else:
                torch.cuda.synchronize()
                t = time.perf_counter()
                step_fn()
                wall.append((time.perf_counter() - t) * 1e3)
        barrier()
        ms = sum(s.elapsed_time(e) for s, e in evs) if use_events else sum(wall)
        ms_local = ms
        if world > 1:
            t = torch.tensor([ms], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms, ms_local

    clocks.mark()                       # samples through warm-up + timed region (same load in both)
    for _ in range(max(3, a.warmup)):
        step_dev()
    eng.synchronize()
    if a.pipe_debug is not None:    # after the warm-up: the LUT scratch then holds valid LUTs for the scan-only mode
        eng.set_option(_lib.FB_OPT_PIPE_DEBUG, a.pipe_debug)
    eng.reset_counters()
    ms_total, ms_rank = timed(step_dev, a.steps, True)
    clk = clocks.stop()
    c = eng.counters()
    value = world * nq * a.steps / (ms_total / 1e3)

    # roofline of the dominant kernel (ADC scan): algorithmic bytes / event time of its launches
    peak, peak_src = peaks()
    piped = c["n_pipe_launches"] > 0
    # pipeline mode: ONE kernel does the ADC scan of chunk c and the LUT build of chunk c+1; its launches
    # carry all the scan bytes (the LUT build rides along on the fp32 pipe), n_chunks + 1 launches per step
    ms_dom = c["ms_pipe"] if piped else c["ms_scan"]
    n_dom = c["n_pipe_launches"] if piped else c["n_scan_launches"]
    scan_gbs = (c["scan_bytes"] / 1e9) / (ms_dom / 1e3) if ms_dom > 0 else None
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "scan_traffic.json")))
        traffic = tj["pipe_dram_bytes_per_launch" if piped else "dram_bytes_per_launch"]
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "ivfadc_pipe_kernel<12,1024,25,...>" if piped else "adc_scan_query_kernel<12,1024>",
                "achieved": scan_gbs, "peak": peak, "unit": "GB/s",
                "frac": (scan_gbs / peak) if scan_gbs else None, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": c["scan_bytes"] / max(1, n_dom),
                "ms_per_launch": ms_dom / max(1, n_dom),
                "stage_ms_per_step": {s: c["ms_" + s] / a.steps for s in ("coarse", "lut", "scan", "pipe", "finalize", "exact")}}
    # the same launches also carry the residual-LUT build (w*m*K*sub*3 individually rounded fp32
    # operations per query); its bound is the fp32 pipe (measured: profiles/r1_microbench_fp32x2.txt)
    lut_ops = 3.0 * a.w * a.m * a.K * (a.d // a.m) * nq * a.steps
    ms_lut_host = ms_dom if piped else c["ms_lut"]
    roofline["lut_build_fp32"] = {"bound": "fp32 pipe", "achieved": lut_ops / (ms_lut_host / 1e3) / 1e12 if ms_lut_host > 0 else None,
                                  "peak": 36.8, "unit": "T rounded fp32 ops/s", "peak_source": "profiles/r1_microbench_fp32x2.txt",
                                  "frac": lut_ops / (ms_lut_host / 1e3) / 1e12 / 36.8 if ms_lut_host > 0 else None,
                                  "note": "same launches as the scan when the pipeline kernel runs: the two fractions add"}
    launches = c["kernel_launches"]
    exact_q = c["exact_path_queries"]

    # end to end through the host-buffer C-ABI call
    for _ in range(2):
        step_e2e()
    ms_e2e, ms_e2e_rank = timed(step_e2e, a.steps, False)
    e2e_value = world * nq * a.steps / (ms_e2e / 1e3)

    # single-query latency through the same host-buffer call (the form config[1] names)
    lat_us = None
    if rank == 0:
        eng.set_option(_lib.FB_OPT_PROFILE, 0)      # no per-kernel events: small calls then replay a captured CUDA graph
        one_q, one_i, one_d = h_q[:1].clone().pin_memory(), h_ids[:1].clone().pin_memory(), h_dist[:1].clone().pin_memory()
        for _ in range(20):
            eng.ivfadc_search_ptr(one_q.data_ptr(), 1, k, w, one_i.data_ptr(), one_d.data_ptr())
        ts = []
        for _ in range(200):
            t0 = time.perf_counter()
            eng.ivfadc_search_ptr(one_q.data_ptr(), 1, k, w, one_i.data_ptr(), one_d.data_ptr())
            ts.append(time.perf_counter() - t0)
        lat_us = float(np.median(ts) * 1e6)

    cpu = None
    if rank == 0 and world == 1 and not a.no_cpu_baseline:
        from oracle import oracle
        pool = None
        if os.path.exists(oracle.REF_SO):       # the reference's own SRF, one backend per core
            try:
                pool = ReferencePool(a, ix)
            except Exception as ex:              # e.g. no room for the table copies: fall back to the oracle port
                print(f"reference pool unavailable ({ex}); using the oracle port", file=sys.stderr)
                pool = None
        if pool is not None:
            cpu = reference_arm(a, ix, h_q.numpy(), a.cpu_seconds, pool)
            pool.close()
            n_chk = min(len(cpu["ids"]), 256)    # and our result for the same queries must be its result
            cpu["gpu_equals_reference_on_sample"] = bool(
                (cpu["ids"][:n_chk] == h_ids.numpy()[:n_chk]).all() and
                (cpu["raw"][:n_chk].view(np.uint32) == h_dist.numpy()[:n_chk].view(np.uint32)).all())
            port = cpu_arm(a, ix, h_q.numpy(), 3.0)
            cpu["oracle_port_all_threads_queries_per_s"] = port["value"]
            cpu = {kk: cpu[kk] for kk in ("value", "unit", "cores", "kind", "sample", "gpu_equals_reference_on_sample",
                                          "oracle_port_all_threads_queries_per_s")}
        else:
            cpu = cpu_arm(a, ix, h_q.numpy(), a.cpu_seconds)
            cpu = {kk: cpu[kk] for kk in ("value", "unit", "cores", "kind", "sample")}

    # what each rank saw (the max over ranks hid which rank / stage was slow)
    mine = {"rank": rank, "gpu": local_rank, "ms_per_step_device": ms_rank / a.steps, "ms_per_step_e2e": ms_e2e_rank / a.steps,
            "rows_scanned_per_query": c["rows_scanned"] / max(1, c["queries"]), "sm_mhz": clk.get("sm_mhz"), "clock_reasons": clk.get("reasons"),
            "stage_ms_per_step": {s_: c["ms_" + s_] / a.steps for s_ in ("coarse", "pipe", "lut", "scan", "exact")},
            "exact_path_queries_per_step": exact_q / a.steps,
            "exact_path_reasons_per_step": {r: c["exact_" + r] / a.steps for r in ("coarse_tie", "coarse_far", "few_rows", "scan_tie", "forced")}}
    per_rank = [mine]
    if world > 1:
        per_rank = [None] * world
        dist.all_gather_object(per_rank, mine)

    secondary = []
    which = a.secondary if a.secondary is not None else ("3,4,5,sigma03,nominal" if world == 1 else "4,5")
    if which != "none" and a.pipe_debug is None:
        import bench_secondary
        pk = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        eng.close()                              # the headline engine's scratch (2 x 0.98 GB of LUTs) is not needed any more
        try:
            secondary = bench_secondary.run(a, lambda: Engine(local_rank), dev, rank, world, dist, vec_t,
                                            (peak, float(pk.get("bf16_tflops", 1590.0))), set(which.split(",")), a.secondary_sample)
        except Exception as ex:                  # a secondary workload must never take the headline line down with it
            import traceback
            traceback.print_exc()
            secondary = [{"name": "secondary workloads", "error": repr(ex)}]

    if rank == 0 and cpu is not None and not a.no_cpu_baseline:
        # (i) the SPI-free port on ONE thread next to the all-cores numbers, and the CPU it ran on
        one = cpu_arm_threads(a, ix, h_q.numpy(), 3.0, 1)
        cpu["oracle_port_single_thread_queries_per_s"] = one["value"]
        cpu["cpu_model"] = cpu_model()
        cpu["host_cores"] = os.cpu_count()

    if rank == 0:
        out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": max(3, a.warmup),
               "ms_per_step": ms_total / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
               "dtype": "f32", "data": "synthetic",
               "config": {"workload": workload_name(a), "parallelism": f"replicated index, queries sharded x{world}",
                          "l2": "flushed (256 MiB write) between timed steps", "index_build_s": round(t_build, 1), "index_upload_s": round(t_load, 1),
                          "index_encode": ({"by": "fb_encode_ivfadc_dev (reference rule: freddy.c:1567-1582, index_utils.c:923-939)",
                                            **ix["encode_report"]} if "encode_report" in ix else "float shortcut (torch)"),
                          "rows_scanned_per_query": c["rows_scanned"] / max(1, c["queries"]),
                          "single_query_latency_us": lat_us,
                          "exact_path_queries_per_step": exact_q / a.steps,
                          "exact_path_reasons_per_step": {r: c["exact_" + r] / a.steps for r in
                                                          ("coarse_tie", "coarse_far", "few_rows", "scan_tie", "forced")},
                          "per_rank": per_rank, "secondary": secondary},
               "clocks": clk, "gpu_launches": launches,
               "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": nq * a.d * 4 * world,
                       "d2h_bytes_per_step": nq * k * 8 * world, "ms_per_step": ms_e2e / a.steps},
               "roofline": roofline}
        if cpu is not None:
            out["cpu_baseline"] = cpu
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
