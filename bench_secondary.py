"""Secondary workloads of bench.py: the other BASELINE.json configs and the generator variants VERDICT r1 asked
for, each measured through the host-buffer C-ABI call (H2D / D2H inside the timed region), with its own
roofline block and a parity check of a query sample against the reference's own SRF (oracle/_ref through the
SPI emulator) or, where that SRF does not exist (analogy is plpgsql over cosine_similarity_bytea), against the
oracle port that is pinned to it.

  config 3   pq_search_in_batch  5k queries x 100k targets, m=12, K=1024         freddy.c:414-675
  config 4   knn_join / ivpq_search_in 5k x 100k, k=5, alpha=100, pvf=20, method 2, 32x32 multi-index
             (N > 1: queries sharded, one all-gather of the per-rank top-k)       ivpq_search_in.c:61-721
  config 5   analogy_3cosadd 1k triples over the 3M table (N > 1: VOCABULARY sharded, one all-gather of the
             per-rank arg-max)                                                     freddy--0.0.1.sql:1270-1288
  sigma03    the headline IVFADC shape on SURVEY 8(d)'s own generator (sigma = 0.3, zipf = 0.7)
  nominal    the headline IVFADC shape on an index whose probed lists are near the nominal N*w/C rows
"""
import os
import time

import numpy as np


def _timed(fn, reps, sync, after_warmup=None):
    fn()                                    # warm-up (buffers sized, kernels loaded)
    sync()
    if after_warmup is not None:
        after_warmup()                      # e.g. reset the engine's counters: stage times then cover the timed runs only
    ts = []
    for _ in range(reps):
        sync()
        t = time.perf_counter()
        fn()
        sync()
        ts.append(time.perf_counter() - t)
    return float(np.median(ts)), ts


def _same(a_ids, a_d, b_ids, b_d):
    return bool((np.asarray(a_ids) == np.asarray(b_ids)).all() and
                (np.asarray(a_d, np.float32).view(np.uint32) == np.asarray(b_d, np.float32).view(np.uint32)).all())


def _bcast_np(arr, shape, dtype, dist, dev, rank):
    """rank 0's numpy array to every rank (raw bytes over NCCL)"""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(arr)).to(dev) if rank == 0 else torch.empty(shape, dtype=dtype, device=dev)
    dist.broadcast(t.view(torch.uint8), 0)
    return t.cpu().numpy()


def run(a, eng_factory, dev, rank, world, dist, vec_t, peaks, which, sample, reps=5):
    """vec_t: the 3M x 300 table as a CUDA tensor on rank 0 (None elsewhere).  Returns a list of dicts (rank 0)."""
    import torch
    from freddy_b200 import _lib
    from freddy_b200.dist import allgather_topk, shard_range
    from freddy_b200.index_build import make_ivpq_index, make_synthetic_index
    from oracle import oracle

    hbm_gbs, bf16_tf = peaks
    have_ref = os.path.exists(oracle.REF_SO)
    threads = os.cpu_count() or 1
    out = []
    N, d, m, K = a.n, a.d, a.m, a.K
    g = torch.Generator(); g.manual_seed(99)
    perm = torch.randperm(N, generator=g)

    def sync():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    eng = eng_factory()
    eng.set_option(_lib.FB_OPT_PROFILE, 1)
    vec = None

    def host_vectors():
        nonlocal vec
        if vec is None:
            if world > 1:
                t = vec_t if rank == 0 else torch.empty(N, d, device=dev)
                dist.broadcast(t, 0)
                vec = t.cpu().numpy()
                del t
            else:
                vec = vec_t.cpu().numpy()
        return vec

    # ------------------------------------------------------------------ config 3 (rank 0 only)
    if "3" in which and rank == 0:
        nq, nt, k = 5000, 100_000, 5
        t0 = time.time()
        ixp = make_synthetic_index(N, d=d, m=m, K=K, C=a.C, n_train=min(100_000, N), n_clusters=1000, sigma=a.sigma, zipf=a.zipf,
                                   kmeans_iters=10, seed=1234, device=dev, with_pq=True, keep_vectors=False)
        t_build = time.time() - t0
        q = vec_t[perm[:nq].to(dev)].cpu().numpy()
        targets = (perm[nq:nq + nt] + 1).numpy().astype(np.int32)
        eng.load_pq_index(ixp)
        res = {}
        sec, _ = _timed(lambda: res.__setitem__("r", eng.pq_search_in_batch(q, k, targets)), reps, sync, eng.reset_counters)
        c = eng.counters()
        ids, dd = res["r"]
        n_runs = reps
        ms_scan = (c["ms_scan"] + c["ms_finalize"]) / n_runs
        lookups = float(nq) * nt * m
        ns = min(sample, nq)
        if have_ref:
            rs = oracle.ReferenceSession()
            rs.load_pq(ixp)
            tt = time.time()
            _, rids, rraw = rs.pq_search_in_batch(q[:ns], np.arange(ns, dtype=np.int32), k, targets, False)
            t_cpu = time.time() - tt
            ok, kind = _same(ids[:ns], dd[:ns], rids, rraw), "reference"
        else:
            tt = time.time()
            rids, rraw = oracle.OracleIndex(ixp, flat_pq=True).pq_search_in_batch(q[:ns], k, targets)
            t_cpu = time.time() - tt
            ok, kind = _same(ids[:ns], dd[:ns], rids, rraw), "port"
        lsu_peak = 148 * 32 * 1.965e9 / 1e12      # one 4-byte shared-memory gather per lane per clock, 148 SMs
        out.append({"name": "config3 pq_search_in_batch 5k x 100k (m=12, K=1024, k=5)", "seconds": sec, "queries_per_s": nq / sec,
                    "gpu_launches": c["kernel_launches"] / n_runs, "index_build_s": round(t_build, 1),
                    "stage_ms": {"subset+lut": c["ms_lut"] / n_runs, "scan": ms_scan, "exact": c["ms_exact"] / n_runs},
                    "roofline": {"bound": "shared-memory gather (LSU)", "achieved": lookups / (ms_scan / 1e3) / 1e12 if ms_scan > 0 else None,
                                 "peak": lsu_peak, "unit": "T lookups/s", "frac": lookups / (ms_scan / 1e3) / 1e12 / lsu_peak if ms_scan > 0 else None,
                                 "peak_source": "148 SMs x 32 lanes x 1.965 GHz (one 4-byte LDS per lane per clock)",
                                 "end_to_end_T_lookups_per_s": lookups / sec / 1e12},
                    "exact_path_queries": c["exact_path_queries"] / n_runs,
                    "equals_reference_on_sample": {"queries": ns, "ok": ok, "kind": kind, "cpu_seconds": t_cpu,
                                                   "cpu_queries_per_s_1thread": ns / t_cpu}})
        del ixp

    # ------------------------------------------------------------------ config 4 (all ranks)
    ivpq = None
    if "4" in which:
        nq, nt, k, alpha, pvf, method, conf = 5000, 100_000, 5, 100, 20, 2, 0.8
        trows = np.sort(perm[nq:nq + nt].numpy())
        t0 = time.time()
        if rank == 0:
            ivpq = make_ivpq_index(vec_t, m=12, K=1024, Kc=32, n_train=100_000, kmeans_iters=10, seed=77, target_rows=trows)
        if world > 1:
            shapes = {"coarse_multi": ((2, 32, d // 2), torch.float32), "ivpq_codebook": ((12, 1024, d // 12), torch.float32),
                      "ivpq_coarse_ids": ((N,), torch.int32), "ivpq_codes": ((N, 12), torch.int16), "stats": ((32 * 32 + 1,), torch.float32)}
            if rank != 0:
                ivpq = {"d": d, "m": 12, "K": 1024, "Kc": 32, "N": N, "ids": np.arange(1, N + 1, dtype=np.int32)}
            for nm, (shp, dt) in shapes.items():
                ivpq[nm] = _bcast_np(ivpq.get(nm), shp, dt, dist, dev, rank)
        t_build = time.time() - t0
        v = host_vectors()
        ids_all = np.asarray(ivpq["ids"], np.int32)
        q_all = v[perm[:nq].numpy()]
        targets = (trows + 1).astype(np.int32)
        eng.load_ivpq_index(ivpq)
        eng.load_vectors(ids_all, v)
        b, e_ = shard_range(nq, rank, world)
        my_q = np.ascontiguousarray(q_all[b:e_])
        res = {}

        def join_step():
            ids, dd = eng.ivpq_search_in(my_q, k, targets, alpha, pvf, method, True, conf)
            if world > 1:
                gi, gd = allgather_topk(torch.from_numpy(ids).to(dev), torch.from_numpy(dd).to(dev), nq)
                res["r"] = (gi.cpu().numpy(), gd.cpu().numpy())
            else:
                res["r"] = (ids, dd)

        sec, _ = _timed(join_step, reps, sync, eng.reset_counters)
        sec = max_over_ranks(sec)
        c = eng.counters()
        n_runs = reps
        if rank == 0:
            ids, dd = res["r"]
            ns = min(sample, nq)
            oi = oracle.OracleIvpq(ivpq, v, ids_all)
            tt = time.time()
            eids, ed, rc, st = oi.search_in(q_all[:ns], k, targets, alpha, pvf, method, True, conf)
            t_cpu = time.time() - tt
            ok, kind = _same(ids[:ns], dd[:ns], eids, ed) and rc == 0, "port (pinned to the real ivpq_search_in SRF in tests/)"
            pairs = st[1] / ns * nq                                   # candidate (query, row) pairs of the whole job
            algo_bytes = pairs * (2 * 12 + 4) + float(nq) * k * pvf * d * 4
            ms_dev = (c["ms_coarse"] + c["ms_lut"] + c["ms_scan"]) / n_runs
            out.append({"name": "config4 knn_join ivpq_search_in 5k x 100k (k=5, alpha=100, pvf=20, method 2, Kc=32, m=12, K=1024)",
                        "seconds": sec, "queries_per_s": nq / sec, "n_gpus": world,
                        "parallelism": "index replicated, queries sharded, one all-gather of per-rank top-k" if world > 1 else "single GPU",
                        "gpu_launches": c["kernel_launches"] / n_runs, "index_build_s": round(t_build, 1),
                        "stage_ms_rank0": {"select": c["ms_coarse"] / n_runs, "lut": c["ms_lut"] / n_runs, "scan+postverify": c["ms_scan"] / n_runs},
                        "candidate_pairs_per_query": st[1] / ns, "rounds": int(st[0]),
                        "roofline": {"bound": "hbm (gather)", "achieved": algo_bytes / (ms_dev / 1e3) / 1e9 * world if ms_dev > 0 else None,
                                     "peak": hbm_gbs * world, "unit": "GB/s",
                                     "frac": algo_bytes / (ms_dev / 1e3) / 1e9 / hbm_gbs if ms_dev > 0 else None,
                                     "algorithmic_bytes": algo_bytes,
                                     "note": "pairs x 28 B of codes + k*pvf gathered 1200-byte vectors per query; latency-bound at this size"},
                        "equals_reference_on_sample": {"queries": ns, "ok": bool(ok), "kind": kind, "cpu_seconds": t_cpu,
                                                       "cpu_queries_per_s_1thread": ns / t_cpu}})

    # ------------------------------------------------------------------ config 5 (all ranks; vocabulary sharded when N > 1)
    if "5" in which:
        nqa = 1000
        v = host_vectors()
        ids_all = np.arange(1, N + 1, dtype=np.int32)
        rows_abc = torch.randint(0, N, (nqa, 3), generator=g).numpy().astype(np.int32)
        res = {}
        if world == 1:
            if ivpq is None:
                eng.load_vectors(ids_all, v)

            def ana_step():
                res["r"] = eng.analogy_3cosadd(ids_all[rows_abc])
        else:
            vb, ve = shard_range(N, rank, world)
            eng.load_vectors(ids_all[vb:ve], np.ascontiguousarray(v[vb:ve]))
            qv = ((v[rows_abc[:, 2]] - v[rows_abc[:, 0]]) + v[rows_abc[:, 1]]).astype(np.float32)   # vec_minus then vec_plus (fp32)
            ex_ids = ids_all[rows_abc]

            def ana_step():
                li, ls = eng.analogy_scan(qv, ex_ids)                       # local arg-max over this rank's rows
                lrow = np.where(li >= 0, li - 1, np.iinfo(np.int32).max).astype(np.int64)   # global table row (ids are row + 1 here)
                t_s = torch.from_numpy(ls).to(dev); t_r = torch.from_numpy(lrow).to(dev); t_i = torch.from_numpy(li).to(dev)
                gs = torch.empty(world, nqa, device=dev); gr = torch.empty(world, nqa, dtype=torch.int64, device=dev)
                gi_ = torch.empty(world, nqa, dtype=torch.int32, device=dev)
                dist.all_gather_into_tensor(gs, t_s); dist.all_gather_into_tensor(gr, t_r); dist.all_gather_into_tensor(gi_, t_i)
                gs = torch.where(gi_ >= 0, gs, torch.full_like(gs, float("-inf")))
                best = gs.max(0).values
                cand = torch.where(gs == best, gr, torch.full_like(gr, 2 ** 62))     # ties: the earliest table row wins
                win = cand.argmin(0)
                res["r"] = (gi_.gather(0, win[None])[0].cpu().numpy(), best.cpu().numpy())

        sec, _ = _timed(ana_step, reps, sync, eng.reset_counters)
        sec = max_over_ranks(sec)
        c = eng.counters()
        n_runs = reps
        if rank == 0:
            got_ids, got_s = res["r"]
            ns = min(sample, nqa)
            tt = time.time()
            erows, es = oracle.analogy_3cosadd(v, rows_abc[:ns], threads=threads)
            t_cpu = time.time() - tt
            ok = bool((got_ids[:ns] == ids_all[erows]).all() and (got_s[:ns].view(np.uint32) == es.view(np.uint32)).all())
            n_loc = (N + world - 1) // world
            flops = 2.0 * ((nqa + 127) // 128 * 128) * ((n_loc + 255) // 256 * 256) * ((d + 15) // 16 * 16)
            ms_gemm = c["ms_scan"] / n_runs
            out.append({"name": "config5 analogy_3cosadd 1k triples, exact over 3M x 300", "seconds": sec, "queries_per_s": nqa / sec, "n_gpus": world,
                        "parallelism": "vocabulary sharded, one all-gather of per-rank arg-max" if world > 1 else "single GPU",
                        "gpu_launches": c["kernel_launches"] / n_runs,
                        "stage_ms_rank0": {"prefilter_gemm": ms_gemm, "exact_rescore": c["ms_finalize"] / n_runs},
                        "prefilter": {"queries": c["prefilter_queries"] / n_runs, "overflow_queries": c["prefilter_overflow_queries"] / n_runs,
                                      "candidates_per_query": c["prefilter_candidates"] / max(1, c["prefilter_queries"])},
                        "roofline": {"bound": "tensor", "kernel": "prefilter_gemm_kernel (tcgen05.mma kind::f16, bf16 x bf16 -> fp32 in TMEM)",
                                     "achieved": flops / (ms_gemm / 1e3) / 1e12 if ms_gemm > 0 else None, "peak": bf16_tf, "unit": "TFLOP/s",
                                     "frac": flops / (ms_gemm / 1e3) / 1e12 / bf16_tf if ms_gemm > 0 and bf16_tf else None,
                                     "peak_source": "MEASURED_PEAKS.json bf16_tflops (burst)", "flops_per_launch_per_gpu": flops,
                                     "table_read_GBps": n_loc * 640 / (ms_gemm / 1e3) / 1e9 if ms_gemm > 0 else None},
                        "equals_reference_on_sample": {"queries": ns, "ok": ok, "kind": "port (analogy_3cosadd is plpgsql over cosine_similarity_bytea; "
                                                       "fo_analogy_3cosadd restates freddy--0.0.1.sql:1270-1288 + core_functions.c:67-81)",
                                                       "cpu_seconds": t_cpu, "cpu_threads": threads,
                                                       "cpu_queries_per_s_all_threads": ns / t_cpu}})
    vec = None

    # ------------------------------------------------------------------ generator variants of the headline shape (rank 0 only)
    def ivfadc_variant(name, sigma, zipf, kmeans_iters, n_train, note, from_centres=False):
        t0 = time.time()
        ix = make_synthetic_index(N, d=d, m=m, K=K, C=a.C, n_train=min(n_train, N), n_clusters=1000, sigma=sigma, zipf=zipf,
                                  kmeans_iters=kmeans_iters, seed=1234, device=dev, keep_vectors=True,
                                  coarse_from_centres=from_centres)
        vt = ix.pop("vectors_t")
        gq = torch.Generator(); gq.manual_seed(4321)
        sel = torch.randperm(N, generator=gq)[:a.batch]
        q = vt[sel.to(dev)].cpu().numpy()
        del vt
        torch.cuda.empty_cache()
        t_build = time.time() - t0
        eng.load_ivfadc_index(ix)
        nq, k, w = len(q), a.k, a.w
        hq = torch.from_numpy(q).pin_memory()
        hi = torch.empty(nq, k, dtype=torch.int32).pin_memory()
        hd = torch.empty(nq, k, dtype=torch.float32).pin_memory()
        sec, _ = _timed(lambda: eng.ivfadc_search_ptr(hq.data_ptr(), nq, k, w, hi.data_ptr(), hd.data_ptr()), 3, sync, eng.reset_counters)
        c = eng.counters()
        n_runs = 3
        ms_dom = c["ms_pipe"] if c["n_pipe_launches"] else c["ms_scan"]
        ns = min(sample, nq)
        if have_ref:
            rs = oracle.ReferenceSession()
            rs.load_ivfadc(ix, w)
            tt = time.time()
            rids, rraw, _ = rs.ivfadc_search(q[:ns], k)
            t_cpu = time.time() - tt
            kind = "reference"
        else:
            tt = time.time()
            rids, rraw, _, _ = oracle.OracleIndex(ix).ivfadc_search(q[:ns], k, w, threads=threads)
            t_cpu = time.time() - tt
            kind = "port"
        ok = _same(hi.numpy()[:ns], hd.numpy()[:ns], rids, rraw)
        gbs = c["scan_bytes"] / 1e9 / (ms_dom / 1e3) if ms_dom > 0 else None
        out.append({"name": name, "generator": {"sigma": sigma, "zipf": zipf, "kmeans_iters": kmeans_iters, "n_train": n_train}, "note": note,
                    "seconds": sec, "queries_per_s": nq / sec, "queries": nq, "index_build_s": round(t_build, 1),
                    "rows_scanned_per_query": c["rows_scanned"] / max(1, c["queries"]),
                    "exact_path_queries": c["exact_path_queries"] / n_runs,
                    "exact_path_reasons": {r: c["exact_" + r] / n_runs for r in ("coarse_tie", "coarse_far", "few_rows", "scan_tie", "forced")},
                    "stage_ms": {s: c["ms_" + s] / n_runs for s in ("coarse", "lut", "scan", "pipe", "finalize", "exact")},
                    "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm_gbs, "unit": "GB/s", "frac": gbs / hbm_gbs if gbs else None,
                                 "note": "algorithmic scan bytes (rows x 28 B) over the event time of the dominant kernel's launches"},
                    "equals_reference_on_sample": {"queries": ns, "ok": ok, "kind": kind, "cpu_seconds": t_cpu,
                                                   "cpu_queries_per_s_1thread": ns / t_cpu if kind == "reference" else None}})

    if rank == 0:
        if "sigma03" in which:
            ivfadc_variant("sigma03: headline shape on SURVEY 8(d)'s generator", 0.3, 0.7, 10, 100_000,
                           "sigma = 0.3 collapses clusters onto few distinct code vectors: many exact-distance ties across the k-th place")
        if "nominal" in which:
            ivfadc_variant("nominal: headline shape, probed lists near N*w/C rows", a.sigma, 0.0, 10, 100_000,
                           "equal cluster sizes, coarse k-means started from the generating centres: one list per cluster, "
                           "about the nominal N/C = 3000 rows each", from_centres=True)
    eng.close()
    return out
