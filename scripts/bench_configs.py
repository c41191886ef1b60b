#!/usr/bin/env python
"""Secondary BASELINE.json configs (bench.py carries the headline one):
  config 3  pq_search_in_batch  5k queries x 100k targets, m=12, K=1024        (freddy.c:414-675)
  config 4  knn_join / ivpq_search_in  5k x 100k, k=5, alpha=100, pvf=20, method 2 (PQ + post verification)
                                                                                  (ivpq_search_in.c:61-721)
  config 5  analogy_3cosadd     1k triples, exact scan over the full 3M vocab    (freddy--0.0.1.sql:1270-1288)
Each prints one JSON line: GPU time through the host-buffer C-ABI call (H2D/D2H included),
the CPU oracle on a bounded sample with all host threads, and a parity check on that sample."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "postgres-word2vec_b200")]
import torch  # noqa: E402

from freddy_b200 import Engine  # noqa: E402
from freddy_b200.index_build import make_synthetic_index  # noqa: E402
from oracle import oracle  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=3_000_000)
ap.add_argument("--which", default="3,4,5")
ap.add_argument("--reps", type=int, default=5)
a = ap.parse_args()
threads = os.cpu_count() or 1

ix = make_synthetic_index(a.n, d=300, m=12, K=1024, C=1000, n_train=100_000, n_clusters=1000, sigma=1.0, zipf=0.35,
                          kmeans_iters=10, seed=1234, device="cuda", with_pq=True, keep_vectors=True)
vec_t = ix.pop("vectors_t")
g = torch.Generator(); g.manual_seed(99)
eng = Engine(0)


def timed(fn, reps):
    fn()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t)
    return float(np.median(ts))


if "3" in a.which.split(","):
    nq, nt, k = 5000, 100_000, 5
    perm = torch.randperm(a.n, generator=g)
    q = vec_t[perm[:nq].cuda()].cpu().numpy()
    targets = (perm[nq:nq + nt] + 1).numpy().astype(np.int32)
    eng.load_pq_index(ix)
    res = {}
    t_gpu = timed(lambda: res.__setitem__("r", eng.pq_search_in_batch(q, k, targets)), a.reps)
    ids, d = res["r"]
    oi = oracle.OracleIndex(ix, flat_pq=True)
    ns = 64
    t0 = time.perf_counter()
    eids, ed = oi.pq_search_in_batch(q[:ns], k, targets)
    t_cpu = time.perf_counter() - t0
    ok = bool((ids[:ns] == eids).all() and (d[:ns].view(np.uint32) == ed.view(np.uint32)).all())
    lookups = nq * nt * 12
    print(json.dumps({"config": "pq_search_in_batch 5k x 100k, m=12, K=1024, k=5", "gpu_seconds_e2e": t_gpu,
                      "gpu_lookups_per_s": lookups / t_gpu, "queries_per_s": nq / t_gpu,
                      "cpu_oracle": {"sample_queries": ns, "seconds": t_cpu, "queries_per_s_1thread": ns / t_cpu, "kind": "port"},
                      "parity_on_sample": ok, "reference_published_s": 14.4}))

if "4" in a.which.split(","):
    from freddy_b200.index_build import make_ivpq_index
    nq, nt, k, alpha, pvf, method, conf = 5000, 100_000, 5, 100, 20, 2, 0.8
    perm = torch.randperm(a.n, generator=g)
    trows = np.sort(perm[nq:nq + nt].numpy())
    ivpq = make_ivpq_index(vec_t, m=12, K=1024, Kc=32, n_train=100_000, kmeans_iters=10, seed=77, target_rows=trows)
    vec = vec_t.cpu().numpy()
    q = vec[perm[:nq].numpy()]
    targets = (trows + 1).astype(np.int32)
    eng.load_ivpq_index(ivpq)
    eng.load_vectors(ivpq["ids"], vec)
    res = {}
    for use_tl in (True,):
        t_gpu = timed(lambda: res.__setitem__("r", eng.ivpq_search_in(q, k, targets, alpha, pvf, method, use_tl, conf)), a.reps)
        ids, d = res["r"]
        oi = oracle.OracleIvpq(ivpq, vec, ivpq["ids"])
        ns = 32
        t0 = time.perf_counter()
        eids, ed, rc, st = oi.search_in(q[:ns], k, targets, alpha, pvf, method, use_tl, conf)
        t_cpu = time.perf_counter() - t0
        same_ids = (ids[:ns] == eids).all(axis=1)
        ok = bool(same_ids.all() and (d[:ns].view(np.uint32) == ed.view(np.uint32)).all())
        print(json.dumps({"config": f"knn_join ivpq_search_in 5k x 100k, k=5, alpha=100, pvf=20, method=2, use_targetlist={use_tl}, m=12 K=1024 Kc=32",
                          "gpu_seconds_e2e": t_gpu, "queries_per_s": nq / t_gpu,
                          "cpu_oracle": {"sample_queries": ns, "seconds": t_cpu, "queries_per_s_1thread": ns / t_cpu,
                                         "rounds": int(st[0]), "candidate_pairs_per_query": st[1] / ns, "kind": "port"},
                          "parity_on_sample": ok, "sample_queries_with_identical_ids": int(same_ids.sum()),
                          "reference_published_s": "8.4-12.2 (PQ+PV, alpha=1000)"}))

if "5" in a.which.split(","):
    nq = 1000
    vec = vec_t.cpu().numpy()
    ids_all = ix["ids"]
    eng.load_vectors(ids_all, vec)
    rows = torch.randint(0, a.n, (nq, 3), generator=g).numpy().astype(np.int32)
    res = {}
    t_gpu = timed(lambda: res.__setitem__("r", eng.analogy_3cosadd(ids_all[rows])), a.reps)
    got_ids, got_s = res["r"]
    ns = min(nq, 2 * threads)
    t0 = time.perf_counter()
    erows, es = oracle.analogy_3cosadd(vec, rows[:ns], threads=threads)
    t_cpu = time.perf_counter() - t0
    ok = bool((got_ids[:ns] == ids_all[erows]).all() and (got_s[:ns].view(np.uint32) == es.view(np.uint32)).all())
    flops = 2.0 * nq * a.n * 300
    print(json.dumps({"config": "analogy_3cosadd 1k triples, exact scan over 3M x 300", "gpu_seconds_e2e": t_gpu,
                      "queries_per_s": nq / t_gpu, "rounded_fp32_ops_per_s": flops / t_gpu,
                      "table_bytes": a.n * 1200, "cpu_oracle": {"sample_queries": ns, "threads": threads, "seconds": t_cpu,
                                                                "queries_per_s": ns / t_cpu, "kind": "port"},
                      "parity_on_sample": ok}))

if "6" in a.which.split(","):
    # k_nearest_neighbour_ivfadc_pv, batch form: ivfadc_search(v, pvf*k) + exact re-rank  (freddy--0.0.1.sql:574-591)
    nq, k, pvf, w = 5000, 5, 20, 10
    vec = vec_t.cpu().numpy()
    ids_all = ix["ids"]
    eng.load_ivfadc_index(ix)
    eng.load_vectors(ids_all, vec)
    perm = torch.randperm(a.n, generator=g)
    q = vec[perm[:nq].numpy()]
    res = {}
    t_gpu = timed(lambda: res.__setitem__("r", eng.ivfadc_search_pv(q, k, pvf, w)), a.reps)
    ids, s = res["r"]
    ns = 32
    t0 = time.perf_counter()
    eids, es = oracle.ivfadc_search_pv(oracle.OracleIndex(ix), vec, ids_all, q[:ns], k, pvf, w, threads=threads)
    t_cpu = time.perf_counter() - t0
    ok = bool((ids[:ns] == eids).all() and (s[:ns].view(np.uint32) == es.view(np.uint32)).all())
    # recall of the exact neighbours on the sample (the README's precision column)
    xids, _ = eng.knn_exact(q[:ns], k)
    prec = float(np.mean([len(set(ids[i]) & set(xids[i])) / k for i in range(ns)]))
    print(json.dumps({"config": f"k_nearest_neighbour_ivfadc_pv batch: {nq} queries, k={k}, pvf={pvf}, w={w}, 3M x 300",
                      "gpu_seconds_e2e": t_gpu, "queries_per_s": nq / t_gpu,
                      "cpu_oracle": {"sample_queries": ns, "threads": threads, "seconds": t_cpu, "queries_per_s": ns / t_cpu, "kind": "port"},
                      "parity_on_sample": ok, "precision_at_k_vs_exact_on_sample": prec}))

if "7" in a.which.split(","):
    # k_nearest_neighbour (exact): cosine_similarity_bytea over all 3M rows, top-k  (freddy--0.0.1.sql:426-439)
    nq, k = 1000, 5
    vec = vec_t.cpu().numpy()
    ids_all = ix["ids"]
    eng.load_vectors(ids_all, vec)
    perm = torch.randperm(a.n, generator=g)
    q = vec[perm[:nq].numpy()]
    res = {}
    t_gpu = timed(lambda: res.__setitem__("r", eng.knn_exact(q, k)), a.reps)
    ids, s = res["r"]
    ns = 4
    t0 = time.perf_counter()
    eids, es = oracle.knn_exact(vec, ids_all, q[:ns], k)
    t_cpu = time.perf_counter() - t0
    ok = bool((ids[:ns] == eids).all() and (s[:ns].view(np.uint32) == es.view(np.uint32)).all())
    print(json.dumps({"config": f"k_nearest_neighbour exact: {nq} queries, k={k}, scan over 3M x 300", "gpu_seconds_e2e": t_gpu,
                      "queries_per_s": nq / t_gpu, "rounded_fp32_ops_per_s": 2.0 * nq * a.n * 300 / t_gpu,
                      "cpu_oracle": {"sample_queries": ns, "threads": 1, "seconds": t_cpu, "queries_per_s": ns / t_cpu,
                                     "kind": "port (numpy, vectorised over rows)"},
                      "parity_on_sample": ok, "reference_published_s_per_query": 8.79}))
