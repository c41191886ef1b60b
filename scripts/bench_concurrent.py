#!/usr/bin/env python
"""What SQL can get today: P backend-like processes, each with its own engine (= its own CUDA context, as a Postgres
backend would have after fork), each issuing single-query fb_ivfadc_search calls (the form `ivfadc_search(bytea, int)`
has, freddy--0.0.1.sql:370-372) against ONE GPU.  Reports total queries/s and p50 / p99 latency per P, and — for
contrast — the same number of queries handed over as one batch call.  

  python scripts/bench_concurrent.py --procs 1,4,16,64 --seconds 4
"""
import argparse
import json
import multiprocessing as mp
import os
import shutil
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "postgres-word2vec_b200")]

KEYS = ("coarse", "residual_codebook", "ids", "coarse_ids", "codes")


def worker(args):
    shm, seconds, k, w, seed, barrier_file, n_procs = args
    from freddy_b200 import Engine
    ix = {kk: np.load(os.path.join(shm, kk + ".npy"), mmap_mode="r") for kk in KEYS}
    ix.update(json.load(open(os.path.join(shm, "meta.json"))))
    q_all = np.load(os.path.join(shm, "queries.npy"))
    rng = np.random.default_rng(seed)
    eng = Engine(0)
    t0 = time.time()
    eng.load_ivfadc_index(ix)
    t_load = time.time() - t0
    q = np.ascontiguousarray(q_all[rng.integers(0, len(q_all), 4096)])
    for i in range(30):
        eng.ivfadc_search(q[i:i + 1], k, w)
    # crude barrier: every process touches a file, then waits until all have
    open(os.path.join(shm, f"ready_{seed}"), "w").close()
    while len([f for f in os.listdir(shm) if f.startswith("ready_")]) < n_procs:
        time.sleep(0.005)
    lat = []
    t_end = time.time() + seconds
    i = 0
    while time.time() < t_end:
        t = time.perf_counter()
        eng.ivfadc_search(q[i % 4096:i % 4096 + 1], k, w)
        lat.append(time.perf_counter() - t)
        i += 1
    eng.close()
    return np.asarray(lat), t_load


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--procs", default="1,4,16,64")
    ap.add_argument("--seconds", type=float, default=4.0)
    ap.add_argument("--n", type=int, default=3_000_000)
    ap.add_argument("--k", type=int, default=5)
    ap.add_argument("--w", type=int, default=10)
    a = ap.parse_args()
    import torch
    from freddy_b200 import Engine
    from freddy_b200.index_build import make_synthetic_index
    ix = make_synthetic_index(a.n, d=300, m=12, K=1024, C=1000, n_train=100_000, n_clusters=1000, sigma=1.0, zipf=0.35,
                              kmeans_iters=10, seed=1234, device="cuda", keep_vectors=True)
    vec = ix.pop("vectors_t")
    g = torch.Generator(); g.manual_seed(4321)
    queries = vec[torch.randperm(a.n, generator=g)[:32768].cuda()].cpu().numpy()
    del vec
    torch.cuda.empty_cache()
    shm = tempfile.mkdtemp(prefix="fb_conc_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        for kk in KEYS:
            np.save(os.path.join(shm, kk + ".npy"), np.ascontiguousarray(ix[kk]))
        np.save(os.path.join(shm, "queries.npy"), queries)
        json.dump({kk: int(ix[kk]) for kk in ("d", "m", "K", "C", "N")}, open(os.path.join(shm, "meta.json"), "w"))
        # the batch form, for contrast: the same engine answering 32768 queries in one call
        eng = Engine(0)
        eng.load_ivfadc_index(ix)
        eng.ivfadc_search(queries, a.k, a.w)
        t = time.perf_counter()
        eng.ivfadc_search(queries, a.k, a.w)
        batch_qps = len(queries) / (time.perf_counter() - t)
        eng.close()
        out = {"workload": f"single-query fb_ivfadc_search calls (k={a.k}, w={a.w}) on N={a.n}, one engine per process, one GPU",
               "mps": os.path.exists("/tmp/nvidia-mps") or bool(os.environ.get("CUDA_MPS_PIPE_DIRECTORY")),
               "one_batch_call_queries_per_s": batch_qps, "runs": []}
        ctx = mp.get_context("spawn")
        for P in [int(x) for x in a.procs.split(",")]:
            for f in os.listdir(shm):
                if f.startswith("ready_"):
                    os.remove(os.path.join(shm, f))
            with ctx.Pool(P) as pool:
                res = pool.map(worker, [(shm, a.seconds, a.k, a.w, s, None, P) for s in range(P)], chunksize=1)
            lat = np.concatenate([r[0] for r in res])
            out["runs"].append({"processes": P, "queries_per_s": float(sum(len(r[0]) for r in res) / a.seconds),
                                "latency_us_p50": float(np.percentile(lat, 50) * 1e6), "latency_us_p99": float(np.percentile(lat, 99) * 1e6),
                                "index_upload_s_per_process": float(np.mean([r[1] for r in res]))})
            print(json.dumps(out["runs"][-1]), flush=True)
        print(json.dumps(out))
    finally:
        shutil.rmtree(shm, ignore_errors=True)


if __name__ == "__main__":
    main()
