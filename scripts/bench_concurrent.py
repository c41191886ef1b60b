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


def sidecar_client(args):
    """a backend without CUDA: posts single queries into the sidecar's shared-memory slots (libfreddy_sidecar.so)"""
    shm, seconds, k, w, seed, name, n_procs = args
    import ctypes as C
    lib = C.CDLL(os.path.join(ROOT, "postgres-word2vec_b200", "libfreddy_sidecar.so"))
    lib.fbsc_client_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.fbsc_client_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    lib.fbsc_client_close.argtypes = [C.c_void_p]
    h = C.c_void_p()
    assert lib.fbsc_client_open(name.encode(), C.byref(h)) == 0
    q_all = np.load(os.path.join(shm, "queries.npy"))
    rng = np.random.default_rng(seed)
    q = np.ascontiguousarray(q_all[rng.integers(0, len(q_all), 4096)])
    ids, dists = np.empty(k, np.int32), np.empty(k, np.float32)
    pi, pd = ids.ctypes.data_as(C.c_void_p), dists.ctypes.data_as(C.c_void_p)
    row = q.strides[0]
    base = q.ctypes.data
    for i in range(30):
        assert lib.fbsc_client_search(h, C.c_void_p(base + i * row), k, w, pi, pd, 10000) == 0
    open(os.path.join(shm, f"ready_{seed}"), "w").close()
    while len([f for f in os.listdir(shm) if f.startswith("ready_")]) < n_procs:
        time.sleep(0.005)
    lat = []
    t_end = time.time() + seconds
    i = 0
    first = None
    while time.time() < t_end:
        t = time.perf_counter()
        rc = lib.fbsc_client_search(h, C.c_void_p(base + (i % 4096) * row), k, w, pi, pd, 10000)
        lat.append(time.perf_counter() - t)
        assert rc == 0, rc
        if first is None:
            first = (int(i % 4096), ids.copy(), dists.copy())
        i += 1
    lib.fbsc_client_close(h)
    return np.asarray(lat), q[first[0]], first[1], first[2]


def sidecar_server(shm, name, k, stop_file, result_file, linger_us):
    from freddy_b200 import Engine
    ix = {kk: np.load(os.path.join(shm, kk + ".npy"), mmap_mode="r") for kk in KEYS}
    ix.update(json.load(open(os.path.join(shm, "meta.json"))))
    eng = Engine(0)
    eng.load_ivfadc_index(ix)
    eng.sidecar_start(name, max_k=max(k, 16), slots=256, max_batch=256, linger_us=linger_us)
    open(os.path.join(shm, "server_up"), "w").close()
    while not os.path.exists(stop_file):
        time.sleep(0.01)
    c = eng.sidecar_stop()
    json.dump(c, open(result_file, "w"))
    eng.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sidecar", type=int, default=1, help="1: also measure the same callers going through one sidecar engine")
    ap.add_argument("--linger-us", type=int, default=0)
    ap.add_argument("--direct", type=int, default=1, help="0: skip the one-engine-per-process runs")
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
        for P in [int(x) for x in a.procs.split(",")] if a.direct else []:
            for f in os.listdir(shm):
                if f.startswith("ready_"):
                    os.remove(os.path.join(shm, f))
            try:
                with ctx.Pool(P) as pool:
                    res = pool.map(worker, [(shm, a.seconds, a.k, a.w, s, None, P) for s in range(P)], chunksize=1)
            except Exception as ex:          # e.g. more processes than the device (or MPS: 48 clients) admits
                out["runs"].append({"processes": P, "error": repr(ex)[:300]})
                print(json.dumps(out["runs"][-1]), flush=True)
                continue
            lat = np.concatenate([r[0] for r in res])
            out["runs"].append({"processes": P, "queries_per_s": float(sum(len(r[0]) for r in res) / a.seconds),
                                "latency_us_p50": float(np.percentile(lat, 50) * 1e6), "latency_us_p99": float(np.percentile(lat, 99) * 1e6),
                                "index_upload_s_per_process": float(np.mean([r[1] for r in res]))})
            print(json.dumps(out["runs"][-1]), flush=True)
        if a.sidecar:
            # the same callers, none of them with a CUDA context: one sidecar process owns the engine
            out["sidecar_runs"] = []
            name = f"/freddy_bench_{os.getpid()}"
            stop_file, result_file = os.path.join(shm, "stop"), os.path.join(shm, "sidecar_counters.json")
            check = Engine(0)
            check.load_ivfadc_index(ix)
            for P in [int(x) for x in a.procs.split(",")]:
                for f in os.listdir(shm):
                    if f.startswith("ready_") or f in ("stop", "server_up", "sidecar_counters.json"):
                        os.remove(os.path.join(shm, f))
                srv = ctx.Process(target=sidecar_server, args=(shm, name, a.k, stop_file, result_file, a.linger_us))
                srv.start()
                while not os.path.exists(os.path.join(shm, "server_up")):
                    time.sleep(0.05)
                    assert srv.is_alive(), "sidecar died"
                with ctx.Pool(P) as pool:
                    res = pool.map(sidecar_client, [(shm, a.seconds, a.k, a.w, s, name, P) for s in range(P)], chunksize=1)
                open(stop_file, "w").close()
                srv.join(60)
                counters = json.load(open(result_file))
                lat = np.concatenate([r[0] for r in res])
                # what a caller got back is what a direct call returns
                same = all((np.array_equal(check.ivfadc_search(r[1][None], a.k, a.w)[0][0], r[2]) and
                            np.array_equal(check.ivfadc_search(r[1][None], a.k, a.w)[1][0].view(np.uint32), r[3].view(np.uint32))) for r in res)
                out["sidecar_runs"].append({"processes": P, "queries_per_s": float(sum(len(r[0]) for r in res) / a.seconds),
                                            "latency_us_p50": float(np.percentile(lat, 50) * 1e6),
                                            "latency_us_p99": float(np.percentile(lat, 99) * 1e6),
                                            "mean_batch": counters["queries"] / max(1, counters["batches"]),
                                            "largest_batch": counters["largest_batch"], "equals_direct_call": bool(same)})
                print(json.dumps({"sidecar": out["sidecar_runs"][-1]}), flush=True)
            check.close()
        print(json.dumps(out))
    finally:
        shutil.rmtree(shm, ignore_errors=True)


if __name__ == "__main__":
    main()
