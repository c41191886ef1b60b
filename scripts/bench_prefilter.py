#!/usr/bin/env python
"""Per-call stage times of the tensor-core pre-filter through the C-ABI (fb_knn_exact / fb_analogy_3cosadd) on the bench's
3M x 300 table: GEMM and re-score CUDA-event times per call, end-to-end seconds, candidates per query."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "postgres-word2vec_b200")]
import torch
from freddy_b200 import Engine, _lib
from freddy_b200.index_build import make_synthetic_index

N = int(os.environ.get("N", 3_000_000))
ix = make_synthetic_index(N, d=300, m=12, K=64, C=100, n_train=20_000, n_clusters=1000, sigma=1.0, zipf=0.35, kmeans_iters=1, seed=1234,
                          device="cuda", keep_vectors=True)
vec = ix.pop("vectors_t").cpu().numpy()
ids = np.arange(1, N + 1, dtype=np.int32)
eng = Engine(0)
eng.load_vectors(ids, vec)
eng.set_option(_lib.FB_OPT_PROFILE, 1)
g = torch.Generator(); g.manual_seed(99)
rows = torch.randint(0, N, (1000, 3), generator=g).numpy().astype(np.int32)
q = vec[rows[:, 0]].copy()
for name, fn in (("analogy_3cosadd 1000", lambda: eng.analogy_3cosadd(ids[rows])), ("knn_exact 1000 k=1", lambda: eng.knn_exact(q, 1)),
                 ("knn_exact 1000 k=5", lambda: eng.knn_exact(q, 5)), ("knn_exact 1 k=5", lambda: eng.knn_exact(q[:1], 5)),
                 ("knn_exact 128 k=5", lambda: eng.knn_exact(q[:128], 5))):
    for lock in (8, 0):
        eng.set_option(_lib.FB_OPT_PREFILTER_LOCKSTEP, lock)
        out = []
        for rep in range(5):
            eng.reset_counters()
            torch.cuda.synchronize()
            t = time.perf_counter()
            fn()
            dt = time.perf_counter() - t
            c = eng.counters()
            out.append((round(dt * 1e3, 3), round(c["ms_scan"], 3), round(c["ms_finalize"], 3), c["prefilter_candidates"] // max(1, c["prefilter_queries"])))
        print(json.dumps({"call": name, "lockstep": lock, "per_call(e2e_ms, gemm_ms, rescore_ms, cand/query)": out}))
