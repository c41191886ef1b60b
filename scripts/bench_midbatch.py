#!/usr/bin/env python
"""Latency of fb_ivfadc_search for mid-sized batches (what the sidecar launches): n = 1..512 queries per call, k=5, w=10, 3M rows,
for several settings of FB_OPT_QSCAN_MIN_QUERIES (from which batch size on a query gets ONE CTA for all its lists instead of
one CTA per (query, list))."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "postgres-word2vec_b200")]


def main():
    import torch
    from freddy_b200 import Engine, _lib
    from freddy_b200.index_build import make_synthetic_index
    ix = make_synthetic_index(3_000_000, d=300, m=12, K=1024, C=1000, n_train=100_000, n_clusters=1000, sigma=1.0, zipf=0.35,
                              kmeans_iters=10, seed=1234, device="cuda", keep_vectors=True)
    vec = ix.pop("vectors_t")
    g = torch.Generator(); g.manual_seed(4321)
    q_all = vec[torch.randperm(3_000_000, generator=g)[:4096].cuda()].cpu().numpy()
    del vec
    torch.cuda.empty_cache()
    eng = Engine(0)
    eng.load_ivfadc_index(ix)
    out = {}
    for qmin in [int(x) for x in (sys.argv[1] if len(sys.argv) > 1 else "64,160,320,1000000").split(",")]:
        eng.set_option(_lib.FB_OPT_QSCAN_MIN_QUERIES, qmin)
        row = {}
        for n in (1, 8, 16, 24, 32, 48, 64, 96, 128, 192, 256, 384, 512):
            ts = []
            for rep in range(24):
                q = np.ascontiguousarray(q_all[(rep * 131) % 3000:(rep * 131) % 3000 + n])
                t = time.perf_counter(); eng.ivfadc_search(q, 5, 10); ts.append(time.perf_counter() - t)
            row[n] = round(float(np.median(ts[4:])) * 1e6, 1)
        out[qmin] = row
        print(json.dumps({"qscan_min": qmin, "latency_us_by_batch": row}), flush=True)
    eng.close()


if __name__ == "__main__":
    main()
