#!/usr/bin/env python
"""Pinning the headline index (3M rows, C=1000, m=12, K=1024): host-side builder against the device kernels
(FB_OPT_DEVICE_BUILD), same layout (checksums), seconds per call.  VERDICT r1 item 8: upload <= 0.2 s."""
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
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 3_000_000
    ix = make_synthetic_index(n, d=300, m=12, K=1024, C=1000, n_train=100_000, n_clusters=1000, sigma=1.0, zipf=0.35,
                              kmeans_iters=10, seed=1234, device="cuda", keep_vectors=True, with_pq=True)
    vec = ix.pop("vectors_t").cpu().numpy()
    torch.cuda.empty_cache()
    out = {"rows": n}
    sums = {}
    for dev in (0, 1):
        e = Engine(0)
        e.set_option(_lib.FB_OPT_DEVICE_BUILD, dev)
        t = []
        for rep in range(3):
            t0 = time.perf_counter(); e.load_ivfadc_index(ix); e.synchronize(); t.append(time.perf_counter() - t0)
        t0 = time.perf_counter(); e.load_pq_index(ix); e.synchronize(); t_pq = time.perf_counter() - t0
        t0 = time.perf_counter(); e.load_vectors(ix["ids"], vec); e.synchronize(); t_vec = time.perf_counter() - t0
        sums[dev] = (e.table_checksum(0), e.table_checksum(1))
        out["device_build" if dev else "host_build"] = {"load_ivfadc_index_s": t, "load_pq_index_s": t_pq, "load_vectors_s": t_vec}
        e.close()
    out["same_layout"] = sums[0] == sums[1]
    print(json.dumps(out))


if __name__ == "__main__":
    main()
