import sys, os, json, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'postgres-word2vec_b200')]
import numpy as np, torch
from freddy_b200 import Engine
from freddy_b200.index_build import make_synthetic_index
from oracle import oracle
ix = make_synthetic_index(3_000_000, d=300, m=12, K=1024, C=1000, n_train=100000, n_clusters=1000, sigma=1.0, zipf=0.35, kmeans_iters=10, seed=1234, device='cuda', keep_vectors=True)
vec = ix.pop('vectors_t')
g = torch.Generator(); g.manual_seed(4321)
q = vec[torch.randperm(3_000_000, generator=g)[:2000].cuda()].cpu().numpy()
e = Engine(0); e.load_ivfadc_index(ix)
for k in (100, 31, 1000):
    e.ivfadc_search(q, k, 10)
    torch.cuda.synchronize(); t0 = time.perf_counter(); ids, d = e.ivfadc_search(q, k, 10); dt = time.perf_counter() - t0
    oi = oracle.OracleIndex(ix)
    eids, ed, rc, _ = oi.ivfadc_search(q[:64], k, 10, threads=64)
    ok = bool((ids[:64] == eids).all() and (d[:64].view(np.uint32) == ed.view(np.uint32)).all())
    print(json.dumps({"config": f"ivfadc_search k={k} w=10, 2000 queries (post-verification candidate fetch), 3M x 300", "seconds_e2e": dt, "queries_per_s": len(q) / dt, "parity_on_64": ok, "exact_path": e.counters()["exact_path_queries"]}))
    e.reset_counters()
