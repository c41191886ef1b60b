import sys, os, time
sys.path[:0]=['/root/repo','/root/repo/postgres-word2vec_b200']
import numpy as np, torch
from freddy_b200 import Engine, _lib
from freddy_b200.index_build import make_synthetic_index
ix = make_synthetic_index(3_000_000, d=300, m=12, K=1024, C=1000, n_train=100_000, n_clusters=1000, sigma=1.0, zipf=0.35, kmeans_iters=10, seed=1234, device="cuda", keep_vectors=True)
vec = ix.pop("vectors_t"); q = vec[12345:12346].cpu().numpy(); del vec
eng = Engine(0); eng.load_ivfadc_index(ix)
eng.set_option(_lib.FB_OPT_CUDA_GRAPHS, int(os.environ.get("GRAPHS","1")))
for _ in range(30): eng.ivfadc_search(q, 5, 10)
ts=[]
for _ in range(300):
    t=time.perf_counter(); eng.ivfadc_search(q, 5, 10); ts.append(time.perf_counter()-t)
print("graphs", os.environ.get("GRAPHS","1"), "median us", np.median(ts)*1e6, "min", min(ts)*1e6)
