import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'postgres-word2vec_b200')]
import numpy as np, torch
from freddy_b200.index_build import make_synthetic_index
for sigma, zipf, ncl in ((1.0, 0.35, 1000), (1.0, 0.0, 1000), (0.6, 0.35, 1000), (2.0, 0.35, 1000), (1.0, 0.35, 4000), (1.0, 0.0, 10000)):
    ix = make_synthetic_index(3_000_000, d=300, m=12, K=1024, C=1000, n_train=100000, n_clusters=ncl, sigma=sigma, zipf=zipf, kmeans_iters=10, seed=1234, device='cuda', keep_vectors=False)
    L = np.bincount(ix['coarse_ids'], minlength=1000).astype(np.float64)
    codes = ix['codes']; cid = ix['coarse_ids']
    h = (codes.astype(np.int64) * np.array([1, 1025, 1025**2, 1025**3, 7, 13, 17, 19, 23, 29, 31, 37], dtype=np.int64)).sum(1) * 1009 + cid
    print(f'sigma={sigma} zipf={zipf} ncl={ncl}: list len min {L.min():.0f} med {np.median(L):.0f} max {L.max():.0f}  CV {L.std()/L.mean():.2f}  size-weighted mean {np.sum(L*L)/L.sum():.0f}  dup code rows ~{len(h)-len(np.unique(h))}', flush=True)
