import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'postgres-word2vec_b200')]
import numpy as np, torch
from freddy_b200.index_build import make_synthetic_index
ix = make_synthetic_index(3_000_000, d=300, m=12, K=1024, C=1000, n_train=100000, n_clusters=1000, kmeans_iters=10, seed=1234, device='cuda', keep_vectors=True)
vec = ix.pop('vectors_t')
print('chunk0 vs chunk1 identical rows:', int((vec[:1000] == vec[262144:262144+1000]).all(1).sum()))
codes = ix['codes']; cid = ix['coarse_ids']
full = np.concatenate([codes.astype(np.int32), cid[:, None]], axis=1)
u, cnt = np.unique(full, axis=0, return_counts=True)
print('unique code rows', len(u), 'of', len(codes), 'max multiplicity', cnt.max())
print('list len min/median/max', np.bincount(cid, minlength=1000).min(), np.median(np.bincount(cid, minlength=1000)), np.bincount(cid, minlength=1000).max())
for p in range(12): print('pos', p, 'distinct codes', len(np.unique(codes[:, p])), 'top code share', np.bincount(codes[:, p]).max() / len(codes))
cb = ix['residual_codebook']
print('codebook norms pos0 (first 8):', np.linalg.norm(cb[0], axis=1)[:8])
co = ix['coarse']
print('coarse pairwise dup:', len(np.unique(co.round(6), axis=0)))
