import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'postgres-word2vec_b200')]
import numpy as np, torch
from freddy_b200 import Engine
from freddy_b200.index_build import make_synthetic_index
from oracle import oracle
ix = make_synthetic_index(3_000_000, d=300, m=12, K=1024, C=1000, n_train=100000, n_clusters=1000, sigma=1.0, zipf=0.35, kmeans_iters=10, seed=1234, device='cuda', keep_vectors=True)
vec = ix.pop('vectors_t')
g = torch.Generator(); g.manual_seed(4321)
sel = torch.randperm(3_000_000, generator=g)[:10000]
q = vec[sel.cuda()].cpu().numpy()
e = Engine(0); e.load_ivfadc_index(ix)
ids, d = e.ivfadc_search(q, 7, 10)
c = e.counters(); print({k: v for k, v in c.items() if k.startswith('exact')})
# how do the ties look: distances of the top-7 for queries whose 5th == 6th
tie = np.nonzero(d[:, 4] == d[:, 5])[0]
print('queries with d5==d6:', len(tie), ' also d6==d7:', int((d[tie, 5] == d[tie, 6]).sum()))
for t in tie[:12]: print(t, d[t], ids[t])
codes = ix['codes']; cid = ix['coarse_ids']
for t in tie[:6]:
    a, b = ids[t, 4] - 1, ids[t, 5] - 1
    print('rows', a, b, 'same list', cid[a] == cid[b], 'same codes', (codes[a] == codes[b]).all())
