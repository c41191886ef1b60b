"""Shared test helpers: small seeded indexes (cached per session)."""
import functools

import numpy as np

from freddy_b200.index_build import make_synthetic_index


@functools.lru_cache(maxsize=None)
def small_index(N=20000, d=48, m=12, K=64, C=40, seed=7, with_pq=False, n_clusters=50):
    ix = make_synthetic_index(N, d=d, m=m, K=K, C=C, n_train=min(N, 20000), n_clusters=n_clusters,
                              kmeans_iters=4, seed=seed, device="cpu", with_pq=with_pq, keep_vectors=True)
    vec = ix.pop("vectors_t").numpy()
    ix["vectors"] = vec
    return ix


def queries_from(ix, n, seed=11, noise=0.0):
    rng = np.random.default_rng(seed)
    sel = rng.choice(ix["N"], size=n, replace=False)
    q = ix["vectors"][sel].copy()
    if noise:
        q += noise * rng.standard_normal(q.shape).astype(np.float32)
    return np.ascontiguousarray(q, np.float32)


def assert_same_topk(got_ids, got_d, exp_ids, exp_d, what=""):
    """ids/ranks bit-exact; distances bit-exact too (same fp32 chain)."""
    bad = np.nonzero((got_ids != exp_ids).any(axis=1))[0]
    assert bad.size == 0, f"{what}: {bad.size} queries differ in ids, first {bad[:5]}: got {got_ids[bad[:2]]} exp {exp_ids[bad[:2]]}"
    np.testing.assert_array_equal(got_d.view(np.uint32), exp_d.view(np.uint32), err_msg=f"{what}: distances differ")


def srf_golden():
    """index + outputs of the reference's own SRFs (tests/golden/make_golden.py)"""
    import os
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "srf_golden.npz"))
    ix = {k: (int(g[k]) if k in ("d", "m", "K", "C", "N") else g[k]) for k in
          ("d", "m", "K", "C", "N", "coarse", "residual_codebook", "ids", "coarse_ids", "codes", "pq_codebook", "pq_codes")}
    return ix, g


def srf_golden_ext():
    """grouping_pq / updateCodebook outputs of the reference's own code on the tiny seeded index
    (tests/golden/make_golden.py::srf_golden_ext)"""
    import os
    return np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "srf_golden_ext.npz"))
