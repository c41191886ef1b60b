"""kNN-join (ivpq_search_in) on the GPU vs the oracle (which is pinned to the real SRF in
tests/test_oracle_vs_reference_srf.py)."""
import numpy as np
import pytest

from helpers import assert_same_topk
from test_oracle_vs_reference_srf import _ivpq_setup

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def setup(oracle_mod):
    from freddy_b200 import Engine
    ivpq, vec, vec_ids, targets, q = _ivpq_setup()
    e = Engine(0)
    e.load_ivpq_index(ivpq)
    e.load_vectors(vec_ids, vec)
    oi = oracle_mod.OracleIvpq(ivpq, vec, vec_ids)
    yield e, oi, targets, q
    e.close()


@pytest.mark.parametrize("method", [0, 1, 2])
@pytest.mark.parametrize("use_tl", [False, True])
def test_knn_join_parity(setup, method, use_tl):
    e, oi, targets, q = setup
    for (k, alpha, pvf, conf) in ((5, 3, 4, 0.8), (5, 1, 2, 0.5), (3, 40, 20, 0.8), (5, 100, 20, 0.8)):
        ids, d = e.ivpq_search_in(q, k, targets, alpha, pvf, method, use_tl, conf)
        eids, ed, rc, st = oi.search_in(q, k, targets, alpha, pvf, method, use_tl, conf)
        assert rc == 0
        assert_same_topk(ids, d, eids, ed, f"method={method} tl={use_tl} k={k} alpha={alpha} pvf={pvf}")


@pytest.mark.parametrize("method", [0, 2])
def test_knn_join_retry_loop(setup, method):
    e, oi, targets, q = setup
    rng = np.random.default_rng(1)
    few = np.sort(rng.choice(targets, 60, replace=False)).astype(np.int32)
    seen_retry = False
    for use_tl in (False, True):
        for (k, alpha, pvf, conf) in ((10, 1, 2, 0.5), (8, 2, 3, 0.8), (70, 1, 1, 0.5)):
            ids, d = e.ivpq_search_in(q, k, few, alpha, pvf, method, use_tl, conf)
            eids, ed, rc, st = oi.search_in(q, k, few, alpha, pvf, method, use_tl, conf)
            assert rc == 0
            seen_retry |= st[0] > 1
            assert_same_topk(ids, d, eids, ed, f"retry method={method} tl={use_tl} k={k}")
    assert seen_retry


def test_knn_join_edge_cases(setup):
    from freddy_b200 import FreddyError, _lib
    e, oi, targets, q = setup
    ids, d = e.ivpq_search_in(q[:3], 4, np.zeros(0, np.int32), 3, 2, 0, False, 0.8)      # no targets at all
    eids, ed, rc, _ = oi.search_in(q[:3], 4, np.zeros(0, np.int32), 3, 2, 0, False, 0.8)
    assert_same_topk(ids, d, eids, ed, "empty target set")
    dup = np.concatenate([targets[:50], targets[:50], [10 ** 7]]).astype(np.int32)          # duplicates + unknown id
    ids, d = e.ivpq_search_in(q, 5, dup, 2, 2, 2, True, 0.6)
    eids, ed, rc, _ = oi.search_in(q, 5, dup, 2, 2, 2, True, 0.6)
    assert_same_topk(ids, d, eids, ed, "duplicate / unknown targets")
    assert FreddyError is not None and _lib is not None


@pytest.mark.parametrize("method", [0, 2])
@pytest.mark.parametrize("use_tl", [False, True])
def test_knn_join_pair_lut_variant(setup, method, use_tl):
    """alpha*k > double_threshold: distances are sums of pair sums (getPrecomputedDistancesDouble,
    index_utils.c:457-475); the oracle's form is pinned to the real SRF in
    test_oracle_vs_reference_srf.py::test_ivpq_search_in_pair_lut_variant"""
    e, oi, targets, q = setup
    for (k, alpha, pvf, conf) in ((5, 3, 4, 0.8), (3, 40, 20, 0.8), (5, 100, 20, 0.8)):
        ids, d = e.ivpq_search_in(q, k, targets, alpha, pvf, method, use_tl, conf, double_threshold=0)
        eids, ed, rc, st = oi.search_in(q, k, targets, alpha, pvf, method, use_tl, conf, 0)
        assert rc == 0
        assert_same_topk(ids, d, eids, ed, f"pair-LUT method={method} tl={use_tl} k={k} alpha={alpha}")


@pytest.mark.parametrize("m,K", [(12, 1024), (30, 32)])
def test_knn_join_d300_multi_index_32x32(oracle_mod, m, K):
    """the shapes of index_creation/config/ivpq_config.json (m=30, K=32) and of BASELINE config 4 (m=12, K=1024):
    d=300, 2 x 32 multi-index, alpha=100, pvf=20, method 2 (PQ + post verification) and the other two methods"""
    import torch
    from freddy_b200 import Engine
    from freddy_b200.index_build import make_ivpq_index
    from helpers import small_index
    ix = small_index(N=60000, d=300, m=12, K=1024, C=100, seed=1, n_clusters=100)
    vec = np.ascontiguousarray(ix["vectors"])
    N = len(vec)
    rng = np.random.default_rng(8)
    trows = np.sort(rng.choice(N, 20000, replace=False))
    ivpq = make_ivpq_index(torch.from_numpy(vec), m=m, K=K, Kc=32, n_train=N, kmeans_iters=3, seed=3, target_rows=trows)
    vec_ids = np.asarray(ivpq["ids"], np.int32)
    targets = (trows + 1).astype(np.int32)
    q = np.ascontiguousarray(vec[rng.choice(N, 160, replace=False)] + 0.01 * rng.standard_normal((160, 300)).astype(np.float32))
    e = Engine(0)
    try:
        e.load_ivpq_index(ivpq)
        e.load_vectors(vec_ids, vec)
        oi = oracle_mod.OracleIvpq(ivpq, vec, vec_ids)
        for method, use_tl, k, alpha, pvf in ((2, True, 5, 100, 20), (2, False, 5, 100, 20), (0, True, 5, 10, 1), (1, False, 3, 4, 1)):
            ids, d = e.ivpq_search_in(q, k, targets, alpha, pvf, method, use_tl, 0.8)
            eids, ed, rc, st = oi.search_in(q, k, targets, alpha, pvf, method, use_tl, 0.8)
            assert rc == 0
            assert_same_topk(ids, d, eids, ed, f"d=300 m={m} K={K} method={method} tl={use_tl}")
    finally:
        e.close()
