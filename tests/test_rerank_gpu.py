"""GPU parity of exact cosine k-NN and post-verification against the oracle's
restatement of the SQL functions: ids identical, similarities bit-identical (cosine_similarity_bytea is a
sequential float4 chain, core_functions.c:67-81)."""
import numpy as np
import pytest

from helpers import queries_from, small_index

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from freddy_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def _same(ids, s, eids, es, what):
    bad = np.nonzero((ids != eids).any(axis=1))[0]
    assert bad.size == 0, f"{what}: {bad.size} queries differ, first {bad[:3]}: got {ids[bad[:2]]} exp {eids[bad[:2]]}"
    np.testing.assert_array_equal(s.view(np.uint32), es.view(np.uint32), err_msg=what)


@pytest.mark.parametrize("k", [1, 5, 32])
def test_knn_exact_full_scan(eng, oracle_mod, k):
    ix = small_index(N=20000, d=48, m=12, K=64, C=40, seed=7)
    vec_ids = np.asarray(ix["ids"], np.int32)
    eng.load_vectors(vec_ids, ix["vectors"])
    q = queries_from(ix, 70, seed=3, noise=0.05)
    ids, s = eng.knn_exact(q, k)
    eids, es = oracle_mod.knn_exact(ix["vectors"], vec_ids, q, k)
    _same(ids, s, eids, es, f"knn_exact k={k}")


def test_knn_exact_d300_and_duplicates(eng, oracle_mod):
    """d=300 rows, a quarter of them exact duplicates: equal similarities are ordered by table row"""
    ix = small_index(N=6000, d=300, m=12, K=256, C=100, seed=4, n_clusters=3)
    v = ix["vectors"].copy()
    v[1500:3000] = v[:1500]
    vec_ids = np.asarray(ix["ids"], np.int32)
    eng.load_vectors(vec_ids, v)
    rng = np.random.default_rng(5)
    q = v[rng.choice(len(v), 40, replace=False)]
    ids, s = eng.knn_exact(q, 6)
    eids, es = oracle_mod.knn_exact(v, vec_ids, q, 6)
    _same(ids, s, eids, es, "knn_exact duplicates")


def test_knn_in_exact_subset(eng, oracle_mod):
    ix = small_index(N=20000, d=48, m=12, K=64, C=40, seed=7)
    vec_ids = np.asarray(ix["ids"], np.int32)[::-1].copy()        # unsorted ids: id -> row through the lookup table
    eng.load_vectors(vec_ids, ix["vectors"])
    q = queries_from(ix, 33, seed=8)
    rng = np.random.default_rng(1)
    targets = rng.choice(np.arange(1, ix["N"] + 300), size=2500, replace=True).astype(np.int32)   # duplicates + unknown ids
    ids, s = eng.knn_exact(q, 7, targets)
    eids, es = oracle_mod.knn_exact(ix["vectors"], vec_ids, q, 7, targets)
    _same(ids, s, eids, es, "knn_in_exact")
    few = targets[:3]
    ids, s = eng.knn_exact(q, 7, few)                              # fewer rows than k: padded with id -1
    eids, es = oracle_mod.knn_exact(ix["vectors"], vec_ids, q, 7, few)
    _same(ids, s, eids, es, "knn_in_exact fewer than k")
    assert (ids[:, 3:] == -1).all()


@pytest.mark.parametrize("k,pvf,w", [(5, 20, 4), (3, 4, 2), (10, 50, 6)])
def test_ivfadc_search_pv(eng, oracle_mod, k, pvf, w):
    ix = small_index(N=20000, d=48, m=12, K=64, C=40, seed=7)
    vec_ids = np.asarray(ix["ids"], np.int32)
    eng.load_ivfadc_index(ix)
    eng.load_vectors(vec_ids, ix["vectors"])
    q = queries_from(ix, 60, seed=6, noise=0.03)
    ids, s = eng.ivfadc_search_pv(q, k, pvf, w)
    eids, es = oracle_mod.ivfadc_search_pv(oracle_mod.OracleIndex(ix), ix["vectors"], vec_ids, q, k, pvf, w)
    _same(ids, s, eids, es, f"ivfadc_pv k={k} pvf={pvf} w={w}")


@pytest.mark.parametrize("k,pvf", [(5, 20), (3, 4)])
def test_pq_search_pv(eng, oracle_mod, k, pvf):
    """k_nearest_neighbour_pq_pv: flat-PQ candidates (pq_search, sentinel 100.0) re-ranked by the exact cosine"""
    ix = small_index(N=20000, d=48, m=12, K=64, C=40, seed=7, with_pq=True)
    vec_ids = np.asarray(ix["ids"], np.int32)
    eng.load_pq_index(ix)
    eng.load_vectors(vec_ids, ix["vectors"])
    q = queries_from(ix, 40, seed=6, noise=0.03)
    ids, s = eng.pq_search_pv(q, k, pvf)
    eids, es = oracle_mod.pq_search_pv(oracle_mod.OracleIndex(ix, flat_pq=True), ix["vectors"], vec_ids, q, k, pvf)
    _same(ids, s, eids, es, f"pq_pv k={k} pvf={pvf}")


def test_pv_readme_shape_and_udf_mirror(eng, oracle_mod):
    from freddy_b200.udf import Session, vec_to_bytea
    ix = small_index(N=60000, d=300, m=12, K=1024, C=100, seed=1, n_clusters=100)
    vec_ids = np.asarray(ix["ids"], np.int32)
    eng.load_ivfadc_index(ix)
    eng.load_vectors(vec_ids, ix["vectors"])
    q = queries_from(ix, 600, seed=2, noise=0.02)                  # >= 512 queries: candidates through the large-k path
    ids, s = eng.ivfadc_search_pv(q, 5, 20, 10)
    eids, es = oracle_mod.ivfadc_search_pv(oracle_mod.OracleIndex(ix), ix["vectors"], vec_ids, q, 5, 20, 10, threads=8)
    _same(ids, s, eids, es, "ivfadc_pv d=300")
    udf = Session(engine=eng)
    udf.set_w(10)
    rows = udf.k_nearest_neighbour_ivfadc_pv(vec_to_bytea(q[0]), 5)
    assert [r[0] for r in rows] == [int(i) for i in eids[0] if i >= 0]
    rows = udf.k_nearest_neighbour(vec_to_bytea(q[1]), 4)
    e2, _ = oracle_mod.knn_exact(ix["vectors"], vec_ids, q[1:2], 4)
    assert [r[0] for r in rows] == [int(i) for i in e2[0]]


def test_edge_cases(eng, oracle_mod):
    from freddy_b200 import FreddyError, _lib
    ix = small_index(N=20000, d=48, m=12, K=64, C=40, seed=7, with_pq=True)
    vec_ids = np.asarray(ix["ids"], np.int32)
    eng.load_ivfadc_index(ix)
    eng.load_pq_index(ix)
    eng.load_vectors(vec_ids, ix["vectors"])
    empty = np.zeros((0, ix["d"]), np.float32)
    assert eng.knn_exact(empty, 5)[0].shape == (0, 5)
    assert eng.ivfadc_search_pv(empty, 5, 20, 3)[0].shape == (0, 5)
    q = queries_from(ix, 3, seed=1)
    with pytest.raises(FreddyError) as ei:
        eng.knn_exact(q, 33)                                       # the exact scan keeps at most 32 keys per query
    assert ei.value.code == _lib.FB_ERR_UNSUPPORTED
    with pytest.raises(FreddyError) as ei:
        eng.ivfadc_search_pv(q, 60, 20, 3)                         # pvf * k beyond the candidate buffer
    assert ei.value.code == _lib.FB_ERR_UNSUPPORTED
    with pytest.raises(FreddyError):
        eng.ivfadc_search_pv(q, 5, 0, 3)
    ids, s = eng.knn_exact(q, 4, np.asarray([10 ** 8], np.int32))  # no target row exists: nothing but padding
    assert (ids == -1).all() and (s == 0).all()
    cids, codes = eng.encode_ivfadc(empty)
    assert cids.shape == (0,) and codes.shape == (0, 12)
    gi, gg = eng.grouping_pq(np.zeros(0, np.int32), np.asarray([5], np.int32))
    assert len(gi) == 0 and len(gg) == 0
