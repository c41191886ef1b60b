"""Tensor-core pre-filter of the exact scans (csrc/prefilter_kernels.cuh): the bf16 tcgen05 GEMM only selects
candidates, the reference's fp32 chain decides — so fb_knn_exact / the analogy scans must return the oracle's ids
and similarity bits with the pre-filter on, exactly as with it off, including when its candidate buffer overflows."""
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


def _vectors(N, d, seed, scale_rows=False):
    rng = np.random.default_rng(seed)
    centres = rng.standard_normal((40, d)).astype(np.float32)
    v = centres[rng.integers(0, 40, N)] + 0.7 * rng.standard_normal((N, d)).astype(np.float32)
    v /= np.linalg.norm(v, axis=1, keepdims=True)
    if scale_rows:   # un-normalised table: the band follows the largest row norm
        v *= rng.uniform(0.2, 3.0, (N, 1)).astype(np.float32)
    return np.ascontiguousarray(v, np.float32)


@pytest.mark.parametrize("d,N,k,scale", [(300, 30000, 5, False), (320, 9000, 1, False), (64, 20000, 32, False),
                                         (16, 5000, 3, True), (17, 4000, 7, True), (100, 70000, 10, False)])
def test_prefilter_equals_fp32_scan_and_oracle(eng, oracle_mod, d, N, k, scale):
    from freddy_b200 import _lib
    v = _vectors(N, d, seed=d + k, scale_rows=scale)
    ids_col = np.arange(1, N + 1, dtype=np.int32)
    eng.load_vectors(ids_col, v)
    rng = np.random.default_rng(3)
    q = v[rng.choice(N, 150, replace=False)] + 0.05 * rng.standard_normal((150, d)).astype(np.float32)
    q = np.ascontiguousarray(q, np.float32)
    eng.set_option(_lib.FB_OPT_PREFILTER, 1)
    eng.reset_counters()
    ids1, s1 = eng.knn_exact(q, k)
    c = eng.counters()
    assert c["prefilter_queries"] == 150, "the tensor-core pre-filter did not run"
    eng.set_option(_lib.FB_OPT_PREFILTER, 0)
    try:
        ids0, s0 = eng.knn_exact(q, k)
    finally:
        eng.set_option(_lib.FB_OPT_PREFILTER, 1)
    _same(ids1, s1, ids0, s0, f"pre-filter vs fp32 scan d={d} k={k}")
    n_chk = 24
    eids, es = oracle_mod.knn_exact(v, ids_col, q[:n_chk], k)
    _same(ids1[:n_chk], s1[:n_chk], eids, es, f"pre-filter vs oracle d={d} k={k}")


def test_prefilter_overflow_falls_back(eng, oracle_mod):
    """6000 identical rows tie at the top of every query: far more candidates than the buffer holds, so the
    queries are handed to the fp32 scan — same result (ties ordered by table row)"""
    N, d = 12000, 300
    v = _vectors(N, d, seed=9)
    v[3000:9000] = v[100]
    ids_col = np.arange(1, N + 1, dtype=np.int32)
    eng.load_vectors(ids_col, v)
    q = np.ascontiguousarray(np.stack([v[100], v[100] * 0.5 + v[7] * 0.5, v[11000]]), np.float32)
    eng.reset_counters()
    ids, s = eng.knn_exact(q, 6)
    c = eng.counters()
    assert c["prefilter_queries"] == 3 and c["prefilter_overflow_queries"] >= 1
    eids, es = oracle_mod.knn_exact(v, ids_col, q, 6)
    _same(ids, s, eids, es, "overflow -> fp32 scan")


def test_analogy_through_prefilter(eng, oracle_mod):
    """analogy_3cosadd: arg-max with the three input rows excluded (k' = 4 in the pre-filter)"""
    from freddy_b200 import _lib
    N, d = 40000, 300
    v = _vectors(N, d, seed=21)
    v[500] = v[17]                     # an exact duplicate of a likely winner: the first table row wins
    ids_col = np.arange(1, N + 1, dtype=np.int32)[::-1].copy()      # unsorted ids
    eng.load_vectors(ids_col, v)
    rng = np.random.default_rng(5)
    rows = rng.integers(0, N, (200, 3)).astype(np.int32)
    rows[0] = (3, 17, 17)              # query = v17 - v3 + v17: rows 17 excluded, its duplicate 500 eligible
    eng.reset_counters()
    got_ids, got_s = eng.analogy_3cosadd(ids_col[rows])
    assert eng.counters()["prefilter_queries"] == 200
    erows, es = oracle_mod.analogy_3cosadd(v, rows, threads=4)
    np.testing.assert_array_equal(got_ids, ids_col[erows])
    np.testing.assert_array_equal(got_s.view(np.uint32), es.view(np.uint32))
    eng.set_option(_lib.FB_OPT_PREFILTER, 0)
    try:
        ids0, s0 = eng.analogy_3cosadd(ids_col[rows])
    finally:
        eng.set_option(_lib.FB_OPT_PREFILTER, 1)
    np.testing.assert_array_equal(got_ids, ids0)
    np.testing.assert_array_equal(got_s.view(np.uint32), s0.view(np.uint32))


def test_prefilter_not_used_outside_its_shapes(eng, oracle_mod):
    """d > 320 does not fit the resident query tile: the fp32 scan answers, the result is the same contract"""
    N, d = 3000, 384
    v = _vectors(N, d, seed=2)
    ids_col = np.arange(1, N + 1, dtype=np.int32)
    eng.load_vectors(ids_col, v)
    q = v[:20].copy()
    eng.reset_counters()
    ids, s = eng.knn_exact(q, 4)
    assert eng.counters()["prefilter_queries"] == 0
    eids, es = oracle_mod.knn_exact(v, ids_col, q, 4)
    _same(ids, s, eids, es, "d=384")


def test_single_query_and_many_queries(eng, oracle_mod):
    """one query tile split over all SMs (nq = 1) and more query tiles than the table has slabs (nq = 3000)"""
    N, d = 150000, 300
    v = _vectors(N, d, seed=33)
    ids_col = np.arange(1, N + 1, dtype=np.int32)
    eng.load_vectors(ids_col, v)
    q1 = v[77:78] + np.float32(0.01)
    ids, s = eng.knn_exact(q1, 5)
    eids, es = oracle_mod.knn_exact(v, ids_col, q1, 5)
    _same(ids, s, eids, es, "single query")
    rng = np.random.default_rng(1)
    q = np.ascontiguousarray(v[rng.choice(N, 3000, replace=False)] * np.float32(1.3), np.float32)
    eng.reset_counters()
    ids, s = eng.knn_exact(q, 3)
    assert eng.counters()["prefilter_queries"] == 3000
    sel = rng.choice(3000, 16, replace=False)
    eids, es = oracle_mod.knn_exact(v, ids_col, q[sel], 3)
    _same(ids[sel], s[sel], eids, es, "3000 queries")
