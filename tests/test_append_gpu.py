"""In-place append to the pinned tables: rows insert_batch adds (freddy.c:1611-1625) are appended on
the device; searches afterwards must equal a fresh upload of the grown tables and the oracle on them."""
import numpy as np
import pytest

from helpers import assert_same_topk, queries_from, small_index

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from freddy_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


def _prefix(ix, n, keys=("ids", "coarse_ids", "codes", "pq_codes")):
    out = dict(ix)
    for k in keys:
        if k in ix:
            out[k] = ix[k][:n]
    out["N"] = n
    return out


@pytest.mark.parametrize("K,d,nq", [(64, 48, 300), (256, 300, 700), (1024, 300, 700)])
def test_append_fine_equals_fresh_upload(eng, oracle_mod, K, d, nq):
    ix = small_index(N=30000, d=d, m=12, K=K, C=40, seed=11, n_clusters=60)
    n0 = 26000
    eng.load_ivfadc_index(_prefix(ix, n0))
    eng.append_fine(ix["ids"][n0:28000], ix["coarse_ids"][n0:28000], ix["codes"][n0:28000])
    eng.append_fine(ix["ids"][28000:], ix["coarse_ids"][28000:], ix["codes"][28000:])
    q = queries_from(ix, nq, seed=5, noise=0.02)
    ids, dd = eng.ivfadc_search(q, 5, 6)
    eids, ed, rc, rows = oracle_mod.OracleIndex(ix).ivfadc_search(q, 5, 6, threads=8)
    assert rc == 0
    assert_same_topk(ids, dd, eids, ed, f"after append K={K}")
    assert (ids >= n0 + 1).any(), "no appended row among the results: the fixture does not test the append"
    ids1, d1 = eng.ivfadc_search(q[:1], 5, 6)                       # single-query path (graph replay after the buffers moved)
    ids1, d1 = eng.ivfadc_search(q[:1], 5, 6)
    assert_same_topk(ids1, d1, eids[:1], ed[:1], "single query after append")
    eng.load_ivfadc_index(ix)
    ids2, d2 = eng.ivfadc_search(q, 5, 6)
    assert_same_topk(ids2, d2, ids, dd, "fresh upload vs append")


def test_append_pq_and_unsorted_ids(eng, oracle_mod):
    ix = dict(small_index(N=20000, d=48, m=12, K=64, C=40, seed=7, with_pq=True))
    rng = np.random.default_rng(2)
    ids = np.asarray(ix["ids"], np.int32).copy()
    ids[19000:] = rng.permutation(np.arange(50_001, 51_001)).astype(np.int32)      # appended ids out of order: the id index re-sorts
    ix["ids"] = ids
    n0 = 18000
    eng.load_pq_index(_prefix(ix, n0))
    eng.append_pq(ids[n0:19000], ix["pq_codes"][n0:19000])                            # ascending beyond the maximum: appended to the index
    eng.append_pq(ids[19000:], ix["pq_codes"][19000:])
    oi = oracle_mod.OracleIndex(ix, flat_pq=True)
    q = queries_from(ix, 60, seed=3)
    got = eng.pq_search(q, 6)
    exp = oi.pq_search(q, 6)
    assert_same_topk(got[0], got[1], exp[0], exp[1], "pq_search after append")
    targets = np.concatenate([ids[17500:19500], rng.choice(ids, 500)]).astype(np.int32)
    got = eng.pq_search_in_batch(q, 5, targets)
    exp = oi.pq_search_in_batch(q, 5, targets)
    assert_same_topk(got[0], got[1], exp[0], exp[1], "pq_search_in_batch after append")


def test_append_ivpq(eng, oracle_mod):
    from freddy_b200 import _lib
    from test_oracle_vs_reference_srf import _ivpq_setup
    ivpq, vec, vec_ids, targets, q = _ivpq_setup()
    n0 = int(ivpq["N"]) - 1500
    part = dict(ivpq)
    for k in ("ids", "ivpq_coarse_ids", "ivpq_codes"):
        part[k] = ivpq[k][:n0]
    part["N"] = n0
    eng.load_ivpq_index(part)
    eng.load_vectors(vec_ids, vec)
    eng.append_pq(ivpq["ids"][n0:], ivpq["ivpq_codes"][n0:], kind=_lib.FB_CB_IVPQ, cells=ivpq["ivpq_coarse_ids"][n0:])
    oi = oracle_mod.OracleIvpq(ivpq, vec, vec_ids)
    for method in (0, 2):
        ids, dd = eng.ivpq_search_in(q, 5, targets, 3, 4, method, True, 0.8)
        eids, ed, rc, _ = oi.search_in(q, 5, targets, 3, 4, method, True, 0.8)
        assert rc == 0
        assert_same_topk(ids, dd, eids, ed, f"join after append, method {method}")


def test_append_vectors_exact_paths(eng, oracle_mod):
    """the word-vector table grows: exact k-NN (tensor-core pre-filter + fp32 chain), analogy and the id lookup see the new rows"""
    ix = small_index(N=20000, d=48, m=12, K=64, C=40, seed=7)
    v = ix["vectors"]
    ids = np.asarray(ix["ids"], np.int32).copy()
    ids[19000:] = np.arange(70_000, 71_000, dtype=np.int32)[::-1]            # appended ids not ascending: the lookup re-sorts
    n0 = 17000
    eng.load_vectors(ids[:n0], v[:n0])
    eng.append_vectors(ids[n0:19000], v[n0:19000])
    eng.append_vectors(ids[19000:], v[19000:])
    q = queries_from(ix, 50, seed=9, noise=0.03)
    got = eng.knn_exact(q, 6)
    exp = oracle_mod.knn_exact(v, ids, q, 6)
    np.testing.assert_array_equal(got[0], exp[0])
    np.testing.assert_array_equal(got[1].view(np.uint32), exp[1].view(np.uint32))
    assert (np.isin(got[0], ids[n0:])).any()
    rng = np.random.default_rng(4)
    rows = rng.integers(0, len(v), (30, 3)).astype(np.int32)
    rows[:10, 2] = rng.integers(n0, len(v), 10)                             # triples that name appended words
    gi, gs = eng.analogy_3cosadd(ids[rows])
    erows, es = oracle_mod.analogy_3cosadd(v, rows, threads=2)
    np.testing.assert_array_equal(gi, ids[erows])
    np.testing.assert_array_equal(gs.view(np.uint32), es.view(np.uint32))
    targets = np.concatenate([ids[16000:18000], ids[19500:]]).astype(np.int32)
    got = eng.knn_exact(q, 4, targets)
    exp = oracle_mod.knn_exact(v, ids, q, 4, targets)
    np.testing.assert_array_equal(got[0], exp[0])
