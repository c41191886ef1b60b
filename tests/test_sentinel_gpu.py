"""Admission against the initial sentinel: the reference admits a row only if
`distance < maxDist`, maxDist starting at the SRF's sentinel (100.0 for pq_search / ivfadc_batch_search,
1000.0 for ivfadc_search / pq_search_in*; freddy.c:90-92, :369, :823-827), so rows at or beyond the sentinel
are never returned: their slots stay (-1, sentinel).  The streaming kernels must do the same."""
import os

import numpy as np
import pytest

from helpers import assert_same_topk, queries_from, small_index


def _scaled(base, s, residual_scale=1.0):
    ix = dict(base)
    for key in ("vectors", "pq_codebook", "coarse"):
        ix[key] = (base[key] * s).astype(np.float32)
    ix["residual_codebook"] = (base["residual_codebook"] * s * residual_scale).astype(np.float32)
    return ix


def _tiny():
    return small_index(N=150, d=24, m=12, K=8, C=64, seed=5, n_clusters=20, with_pq=True)


def _real_batch_search(oracle_mod, ix, order, k):
    """the reference's own ivfadc_batch_search SRF (oracle/_ref) on this index"""
    vec_ids = np.asarray(ix["ids"], np.int32)
    rs = oracle_mod.ReferenceSession()
    rs.load_ivfadc(ix, 1)
    rs.load_vectors_table(ix["vectors"], vec_ids)
    r = rs.ivfadc_batch_search(np.asarray(order, np.int32), k)
    return (np.asarray(r[0], np.int32), np.asarray(r[1], np.int32).reshape(-1, k),
            np.asarray(r[2], np.float32).reshape(-1, k))


def test_reference_batch_search_counts_admissions(oracle_mod):
    """CPU: ivfadc_batch_search counts ADMISSIONS (`distance < maxDist`, maxDist from 100.0), not rows: where the
    w = 1 search with the 1000.0 sentinel would return rows at >= 100, the batch SRF never returns them and goes
    on probing lists until k rows below 100 were admitted (freddy.c:966-981)"""
    if not os.path.exists(oracle_mod.REF_SO):
        pytest.skip("oracle/_ref not built")
    ix = _scaled(_tiny(), 5.0, 6.0)
    order = np.asarray(ix["ids"], np.int32)[:40]
    eids, ed, rc, _ = oracle_mod.OracleIndex(ix).ivfadc_search(ix["vectors"][order - 1], 12, 1)
    assert rc == 0
    assert (ed >= np.float32(100.0)).any(), "the fixture must reach beyond the sentinel"
    oq, rids, rd = _real_batch_search(oracle_mod, ix, order, 12)
    np.testing.assert_array_equal(oq, order)
    assert (rd < np.float32(100.0)).all() and (rids >= 0).all()
    clipped = (ed >= np.float32(100.0)).any(axis=1)
    same = (rids == eids).all(axis=1)
    assert same[~clipped].all() and not same[clipped].any()


@pytest.fixture(scope="module")
def eng():
    from freddy_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.gpu
@pytest.mark.parametrize("force_exact", [0, 1])
def test_pq_search_sentinel_100(eng, oracle_mod, force_exact):
    """pq_search on vectors of norm 9: most cross-cluster ADC distances are >= 100 and must not be returned"""
    from freddy_b200 import _lib
    ix = _scaled(_tiny(), 9.0)
    eng.load_pq_index(ix)
    oi = oracle_mod.OracleIndex(ix, flat_pq=True)
    q = queries_from(ix, 40, seed=13)
    eids, ed = oi.pq_search(q, 30)
    assert (eids == -1).any() and (eids >= 0).any()
    eng.set_option(_lib.FB_OPT_FORCE_EXACT_PATH, force_exact)
    try:
        ids, d = eng.pq_search(q, 30)
    finally:
        eng.set_option(_lib.FB_OPT_FORCE_EXACT_PATH, 0)
    assert_same_topk(ids, d, eids, ed, "pq_search beyond the sentinel")
    # pq_search_in_batch has the 1000.0 sentinel: the same rows are all admitted there
    targets = np.asarray(ix["ids"], np.int32)
    ids, d = eng.pq_search_in_batch(q, 30, targets)
    eids, ed = oi.pq_search_in_batch(q, 30, targets)
    assert (eids >= 0).all()
    assert_same_topk(ids, d, eids, ed, "pq_search_in_batch, sentinel 1000")


@pytest.mark.gpu
@pytest.mark.parametrize("qscan_min", [0, 1 << 30])
def test_ivfadc_batch_search_sentinel_100(eng, oracle_mod, qscan_min):
    from freddy_b200 import _lib
    if not os.path.exists(oracle_mod.REF_SO):
        pytest.skip("oracle/_ref not built")
    ix = _scaled(_tiny(), 5.0, 6.0)
    vec_ids = np.asarray(ix["ids"], np.int32)
    eng.load_ivfadc_index(ix)
    eng.load_vectors(vec_ids, ix["vectors"])
    order = vec_ids[:40]
    eq, eids, ed = _real_batch_search(oracle_mod, ix, order, 12)
    eng.set_option(_lib.FB_OPT_QSCAN_MIN_QUERIES, qscan_min)
    try:
        oq, ids, d = eng.ivfadc_batch_search(order, 12)
    finally:
        eng.set_option(_lib.FB_OPT_QSCAN_MIN_QUERIES, 64)
    np.testing.assert_array_equal(oq, eq)
    assert_same_topk(ids, d, eids, ed, "ivfadc_batch_search beyond the sentinel")


@pytest.mark.gpu
def test_pipeline_kernel_sentinel(eng, oracle_mod):
    """the warp-specialised pipeline kernel (>= 512 queries, d=300, m=12) with sentinel 100.0 through
    ivfadc_batch_search, against the reference's own SRF"""
    if not os.path.exists(oracle_mod.REF_SO):
        pytest.skip("oracle/_ref not built")
    base = small_index(N=3000, d=300, m=12, K=256, C=40, seed=5, n_clusters=50)
    ix = dict(base)
    for key in ("vectors", "coarse"):
        ix[key] = (base[key] * 4.0).astype(np.float32)
    ix["residual_codebook"] = (base["residual_codebook"] * 4.0 * 5.0).astype(np.float32)
    vec_ids = np.asarray(ix["ids"], np.int32)
    eng.load_ivfadc_index(ix)
    eng.load_vectors(vec_ids, ix["vectors"])
    order = vec_ids[:700]
    eq, eids, ed = _real_batch_search(oracle_mod, ix, order, 8)
    eng.reset_counters()
    oq, ids, d = eng.ivfadc_batch_search(order, 8)
    c = eng.counters()
    assert c["n_pipe_launches"] >= 2
    assert c["exact_path_queries"] > 0, "some queries must need further rounds"
    np.testing.assert_array_equal(oq, eq)
    assert_same_topk(ids, d, eids, ed, "pipeline kernel beyond the sentinel")
