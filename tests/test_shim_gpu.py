"""The drop-in exercised at the reference's own plugin boundary: the Postgres-side shim
(postgres-word2vec_b200/shim/freddy_shim.c), compiled against the stub Postgres headers and the
in-memory SPI emulator, is called through the fmgr / value-per-call SRF protocol exactly like the
reference's SRFs in test_oracle_vs_reference_srf.py — but its bodies run on the GPU through the
C-ABI.  Results (ids, fp32 distances, "%f" text) must equal the oracle's."""
import os

import numpy as np
import pytest

from helpers import queries_from, small_index

pytestmark = pytest.mark.gpu


def _same(a, b):
    np.testing.assert_array_equal(np.asarray(a).view(np.uint32), np.asarray(b).view(np.uint32))


@pytest.fixture()
def shim(oracle_mod):
    if not os.path.exists(oracle_mod.SHIM_SO):
        pytest.skip("oracle/_ref/libfreddy_shim_emul.so not built")
    return lambda: oracle_mod.ReferenceSession(lib_path=oracle_mod.SHIM_SO)


def test_shim_ivfadc_search(shim, oracle_mod):
    ix = small_index()
    q = queries_from(ix, 25, seed=21, noise=0.02)
    for k, w in ((5, 4), (10, 3), (40, 2)):
        s = shim()
        s.load_ivfadc(ix, w)
        ids, raw, txt = s.ivfadc_search(q, k)
        oids, od, rc, _ = oracle_mod.OracleIndex(ix).ivfadc_search(q, k, w)
        assert rc == 0
        np.testing.assert_array_equal(ids, oids)
        _same(raw, od)
        L = oracle_mod.lib()
        _same(txt.ravel(), np.array([L.fo_round_through_text(float(x)) for x in od.ravel()], np.float32))


def test_shim_flat_pq_and_batch(shim, oracle_mod):
    ix = small_index(N=20000, d=48, m=12, K=64, C=40, seed=7, with_pq=True)
    q = queries_from(ix, 10, seed=13)
    s = shim()
    s.load_pq(ix)
    oi = oracle_mod.OracleIndex(ix, flat_pq=True)
    ids, raw = s.pq_search(q[:3], 5)
    oids, od = oi.pq_search(q[:3], 5)
    np.testing.assert_array_equal(ids, oids)
    _same(raw, od)
    rng = np.random.default_rng(0)
    targets = rng.choice(np.arange(1, ix["N"] + 500), size=2000, replace=True).astype(np.int32)
    ids, raw = s.pq_search_in(q[0], 6, targets)
    oids, od = oi.pq_search_in_batch(q[:1], 6, targets)
    np.testing.assert_array_equal(ids, oids[0])
    _same(raw, od[0])
    qids = np.arange(100, 100 + len(q), dtype=np.int32)
    rq, ids, raw = s.pq_search_in_batch(q, qids, 6, targets, True)
    oids, od = oi.pq_search_in_batch(q, 6, targets)
    np.testing.assert_array_equal(rq, np.repeat(qids[:, None], 6, 1))
    np.testing.assert_array_equal(ids, oids)
    _same(raw, od)
    # ivfadc_batch_search: vectors fetched by id from the normalized table
    s = shim()
    s.load_ivfadc(ix, 3)
    vec_ids = np.asarray(ix["ids"], np.int32)
    s.load_vectors_table(ix["vectors"], vec_ids)
    want = rng.choice(vec_ids, 15, replace=False).astype(np.int32)
    rq, ids, raw = s.ivfadc_batch_search(want, 5)
    order = np.sort(want)
    np.testing.assert_array_equal(rq, order)
    oids, od, rc, _ = oracle_mod.OracleIndex(ix).ivfadc_search(ix["vectors"][order - 1], 5, 1)
    np.testing.assert_array_equal(ids, oids)
    _same(raw, np.where(oids == -1, np.float32(100.0), od))


def test_shim_knn_join(shim, oracle_mod):
    from test_oracle_vs_reference_srf import _ivpq_setup
    ivpq, vec, vec_ids, targets, q = _ivpq_setup()
    oi = oracle_mod.OracleIvpq(ivpq, vec, vec_ids)
    qids = np.arange(500, 500 + len(q), dtype=np.int32)
    for method in (0, 1, 2):
        s = shim()
        s.load_ivpq(ivpq, vec, vec_ids)
        rq, ids, raw = s.ivpq_search_in(q, qids, 5, targets, 3, 4, method, True, 0.8, 10_000_000)
        oids, od, rc, _ = oi.search_in(q, 5, targets, 3, 4, method, True, 0.8)
        assert rc == 0
        np.testing.assert_array_equal(rq, np.repeat(qids[:, None], 5, 1))
        np.testing.assert_array_equal(ids, oids)
        _same(raw, od)


def test_shim_grouping_pq(shim, oracle_mod):
    """the seventh SRF behind the shim: grouping_pq(int[], int[]) through the fmgr / SRF protocol"""
    ix = small_index(N=20000, d=48, m=12, K=64, C=40, seed=7, with_pq=True)
    vec_ids = np.asarray(ix["ids"], np.int32)
    s = shim()
    s.load_pq(ix)
    s.load_vectors_table(ix["vectors"], vec_ids)
    rng = np.random.default_rng(5)
    ids = rng.choice(np.arange(1, ix["N"] + 100), size=1500, replace=True).astype(np.int32)
    groups = np.asarray([77, 5, 1234, 19000], np.int32)
    got_i, got_g = s.grouping_pq(ids, groups)
    want_i, want_g, rc = oracle_mod.OracleIndex(ix, flat_pq=True).grouping_pq(ix["vectors"], vec_ids, ids, groups)
    assert rc == len(want_i)
    np.testing.assert_array_equal(got_i, want_i)
    np.testing.assert_array_equal(got_g, want_g)
    with pytest.raises(RuntimeError):
        s.grouping_pq(ids, np.asarray([5, 10 ** 8], np.int32))
