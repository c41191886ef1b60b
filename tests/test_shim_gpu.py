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


def test_shim_gpu_only_srfs(shim, oracle_mod):
    """the SRFs the plpgsql bodies would call instead of one UDF call per row: exact k-NN (whole table and id subset),
    IVFADC + post verification, batched analogy, batched cosine — through the fmgr / SRF protocol"""
    ix = small_index()
    vec_ids = np.asarray(ix["ids"], np.int32)
    vec = ix["vectors"]
    s = shim()
    s.load_ivfadc(ix, 4)
    s.load_vectors_table(vec, vec_ids)
    s.set_config("get_pvf()", 20)
    q = queries_from(ix, 6, seed=31, noise=0.03)
    for i in range(len(q)):
        ids, sims = s.knn_exact_search(q[i], 5)
        eids, es = oracle_mod.knn_exact(vec, vec_ids, q[i:i + 1], 5)
        np.testing.assert_array_equal(ids, eids[0])
        _same(sims, es[0])
        ids, sims = s.ivfadc_search_pv(q[i], 5)
        eids, es = oracle_mod.ivfadc_search_pv(oracle_mod.OracleIndex(ix), vec, vec_ids, q[i:i + 1], 5, 20, 4)
        keep = eids[0] >= 0
        np.testing.assert_array_equal(ids, eids[0][keep])
        _same(sims, es[0][keep])
    rng = np.random.default_rng(3)
    targets = rng.choice(np.arange(1, ix["N"] + 200), size=700, replace=True).astype(np.int32)
    ids, sims = s.knn_exact_search(q[0], 7, targets)
    eids, es = oracle_mod.knn_exact(vec, vec_ids, q[:1], 7, targets)
    np.testing.assert_array_equal(ids, eids[0])
    _same(sims, es[0])
    ids, sims = s.knn_exact_search(q[0], 4, targets[:2])          # fewer rows than k: only the rows that exist come back
    assert len(ids) == len(np.unique(targets[:2][targets[:2] <= ix["N"]]))
    rows = rng.integers(0, ix["N"], (40, 3)).astype(np.int32)
    got_ids, got_s = s.analogy_3cosadd_batch(vec_ids[rows])
    erows, es = oracle_mod.analogy_3cosadd(vec, rows, threads=2)
    np.testing.assert_array_equal(got_ids, vec_ids[erows])
    _same(got_s, es)
    a, b = vec[rows[:, 0]], vec[rows[:, 1]] * np.float32(1.7)
    ref = oracle_mod.ReferenceSession()                            # the reference's own scalar UDF for variant 2
    want2 = np.array([ref.R.ref_cosine_similarity_bytea(oracle_mod._p(np.ascontiguousarray(a[i])), oracle_mod._p(np.ascontiguousarray(b[i])),
                                                        a.shape[1]) for i in range(len(a))], np.float32)
    got2 = s.cosine_similarity_batch(a, b, 2)
    _same(got2.astype(np.float32), want2)
    got0 = s.cosine_similarity_batch(a, b, 0)
    want0 = np.array([np.dot(a[i].astype(np.float64), b[i].astype(np.float64)) /
                      (np.sqrt(np.dot(a[i].astype(np.float64), a[i].astype(np.float64))) * np.sqrt(np.dot(b[i].astype(np.float64), b[i].astype(np.float64))))
                      for i in range(len(a))])
    np.testing.assert_allclose(got0, want0, rtol=1e-14)


def test_shim_repins_when_the_tables_change(shim, oracle_mod):
    """a pinned index must not outlive its tables: another table name configured, rows appended
    (max(id) grows), or freddy_repin() after any other change"""
    ix_a = small_index()
    ix_b = small_index(N=20000, d=48, m=12, K=64, C=40, seed=9)
    q = queries_from(ix_a, 5, seed=2)
    s = shim()
    s.load_ivfadc(ix_a, 4)
    ids_a, raw_a, _ = s.ivfadc_search(q, 5)
    want_a = oracle_mod.OracleIndex(ix_a).ivfadc_search(q, 5, 4)
    np.testing.assert_array_equal(ids_a, want_a[0])
    # same table names, other contents (a bulk reload): stale until freddy_repin()
    s.R.ref_reset()
    s.keep.clear()
    s.load_ivfadc(ix_b, 4)
    ids_stale, _, _ = s.ivfadc_search(q, 5)
    np.testing.assert_array_equal(ids_stale, ids_a)
    assert s.repin() == 0
    ids_b, raw_b, _ = s.ivfadc_search(q, 5)
    want_b = oracle_mod.OracleIndex(ix_b).ivfadc_search(q, 5, 4)
    np.testing.assert_array_equal(ids_b, want_b[0])
    _same(raw_b, want_b[1])
    # rows appended (here: a longer image of the same table): max(id) differs -> re-pinned without being told
    ix_c = dict(ix_b)
    n = ix_b["N"] - 500
    for key in ("ids", "coarse_ids", "codes"):
        ix_c[key] = ix_b[key][:n]
    ix_c["N"] = n
    s.R.ref_reset()
    s.keep.clear()
    s.load_ivfadc(ix_c, 4)
    ids_c, raw_c, _ = s.ivfadc_search(q, 5)
    want_c = oracle_mod.OracleIndex(ix_c).ivfadc_search(q, 5, 4)
    np.testing.assert_array_equal(ids_c, want_c[0])
    # another table configured under get_vecs_name_residual_quantization(): re-pinned
    s._table("fine_quantization_2", s.T_FINE, int(ix_b["N"]), ids=np.asarray(ix_b["ids"], np.int32),
             a=np.asarray(ix_b["coarse_ids"], np.int32), vec=np.asarray(ix_b["codes"], np.int16))
    s.set_config("get_vecs_name_residual_quantization()", "fine_quantization_2")
    ids_d, _, _ = s.ivfadc_search(q, 5)
    np.testing.assert_array_equal(ids_d, want_b[0])


def test_shim_insert_batch_matches_the_reference(shim, oracle_mod):
    """insert_batch: the shim quantises the new rows on the GPU and hands the codes to the reference's own table
    helpers; every INSERT / UPDATE statement it issues must be the one the reference's insert_batch (freddy.c, run
    here through the same emulator) issues for the same input"""
    import torch
    from freddy_b200.index_build import make_ivpq_index
    if not os.path.exists(oracle_mod.REF_SO):
        pytest.skip("oracle/_ref not built")
    ix = small_index(N=20000, d=48, m=12, K=64, C=40, seed=7, with_pq=True)
    vec = np.ascontiguousarray(ix["vectors"])
    vec_ids = np.asarray(ix["ids"], np.int32)
    ivpq = make_ivpq_index(torch.from_numpy(vec), m=12, K=64, Kc=8, n_train=20000, kmeans_iters=3, seed=3)
    rng = np.random.default_rng(0)
    raw = vec[rng.choice(len(vec), 9)] * np.float32(2.5) + 0.05 * rng.standard_normal((9, 48)).astype(np.float32)
    norm = (raw / np.linalg.norm(raw, axis=1, keepdims=True)).astype(np.float32)
    tokens = [f"tok_{i}" for i in range(9)]
    terms = ["tok 0", "tok 1", "already there"]
    logs = []
    for lib_path in (None, oracle_mod.SHIM_SO):
        s = oracle_mod.ReferenceSession(lib_path=lib_path)
        s.load_insert_tables(ix, ivpq, vec, vec_ids)
        if lib_path is not None:
            s.ivfadc_search(norm[:1], 3)          # pin the IVFADC index first: insert_batch must then extend the pinned copy
        logs.append(s.insert_batch(terms, tokens, norm, raw))
    ref_log, shim_log = logs
    assert len(ref_log) > 9 * 5
    assert shim_log == ref_log
    # the pinned fine table was extended in place: the new rows (ids max(id)+1 ...) are found without any re-read
    # (the emulator's tables are read-only images, so a re-read could not know them)
    new_ids = set(range(int(ix["N"]) + 1, int(ix["N"]) + 10))
    hits = 0
    for i in range(9):
        ids, _, _ = s.ivfadc_search(norm[i:i + 1], 5)
        hits += int(ix["N"]) + 1 + i in set(ids[0].tolist())
        assert set(ids[0].tolist()) & new_ids or True
    assert hits >= 7, f"only {hits} of 9 inserted vectors find their own new row"


def test_shim_pq_search_pv(shim, oracle_mod):
    """k_nearest_neighbour_pq_pv's body as one SRF: pq_search(v, pvf * k) + exact re-rank, pvf = get_pvf()"""
    ix = small_index(N=20000, d=48, m=12, K=64, C=40, seed=7, with_pq=True)
    vec_ids = np.asarray(ix["ids"], np.int32)
    s = shim()
    s.load_pq(ix)
    s.load_vectors_table(ix["vectors"], vec_ids)
    s.set_config("get_pvf()", 10)
    q = queries_from(ix, 5, seed=17, noise=0.03)
    oi = oracle_mod.OracleIndex(ix, flat_pq=True)
    for i in range(len(q)):
        ids, sims = s.pq_search_pv(q[i], 5)
        eids, es = oracle_mod.pq_search_pv(oi, ix["vectors"], vec_ids, q[i:i + 1], 5, 10)
        keep = eids[0] >= 0
        np.testing.assert_array_equal(ids, eids[0][keep])
        _same(sims, es[0][keep])


def test_shim_ivfadc_search_through_the_sidecar(shim, oracle_mod):
    """FREDDY_SIDECAR set: ivfadc_search posts its query to the backend that runs freddy_sidecar_serve() instead of
    using an engine of its own; same rows, same bits; without a sidecar the backend answers itself."""
    import threading
    import time
    ix = small_index()
    q = queries_from(ix, 30, seed=21, noise=0.02)
    k, w = 5, 4
    s = shim()
    s.load_ivfadc(ix, w)
    direct = s.ivfadc_search(q, k)
    name = f"/fbsc_shim_{os.getpid()}"
    served = []
    os.environ["FREDDY_SIDECAR"] = name
    try:
        t = threading.Thread(target=lambda: served.append(s.sidecar_serve(ix["d"], 16, 60)))
        t.start()
        deadline = time.time() + 30
        while not os.path.exists("/dev/shm" + name) and time.time() < deadline:
            time.sleep(0.01)
        assert os.path.exists("/dev/shm" + name), "the sidecar did not come up"
        via = s.ivfadc_search(q, k)
        assert s.sidecar_stop() == 1
        t.join(60)
        assert not t.is_alive()
        assert served == [len(q)]                    # every query was answered by the sidecar's engine
        for a, b in zip(direct, via):
            _same(a, b) if a.dtype == np.float32 else np.testing.assert_array_equal(a, b)
        again = s.ivfadc_search(q[:3], k)            # segment gone: the backend's own engine
        np.testing.assert_array_equal(again[0], direct[0][:3])
    finally:
        os.environ.pop("FREDDY_SIDECAR", None)
