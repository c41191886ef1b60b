"""Pins the oracle's DRIVER restatements against the reference extension's own SRFs:
freddy.c is compiled unmodified (oracle/Makefile `ref`) and run in-process on
in-memory tables through the SPI/fmgr emulator oracle/pg_emul.c.  ids and raw fp32
distances must agree bit for bit; the SRF's "%f" text output must equal
fo_round_through_text of the oracle's distance."""
import numpy as np
import pytest

from helpers import queries_from, small_index


@pytest.fixture(scope="module")
def ref(oracle_mod):
    if oracle_mod.ref_lib() is None:
        pytest.skip("oracle/_ref not built (reference sources absent and no prebuilt .so)")
    return oracle_mod.ReferenceSession


def _same(a, b):
    np.testing.assert_array_equal(np.asarray(a).view(np.uint32), np.asarray(b).view(np.uint32))


@pytest.mark.parametrize("k,w", [(5, 4), (1, 1), (10, 3), (7, 10)])
def test_ivfadc_search_driver(ref, oracle_mod, k, w):
    ix = small_index()
    q = queries_from(ix, 40, seed=21, noise=0.02)
    s = ref()
    s.load_ivfadc(ix, w)
    rids, rraw, rtxt = s.ivfadc_search(q, k)
    oids, od, rc, _ = oracle_mod.OracleIndex(ix).ivfadc_search(q, k, w)
    assert rc == 0
    np.testing.assert_array_equal(oids, rids)
    _same(od, rraw)
    L = oracle_mod.lib()
    _same(np.array([L.fo_round_through_text(float(x)) for x in od.ravel()], np.float32), rtxt.ravel())


def test_ivfadc_search_ties_and_reprobe(ref, oracle_mod):
    # duplicate-heavy table: ties straddle the k-th place constantly
    ix = small_index(N=6000, d=8, m=2, K=4, C=8, seed=3, n_clusters=5)
    q = queries_from(ix, 60, seed=9)
    for k, w in ((5, 2), (3, 8), (20, 3)):
        s = ref()
        s.load_ivfadc(ix, w)
        rids, rraw, _ = s.ivfadc_search(q, k)
        oids, od, rc, _ = oracle_mod.OracleIndex(ix).ivfadc_search(q, k, w)
        assert rc == 0
        np.testing.assert_array_equal(oids, rids)
        _same(od, rraw)
    # lists shorter than k: the re-probe loop with its blacklist (freddy.c:262-293)
    ix = small_index(N=150, d=24, m=12, K=8, C=64, seed=5, n_clusters=20)
    q = queries_from(ix, 30, seed=2)
    s = ref()
    s.load_ivfadc(ix, 2)
    rids, rraw, _ = s.ivfadc_search(q, 12)
    oids, od, rc, _ = oracle_mod.OracleIndex(ix).ivfadc_search(q, 12, 2)
    assert rc == 0
    np.testing.assert_array_equal(oids, rids)
    _same(od, rraw)


def test_flat_pq_drivers(ref, oracle_mod):
    ix = small_index(N=20000, d=48, m=12, K=64, C=40, seed=7, with_pq=True)
    q = queries_from(ix, 12, seed=13)
    s = ref()
    s.load_pq(ix)
    oi = oracle_mod.OracleIndex(ix, flat_pq=True)
    rids, rraw = s.pq_search(q[:4], 5)
    oids, od = oi.pq_search(q[:4], 5)
    np.testing.assert_array_equal(oids, rids)
    _same(od, rraw)
    rng = np.random.default_rng(0)
    targets = rng.choice(np.arange(1, ix["N"] + 500), size=2000, replace=True).astype(np.int32)
    rids, rraw = s.pq_search_in(q[0], 6, targets)
    oids, od = oi.pq_search_in_batch(q[:1], 6, targets)
    np.testing.assert_array_equal(oids[0], rids)
    _same(od[0], rraw)
    qids = np.arange(100, 100 + len(q), dtype=np.int32)
    for tl in (False, True):
        rq, rids, rraw = s.pq_search_in_batch(q, qids, 6, targets, tl)
        oids, od = oi.pq_search_in_batch(q, 6, targets, use_target_lists=tl)
        np.testing.assert_array_equal(rq, np.repeat(qids[:, None], 6, 1))
        np.testing.assert_array_equal(oids, rids)
        _same(od, rraw)
    few = targets[:3]   # fewer rows than k: padded with id -1 / sentinel
    rq, rids, rraw = s.pq_search_in_batch(q, qids, 6, few, False)
    oids, od = oi.pq_search_in_batch(q, 6, few)
    np.testing.assert_array_equal(oids, rids)
    _same(od, rraw)


def test_pq_search_in_batch_unsorted_and_repeated_ids(oracle_mod, ref):
    """`WHERE id IN (...)` on a table whose id column is neither ascending nor unique: rows come in TABLE order and
    every row carrying a listed id is a candidate — the oracle's general row selection against the real SRF"""
    ix = dict(small_index(N=20000, d=48, m=12, K=64, C=40, seed=7, with_pq=True))
    rng = np.random.default_rng(4)
    ids = rng.permutation(np.arange(1, ix["N"] + 1)).astype(np.int32)
    ids[100:140] = ids[5000:5040]
    ix["ids"] = ids
    s = ref()
    s.load_pq(ix)
    oi = oracle_mod.OracleIndex(ix, flat_pq=True)
    targets = np.concatenate([ids[90:150], ids[4990:5050], rng.choice(ids, 1500), [10 ** 8, -5]]).astype(np.int32)
    q = queries_from(ix, 9, seed=2)
    qids = np.arange(len(q), dtype=np.int32)
    rq, rids, rraw = s.pq_search_in_batch(q, qids, 5, targets, False)
    oids, od = oi.pq_search_in_batch(q, 5, targets)
    np.testing.assert_array_equal(oids, rids)
    _same(od, rraw)


def _ivpq_setup(N=12000, d=48, m=12, K=64, Kc=8, nt=3000, seed=5):
    import torch
    from freddy_b200.index_build import make_ivpq_index
    ix = small_index(N=20000, d=d, m=12, K=64, C=40, seed=7)
    vec = np.ascontiguousarray(ix["vectors"][:N])
    rng = np.random.default_rng(seed)
    trows = np.sort(rng.choice(N, nt, replace=False))
    ivpq = make_ivpq_index(torch.from_numpy(vec), m=m, K=K, Kc=Kc, n_train=N, kmeans_iters=4, seed=3, target_rows=trows)
    vec_ids = np.arange(1, N + 1, dtype=np.int32)
    targets = (trows + 1).astype(np.int32)
    q = vec[rng.choice(N, 30, replace=False)] + 0.02 * rng.standard_normal((30, d)).astype(np.float32)
    return ivpq, vec, vec_ids, targets, np.ascontiguousarray(q, np.float32)


@pytest.mark.parametrize("method", [0, 1, 2])
@pytest.mark.parametrize("use_tl", [False, True])
def test_ivpq_search_in_driver(ref, oracle_mod, method, use_tl):
    """kNN-join driver (ivpq_search_in.c) incl. confidence stop, PV buffer, retry loop"""
    ivpq, vec, vec_ids, targets, q = _ivpq_setup()
    oi = oracle_mod.OracleIvpq(ivpq, vec, vec_ids)
    qids = np.arange(500, 500 + len(q), dtype=np.int32)
    for (k, alpha, pvf, conf) in ((5, 3, 4, 0.8), (5, 1, 2, 0.5), (3, 40, 20, 0.8)):
        s = ref()
        s.load_ivpq(ivpq, vec, vec_ids)
        rq, rids, rraw = s.ivpq_search_in(q, qids, k, targets, alpha, pvf, method, use_tl, conf, 10_000_000)
        oids, od, rc, st = oi.search_in(q, k, targets, alpha, pvf, method, use_tl, conf)
        assert rc == 0
        np.testing.assert_array_equal(rq, np.repeat(qids[:, None], k, 1))
        np.testing.assert_array_equal(oids, rids, err_msg=f"method={method} tl={use_tl} k={k} alpha={alpha} rounds={st[0]}")
        _same(od, rraw)


@pytest.mark.parametrize("method", [0, 2])
def test_ivpq_search_in_retry_loop(ref, oracle_mod, method):
    """few targets: the first rounds find < k candidates, alpha doubles until enough cells are probed
    (ivpq_search_in.c:639-680), incl. the target-list skip rule (:553-557)"""
    ivpq, vec, vec_ids, targets, q = _ivpq_setup()
    rng = np.random.default_rng(1)
    few = np.sort(rng.choice(targets, 60, replace=False)).astype(np.int32)
    oi = oracle_mod.OracleIvpq(ivpq, vec, vec_ids)
    qids = np.arange(len(q), dtype=np.int32)
    seen_retry = False
    for use_tl in (False, True):
        for (k, alpha, pvf, conf) in ((10, 1, 2, 0.5), (8, 2, 3, 0.8), (70, 1, 1, 0.5)):
            s = ref()
            s.load_ivpq(ivpq, vec, vec_ids)
            rq, rids, rraw = s.ivpq_search_in(q, qids, k, few, alpha, pvf, method, use_tl, conf, 10_000_000)
            oids, od, rc, st = oi.search_in(q, k, few, alpha, pvf, method, use_tl, conf)
            assert rc == 0
            seen_retry |= st[0] > 1
            np.testing.assert_array_equal(oids, rids, err_msg=f"method={method} tl={use_tl} k={k} rounds={st[0]}")
            _same(od, rraw)
    assert seen_retry


def test_ivfadc_batch_search_equals_w1_search(ref, oracle_mod):
    """ivfadc_batch_search (freddy.c:677-1024) == per fetched vector ivfadc_search with w = 1
    (one list per round until k rows were seen), sentinel 100.0, vectors fetched in table order"""
    ix = small_index()
    vec_ids = np.asarray(ix["ids"], np.int32)
    rng = np.random.default_rng(8)
    qids = rng.choice(vec_ids, 25, replace=False).astype(np.int32)
    qids = np.concatenate([qids, qids[:3], [10 ** 8]]).astype(np.int32)          # duplicates + unknown id
    for k in (5, 1, 12):
        s = ref()
        s.load_ivfadc(ix, 3)
        s.load_vectors_table(ix["vectors"], vec_ids)
        rq, rids, rraw = s.ivfadc_batch_search(qids, k)
        order = np.sort(np.unique(qids[qids < 10 ** 8]))
        np.testing.assert_array_equal(rq, order)
        q = ix["vectors"][order - 1]
        oids, od, rc, _ = oracle_mod.OracleIndex(ix).ivfadc_search(q, k, 1)
        assert rc == 0
        np.testing.assert_array_equal(oids, rids)
        od = np.where(oids == -1, np.float32(100.0), od)
        _same(od, rraw)
    # lists shorter than k: several rounds
    ix = small_index(N=150, d=24, m=12, K=8, C=64, seed=5, n_clusters=20)
    vec_ids = np.asarray(ix["ids"], np.int32)
    s = ref()
    s.load_ivfadc(ix, 3)
    s.load_vectors_table(ix["vectors"], vec_ids)
    qids = vec_ids[::7].copy()
    rq, rids, rraw = s.ivfadc_batch_search(qids, 12)
    oids, od, rc, _ = oracle_mod.OracleIndex(ix).ivfadc_search(ix["vectors"][rq - 1], 12, 1)
    assert rc == 0
    np.testing.assert_array_equal(oids, rids)
    _same(np.where(oids == -1, np.float32(100.0), od), rraw)


def test_grouping_pq_against_the_real_srf():
    """fo_grouping_pq vs the reference's own grouping_pq SRF (freddy.c:1178-1401) through the emulator:
    ids in table order, group assignment incl. ties between groups (first group in ascending id order wins),
    duplicate / unknown input ids, and the "Group ids do not exist" error"""
    from helpers import small_index
    from oracle import oracle
    ix = small_index(N=20000, d=48, m=12, K=64, C=40, seed=7, with_pq=True)
    vec_ids = np.asarray(ix["ids"], np.int32)
    vectors = ix["vectors"].copy()
    vectors[500] = vectors[100]                                   # groups 101 and 501 are the same vector: ties
    rs = oracle.ReferenceSession()
    rs.load_pq(ix)
    rs.load_vectors_table(vectors, vec_ids)
    oi = oracle.OracleIndex(ix, flat_pq=True)
    rng = np.random.default_rng(2)
    ids = rng.choice(np.arange(1, ix["N"] + 200), size=3000, replace=True).astype(np.int32)
    for groups in ([501, 101, 7, 9000], [42], [19999, 3, 250, 251, 252, 17, 101]):
        g = np.asarray(groups, np.int32)
        want_i, want_g = rs.grouping_pq(ids, g)
        got_i, got_g, rc = oi.grouping_pq(vectors, vec_ids, ids, g)
        assert rc == len(want_i)
        np.testing.assert_array_equal(got_i, want_i)
        np.testing.assert_array_equal(got_g, want_g)
    assert 501 not in set(oi.grouping_pq(vectors, vec_ids, ids, [501, 101])[1].tolist())
    with pytest.raises(RuntimeError):
        rs.grouping_pq(ids, np.asarray([5, 10 ** 8], np.int32))
    assert oi.grouping_pq(vectors, vec_ids, ids, [5, 10 ** 8])[2] == -1
    assert oi.grouping_pq(vectors, vec_ids, ids, [5, 5])[2] == -1


@pytest.mark.parametrize("method", [0, 2])
@pytest.mark.parametrize("use_tl", [False, True])
def test_ivpq_search_in_pair_lut_variant(ref, oracle_mod, method, use_tl):
    """alpha*k > double_threshold: the reference switches to getPrecomputedDistancesDouble (index_utils.c:457-475),
    whose distances are sums of PAIR sums — different fp32 roundings from the single-position chain"""
    ivpq, vec, vec_ids, targets, q = _ivpq_setup(m=12, K=64)        # K*K = 4096 fits the reference's int16 pair codes
    oi = oracle_mod.OracleIvpq(ivpq, vec, vec_ids)
    qids = np.arange(500, 500 + len(q), dtype=np.int32)
    differs = False
    for (k, alpha, pvf, conf) in ((5, 3, 4, 0.8), (3, 40, 20, 0.8)):
        s = ref()
        s.load_ivpq(ivpq, vec, vec_ids)
        rq, rids, rraw = s.ivpq_search_in(q, qids, k, targets, alpha, pvf, method, use_tl, conf, 0)
        oids, od, rc, st = oi.search_in(q, k, targets, alpha, pvf, method, use_tl, conf, 0)
        assert rc == 0
        np.testing.assert_array_equal(oids, rids, err_msg=f"pair-LUT method={method} tl={use_tl} k={k} alpha={alpha}")
        _same(od, rraw)
        _, od1, _, _ = oi.search_in(q, k, targets, alpha, pvf, method, use_tl, conf)
        differs |= bool((od1.view(np.uint32) != od.view(np.uint32)).any())
    if method == 0:
        assert differs, "the pair-sum chain should round differently from the single-position chain somewhere"
