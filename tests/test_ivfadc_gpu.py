"""GPU parity: the CUDA engine (through the C-ABI) against the CPU oracle on the
same seeded indexes.  Bar: ids/ranks AND distances bit-exact (same fp32 chain)."""
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


def _run_both(eng, oracle_mod, ix, q, k, w, force_exact=False, qscan_min=64):
    """qscan_min: 0 = always the one-CTA-per-query scan, 1<<30 = always one CTA per (query, list)"""
    from freddy_b200 import _lib
    eng.load_ivfadc_index(ix)
    eng.set_option(_lib.FB_OPT_FORCE_EXACT_PATH, 1 if force_exact else 0)
    eng.set_option(_lib.FB_OPT_QSCAN_MIN_QUERIES, qscan_min)
    eng.reset_counters()
    ids, d = eng.ivfadc_search(q, k, w)
    eng.set_option(_lib.FB_OPT_FORCE_EXACT_PATH, 0)
    eng.set_option(_lib.FB_OPT_QSCAN_MIN_QUERIES, 64)
    oi = oracle_mod.OracleIndex(ix)
    eids, ed, rc, rows = oi.ivfadc_search(q, k, w, threads=4)
    assert rc == 0
    return ids, d, eids, ed, rows


@pytest.mark.parametrize("qscan_min", [0, 1 << 30])
@pytest.mark.parametrize("k,w", [(5, 4), (1, 1), (10, 3), (31, 8)])
def test_parity_small(eng, oracle_mod, k, w, qscan_min):
    ix = small_index()
    q = queries_from(ix, 300)
    ids, d, eids, ed, rows = _run_both(eng, oracle_mod, ix, q, k, w, qscan_min=qscan_min)
    assert_same_topk(ids, d, eids, ed, f"k={k} w={w}")
    c = eng.counters()
    assert c["rows_scanned"] == rows
    assert c["queries"] == 300


def test_parity_general_kernel(eng, oracle_mod):
    ix = small_index()
    q = queries_from(ix, 120, seed=5)
    ids, d, eids, ed, _ = _run_both(eng, oracle_mod, ix, q, 5, 4, force_exact=True)
    assert_same_topk(ids, d, eids, ed, "forced general kernel")
    assert eng.counters()["exact_path_queries"] == 120


@pytest.mark.parametrize("overlap", [0, 1])
def test_parity_noisy_queries_and_chunks(eng, oracle_mod, overlap):
    """several pipeline chunks (last one ragged); overlap=1: LUT build of chunk c+1 runs on a second
    stream concurrently with the scan of chunk c (two alternating LUT buffers)"""
    from freddy_b200 import _lib
    ix = small_index()
    q = queries_from(ix, 457, seed=3, noise=0.05)
    eng.set_option(_lib.FB_OPT_QUERY_CHUNK, 100)   # 5 chunks, last one ragged
    eng.set_option(_lib.FB_OPT_OVERLAP, overlap)
    eng.set_option(_lib.FB_OPT_LUT_CTAS_PER_SM, 1 if overlap else 0)
    try:
        for _ in range(3):
            ids, d, eids, ed, _ = _run_both(eng, oracle_mod, ix, q, 7, 5)
            assert_same_topk(ids, d, eids, ed, f"chunked overlap={overlap}")
    finally:
        eng.set_option(_lib.FB_OPT_QUERY_CHUNK, 2048)
        eng.set_option(_lib.FB_OPT_OVERLAP, 0)
        eng.set_option(_lib.FB_OPT_LUT_CTAS_PER_SM, 0)


def test_ties_everywhere(eng, oracle_mod):
    """m=2 (or 4) positions x K=4 codes: every list holds only 16 (256) distinct
    code vectors, so distance ties straddle the k-th place all the time."""
    general = 0
    for (d, m) in ((8, 2), (16, 4)):
        ix = small_index(N=6000, d=d, m=m, K=4, C=8, seed=3, n_clusters=5)
        q = queries_from(ix, 200, seed=9)
        for k, w in ((5, 2), (3, 8), (20, 3)):
            ids, d_, eids, ed, _ = _run_both(eng, oracle_mod, ix, q, k, w, qscan_min=(1 << 30) if k == 3 else 0)
            assert_same_topk(ids, d_, eids, ed, f"ties m={m} k={k} w={w}")
            general += eng.counters()["exact_path_queries"]
    # short tie groups are replayed inside the merging warp (warp_emit_topk); groups that do not fit
    # the warp's 32 keys (m=2: about 47 duplicates per distinct code vector) still reach the general kernel
    assert general > 0


def test_coarse_ties_replayed_in_the_selection_warp(eng, oracle_mod):
    """equal centroids: the reference keeps the LATEST of the tied entries when a nearer centroid with a higher id
    arrives afterwards (insert-before-equal, freddy.c:272-283) — not the lowest ids.  The coarse kernel replays that
    in the warp (no general kernel) unless more than 31 - w centroids tie."""
    ix = small_index(N=20000, d=48, m=12, K=64, C=40, seed=7)
    coarse = ix["coarse"].copy()
    coarse[[0, 1, 2]] = coarse[20]                  # a four-way tie whose members mostly precede the nearer centroids
    coarse[[5, 30]] = coarse[11]                    # and a three-way one
    ix = dict(ix, coarse=coarse)
    q = queries_from(ix, 600, seed=5, noise=0.01)
    straddled = 0
    for k, w in ((5, 1), (5, 2), (5, 3), (3, 5), (8, 9)):
        for qscan_min in (0, 1 << 30):
            ids, d_, eids, ed, _ = _run_both(eng, oracle_mod, ix, q, k, w, qscan_min=qscan_min)
            assert_same_topk(ids, d_, eids, ed, f"coarse ties k={k} w={w}")
            assert eng.counters()["exact_coarse_tie"] == 0
        dist = ((q[:, None, :] - coarse[None]) ** 2).sum(-1)
        srt = np.sort(dist, axis=1)
        straddled += int((srt[:, w - 1] == srt[:, w]).sum())
    assert straddled > 100, "the fixture must put ties across the w-th place"
    # 36 equal centroids: more than the warp's 32 keys can settle, the general kernel takes over — same answer
    coarse = ix["coarse"].copy()
    coarse[:36] = coarse[37]
    ix2 = dict(ix, coarse=coarse)
    ids, d_, eids, ed, _ = _run_both(eng, oracle_mod, ix2, q[:100], 5, 3)
    assert_same_topk(ids, d_, eids, ed, "coarse ties beyond the warp list")
    assert eng.counters()["exact_coarse_tie"] > 0


def test_reprobe_loop(eng, oracle_mod):
    """lists much shorter than k: the reference re-probes with a blacklist (freddy.c:262)"""
    ix = small_index(N=150, d=24, m=12, K=8, C=64, seed=5, n_clusters=20)
    q = queries_from(ix, 60, seed=2)
    ids, d, eids, ed, _ = _run_both(eng, oracle_mod, ix, q, 12, 2)
    assert_same_topk(ids, d, eids, ed, "re-probe")
    assert eng.counters()["exact_path_queries"] > 0


def test_large_k_and_w(eng, oracle_mod):
    ix = small_index()
    q = queries_from(ix, 40, seed=8)
    ids, d, eids, ed, _ = _run_both(eng, oracle_mod, ix, q, 100, 6)   # k > 31: general kernel
    assert_same_topk(ids, d, eids, ed, "k=100")
    ids, d, eids, ed, _ = _run_both(eng, oracle_mod, ix, q, 5, 36)    # w > 31: general kernel
    assert_same_topk(ids, d, eids, ed, "w=36")


def test_generic_m(eng, oracle_mod):
    ix = small_index(N=8000, d=40, m=10, K=32, C=16, seed=11)         # m=10: runtime-m scan kernel
    q = queries_from(ix, 100, seed=4)
    ids, d, eids, ed, _ = _run_both(eng, oracle_mod, ix, q, 5, 3)
    assert_same_topk(ids, d, eids, ed, "m=10")


@pytest.mark.parametrize("packed", [1, 0])
def test_readme_shape(eng, oracle_mod, packed):
    """d=300, m=12, K=1024 (README.md:125-128) on a 60k-row table; LUT build with the
    packed f32x2 and the scalar forms of the same rounded operations"""
    from freddy_b200 import _lib
    ix = small_index(N=60000, d=300, m=12, K=1024, C=100, seed=1, n_clusters=100)
    q = queries_from(ix, 150, seed=6)
    eng.set_option(_lib.FB_OPT_PACKED_FP32, packed)
    try:
        for qmin in (0, 1 << 30):
            for w in (10, 3, 7):       # even / odd numbers of chains per thread
                ids, d, eids, ed, _ = _run_both(eng, oracle_mod, ix, q, 5, w, qscan_min=qmin)
                assert_same_topk(ids, d, eids, ed, f"d=300 K=1024 qscan_min={qmin} w={w} packed={packed}")
    finally:
        eng.set_option(_lib.FB_OPT_PACKED_FP32, 1)


def test_edge_cases(eng, oracle_mod):
    from freddy_b200 import FreddyError, _lib
    ix = small_index()
    eng.load_ivfadc_index(ix)
    ids, d = eng.ivfadc_search(np.zeros((0, ix["d"]), np.float32), 5, 3)
    assert ids.shape == (0, 5)
    with pytest.raises(FreddyError) as ei:
        eng.ivfadc_search(queries_from(ix, 2), 5, ix["C"] + 1)
    assert ei.value.code == _lib.FB_ERR_REFERENCE_UB
    with pytest.raises(FreddyError):
        eng.ivfadc_search(queries_from(ix, 2), 0, 3)


def test_flat_pq(eng, oracle_mod):
    ix = small_index(N=20000, d=48, m=12, K=64, C=40, seed=7, with_pq=True)
    eng.load_pq_index(ix)
    oi = oracle_mod.OracleIndex(ix, flat_pq=True)
    q = queries_from(ix, 40, seed=13)
    ids, d = eng.pq_search(q, 5)
    eids, ed = oi.pq_search(q, 5)
    assert_same_topk(ids, d, eids, ed, "pq_search")
    rng = np.random.default_rng(0)
    targets = rng.choice(np.arange(1, ix["N"] + 500), size=3000, replace=True).astype(np.int32)  # dups + unknown ids
    for tl in (False, True):
        ids, d = eng.pq_search_in_batch(q, 6, targets, use_target_lists=tl)
        eids, ed = oi.pq_search_in_batch(q, 6, targets, use_target_lists=tl)
        assert_same_topk(ids, d, eids, ed, "pq_search_in_batch")
    few = targets[:3]
    ids, d = eng.pq_search_in_batch(q, 6, few)
    eids, ed = oi.pq_search_in_batch(q, 6, few)
    assert_same_topk(ids, d, eids, ed, "fewer targets than k")


def test_flat_pq_in_batch_shapes(eng, oracle_mod):
    """pq_search_in_batch: the target subset is built on the device (id -> rows, bitmap, scan, gather); unsorted and
    repeated ids in the table (every matching row is a candidate, once, in table order), many queries (one CTA per
    query, resident LUT) and few queries (the table cut into segments), empty and unknown target lists"""
    ix = dict(small_index(N=20000, d=48, m=12, K=64, C=40, seed=7, with_pq=True))
    rng = np.random.default_rng(4)
    ids = rng.permutation(np.arange(1, ix["N"] + 1)).astype(np.int32)
    ids[100:140] = ids[5000:5040]                       # repeated ids: both rows are selected by `id IN (...)`
    ix["ids"] = ids
    eng.load_pq_index(ix)
    oi = oracle_mod.OracleIndex(ix, flat_pq=True)
    targets = np.concatenate([ids[90:150], ids[4990:5050], rng.choice(ids, 4000), [10 ** 8, -5]]).astype(np.int32)
    from freddy_b200 import _lib
    for nq in (1, 7, 400):
        q = queries_from(ix, nq, seed=nq)
        exp = oi.pq_search_in_batch(q, 5, targets)
        for placement in (1, 2, 0):                     # 2: the subset's rows are placed bank-aware (what large batches get)
            eng.set_option(_lib.FB_OPT_SUBSET_PLACEMENT, placement)
            got = eng.pq_search_in_batch(q, 5, targets)
            assert_same_topk(got[0], got[1], exp[0], exp[1], f"pq_search_in_batch nq={nq} placement={placement}")
    eng.set_option(_lib.FB_OPT_SUBSET_PLACEMENT, 1)
    q = queries_from(ix, 9, seed=2)
    for t in (np.zeros(0, np.int32), np.asarray([10 ** 8], np.int32)):
        got = eng.pq_search_in_batch(q, 4, t)
        assert (got[0] == -1).all() and (got[1] == np.float32(1000.0)).all()
    got = eng.pq_search(q, 12)                          # the whole table, few queries: segmented scan + finalize
    exp = oi.pq_search(q, 12)
    assert_same_topk(got[0], got[1], exp[0], exp[1], "pq_search few queries")
    q = queries_from(ix, 700, seed=3)
    got = eng.pq_search(q, 3)                           # the whole table, one CTA per query
    exp = oi.pq_search(q, 3)
    assert_same_topk(got[0], got[1], exp[0], exp[1], "pq_search many queries")


def test_against_reference_srf_golden(eng):
    """the CUDA engine directly against committed outputs of the reference's own SRFs
    (freddy.c run through the emulator, tests/golden/make_golden.py)"""
    from helpers import srf_golden
    ix, g = srf_golden()
    eng.load_ivfadc_index(ix)
    for k, w, tag in ((5, 3, "ivfadc_k5_w3"), (12, 1, "ivfadc_k12_w1")):
        ids, d = eng.ivfadc_search(g["queries"], k, w)
        assert_same_topk(ids, d, g[tag + "_ids"], g[tag + "_dist"], tag)
    eng.load_pq_index(ix)
    ids, d = eng.pq_search(g["queries"][:6], 4)
    assert_same_topk(ids, d, g["pq_search_k4_ids"], g["pq_search_k4_dist"], "pq_search")
    ids, d = eng.pq_search_in_batch(g["queries"], 5, g["targets"])
    assert_same_topk(ids, d, g["pq_in_k5_ids"], g["pq_in_k5_dist"], "pq_search_in_batch")


def test_ivfadc_batch_search(eng, oracle_mod):
    """fb_ivfadc_batch_search: vectors fetched by id in table order, one list per round, sentinel 100.0
    (equivalence with the w = 1 search is pinned against the real SRF in test_oracle_vs_reference_srf.py)"""
    for ix, k in ((small_index(), 5), (small_index(N=150, d=24, m=12, K=8, C=64, seed=5, n_clusters=20), 12)):
        vec_ids = np.asarray(ix["ids"], np.int32)
        eng.load_ivfadc_index(ix)
        eng.load_vectors(vec_ids, ix["vectors"])
        rng = np.random.default_rng(2)
        qids = rng.choice(vec_ids, min(40, len(vec_ids)), replace=False).astype(np.int32)
        qids = np.concatenate([qids, qids[:2], [10 ** 8]]).astype(np.int32)
        oq, ids, d = eng.ivfadc_batch_search(qids, k)
        order = np.sort(np.unique(qids[qids < 10 ** 8]))
        np.testing.assert_array_equal(oq, order)
        eids, ed, rc, _ = oracle_mod.OracleIndex(ix).ivfadc_search(ix["vectors"][order - 1], k, 1)
        assert rc == 0
        ed = np.where(eids == -1, np.float32(100.0), ed)
        assert_same_topk(ids, d, eids, ed, "ivfadc_batch_search")


def test_small_calls_replay_a_cuda_graph(eng, oracle_mod):
    """host-buffer calls with <= 16 queries: the first call runs the ordinary path, the second captures the
    whole call as a CUDA graph, later ones replay it — all must give the oracle's bits, also after the index
    or an option changed (re-capture), and for queries that need the general kernel (ties)"""
    from freddy_b200 import _lib
    for ix in (small_index(), small_index(N=6000, d=16, m=4, K=4, C=8, seed=3, n_clusters=5)):
        eng.load_ivfadc_index(ix)
        oi = oracle_mod.OracleIndex(ix)
        for nq, k, w in ((1, 5, 4), (7, 3, 2), (16, 10, 3)):
            for rep in range(4):
                q = queries_from(ix, nq, seed=20 + rep)
                ids, d = eng.ivfadc_search(q, k, w)
                eids, ed, rc, _ = oi.ivfadc_search(q, k, w)
                assert rc == 0
                assert_same_topk(ids, d, eids, ed, f"graph nq={nq} k={k} w={w} rep={rep}")
            eng.set_option(_lib.FB_OPT_CUDA_GRAPHS, 0)
            q = queries_from(ix, nq, seed=99)
            a = eng.ivfadc_search(q, k, w)
            eng.set_option(_lib.FB_OPT_CUDA_GRAPHS, 1)
            for _ in range(3):
                b = eng.ivfadc_search(q, k, w)
                np.testing.assert_array_equal(a[0], b[0])
                np.testing.assert_array_equal(a[1].view(np.uint32), b[1].view(np.uint32))
