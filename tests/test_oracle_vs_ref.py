"""Pins the oracle: kernel-level restatements vs the reference's own compiled
index_utils.c (oracle/_ref/libfreddy_ref.so), bit for bit, and vs the golden
vectors generated from it (tests/golden/make_golden.py)."""
import ctypes as C
import json
import os

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.fixture(scope="module")
def libs(oracle_mod):
    R = oracle_mod.ref_lib()
    if R is None:
        pytest.skip("oracle/_ref not built (reference sources absent and no prebuilt .so)")
    return oracle_mod.lib(), R


def test_square_distance_bitexact(libs):
    L, R = libs
    rng = np.random.default_rng(0)
    for n in (1, 2, 25, 150, 300, 301):
        for _ in range(50):
            a = rng.standard_normal(n).astype(np.float32)
            b = rng.standard_normal(n).astype(np.float32)
            x, y = L.fo_square_distance(_p(a), _p(b), n), R.squareDistance(_p(a), _p(b), n)
            assert np.float32(x).view(np.uint32) == np.float32(y).view(np.uint32)


def test_update_topk_matches_reference_including_ties(libs, oracle_mod):
    L, R = libs
    rng = np.random.default_rng(1)
    for trial in range(300):
        k = int(rng.integers(1, 9))
        n = int(rng.integers(1, 40))
        # few distinct distances => many ties
        d = rng.integers(0, 6, size=n).astype(np.float32) / 4
        a = (oracle_mod.TopKEntry * k)()
        b = (oracle_mod.TopKEntry * k)()
        L.fo_init_topk(a, k, 1000.0)
        L.fo_init_topk(b, k, 1000.0)
        for i in range(n):
            if d[i] < a[k - 1].distance:
                L.fo_update_topk(a, float(d[i]), i, k)
            if d[i] < b[k - 1].distance:
                R.updateTopK(b, float(d[i]), i, k, 0)
        assert [(e.id, e.distance) for e in a] == [(e.id, e.distance) for e in b]


def test_survey_tie_example(libs, oracle_mod):
    # (0,3)(1,1)(2,3)(3,2)(4,1)(5,.5)(6,3), k=5
    L, _ = libs
    k = 5
    tk = (oracle_mod.TopKEntry * k)()
    L.fo_init_topk(tk, k, 1000.0)
    for i, dist in enumerate([3, 1, 3, 2, 1, .5, 3]):
        if dist < tk[k - 1].distance:
            L.fo_update_topk(tk, dist, i, k)
    assert [(e.id, e.distance) for e in tk] == [(5, .5), (4, 1.0), (1, 1.0), (3, 2.0), (2, 3.0)]


class CodebookEntry(C.Structure):  # index_utils.h:51-55
    _fields_ = [("pos", C.c_int), ("code", C.c_int), ("vector", C.c_void_p)]


def test_precomputed_distances_and_adc_bitexact(libs):
    L, R = libs
    rng = np.random.default_rng(2)
    for (m, K, sub) in ((12, 64, 25), (4, 16, 3), (30, 32, 10)):
        cb = rng.standard_normal((m, K, sub)).astype(np.float32)
        q = rng.standard_normal(m * sub).astype(np.float32)
        mine = np.empty(m * K, np.float32)
        L.fo_precomputed_distances(_p(mine), m, K, sub, _p(q), _p(cb))
        # reference walks (pos, code, vector*) rows, here in shuffled row order
        order = rng.permutation(m * K)
        ents = (CodebookEntry * (m * K))()
        for slot, idx in enumerate(order):
            p, c = divmod(int(idx), K)
            ents[slot].pos, ents[slot].code = p, c
            ents[slot].vector = cb[p, c].ctypes.data
        ref = np.empty(m * K, np.float32)
        R.getPrecomputedDistances(_p(ref), m, K, sub, _p(q), ents)
        np.testing.assert_array_equal(mine.view(np.uint32), ref.view(np.uint32))
        for _ in range(100):
            codes = rng.integers(0, K, size=m).astype(np.int16)
            x = L.fo_pq_distance_int16(_p(mine), _p(codes), m, K)
            y = R.computePQDistanceInt16(_p(ref), _p(codes), m, K)
            assert np.float32(x).view(np.uint32) == np.float32(y).view(np.uint32)


def test_golden_vectors(oracle_mod):
    """fixtures produced by tests/golden/make_golden.py from oracle/_ref (the real reference code)"""
    path = os.path.join(HERE, "golden", "kernels_golden.json")
    g = json.load(open(path))
    L = oracle_mod.lib()
    for case in g["square_distance"]:
        a = np.array(case["a_bits"], np.uint32).view(np.float32)
        b = np.array(case["b_bits"], np.uint32).view(np.float32)
        got = np.float32(L.fo_square_distance(_p(a), _p(b), len(a))).view(np.uint32)
        assert int(got) == case["out_bits"]
    for case in g["topk"]:
        k = case["k"]
        tk = (oracle_mod.TopKEntry * k)()
        L.fo_init_topk(tk, k, 1000.0)
        for i, dist in enumerate(case["stream"]):
            if dist < tk[k - 1].distance:
                L.fo_update_topk(tk, dist, i, k)
        assert [[e.id, e.distance] for e in tk] == case["out"]
    for case in g["lut_adc"]:
        m, K, sub = case["m"], case["K"], case["sub"]
        cb = np.array(case["cb_bits"], np.uint32).view(np.float32)
        q = np.array(case["q_bits"], np.uint32).view(np.float32)
        lut = np.empty(m * K, np.float32)
        L.fo_precomputed_distances(_p(lut), m, K, sub, _p(q), _p(cb))
        assert lut.view(np.uint32).tolist() == case["lut_bits"]
        for codes, out in zip(case["codes"], case["adc_bits"]):
            c = np.array(codes, np.int16)
            got = np.float32(L.fo_pq_distance_int16(_p(lut), _p(c), m, K)).view(np.uint32)
            assert int(got) == out


def test_srf_golden_vectors(oracle_mod):
    """the oracle's drivers against committed outputs of the reference's own SRFs"""
    from helpers import srf_golden
    ix, g = srf_golden()
    oi = oracle_mod.OracleIndex(ix)
    for k, w, tag in ((5, 3, "ivfadc_k5_w3"), (12, 1, "ivfadc_k12_w1")):
        ids, d, rc, _ = oi.ivfadc_search(g["queries"], k, w)
        assert rc == 0
        np.testing.assert_array_equal(ids, g[tag + "_ids"])
        np.testing.assert_array_equal(d.view(np.uint32), g[tag + "_dist"].view(np.uint32))
    L = oracle_mod.lib()
    txt = np.array([L.fo_round_through_text(float(x)) for x in g["ivfadc_k5_w3_dist"].ravel()], np.float32)
    np.testing.assert_array_equal(txt.view(np.uint32), g["ivfadc_k5_w3_text"].ravel().view(np.uint32))
    op = oracle_mod.OracleIndex(ix, flat_pq=True)
    ids, d = op.pq_search(g["queries"][:6], 4)
    np.testing.assert_array_equal(ids, g["pq_search_k4_ids"])
    np.testing.assert_array_equal(d.view(np.uint32), g["pq_search_k4_dist"].view(np.uint32))
    ids, d = op.pq_search_in_batch(g["queries"], 5, g["targets"])
    np.testing.assert_array_equal(ids, g["pq_in_k5_ids"])
    np.testing.assert_array_equal(d.view(np.uint32), g["pq_in_k5_dist"].view(np.uint32))


def test_encode_against_reference_update_codebook():
    """fo_encode's per-position assignment vs the nearestCentroids the reference's own updateCodebook
    (index_utils.c:908-957) computes, incl. duplicate codewords (first minimum in table order wins)"""
    from oracle import oracle
    if oracle.ref_lib() is None:
        pytest.skip("oracle/_ref not built")
    rng = np.random.default_rng(3)
    m, K, sub = 6, 32, 5
    cb = rng.standard_normal((m, K, sub)).astype(np.float32) * 0.3
    cb[:, 7] = cb[:, 3]                                        # duplicate codeword: ties
    v = rng.standard_normal((300, m * sub)).astype(np.float32) * 0.3
    v[:40] = cb[:, 7].reshape(-1)[None, :]                     # vectors exactly on the duplicated codeword
    want = oracle.reference_update_codebook_assignments(v, cb)
    _, got, rc = oracle.encode(v, cb)
    assert rc == 0
    np.testing.assert_array_equal(got, want)
    assert (got[:40] == 3).all()


def test_sql_level_oracles_are_built_from_pinned_pieces():
    """k_nearest_neighbour / k_nearest_neighbour_ivfadc_pv are plpgsql in the reference; their oracle restatement
    (oracle.knn_exact, oracle.ivfadc_search_pv) must equal the composition of the reference's OWN compiled pieces:
    the real ivfadc_search SRF (candidates) and the real cosine_similarity_bytea (similarities), ordered by
    (similarity desc, table row asc) and cut at k."""
    from helpers import queries_from, small_index
    from oracle import oracle
    if oracle.ref_lib() is None:
        pytest.skip("oracle/_ref not built")
    ix = small_index(N=20000, d=48, m=12, K=64, C=40, seed=7)
    vec_ids = np.asarray(ix["ids"], np.int32)
    vectors = ix["vectors"]
    q = queries_from(ix, 12, seed=31, noise=0.03)
    rs = oracle.ReferenceSession()
    R = rs.R

    def ref_sim(a, b):
        a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
        return np.float32(R.ref_cosine_similarity_bytea(a.ctypes.data, b.ctypes.data, len(a)))

    # (1) the vectorised float4 chain == the reference's scalar loop, bit for bit
    sims = oracle.cosine_similarity_bytea_many(q[0], vectors[:400])
    want = np.array([ref_sim(q[0], v) for v in vectors[:400]], np.float32)
    np.testing.assert_array_equal(sims.view(np.uint32), want.view(np.uint32))
    # (2) post-verification: real SRF candidates + real similarities
    k, pvf, w = 4, 5, 3
    rs.load_ivfadc(ix, w)
    cand, _, _ = rs.ivfadc_search(q, k * pvf)
    got_ids, got_s = oracle.ivfadc_search_pv(oracle.OracleIndex(ix), vectors, vec_ids, q, k, pvf, w, threads=2)
    row_of = {int(i): r for r, i in enumerate(vec_ids)}
    for qi in range(len(q)):
        rows = [row_of[int(c)] for c in cand[qi] if int(c) in row_of]
        scored = sorted(((-float(ref_sim(q[qi], vectors[r])), r) for r in rows))[:k]
        assert [int(vec_ids[r]) for _, r in scored] == [int(i) for i in got_ids[qi] if i >= 0]
        np.testing.assert_array_equal(np.array([-s for s, _ in scored], np.float32).view(np.uint32),
                                      got_s[qi][:len(scored)].view(np.uint32))
    # (2b) the flat-PQ twin (k_nearest_neighbour_pq_pv): real pq_search SRF candidates + real similarities
    ixp = small_index(N=20000, d=48, m=12, K=64, C=40, seed=7, with_pq=True)
    rs2 = oracle.ReferenceSession()
    rs2.load_pq(ixp)
    cand, _ = rs2.pq_search(q[:6], k * pvf)
    got_ids, got_s = oracle.pq_search_pv(oracle.OracleIndex(ixp, flat_pq=True), ixp["vectors"], vec_ids, q[:6], k, pvf)
    for qi in range(6):
        rows = [row_of[int(c)] for c in cand[qi] if int(c) in row_of]
        scored = sorted(((-float(ref_sim(q[qi], ixp["vectors"][r])), r) for r in rows))[:k]
        assert [int(vec_ids[r]) for _, r in scored] == [int(i) for i in got_ids[qi] if i >= 0]
        np.testing.assert_array_equal(np.array([-s for s, _ in scored], np.float32).view(np.uint32),
                                      got_s[qi][:len(scored)].view(np.uint32))
    # (3) exact k-NN over a slice of the table
    sub_ids = vec_ids[:600]
    e_ids, e_s = oracle.knn_exact(vectors[:600], sub_ids, q[:3], 5)
    for qi in range(3):
        scored = sorted(((-float(ref_sim(q[qi], vectors[r])), r) for r in range(600)))[:5]
        assert [int(sub_ids[r]) for _, r in scored] == e_ids[qi].tolist()


def test_oracle_against_committed_extended_golden():
    """no /root/reference needed: grouping_pq and the quantisation rule against outputs of the reference's own
    code committed in tests/golden/srf_golden_ext.npz"""
    from helpers import srf_golden_ext
    from oracle import oracle
    g = srf_golden_ext()
    ix = {"d": int(g["d"]), "m": int(g["m"]), "K": int(g["K"]), "C": 0, "N": int(g["N"]), "ids": g["ids"],
          "pq_codebook": g["pq_codebook"], "pq_codes": g["pq_codes"]}
    oi = oracle.OracleIndex(ix, flat_pq=True)
    ids, groups, rc = oi.grouping_pq(g["vectors"], g["ids"], g["grouping_in_ids"], g["grouping_groups"])
    assert rc == len(g["grouping_out_ids"])
    np.testing.assert_array_equal(ids, g["grouping_out_ids"])
    np.testing.assert_array_equal(groups, g["grouping_out_groups"])
    assert 701 not in set(groups.tolist()) and 41 in set(groups.tolist())      # identical vectors: the lower id wins
    _, codes, rc = oracle.encode(g["encode_rows"], g["pq_codebook"])
    assert rc == 0
    np.testing.assert_array_equal(codes, g["encode_pq_codes"])


def test_create_statistics_restatement(oracle_mod):
    """create_statistics is plpgsql (freddy--0.0.1.sql:150-171): no compiled reference to run, so the restatement is
    pinned to a hand-worked case of the SQL and to the rule the index builder already used (float8 division, float4 column)"""
    ids = np.arange(1, 11)
    cells = np.array([0, 1, 1, 2, 0, 3, 3, 3, 1, 0])
    st = oracle_mod.create_statistics(ids, cells, None, 4)
    np.testing.assert_array_equal(st, np.asarray([0.3, 0.3, 0.1, 0.3, 10.0], np.float32))
    # user column holds word 1 twice, word 2 once, word 10 once and two words the index does not know: JOIN has 4 rows
    st = oracle_mod.create_statistics(ids, cells, [1, 1, 2, 99, -4, 10], 4)
    np.testing.assert_array_equal(st, np.asarray([0.75, 0.25, 0.0, 0.0, 4.0], np.float32))
    # an index with a repeated id: both of its rows join
    st = oracle_mod.create_statistics([5, 5, 6], [2, 0, 1], [5], 3)
    np.testing.assert_array_equal(st, np.asarray([0.5, 0.0, 0.5, 2.0], np.float32))
    rng = np.random.default_rng(1)
    cells = rng.integers(0, 49, 5000)
    st = oracle_mod.create_statistics(np.arange(5000), cells, None, 49)
    exp = (np.bincount(cells, minlength=49).astype(np.float64) / 5000.0).astype(np.float32)
    np.testing.assert_array_equal(st[:49].view(np.uint32), exp.view(np.uint32))
