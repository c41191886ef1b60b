"""cosine_similarity*, vec_*_bytea and analogy_3cosadd: oracle vs the reference's own
compiled functions (CPU), CUDA engine vs oracle (GPU)."""
import ctypes as C

import numpy as np
import pytest

from helpers import small_index


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _datums(v):
    """float4 values as a Datum[] (what getArray hands to cosine_similarity_simple)"""
    return np.ascontiguousarray(v.view(np.uint32).astype(np.uint64))


def test_oracle_vector_udfs_match_reference(oracle_mod):
    R = oracle_mod.ref_lib()
    if R is None:
        pytest.skip("oracle/_ref not built")
    L = oracle_mod.lib()
    rng = np.random.default_rng(3)
    for d in (1, 7, 300):
        for _ in range(40):
            a = rng.standard_normal(d).astype(np.float32)
            b = rng.standard_normal(d).astype(np.float32)
            da, db = _datums(a), _datums(b)
            assert L.fo_cosine_similarity(_p(a), _p(b), d) == R.cosine_similarity_simple(_p(da), _p(db), d)
            assert L.fo_cosine_similarity_norm(_p(a), _p(b), d) == R.cosine_similarity_simple_norm(_p(da), _p(db), d)
            x, y = L.fo_cosine_similarity_bytea(_p(a), _p(b), d), R.ref_cosine_similarity_bytea(_p(a), _p(b), d)
            assert np.float32(x).view(np.uint32) == np.float32(y).view(np.uint32)
            o1, o2 = np.empty(d, np.float32), np.empty(d, np.float32)
            L.fo_vec_minus(_p(a), _p(b), d, _p(o1)); R.ref_vec_minus_bytea(_p(a), _p(b), d, _p(o2))
            np.testing.assert_array_equal(o1.view(np.uint32), o2.view(np.uint32))
            L.fo_vec_plus(_p(a), _p(b), d, _p(o1)); R.ref_vec_plus_bytea(_p(a), _p(b), d, _p(o2))
            np.testing.assert_array_equal(o1.view(np.uint32), o2.view(np.uint32))
            L.fo_vec_normalize(_p(a), d, _p(o1)); R.ref_vec_normalize_bytea(_p(a), d, _p(o2))
            np.testing.assert_array_equal(o1.view(np.uint32), o2.view(np.uint32))
    z = np.zeros(5, np.float32)
    assert L.fo_cosine_similarity(_p(z), _p(z), 5) == 0.0 == R.cosine_similarity_simple(_p(_datums(z)), _p(_datums(z)), 5)


@pytest.fixture(scope="module")
def eng():
    from freddy_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.gpu
def test_gpu_vector_udfs(eng, oracle_mod):
    L = oracle_mod.lib()
    rng = np.random.default_rng(4)
    n, d = 500, 300
    a = rng.standard_normal((n, d)).astype(np.float32)
    b = rng.standard_normal((n, d)).astype(np.float32)
    a[3] = 0
    exp = {0: [L.fo_cosine_similarity(_p(a[i]), _p(b[i]), d) for i in range(n)],
           1: [L.fo_cosine_similarity_norm(_p(a[i]), _p(b[i]), d) for i in range(n)],
           2: [float(np.float32(L.fo_cosine_similarity_bytea(_p(a[i]), _p(b[i]), d))) for i in range(n)]}
    for variant in (0, 1, 2):
        got = eng.cosine_similarity(a, b, variant)
        np.testing.assert_array_equal(got.view(np.uint64), np.array(exp[variant], np.float64).view(np.uint64))
    o = np.empty(d, np.float32)
    for op, fn in ((0, L.fo_vec_minus), (1, L.fo_vec_plus)):
        got = eng.vec_op(op, a, b)
        for i in range(0, n, 37):
            fn(_p(a[i]), _p(b[i]), d, _p(o))
            np.testing.assert_array_equal(got[i].view(np.uint32), o.view(np.uint32))
    got = eng.vec_op(2, a[4:])
    for i in range(0, n - 4, 41):
        L.fo_vec_normalize(_p(a[4 + i]), d, _p(o))
        np.testing.assert_array_equal(got[i].view(np.uint32), o.view(np.uint32))


@pytest.mark.gpu
@pytest.mark.parametrize("N,d", [(5000, 48), (20011, 300)])
def test_gpu_analogy_3cosadd(eng, oracle_mod, N, d):
    ix = small_index(N=20000, d=48, m=12, K=64, C=40, seed=7) if d == 48 else None
    rng = np.random.default_rng(N)
    if ix is not None:
        vec = np.ascontiguousarray(ix["vectors"][:N])
    else:
        vec = rng.standard_normal((N, d)).astype(np.float32)
        vec /= np.linalg.norm(vec, axis=1, keepdims=True)
    vec[N // 2] = vec[N // 3]            # exact duplicate rows: equal scores, the earlier row must win
    ids = np.arange(1, N + 1, dtype=np.int32) * 3   # ids != rows
    eng.load_vectors(ids, vec)
    nq = 70
    rows = rng.integers(0, N, size=(nq, 3)).astype(np.int32)
    rows[0] = (N // 3, N // 3 + 1, N // 3 + 2)
    got_ids, got_s = eng.analogy_3cosadd(ids[rows])
    exp_rows, exp_s = oracle_mod.analogy_3cosadd(vec, rows, threads=4)
    np.testing.assert_array_equal(got_ids, ids[exp_rows])
    np.testing.assert_array_equal(got_s.view(np.uint32), exp_s.view(np.uint32))
    # explicit-query form (vocabulary-sharded building block)
    q = np.stack([(vec[c] - vec[a]) + vec[b] for a, b, c in rows]).astype(np.float32)
    got2, s2 = eng.analogy_scan(q, ids[rows])
    np.testing.assert_array_equal(got2, got_ids)
    np.testing.assert_array_equal(s2.view(np.uint32), got_s.view(np.uint32))
