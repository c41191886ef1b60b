"""GPU parity of grouping_pq against the oracle, which is pinned to the reference's own
SRF in test_oracle_vs_reference_srf.py::test_grouping_pq_against_the_real_srf."""
import numpy as np
import pytest

from helpers import small_index

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from freddy_b200 import Engine
    e = Engine(0)
    yield e
    e.close()


@pytest.mark.parametrize("shape", [dict(N=20000, d=48, m=12, K=64, C=40, seed=7),
                                   dict(N=20000, d=40, m=10, K=32, C=16, seed=11),
                                   dict(N=30000, d=300, m=12, K=1024, C=100, seed=1, n_clusters=100)])
def test_grouping_pq(eng, oracle_mod, shape):
    from freddy_b200 import FreddyError
    ix = small_index(with_pq=True, **shape)
    vec_ids = np.asarray(ix["ids"], np.int32)
    vectors = ix["vectors"].copy()
    vectors[500] = vectors[100]                                   # two identical group vectors: the lower id wins
    eng.load_pq_index(ix)
    eng.load_vectors(vec_ids, vectors)
    oi = oracle_mod.OracleIndex(ix, flat_pq=True)
    rng = np.random.default_rng(3)
    ids = rng.choice(np.arange(1, ix["N"] + 300), size=9000, replace=True).astype(np.int32)
    for groups in ([501, 101, 7, 9000], [42], list(range(3, 400, 11)) + [101, 501]):
        g = np.asarray(groups, np.int32)
        got_i, got_g = eng.grouping_pq(ids, g)
        want_i, want_g, rc = oi.grouping_pq(vectors, vec_ids, ids, g)
        assert rc == len(want_i)
        np.testing.assert_array_equal(got_i, want_i)
        np.testing.assert_array_equal(got_g, want_g)
    got_i, got_g = eng.grouping_pq(np.asarray([10 ** 8], np.int32), np.asarray([5], np.int32))   # no such rows
    assert len(got_i) == 0
    with pytest.raises(FreddyError) as ei:
        eng.grouping_pq(ids, np.asarray([5, 10 ** 8], np.int32))
    assert "Group ids do not exist" in str(ei.value)


def test_against_reference_golden_ext(eng):
    """the CUDA engine directly against committed outputs of the reference's own grouping_pq SRF and
    updateCodebook (tests/golden/srf_golden_ext.npz)"""
    from freddy_b200 import _lib
    from helpers import srf_golden_ext
    g = srf_golden_ext()
    ix = {"d": int(g["d"]), "m": int(g["m"]), "K": int(g["K"]), "N": int(g["N"]), "ids": g["ids"],
          "pq_codebook": g["pq_codebook"], "pq_codes": g["pq_codes"]}
    eng.load_pq_index(ix)
    eng.load_vectors(np.asarray(g["ids"], np.int32), g["vectors"])
    ids, groups = eng.grouping_pq(g["grouping_in_ids"], g["grouping_groups"])
    np.testing.assert_array_equal(ids, g["grouping_out_ids"])
    np.testing.assert_array_equal(groups, g["grouping_out_groups"])
    eng.load_codebook(_lib.FB_CB_PQ, g["pq_codebook"])
    np.testing.assert_array_equal(eng.encode_pq(g["encode_rows"]), g["encode_pq_codes"])
