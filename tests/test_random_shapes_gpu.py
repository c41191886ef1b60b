"""Randomised shapes: every (d, m, K, C, N, k, w) drawn here must agree with the oracle bit for
bit — or fail with FB_ERR_REFERENCE_UB exactly where the oracle reports that the reference would
run into undefined behaviour (fewer than w unprobed lists left)."""
import numpy as np
import pytest

from helpers import assert_same_topk

pytestmark = pytest.mark.gpu


def _case(rng):
    m = int(rng.choice([1, 2, 3, 5, 8, 12, 13, 16]))
    sub = int(rng.integers(1, 8))
    K = int(rng.choice([4, 8, 32, 64, 256]))
    C = int(rng.choice([1, 2, 7, 40]))
    N = int(rng.choice([10, 60, 700, 5000]))
    k = int(rng.choice([1, 2, 5, 17, 30, 31, 35]))
    w = int(rng.integers(1, min(C, 33) + 1))
    return dict(d=m * sub, m=m, K=K, C=C, N=N, k=k, w=w)


@pytest.mark.parametrize("seed", range(24))
def test_random_shape(seed, oracle_mod):
    from freddy_b200 import Engine, FreddyError, _lib
    from freddy_b200.index_build import make_synthetic_index
    rng = np.random.default_rng(1000 + seed)
    c = _case(rng)
    ix = make_synthetic_index(c["N"], d=c["d"], m=c["m"], K=c["K"], C=c["C"], n_train=c["N"], n_clusters=6, sigma=0.6,
                              kmeans_iters=3, seed=seed, device="cpu", keep_vectors=True)
    vec = ix.pop("vectors_t").numpy()
    nq = min(50, c["N"])
    q = vec[rng.choice(c["N"], nq, replace=False)] + 0.05 * rng.standard_normal((nq, c["d"])).astype(np.float32)
    q = np.ascontiguousarray(q, np.float32)
    eids, ed, rc, _ = oracle_mod.OracleIndex(ix).ivfadc_search(q, c["k"], c["w"])
    e = Engine(0)
    try:
        e.load_ivfadc_index(ix)
        for qmin in (0, 1 << 30):
            e.set_option(_lib.FB_OPT_QSCAN_MIN_QUERIES, qmin)
            if rc != 0:
                with pytest.raises(FreddyError) as ei:
                    e.ivfadc_search(q, c["k"], c["w"])
                assert ei.value.code == _lib.FB_ERR_REFERENCE_UB, c
            else:
                ids, d = e.ivfadc_search(q, c["k"], c["w"])
                assert_same_topk(ids, d, eids, ed, f"{c} qscan_min={qmin}")
    finally:
        e.close()
