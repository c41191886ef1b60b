"""GPU parity of the quantisation of new rows (fb_encode_ivfadc / fb_encode_pq) against
the oracle's restatement of insert_batch's assignment loops (freddy.c:1567-1582, index_utils.c:923-939; the
per-position loop is pinned to the reference's compiled updateCodebook in test_oracle_vs_ref.py)."""
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
                                   dict(N=20000, d=300, m=12, K=1024, C=100, seed=1, n_clusters=100)])
def test_encode_ivfadc_reproduces_the_index(eng, oracle_mod, shape):
    from freddy_b200 import _lib
    ix = small_index(with_pq=True, **shape)
    eng.load_coarse(ix["coarse"])
    eng.load_codebook(_lib.FB_CB_RESIDUAL, ix["residual_codebook"])
    rng = np.random.default_rng(0)
    v = ix["vectors"][rng.choice(ix["N"], 5000, replace=False)].copy()
    v[:50] += 0.01 * rng.standard_normal((50, ix["d"])).astype(np.float32)
    cids, codes = eng.encode_ivfadc(v)
    ecids, ecodes, rc = oracle_mod.encode(v, ix["residual_codebook"], ix["coarse"])
    assert rc == 0
    np.testing.assert_array_equal(cids, ecids)
    np.testing.assert_array_equal(codes, ecodes)
    # flat PQ on the raw vectors
    eng.load_codebook(_lib.FB_CB_PQ, ix["pq_codebook"])
    pq = eng.encode_pq(v)
    _, epq, rc = oracle_mod.encode(v, ix["pq_codebook"])
    assert rc == 0
    np.testing.assert_array_equal(pq, epq)


def test_encode_ties_and_far_rows(eng, oracle_mod):
    from freddy_b200 import FreddyError, _lib
    rng = np.random.default_rng(4)
    d, m, K, C = 24, 6, 16, 9
    coarse = rng.standard_normal((C, d)).astype(np.float32) * 0.2
    coarse[5] = coarse[2]                                       # duplicate centroid: the first one wins
    cb = rng.standard_normal((m, K, d // m)).astype(np.float32) * 0.2
    cb[:, 9] = cb[:, 4]                                         # duplicate codewords
    eng.load_coarse(coarse)
    eng.load_codebook(_lib.FB_CB_RESIDUAL, cb)
    v = rng.standard_normal((700, d)).astype(np.float32) * 0.2
    v[:100] = coarse[2] + cb[:, 4].reshape(-1)                  # exactly centroid 2 + codeword 4
    cids, codes = eng.encode_ivfadc(v)
    ecids, ecodes, rc = oracle_mod.encode(v, cb, coarse)
    assert rc == 0
    np.testing.assert_array_equal(cids, ecids)
    np.testing.assert_array_equal(codes, ecodes)
    assert (cids[:100] == 2).all()
    far = np.full((3, d), 50.0, np.float32)                     # every coarse distance >= 100: undefined in the reference
    with pytest.raises(FreddyError) as ei:
        eng.encode_ivfadc(far)
    assert ei.value.code == _lib.FB_ERR_REFERENCE_UB
    assert oracle_mod.encode(far, cb, coarse)[2] != 0


def test_encode_device_pointers_and_index_builder(eng, oracle_mod):
    """the device-pointer form (index build: rows already in HBM) gives the host form's codes; the synthetic index
    builder with encoder= produces an index whose every row is quantised by the reference's rule"""
    import torch
    from freddy_b200 import FreddyError, _lib
    from freddy_b200.index_build import make_synthetic_index
    ix = small_index(N=20000, d=300, m=12, K=1024, C=100, seed=1, n_clusters=100, with_pq=True)
    eng.load_coarse(ix["coarse"])
    eng.load_codebook(_lib.FB_CB_RESIDUAL, ix["residual_codebook"])
    eng.load_codebook(_lib.FB_CB_PQ, ix["pq_codebook"])
    v = np.ascontiguousarray(ix["vectors"][:18001])                 # two device chunks, the second ragged
    cids, codes = eng.encode_ivfadc(v)
    tv = torch.from_numpy(v).cuda()
    tc = torch.empty(len(v), dtype=torch.int32, device="cuda")
    tk = torch.empty(len(v), 12, dtype=torch.int16, device="cuda")
    torch.cuda.synchronize()
    eng.encode_ivfadc_dev(tv.data_ptr(), len(v), tc.data_ptr(), tk.data_ptr())
    eng.synchronize()
    np.testing.assert_array_equal(tc.cpu().numpy(), cids)
    np.testing.assert_array_equal(tk.cpu().numpy(), codes)
    eng.encode_pq_dev(tv.data_ptr(), len(v), tk.data_ptr())
    eng.synchronize()
    np.testing.assert_array_equal(tk.cpu().numpy(), eng.encode_pq(v))
    far = torch.full((5, 300), 50.0, device="cuda")
    with pytest.raises(FreddyError) as ei:
        eng.encode_ivfadc_dev(far.data_ptr(), 5, tc.data_ptr(), tk.data_ptr())
    assert ei.value.code == _lib.FB_ERR_REFERENCE_UB
    ix2 = make_synthetic_index(30000, d=48, m=12, K=64, C=40, n_train=20000, n_clusters=50, kmeans_iters=3, seed=5, device="cuda",
                               with_pq=True, keep_vectors=True, encoder=eng)
    vec = ix2.pop("vectors_t").cpu().numpy()
    ecids, ecodes, rc = oracle_mod.encode(vec, ix2["residual_codebook"], ix2["coarse"])
    assert rc == 0
    np.testing.assert_array_equal(ix2["coarse_ids"], ecids)
    np.testing.assert_array_equal(ix2["codes"], ecodes)
    _, epq, rc = oracle_mod.encode(vec, ix2["pq_codebook"])
    np.testing.assert_array_equal(ix2["pq_codes"], epq)
