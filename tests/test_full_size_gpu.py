"""Parity at BASELINE.json's full size (3M x 300, m=12, K=1024, C=1000, k=5, w=10 — the index bench.py times).
The oracle finishes a sample of the queries in seconds; the whole batch is checked through properties that do
not need it: the three CUDA paths (pipeline kernel, separate LUT/scan kernels, literal general kernel) must agree
bit for bit, results are independent of the stored row placement and of chunking, repeated runs are identical,
distances ascend, every returned row lies in one of the probed lists and re-computing its ADC distance from the
tables gives the returned bits."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

N, D, M, K, C, KNN, W = 3_000_000, 300, 12, 1024, 1000, 5, 10


@pytest.fixture(scope="module")
def big():
    import torch
    from freddy_b200 import Engine
    from freddy_b200.index_build import make_synthetic_index
    ix = make_synthetic_index(N, d=D, m=M, K=K, C=C, n_train=100_000, n_clusters=1000, sigma=1.0, zipf=0.35,
                              kmeans_iters=10, seed=1234, device="cuda", keep_vectors=True)
    vec = ix.pop("vectors_t")
    g = torch.Generator(); g.manual_seed(4321)
    sel = torch.randperm(N, generator=g)[:3000]
    q = vec[sel.to(vec.device)].cpu().numpy()
    del vec
    torch.cuda.empty_cache()
    e = Engine(0)
    e.load_ivfadc_index(ix)
    yield ix, q, e
    e.close()


def _run(e, q, **opts):
    from freddy_b200 import _lib
    names = {"pipeline": _lib.FB_OPT_PIPELINE, "force_exact": _lib.FB_OPT_FORCE_EXACT_PATH, "pipe_chunk": _lib.FB_OPT_PIPE_CHUNK,
             "ramp": _lib.FB_OPT_PIPE_RAMP}
    defaults = {"pipeline": 1, "force_exact": 0, "pipe_chunk": 2048, "ramp": 0}
    for k_, v in {**defaults, **opts}.items():
        e.set_option(names[k_], v)
    try:
        e.reset_counters()
        ids, d = e.ivfadc_search(q, KNN, W)
        return ids, d, e.counters()
    finally:
        for k_, v in defaults.items():
            e.set_option(names[k_], v)


def test_full_size_paths_agree_and_match_the_oracle(big, oracle_mod):
    ix, q, e = big
    ids, d, c = _run(e, q)
    assert c["n_pipe_launches"] >= 2 and c["queries"] == len(q)
    # oracle (freddy.c:247-378 restated, pinned to the reference) on a sample
    ns = 48
    eids, ed, rc, rows = oracle_mod.OracleIndex(ix).ivfadc_search(q[:ns], KNN, W, threads=16)
    assert rc == 0
    np.testing.assert_array_equal(ids[:ns], eids)
    np.testing.assert_array_equal(d[:ns].view(np.uint32), ed.view(np.uint32))
    # idempotence, chunking, ramp
    for opts in ({}, {"pipe_chunk": 700}, {"pipe_chunk": 1024, "ramp": 1}):
        i2, d2, _ = _run(e, q, **opts)
        np.testing.assert_array_equal(i2, ids)
        np.testing.assert_array_equal(d2.view(np.uint32), d.view(np.uint32))
    # separate LUT / scan kernels
    i3, d3, c3 = _run(e, q, pipeline=0)
    assert c3["n_pipe_launches"] == 0 and c3["rows_scanned"] == c["rows_scanned"]
    np.testing.assert_array_equal(i3, ids)
    np.testing.assert_array_equal(d3.view(np.uint32), d.view(np.uint32))
    # literal general kernel (reference control flow) on a slice
    i4, d4, c4 = _run(e, q[:256], force_exact=1)
    assert c4["exact_path_queries"] == 256
    np.testing.assert_array_equal(i4, ids[:256])
    np.testing.assert_array_equal(d4.view(np.uint32), d[:256].view(np.uint32))


def test_full_size_result_properties(big):
    ix, q, e = big
    ids, d, _ = _run(e, q)
    assert (ids >= 1).all() and (np.diff(d, axis=1) >= 0).all()
    assert all(len(set(r)) == KNN for r in ids)
    # every winner lies in one of the w nearest lists, and its ADC distance recomputed on the host from the
    # tables (fp32, reference order of operations) is the returned value
    coarse, cb = ix["coarse"], ix["residual_codebook"]
    rows = ids - 1                                               # ids are 1-based table positions
    for qi in range(0, len(q), 97):
        cd = np.zeros(C, np.float32)
        for i in range(D):                                       # squareDistance: sequential fp32
            t = (q[qi, i] - coarse[:, i]).astype(np.float32)
            cd = (cd + (t * t).astype(np.float32)).astype(np.float32)
        probed = set(np.argsort(cd, kind="stable")[:W].tolist())
        for r, dist in zip(rows[qi], d[qi]):
            cid = int(ix["coarse_ids"][r])
            assert cid in probed
            res = (q[qi] - coarse[cid]).astype(np.float32)
            acc = np.float32(0)
            for p in range(M):
                cw = cb[p, int(ix["codes"][r, p])]
                s = np.float32(0)
                for i in range(D // M):
                    t = np.float32(res[p * (D // M) + i] - cw[i])
                    s = np.float32(s + np.float32(t * t))
                acc = np.float32(acc + s)
            assert acc.view(np.uint32) == dist.view(np.uint32)


def test_full_size_placement_does_not_change_results(big):
    from freddy_b200 import Engine, _lib
    ix, q, e = big
    ids, d, _ = _run(e, q[:1500])
    e2 = Engine(0)
    try:
        e2.set_option(_lib.FB_OPT_PLACEMENT_WINDOW, 0)          # rows stored in arrival order
        e2.load_ivfadc_index(ix)
        i2, d2 = e2.ivfadc_search(q[:1500], KNN, W)
    finally:
        e2.close()
    np.testing.assert_array_equal(i2, ids[:1500])
    np.testing.assert_array_equal(d2.view(np.uint32), d[:1500].view(np.uint32))
