"""GPU parity of the warp-specialised pipeline kernel (pipeline_kernels.cuh): LUT build of chunk
c+1 and ADC scan of chunk c inside one launch.  Bar: ids, ranks and distance bits equal to the
oracle's (freddy.c:247-378), and equal to the separate-kernel path of the same engine."""
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


def _search(eng, q, k, w, pipeline, pipe_chunk=2048, shape=0, ramp=0):
    from freddy_b200 import _lib
    eng.set_option(_lib.FB_OPT_PIPELINE, pipeline)
    eng.set_option(_lib.FB_OPT_PIPE_CHUNK, pipe_chunk)
    eng.set_option(_lib.FB_OPT_PIPE_SHAPE, shape)
    eng.set_option(_lib.FB_OPT_PIPE_RAMP, ramp)
    eng.reset_counters()
    try:
        ids, d = eng.ivfadc_search(q, k, w)
        c = eng.counters()
    finally:
        eng.set_option(_lib.FB_OPT_PIPELINE, 1)
        eng.set_option(_lib.FB_OPT_PIPE_CHUNK, 2048)
        eng.set_option(_lib.FB_OPT_PIPE_SHAPE, 0)
        eng.set_option(_lib.FB_OPT_PIPE_RAMP, 0)
    return ids, d, c


@pytest.mark.parametrize("K", [1024, 256])
def test_pipeline_parity(eng, oracle_mod, K):
    ix = small_index(N=60000, d=300, m=12, K=K, C=100, seed=1, n_clusters=100)
    eng.load_ivfadc_index(ix)
    oi = oracle_mod.OracleIndex(ix)
    q = queries_from(ix, 1300, seed=6, noise=0.02)
    cases = ((5, 10, 1024, 0, 0), (5, 10, 300, 1, 1), (1, 1, 1024, 2, 0), (7, 3, 200, 0, 1), (30, 7, 1024, 1, 0),
             (5, 16, 500, 2, 1))
    for k, w, chunk, shape, ramp in cases:
        ids, d, c = _search(eng, q, k, w, 1, chunk, shape if K == 1024 else 0, ramp)
        assert c["n_pipe_launches"] >= 2, "pipeline kernel did not run"
        eids, ed, rc, rows = oi.ivfadc_search(q, k, w, threads=8)
        assert rc == 0
        assert_same_topk(ids, d, eids, ed, f"pipeline K={K} k={k} w={w} chunk={chunk} shape={shape} ramp={ramp}")
        assert c["rows_scanned"] == rows
        ids2, d2, c2 = _search(eng, q, k, w, 0)
        assert c2["n_pipe_launches"] == 0
        assert_same_topk(ids2, d2, ids, d, "separate kernels vs pipeline")


def test_pipeline_ties_and_few_rows(eng, oracle_mod):
    """duplicate-heavy table (ties across the k-th place) and lists shorter than k: flagged queries
    leave the pipeline for the general kernel exactly as they leave the separate-kernel path"""
    ix = small_index(N=4000, d=300, m=12, K=256, C=100, seed=4, n_clusters=3)
    # collapse the codes so that many rows share a distance
    ix = dict(ix)
    codes = ix["codes"].copy()
    codes[:, 2:] = codes[:, 2:] % 2
    ix["codes"] = codes
    eng.load_ivfadc_index(ix)
    oi = oracle_mod.OracleIndex(ix)
    q = queries_from(ix, 700, seed=2, noise=0.01)
    flagged = 0
    for k, w in ((5, 10), (20, 2), (30, 1)):
        ids, d, c = _search(eng, q, k, w, 1, 256)
        assert c["n_pipe_launches"] >= 2
        eids, ed, rc, _ = oi.ivfadc_search(q, k, w, threads=8)
        assert rc == 0
        assert_same_topk(ids, d, eids, ed, f"pipeline ties k={k} w={w}")
        flagged += c["exact_path_queries"]
    assert flagged > 0


def test_pipeline_repeatable(eng):
    """dynamic query->CTA assignment must not change results"""
    ix = small_index(N=60000, d=300, m=12, K=1024, C=100, seed=1, n_clusters=100)
    eng.load_ivfadc_index(ix)
    q = queries_from(ix, 2000, seed=9)
    ref = None
    for _ in range(4):
        ids, d, _ = _search(eng, q, 5, 10, 1, 512)
        if ref is None:
            ref = (ids.copy(), d.copy())
        np.testing.assert_array_equal(ids, ref[0])
        np.testing.assert_array_equal(d.view(np.uint32), ref[1].view(np.uint32))


def test_host_buffer_call_matches_device_path(eng):
    """fb_ivfadc_search (host buffers) and fb_ivfadc_search_dev (device pointers) give the same bits"""
    import torch
    ix = small_index(N=60000, d=300, m=12, K=1024, C=100, seed=1, n_clusters=100)
    eng.load_ivfadc_index(ix)
    q = queries_from(ix, 5000, seed=12, noise=0.02)
    for _ in range(2):
        ids, d = eng.ivfadc_search(q, 5, 10)
        dq = torch.from_numpy(q).cuda()
        oi = torch.empty(len(q), 5, dtype=torch.int32, device="cuda")
        od = torch.empty(len(q), 5, dtype=torch.float32, device="cuda")
        eng.ivfadc_search_dev(dq.data_ptr(), len(q), 5, 10, oi.data_ptr(), od.data_ptr())
        eng.synchronize()
        np.testing.assert_array_equal(ids, oi.cpu().numpy())
        np.testing.assert_array_equal(d.view(np.uint32), od.cpu().numpy().view(np.uint32))
        # pinned host buffers: the coarse kernel reads the queries through the mapped pointer (zero-copy upload)
        hq = torch.from_numpy(q).pin_memory()
        hi = torch.empty(len(q), 5, dtype=torch.int32).pin_memory()
        hd = torch.empty(len(q), 5, dtype=torch.float32).pin_memory()
        eng.ivfadc_search_ptr(hq.data_ptr(), len(q), 5, 10, hi.data_ptr(), hd.data_ptr())
        np.testing.assert_array_equal(hi.numpy(), ids)
        np.testing.assert_array_equal(hd.numpy().view(np.uint32), d.view(np.uint32))


def test_byte_code_table_k256(eng, oracle_mod):
    """K = 256 (index_creation/config/ivfadc_config.json): the scan kernels read the true uint8 image of the codes
    (16 bytes per row); results equal the 16-bit path and the oracle, scan bytes are accounted at m + 4 per row"""
    from freddy_b200 import _lib
    ix = small_index(N=60000, d=300, m=12, K=256, C=100, seed=1, n_clusters=100)
    eng.load_ivfadc_index(ix)
    oi = oracle_mod.OracleIndex(ix)
    for nq, k, w in ((1300, 5, 10), (200, 7, 4), (3, 5, 6)):     # pipeline kernel / one CTA per query / one CTA per (query, list)
        q = queries_from(ix, nq, seed=nq, noise=0.02)
        eids, ed, rc, rows = oi.ivfadc_search(q, k, w, threads=8)
        assert rc == 0
        eng.set_option(_lib.FB_OPT_BYTE_CODES, 1)
        ids, d, c = _search(eng, q, k, w, 1)
        assert_same_topk(ids, d, eids, ed, f"byte codes nq={nq}")
        assert c["rows_scanned"] == rows
        if nq >= 64:
            assert c["scan_bytes"] == rows * (12 + 4)
        eng.set_option(_lib.FB_OPT_BYTE_CODES, 0)
        try:
            ids2, d2, c2 = _search(eng, q, k, w, 1)
        finally:
            eng.set_option(_lib.FB_OPT_BYTE_CODES, 1)
        assert_same_topk(ids2, d2, eids, ed, f"16-bit units nq={nq}")
        assert c2["scan_bytes"] == rows * (2 * 12 + 4)
