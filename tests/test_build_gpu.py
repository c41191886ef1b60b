"""Pinning a table on the device (FB_OPT_DEVICE_BUILD, build_kernels.cuh): the kernels that group rows by list, place
them and pack them produce the SAME layout as the host-side builder (checksums over slots and code units), the same
search results, the same argument errors; create_statistics (freddy--0.0.1.sql:150-171) on the device equals its
restatement in oracle/oracle.py."""
import numpy as np
import pytest

from helpers import small_index, queries_from, assert_same_topk

pytestmark = pytest.mark.gpu


def _engine(device_build, window=None):
    from freddy_b200 import Engine, _lib
    e = Engine(0)
    e.set_option(_lib.FB_OPT_DEVICE_BUILD, 1 if device_build else 0)
    if window is not None:
        e.set_option(_lib.FB_OPT_PLACEMENT_WINDOW, window)
    return e


@pytest.mark.parametrize("shape,window", [
    (dict(N=20000, d=48, m=12, K=64, C=40, seed=7), None),             # byte-code image too (K <= 256)
    (dict(N=20000, d=48, m=12, K=64, C=40, seed=7), 0),                # arrival order
    (dict(N=20000, d=48, m=12, K=64, C=40, seed=7), 2),
    (dict(N=20000, d=48, m=12, K=64, C=40, seed=7), 1000),             # window wider than most lists
    (dict(N=30000, d=300, m=12, K=1024, C=100, seed=1, n_clusters=100), None),
    (dict(N=3000, d=48, m=12, K=64, C=400, seed=3), None),             # most lists shorter than a block, some empty
    (dict(N=6000, d=96, m=48, K=16, C=7, seed=5), 64),                 # m > 16: no byte image
])
def test_device_build_gives_the_host_layout(shape, window):
    ix = small_index(with_pq=True, **shape)
    sums, results = [], []
    q = queries_from(ix, 64, noise=0.02)
    for dev in (False, True):
        e = _engine(dev, window)
        e.load_ivfadc_index(ix)
        e.load_pq_index(ix)
        sums.append((e.table_checksum(0), e.table_checksum(1)))
        results.append(e.ivfadc_search(q, 5, 4) + e.pq_search(q, 5))
        e.close()
    assert sums[0] == sums[1]
    for a, b in zip(results[0], results[1]):
        np.testing.assert_array_equal(a, b)


def test_device_build_unsorted_repeated_ids_and_subsets():
    """the id index is sorted on the device: `WHERE id IN (...)` over a table whose ids are neither sorted nor unique"""
    ix = small_index(N=8000, d=48, m=12, K=64, C=20, seed=11, with_pq=True)
    rng = np.random.default_rng(0)
    ids = rng.permutation(8000).astype(np.int32)
    ids[100:140] = ids[10:50]                                   # repeated ids
    ix = dict(ix, ids=ids)
    q = queries_from(ix, 16, noise=0.02)
    targets = np.concatenate([ids[rng.choice(8000, 900, replace=False)], np.asarray([ids[10], ids[10], 999999], np.int32)])
    out = []
    for dev in (False, True):
        e = _engine(dev)
        e.load_pq_index(ix)
        out.append(e.pq_search_in_batch(q, 5, targets))
        e.close()
    for a, b in zip(out[0], out[1]):
        np.testing.assert_array_equal(a, b)


def test_device_build_argument_errors_match():
    from freddy_b200 import FreddyError
    ix = small_index(N=5000, d=48, m=12, K=64, C=20, seed=2)
    msgs = []
    for dev in (False, True):
        e = _engine(dev)
        bad = dict(ix, coarse_ids=ix["coarse_ids"].copy())
        bad["coarse_ids"][[4000, 1234]] = [77, -1]
        with pytest.raises(FreddyError) as e1:
            e.load_ivfadc_index(bad)
        bad = dict(ix, codes=ix["codes"].copy())
        bad["codes"][3000, 5] = 64
        bad["codes"][2999, 7] = -3
        with pytest.raises(FreddyError) as e2:
            e.load_ivfadc_index(bad)
        msgs.append((str(e1.value), str(e2.value)))
        e.close()
    assert msgs[0] == msgs[1]
    assert "row 1234" in msgs[0][0] and "row 2999 pos 7" in msgs[0][1]


def test_statistics_on_the_device(oracle_mod):
    from freddy_b200.index_build import make_ivpq_index
    ix = small_index(N=12000, d=48, m=12, K=64, C=20, seed=9)
    import torch
    ivpq = make_ivpq_index(torch.from_numpy(ix["vectors"]), m=12, K=32, Kc=8, n_train=12000, kmeans_iters=3, seed=1)
    cells = 64
    e = _engine(True)
    e.load_ivpq_index(dict(ivpq, stats=None))                   # statistics over all rows, computed on the device
    got_all = e.ivpq_statistics(None)
    exp_all = oracle_mod.create_statistics(ivpq["ids"], ivpq["ivpq_coarse_ids"], None, cells)
    np.testing.assert_array_equal(got_all.view(np.uint32), exp_all.view(np.uint32))
    rng = np.random.default_rng(3)
    listed = np.concatenate([rng.choice(ivpq["ids"], 1777, replace=False), ivpq["ids"][:300], ivpq["ids"][:100],
                             np.asarray([10 ** 8, -5])]).astype(np.int32)       # repeats count, unknown ids do not
    got = e.ivpq_statistics(listed, install=True)
    exp = oracle_mod.create_statistics(ivpq["ids"], ivpq["ivpq_coarse_ids"], listed, cells)
    np.testing.assert_array_equal(got.view(np.uint32), exp.view(np.uint32))
    assert got[cells] == len(listed) - 2                        # ids are unique in this table: every known listed id matches one row
    # the join with installed statistics equals a second engine that was handed the same statistics
    q = queries_from(ix, 24, noise=0.02)
    targets = rng.choice(ivpq["ids"], 3000, replace=False).astype(np.int32)
    e.load_vectors(ivpq["ids"], ix["vectors"])
    a = e.ivpq_search_in(q, 5, targets, 10, 4, 2, True, 0.8)
    e2 = _engine(False)
    e2.load_ivpq_index(dict(ivpq, stats=exp))
    e2.load_vectors(ivpq["ids"], ix["vectors"])
    b = e2.ivpq_search_in(q, 5, targets, 10, 4, 2, True, 0.8)
    for x, y in zip(a, b):
        np.testing.assert_array_equal(x, y)
    from freddy_b200 import FreddyError
    with pytest.raises(FreddyError):
        e.ivpq_statistics(np.asarray([10 ** 8], np.int32))      # nothing matches: the reference divides by zero
    e.close(); e2.close()
