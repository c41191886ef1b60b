"""Host logic of the bank-conflict-aware row placement (engine.cu: place_rows_of_list), no GPU:
the slot order is a permutation, and the modelled gather cost (max over the 32 shared-memory banks of
the distinct codes per position, the rule scripts/microbench_smem.cu measured) goes down."""
import ctypes as C

import numpy as np

from freddy_b200 import _lib


def gather_cost(codes):
    """mean data-pipe cycles per warp gather for rows laid out 32 per block in the given order"""
    n, m = codes.shape
    tot, cnt = 0, 0
    for b in range(0, n - 31, 32):
        blk = codes[b:b + 32]
        for p in range(m):
            uniq = np.unique(blk[:, p])
            tot += np.bincount(uniq % 32, minlength=32).max()
            cnt += 1
    return tot / cnt


def _order(codes, K, window):
    lib = _lib.load()
    codes = np.ascontiguousarray(codes, np.int16)
    out = np.empty(len(codes), np.int32)
    rc = lib.fb_placement_order(codes.ctypes.data_as(C.c_void_p), len(codes), codes.shape[1], K, window,
                                out.ctypes.data_as(C.c_void_p))
    assert rc == 0
    return out


def test_placement_is_a_permutation_and_cuts_conflicts():
    rng = np.random.default_rng(0)
    codes = rng.integers(0, 1024, size=(4000, 12)).astype(np.int16)
    base = gather_cost(codes)
    order = _order(codes, 1024, 128)
    assert sorted(order.tolist()) == list(range(len(codes)))
    placed = gather_cost(codes[order])
    assert base > 3.2                      # 32 random codes over 32 banks: about 3.5
    assert placed < 0.7 * base, (base, placed)
    # window <= 1 and short lists keep arrival order
    assert (_order(codes, 1024, 0) == np.arange(len(codes))).all()
    assert (_order(codes[:20], 1024, 128) == np.arange(20)).all()


def test_placement_duplicates_and_small_K():
    rng = np.random.default_rng(1)
    codes = rng.integers(0, 4, size=(1000, 12)).astype(np.int16)       # K=4: every gather is conflict-free
    order = _order(codes, 4, 64)
    assert sorted(order.tolist()) == list(range(1000))
    lib = _lib.load()
    bad = np.full((40, 12), 7, np.int16)
    out = np.empty(40, np.int32)
    assert lib.fb_placement_order(bad.ctypes.data_as(C.c_void_p), 40, 12, 4, 64, out.ctypes.data_as(C.c_void_p)) != 0
