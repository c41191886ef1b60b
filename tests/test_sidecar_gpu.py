"""The sidecar over the real engine (fb_sidecar_start / fb_sidecar_stop): concurrent single-query callers that hold no
CUDA context get, bit for bit, what a direct fb_ivfadc_search call returns, and their requests share launches."""
import ctypes as C
import os
import threading

import numpy as np
import pytest

from helpers import small_index, queries_from

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_sidecar_answers_equal_direct_calls():
    from freddy_b200 import Engine
    ix = small_index(N=20000, d=48, m=12, K=64, C=40, seed=7)
    q = queries_from(ix, 400, noise=0.02)
    eng = Engine(0)
    eng.load_ivfadc_index(ix)
    exp = {(k, w): eng.ivfadc_search(q, k, w) for (k, w) in ((5, 4), (3, 2))}
    name = f"/fbsc_gpu_{os.getpid()}"
    eng.sidecar_start(name, max_k=8, slots=16, max_batch=16)
    lib = C.CDLL(os.path.join(ROOT, "postgres-word2vec_b200", "libfreddy_sidecar.so"))
    lib.fbsc_client_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.fbsc_client_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    lib.fbsc_client_close.argtypes = [C.c_void_p]
    bad = []

    def caller(t):
        h = C.c_void_p()
        assert lib.fbsc_client_open(name.encode(), C.byref(h)) == 0
        k, w = ((5, 4), (3, 2))[t % 2]
        ids, dists = np.empty(k, np.int32), np.empty(k, np.float32)
        for i in range(t, 400, 8):
            qi = np.ascontiguousarray(q[i])
            rc = lib.fbsc_client_search(h, qi.ctypes.data_as(C.c_void_p), k, w, ids.ctypes.data_as(C.c_void_p),
                                        dists.ctypes.data_as(C.c_void_p), 20000)
            if rc != 0 or not np.array_equal(ids, exp[(k, w)][0][i]) or \
                    not np.array_equal(dists.view(np.uint32), exp[(k, w)][1][i].view(np.uint32)):
                bad.append((t, i, rc))
        lib.fbsc_client_close(h)

    threads = [threading.Thread(target=caller, args=(t,)) for t in range(8)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    c = eng.sidecar_stop()
    assert not bad, bad[:5]
    assert c["queries"] == 400
    assert c["batches"] < 400 and c["largest_batch"] > 1
    # the engine is the owner's again
    ids, _ = eng.ivfadc_search(q[:4], 5, 4)
    np.testing.assert_array_equal(ids, exp[(5, 4)][0][:4])
    eng.close()
