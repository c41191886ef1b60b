"""The CUDA-free half of the sidecar (include/freddy_sidecar.h, postgres-word2vec_b200/sidecar/freddy_sidecar.c):
request slots in shared memory, batching of whatever is pending, error paths.  The batch function here is a Python
stub, so no GPU is needed; tests/test_sidecar_gpu.py runs the same protocol over fb_ivfadc_search."""
import ctypes as C
import multiprocessing as mp
import os
import threading
import time

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "postgres-word2vec_b200", "libfreddy_sidecar.so")
BATCH_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(C.c_float), C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_float))


def _lib():
    lib = C.CDLL(LIB)
    lib.fbsc_server_create.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]
    lib.fbsc_server_run.argtypes = [C.c_void_p, BATCH_FN, C.c_void_p, C.c_int, C.c_int]
    lib.fbsc_server_stop.argtypes = [C.c_void_p]
    lib.fbsc_server_stop.restype = None
    lib.fbsc_server_destroy.argtypes = [C.c_void_p]
    lib.fbsc_server_destroy.restype = None
    lib.fbsc_server_counters.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    lib.fbsc_server_counters.restype = None
    lib.fbsc_client_open.argtypes = [C.c_char_p, C.POINTER(C.c_void_p)]
    lib.fbsc_client_search.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    lib.fbsc_client_close.argtypes = [C.c_void_p]
    lib.fbsc_client_close.restype = None
    lib.fbsc_client_dim.argtypes = [C.c_void_p]
    return lib


D = 16


def _stub(ctx, q, nq, k, w, ids, dists):
    """ids[i][j] = round(q[i][0]) * 100 + j + w, dists[i][j] = q[i][1] + j; k == 7 is 'an engine error'"""
    if k == 7:
        return -42
    time.sleep(0.0005)                              # a launch takes time: requests pile up meanwhile
    for i in range(nq):
        for j in range(k):
            ids[i * k + j] = int(round(q[i * D])) * 100 + j + w
            dists[i * k + j] = q[i * D + 1] + j
    return 0


def _client(name, seed, n, out_q):
    lib = _lib()
    h = C.c_void_p()
    rc = lib.fbsc_client_open(name, C.byref(h))
    if rc != 0:
        out_q.put(("open", rc)); return
    assert lib.fbsc_client_dim(h) == D
    rng = np.random.default_rng(seed)
    bad = 0
    for it in range(n):
        k = int(rng.integers(1, 6))
        w = int(rng.integers(1, 4))
        q = rng.standard_normal(D).astype(np.float32)
        q[0] = seed * 50 + it % 50
        ids, dists = np.empty(k, np.int32), np.empty(k, np.float32)
        rc = lib.fbsc_client_search(h, q.ctypes.data_as(C.c_void_p), k, w, ids.ctypes.data_as(C.c_void_p),
                                    dists.ctypes.data_as(C.c_void_p), 20000)
        exp_ids = int(round(float(q[0]))) * 100 + np.arange(k) + w
        exp_d = (q[1] + np.arange(k, dtype=np.float32)).astype(np.float32)
        if rc != 0 or not (ids == exp_ids).all() or not np.allclose(dists, exp_d):
            bad += 1
    lib.fbsc_client_close(h)
    out_q.put(("done", bad))


@pytest.mark.timeout(120)
def test_sidecar_batches_concurrent_single_query_callers():
    lib = _lib()
    name = f"/fbsc_test_{os.getpid()}".encode()
    srv = C.c_void_p()
    assert lib.fbsc_server_create(name, D, 8, 6, C.byref(srv)) == 0      # 6 slots for 8 callers: slot contention too
    ctx = mp.get_context("spawn")
    out_q = ctx.Queue()
    procs = [ctx.Process(target=_client, args=(name, s, 150, out_q)) for s in range(8)]
    for p in procs:
        p.start()
    cb = BATCH_FN(_stub)
    t = threading.Thread(target=lambda: lib.fbsc_server_run(srv, cb, None, 64, 0))
    t.start()
    res = [out_q.get(timeout=100) for _ in procs]
    for p in procs:
        p.join(10)
    # in-process caller: an engine error reaches the caller, an oversized k is refused locally
    h = C.c_void_p()
    assert lib.fbsc_client_open(name, C.byref(h)) == 0
    q = np.zeros(D, np.float32)
    ids, dists = np.empty(9, np.int32), np.empty(9, np.float32)
    assert lib.fbsc_client_search(h, q.ctypes.data_as(C.c_void_p), 7, 1, ids.ctypes.data_as(C.c_void_p), dists.ctypes.data_as(C.c_void_p), 5000) == -42
    assert lib.fbsc_client_search(h, q.ctypes.data_as(C.c_void_p), 9, 1, ids.ctypes.data_as(C.c_void_p), dists.ctypes.data_as(C.c_void_p), 5000) == -1
    lib.fbsc_server_stop(srv)
    t.join(10)
    assert not t.is_alive()
    b, n, big = C.c_int64(), C.c_int64(), C.c_int64()
    lib.fbsc_server_counters(srv, C.byref(b), C.byref(n), C.byref(big))
    assert all(r == ("done", 0) for r in res), res
    assert n.value == 8 * 150 + 1
    assert big.value > 1 and b.value < n.value, "requests that arrived during a launch must share the next batch"
    lib.fbsc_server_destroy(srv)
    # the segment is gone: a waiting caller learns it, a new one cannot open it
    assert lib.fbsc_client_search(h, q.ctypes.data_as(C.c_void_p), 3, 1, ids.ctypes.data_as(C.c_void_p), dists.ctypes.data_as(C.c_void_p), 2000) == -3
    lib.fbsc_client_close(h)
    h2 = C.c_void_p()
    assert lib.fbsc_client_open(name, C.byref(h2)) == -3


@pytest.mark.timeout(60)
def test_sidecar_busy_slots_remote_stop():
    """one slot, no server yet: a second caller gets FBSC_ERR_BUSY after its timeout; a client can ask the server to stop"""
    lib = _lib()
    lib.fbsc_client_request_stop = lib.fbsc_client_request_stop
    lib.fbsc_client_request_stop.argtypes = [C.c_void_p]
    name = f"/fbsc_busy_{os.getpid()}".encode()
    srv = C.c_void_p()
    assert lib.fbsc_server_create(name, D, 4, 1, C.byref(srv)) == 0
    ha, hb = C.c_void_p(), C.c_void_p()
    assert lib.fbsc_client_open(name, C.byref(ha)) == 0 and lib.fbsc_client_open(name, C.byref(hb)) == 0
    q = np.zeros(D, np.float32)
    q[0], q[1] = 3.0, 0.5
    ids_a, d_a = np.empty(2, np.int32), np.empty(2, np.float32)
    rc_a = []
    ta = threading.Thread(target=lambda: rc_a.append(lib.fbsc_client_search(
        ha, q.ctypes.data_as(C.c_void_p), 2, 1, ids_a.ctypes.data_as(C.c_void_p), d_a.ctypes.data_as(C.c_void_p), 0)))
    ta.start()
    time.sleep(0.2)                                   # A holds the only slot, nobody serves yet
    ids_b, d_b = np.empty(2, np.int32), np.empty(2, np.float32)
    t0 = time.time()
    assert lib.fbsc_client_search(hb, q.ctypes.data_as(C.c_void_p), 2, 1, ids_b.ctypes.data_as(C.c_void_p),
                                  d_b.ctypes.data_as(C.c_void_p), 100) == -4
    assert 0.05 < time.time() - t0 < 5.0
    cb = BATCH_FN(_stub)
    ts = threading.Thread(target=lambda: lib.fbsc_server_run(srv, cb, None, 8, 0))
    ts.start()
    ta.join(20)
    assert rc_a == [0] and ids_a.tolist() == [301, 302] and np.allclose(d_a, [0.5, 1.5])
    assert lib.fbsc_client_request_stop(hb) == 0      # what freddy_sidecar_stop() does from another backend
    ts.join(20)
    assert not ts.is_alive()
    lib.fbsc_client_close(ha); lib.fbsc_client_close(hb)
    lib.fbsc_server_destroy(srv)
