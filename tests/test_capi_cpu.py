"""No-GPU checks of the C-ABI library: it loads, exports every symbol
include/freddy_b200.h declares, and fails loudly (no CPU fallback) without a device."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from freddy_b200 import _lib
    if not os.path.exists(_lib.LIB_PATH):
        import subprocess
        subprocess.run(["make", "-C", os.path.join(ROOT, "postgres-word2vec_b200", "csrc")], check=True)
    return _lib.load()


def test_header_symbols_all_exported(lib):
    from freddy_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "freddy_b200.h")).read()
    declared = set(re.findall(r"\b(fb_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"fb_engine", "fb_counters"}
    assert declared, "no declarations parsed"
    assert declared == set(_lib.SIGNATURES), (declared ^ set(_lib.SIGNATURES))
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_sidecar_header_symbols_exported_by_both_libraries(lib):
    """include/freddy_sidecar.h: the CUDA-free library a backend links (libfreddy_sidecar.so) and the engine library
    (which runs the server loop) both export every declared entry point"""
    hdr = open(os.path.join(ROOT, "include", "freddy_sidecar.h")).read()
    declared = set(re.findall(r"\b(fbsc_[a-z0-9_]+)\s*\(", hdr)) - {"fbsc_batch_fn"}
    assert len(declared) >= 11, declared
    side = C.CDLL(os.path.join(ROOT, "postgres-word2vec_b200", "libfreddy_sidecar.so"))
    for name in declared:
        assert hasattr(side, name), f"{name} missing from libfreddy_sidecar.so"
        assert hasattr(lib, name), f"{name} missing from libfreddy_b200.so"
    import subprocess
    needed = subprocess.run(["readelf", "-d", os.path.join(ROOT, "postgres-word2vec_b200", "libfreddy_sidecar.so")],
                            capture_output=True, text=True).stdout
    assert "cuda" not in needed.lower(), "the backend-side library must not depend on CUDA"


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = lib.fb_create(0, C.byref(h))
    assert rc != 0 and not h.value
    assert b"no CUDA device" in lib.fb_last_error(None)


def test_round_through_text(lib, oracle_mod):
    L = oracle_mod.lib()
    for v in (0.0, 0.1234567, 1.9999996, 1000.0, 3.4e-7):
        assert lib.fb_round_through_text(v) == L.fo_round_through_text(v)


def test_option_and_status_constants_match_the_header():
    """the ctypes mirror must use the header's numbers (FB_OPT_*, FB_ERR_*, FB_CB_*) and the counters struct must
    have the header's fields in the header's order"""
    from freddy_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "freddy_b200.h")).read()
    consts = {m.group(1): int(m.group(2)) for m in re.finditer(r"\b(FB_[A-Z0-9_]+)\s*=\s*(-?\d+)", hdr)}
    assert len(consts) > 15
    checked = 0
    for name, val in consts.items():
        if hasattr(_lib, name):
            assert getattr(_lib, name) == val, f"{name}: header {val}, _lib.py {getattr(_lib, name)}"
            checked += 1
    assert checked >= 15, checked
    body = re.search(r"typedef struct \{(.*?)\} fb_counters;", hdr, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = []
    for decl in body.split(";"):
        decl = decl.strip()
        if decl:
            fields += [f.strip() for f in decl.split(None, 1)[1].split(",")]
    assert fields == [n for n, _ in _lib.Counters._fields_], (fields, [n for n, _ in _lib.Counters._fields_])


def test_bench_reference_pool_smoke(oracle_mod):
    """bench.py's CPU arm: the reference's own SRF in spawned single-threaded backends over memory-mapped tables
    gives the oracle port's bits (tiny index; skipped when oracle/_ref is not built)"""
    import argparse
    import sys
    if not os.path.exists(oracle_mod.REF_SO):
        pytest.skip("oracle/_ref not built")
    sys.path.insert(0, ROOT)
    import bench
    from helpers import queries_from, small_index
    ix = small_index()
    a = argparse.Namespace(k=5, w=4, d=ix["d"])
    pool = bench.ReferencePool(a, ix)
    try:
        pool.procs = min(pool.procs, 4)
        q = queries_from(ix, 37, seed=3)
        ids, raw, dt = pool.run(q)
    finally:
        pool.close()
    eids, ed, rc, _ = oracle_mod.OracleIndex(ix).ivfadc_search(q, 5, 4)
    assert rc == 0 and dt > 0
    import numpy as np
    np.testing.assert_array_equal(ids, eids)
    np.testing.assert_array_equal(raw.view(np.uint32), ed.view(np.uint32))


def test_shim_defines_every_symbol_the_sql_script_binds(oracle_mod):
    """CREATE EXTENSION resolves every `AS '$libdir/freddy', '<symbol>'` of freddy--0.0.1.sql in ONE library.  The
    deployment drops freddy.c / ivpq_search_in.c for the shim, so the shim (+ the reference files that stay:
    index_utils.c, core_functions.c, cosine_similarity.c, output_utils.c) must define all 23 of them.
    tests/golden/sql_symbols.json is the list parsed from the reference's SQL script; when /root/reference is present
    the list is re-derived and must match."""
    import json
    import subprocess
    if not os.path.exists(oracle_mod.SHIM_SO):
        pytest.skip("oracle/_ref/libfreddy_shim_emul.so not built")
    want = json.load(open(os.path.join(ROOT, "tests", "golden", "sql_symbols.json")))["symbols"]
    sql = "/root/reference/freddy_extension/freddy--0.0.1.sql"
    if os.path.exists(sql):
        live = sorted(set(re.findall(r"AS '\$libdir/freddy', '([a-z0-9_]+)'", open(sql).read())))
        assert live == want
    assert len(want) == 23
    out = subprocess.run(["nm", "-D", "--defined-only", oracle_mod.SHIM_SO], capture_output=True, text=True, check=True).stdout
    defined = {line.split()[-1] for line in out.splitlines() if line.strip()}
    missing = [s for s in want if s not in defined]
    assert not missing, f"symbols the SQL script binds but the shim library lacks: {missing}"
    for extra in ("knn_exact_search", "knn_in_exact_search", "ivfadc_search_pv", "pq_search_pv", "analogy_3cosadd_batch",
                  "cosine_similarity_batch", "freddy_repin", "freddy_sidecar_serve", "freddy_sidecar_stop"):
        assert extra in defined, extra
