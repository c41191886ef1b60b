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
