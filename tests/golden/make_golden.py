"""Generates tests/golden/kernels_golden.json from the reference's OWN compiled
arithmetic (oracle/_ref/libfreddy_ref.so = /root/reference/freddy_extension/
index_utils.c built unmodified, see oracle/Makefile).  Run in the build container,
where /root/reference exists:   python tests/golden/make_golden.py
The JSON is committed so the GPU box (no /root/reference) can still pin the oracle.
"""
import ctypes as C
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle import oracle  # noqa: E402


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class CodebookEntry(C.Structure):
    _fields_ = [("pos", C.c_int), ("code", C.c_int), ("vector", C.c_void_p)]


def main():
    oracle.build()
    R = oracle.ref_lib()
    assert R is not None, "build oracle/_ref first (make -C oracle ref)"
    rng = np.random.default_rng(20260925)
    out = {"source": "oracle/_ref/libfreddy_ref.so (reference index_utils.c, gcc -O2 -ffp-contract=off)",
           "square_distance": [], "topk": [], "lut_adc": []}
    for n in (1, 3, 25, 150, 300):
        for _ in range(4):
            a = rng.standard_normal(n).astype(np.float32)
            b = rng.standard_normal(n).astype(np.float32)
            r = np.float32(R.squareDistance(_p(a), _p(b), n))
            out["square_distance"].append({"a_bits": a.view(np.uint32).tolist(), "b_bits": b.view(np.uint32).tolist(),
                                           "out_bits": int(r.view(np.uint32))})
    for _ in range(40):
        k = int(rng.integers(1, 8))
        n = int(rng.integers(1, 30))
        stream = (rng.integers(0, 6, size=n) / 4).tolist()
        tk = (oracle.TopKEntry * k)()
        for i in range(k):
            tk[i].id, tk[i].distance = -1, 1000.0
        for i, dist in enumerate(stream):
            if dist < tk[k - 1].distance:
                R.updateTopK(tk, float(dist), i, k, 0)
        out["topk"].append({"k": k, "stream": stream, "out": [[e.id, e.distance] for e in tk]})
    for (m, K, sub) in ((12, 16, 25), (4, 8, 3)):
        cb = rng.standard_normal((m, K, sub)).astype(np.float32)
        q = rng.standard_normal(m * sub).astype(np.float32)
        ents = (CodebookEntry * (m * K))()
        for idx in range(m * K):
            p, c = divmod(idx, K)
            ents[idx].pos, ents[idx].code, ents[idx].vector = p, c, cb[p, c].ctypes.data
        lut = np.empty(m * K, np.float32)
        R.getPrecomputedDistances(_p(lut), m, K, sub, _p(q), ents)
        codes = rng.integers(0, K, size=(8, m)).astype(np.int16)
        adc = [int(np.float32(R.computePQDistanceInt16(_p(lut), _p(codes[i]), m, K)).view(np.uint32)) for i in range(8)]
        out["lut_adc"].append({"m": m, "K": K, "sub": sub, "cb_bits": cb.view(np.uint32).ravel().tolist(),
                               "q_bits": q.view(np.uint32).tolist(), "lut_bits": lut.view(np.uint32).tolist(),
                               "codes": codes.tolist(), "adc_bits": adc})
    with open(os.path.join(HERE, "kernels_golden.json"), "w") as f:
        json.dump(out, f)
    print("wrote kernels_golden.json")


if __name__ == "__main__":
    main()
