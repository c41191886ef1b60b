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
    srf_golden()


def srf_golden():
    """outputs of the reference's OWN SRFs (freddy.c compiled unmodified, run through
    oracle/pg_emul.c) on a tiny seeded index -> tests/golden/srf_golden.npz"""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "postgres-word2vec_b200"))
    from freddy_b200.index_build import make_synthetic_index
    ix = make_synthetic_index(3000, d=24, m=6, K=16, C=12, n_train=3000, n_clusters=8, sigma=0.5, kmeans_iters=4,
                              seed=99, device="cpu", with_pq=True, keep_vectors=True)
    vec = ix.pop("vectors_t").numpy()
    rng = np.random.default_rng(5)
    q = vec[rng.choice(len(vec), 24, replace=False)] + 0.01 * rng.standard_normal((24, 24)).astype(np.float32)
    q = np.ascontiguousarray(q, np.float32)
    S = oracle.ReferenceSession()
    S.load_ivfadc(ix, 3)
    ids_a, raw_a, txt_a = S.ivfadc_search(q, 5)
    S = oracle.ReferenceSession()
    S.load_ivfadc(ix, 1)
    ids_b, raw_b, _ = S.ivfadc_search(q, 12)
    S = oracle.ReferenceSession()
    S.load_pq(ix)
    pq_ids, pq_raw = S.pq_search(q[:6], 4)
    targets = rng.choice(np.arange(1, 3200), size=300, replace=True).astype(np.int32)
    _, in_ids, in_raw = S.pq_search_in_batch(q, np.arange(len(q), dtype=np.int32), 5, targets, False)
    np.savez_compressed(
        os.path.join(HERE, "srf_golden.npz"),
        d=ix["d"], m=ix["m"], K=ix["K"], C=ix["C"], N=ix["N"], coarse=ix["coarse"],
        residual_codebook=ix["residual_codebook"], ids=ix["ids"], coarse_ids=ix["coarse_ids"], codes=ix["codes"],
        pq_codebook=ix["pq_codebook"], pq_codes=ix["pq_codes"], queries=q, targets=targets,
        ivfadc_k5_w3_ids=ids_a, ivfadc_k5_w3_dist=raw_a, ivfadc_k5_w3_text=txt_a,
        ivfadc_k12_w1_ids=ids_b, ivfadc_k12_w1_dist=raw_b,
        pq_search_k4_ids=pq_ids, pq_search_k4_dist=pq_raw, pq_in_k5_ids=in_ids, pq_in_k5_dist=in_raw)
    print("wrote srf_golden.npz")


def srf_golden_ext():
    """second fixture (tests/golden/srf_golden_ext.npz): the reference's own grouping_pq SRF and updateCodebook
    assignments on the same tiny seeded index, plus the word-vector table they read"""
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(HERE)), "postgres-word2vec_b200"))
    from freddy_b200.index_build import make_synthetic_index
    ix = make_synthetic_index(3000, d=24, m=6, K=16, C=12, n_train=3000, n_clusters=8, sigma=0.5, kmeans_iters=4,
                              seed=99, device="cpu", with_pq=True, keep_vectors=True)
    vec = ix.pop("vectors_t").numpy().copy()
    vec[700] = vec[40]                                           # two identical group vectors (ids 41 and 701)
    vec_ids = np.asarray(ix["ids"], np.int32)
    rng = np.random.default_rng(11)
    S = oracle.ReferenceSession()
    S.load_pq(ix)
    S.load_vectors_table(vec, vec_ids)
    in_ids = rng.choice(np.arange(1, 3100), size=900, replace=True).astype(np.int32)
    groups = np.asarray([701, 41, 7, 2500, 1234], np.int32)
    g_ids, g_groups = S.grouping_pq(in_ids, groups)
    new_rows = (vec[rng.choice(len(vec), 200, replace=False)] +
                0.02 * rng.standard_normal((200, vec.shape[1])).astype(np.float32)).astype(np.float32)
    assign = oracle.reference_update_codebook_assignments(new_rows, ix["pq_codebook"])
    np.savez_compressed(
        os.path.join(HERE, "srf_golden_ext.npz"),
        d=ix["d"], m=ix["m"], K=ix["K"], N=ix["N"], ids=ix["ids"], pq_codebook=ix["pq_codebook"], pq_codes=ix["pq_codes"],
        vectors=vec, grouping_in_ids=in_ids, grouping_groups=groups, grouping_out_ids=g_ids, grouping_out_groups=g_groups,
        encode_rows=new_rows, encode_pq_codes=assign)
    print("wrote srf_golden_ext.npz")


if __name__ == "__main__":
    main()
    srf_golden_ext()
