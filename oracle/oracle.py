"""ctypes access to the CPU oracle — TEST INFRASTRUCTURE ONLY.

  libfreddy_oracle.so      our plain-C restatement (oracle/freddy_oracle.c)
  _ref/libfreddy_ref.so    the reference's own index_utils.c / cosine_similarity.c,
                           compiled unmodified (oracle/Makefile `ref`)

Nothing under postgres-word2vec_b200/ imports this module; only tests/,
bench.py's CPU-baseline legs and __graft_entry__.smoke() do.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_SO = os.path.join(_HERE, "libfreddy_oracle.so")
REF_SO = os.path.join(_HERE, "_ref", "libfreddy_ref.so")
SHIM_SO = os.path.join(_HERE, "_ref", "libfreddy_shim_emul.so")


def build(quiet=True):
    subprocess.run(["make", "-C", _HERE, "all"], check=True,
                   stdout=subprocess.DEVNULL if quiet else None, stderr=subprocess.STDOUT if quiet else None)


class TopKEntry(C.Structure):
    _fields_ = [("id", C.c_int), ("distance", C.c_float)]


class FoIndex(C.Structure):
    _fields_ = [("d", C.c_int), ("m", C.c_int), ("K", C.c_int), ("C", C.c_int), ("N", C.c_int),
                ("coarse", C.c_void_p), ("codebook", C.c_void_p), ("ids", C.c_void_p),
                ("coarse_ids", C.c_void_p), ("codes", C.c_void_p),
                ("list_offsets", C.c_void_p), ("list_rows", C.c_void_p),
                ("list_codes", C.c_void_p), ("list_ids", C.c_void_p)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(ORACLE_SO):
            build()
        L = C.CDLL(ORACLE_SO)
        L.fo_square_distance.restype = C.c_float
        L.fo_square_distance.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.fo_update_topk.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int]
        L.fo_init_topk.argtypes = [C.c_void_p, C.c_int, C.c_float]
        L.fo_precomputed_distances.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.fo_pq_distance_int16.restype = C.c_float
        L.fo_pq_distance_int16.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
        L.fo_round_through_text.restype = C.c_float
        L.fo_round_through_text.argtypes = [C.c_float]
        L.fo_index_prepare.argtypes = [C.POINTER(FoIndex)]
        L.fo_index_release.argtypes = [C.POINTER(FoIndex)]
        L.fo_ivfadc_search.argtypes = [C.POINTER(FoIndex), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.fo_pq_search.argtypes = [C.POINTER(FoIndex), C.c_void_p, C.c_int, C.c_void_p]
        L.fo_pq_search_in.argtypes = [C.POINTER(FoIndex), C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p]
        L.fo_pq_search_in_batch.argtypes = [C.POINTER(FoIndex), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int,
                                            C.c_int, C.c_void_p]
        L.fo_ivfadc_search_many.argtypes = [C.POINTER(FoIndex), C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                            C.c_void_p, C.c_void_p]
        L.fo_cosine_similarity.restype = C.c_double
        L.fo_cosine_similarity.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.fo_cosine_similarity_norm.restype = C.c_double
        L.fo_cosine_similarity_norm.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.fo_cosine_similarity_bytea.restype = C.c_float
        L.fo_cosine_similarity_bytea.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        L.fo_vec_minus.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.fo_vec_plus.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.fo_vec_normalize.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        L.fo_grouping_pq.argtypes = [C.POINTER(FoIndex), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                     C.c_void_p, C.c_void_p]
        L.fo_encode.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.fo_analogy_3cosadd_many.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        _lib = L
    return _lib


class FoIvpqIndex(C.Structure):
    _fields_ = [("d", C.c_int), ("m", C.c_int), ("K", C.c_int), ("Kc", C.c_int), ("N", C.c_int),
                ("coarse_multi", C.c_void_p), ("codebook", C.c_void_p), ("ids", C.c_void_p), ("coarse_ids", C.c_void_p),
                ("codes", C.c_void_p), ("stats", C.c_void_p), ("Nv", C.c_int), ("vec_ids", C.c_void_p),
                ("vectors", C.c_void_p)]


class OracleIvpq:
    """oracle kNN-join (fo_ivpq_search_in) over an IVPQ index + the normalized vector table"""

    def __init__(self, ivpq, vectors, vec_ids):
        self.keep = [np.ascontiguousarray(ivpq["coarse_multi"], np.float32), np.ascontiguousarray(ivpq["ivpq_codebook"], np.float32),
                     np.ascontiguousarray(ivpq["ids"], np.int32), np.ascontiguousarray(ivpq["ivpq_coarse_ids"], np.int32),
                     np.ascontiguousarray(ivpq["ivpq_codes"], np.int16), np.ascontiguousarray(ivpq["stats"], np.float32),
                     np.ascontiguousarray(vec_ids, np.int32), np.ascontiguousarray(vectors, np.float32)]
        ix = FoIvpqIndex()
        ix.d, ix.m, ix.K, ix.Kc, ix.N = int(ivpq["d"]), int(ivpq["m"]), int(ivpq["K"]), int(ivpq["Kc"]), int(ivpq["N"])
        (ix.coarse_multi, ix.codebook, ix.ids, ix.coarse_ids, ix.codes, ix.stats, ix.vec_ids, ix.vectors) = [_p(a) for a in self.keep]
        ix.Nv = len(self.keep[6])
        self.ix = ix
        L = lib()
        L.fo_ivpq_search_in.argtypes = [C.POINTER(FoIvpqIndex), C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_void_p]

    def search_in(self, queries, k, targets, alpha, pvf, method, use_targetlist, confidence, dbl_threshold=10_000_000):
        q = np.ascontiguousarray(queries, np.float32).reshape(-1, self.ix.d)
        t = np.ascontiguousarray(targets, np.int32)
        nq = len(q)
        tk = (TopKEntry * (nq * k))()
        st = (C.c_int64 * 2)()
        rc = lib().fo_ivpq_search_in(C.byref(self.ix), _p(q), nq, k, _p(t), len(t), alpha, pvf, method,
                                     1 if use_targetlist else 0, confidence, dbl_threshold, tk, st)
        a = np.frombuffer(tk, dtype=[("id", np.int32), ("distance", np.float32)]).reshape(nq, k)
        return a["id"].copy(), a["distance"].copy(), rc, (st[0], st[1])


def analogy_3cosadd(vectors, rows_abc, threads=1):
    """oracle: winning table row and score per (a, b, c) row triple"""
    v = np.ascontiguousarray(vectors, np.float32)
    t = np.ascontiguousarray(rows_abc, np.int32).reshape(-1, 3)
    rows, scores = np.empty(len(t), np.int32), np.empty(len(t), np.float32)
    lib().fo_analogy_3cosadd_many(_p(v), v.shape[0], v.shape[1], _p(t), len(t), threads, _p(rows), _p(scores))
    return rows, scores


def ref_lib():
    """The reference's own compiled kernels, or None when not built/shipped."""
    if not os.path.exists(REF_SO):
        return None
    R = C.CDLL(REF_SO)
    R.squareDistance.restype = C.c_float
    R.squareDistance.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    R.updateTopK.argtypes = [C.c_void_p, C.c_float, C.c_int, C.c_int, C.c_int]
    R.computePQDistanceInt16.restype = C.c_float
    R.computePQDistanceInt16.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int]
    R.getPrecomputedDistances.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    R.cosine_similarity_simple.restype = C.c_double
    R.cosine_similarity_simple.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    R.cosine_similarity_simple_norm.restype = C.c_double
    R.cosine_similarity_simple_norm.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    R.ref_cosine_similarity_bytea.restype = C.c_float
    R.ref_cosine_similarity_bytea.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
    R.ref_vec_minus_bytea.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    R.ref_vec_plus_bytea.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    R.ref_vec_normalize_bytea.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    return R


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class ReferenceSession:
    """The reference extension's OWN SRFs (freddy.c, ivpq_search_in.c, core_functions.c
    compiled unmodified into oracle/_ref/libfreddy_ref.so) running on in-memory copies of
    the index tables through the SPI/fmgr emulator oracle/pg_emul.c.  Plays one Postgres
    backend with the tables `init()` names (freddy--0.0.1.sql:5-19)."""

    T_COARSE, T_CODEBOOK, T_FINE, T_PQ, T_VECS, T_STATS = 1, 2, 3, 4, 5, 6

    def __init__(self, lib_path=None):
        """lib_path=None: the reference's own SRFs.  lib_path=SHIM_SO: OUR Postgres-side shim
        (postgres-word2vec_b200/shim/freddy_shim.c -> libfreddy_b200.so) behind the same emulated
        fmgr/SPI boundary."""
        path = lib_path or REF_SO
        if not os.path.exists(path):
            raise RuntimeError(f"{path} not built (reference sources absent)")
        R = C.CDLL(path)
        if lib_path is not None:
            R.freddy_shim_reset()
        R.ref_register_table.argtypes = [C.c_char_p, C.c_int, C.c_int64, C.c_void_p, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_int, C.c_void_p]
        R.ref_set_config.argtypes = [C.c_char_p, C.c_char_p]
        R.ref_last_error.restype = C.c_char_p
        R.ref_ivfadc_search.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        R.ref_pq_search.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        R.ref_pq_search_in.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        R.ref_pq_search_in_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                             C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        R.ref_cosine_similarity_bytea.restype = C.c_float
        R.ref_cosine_similarity_bytea.argtypes = [C.c_void_p, C.c_void_p, C.c_int]
        for f in (R.ref_vec_minus_bytea, R.ref_vec_plus_bytea):
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        R.ref_vec_normalize_bytea.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        self.R = R
        self.keep = []
        R.ref_reset()
        self.d = None

    def _err(self, what, n):
        raise RuntimeError(f"{what} failed ({n}): {self.R.ref_last_error().decode()}")

    def set_config(self, fn, value):
        self.R.ref_set_config(fn.encode(), str(value).encode())

    def _table(self, name, kind, n, ids=None, a=None, b=None, vec=None, freq=None):
        arrs = [None if x is None else np.ascontiguousarray(x) for x in (ids, a, b, vec, freq)]
        self.keep.append(arrs)
        ids, a, b, vec, freq = arrs
        vb = 0 if vec is None else vec.dtype.itemsize * int(np.prod(vec.shape[1:]))
        pp = lambda x: None if x is None else _p(x)
        rc = self.R.ref_register_table(name.encode(), kind, n, pp(ids), pp(a), pp(b), pp(vec), vb, pp(freq))
        assert rc == 0

    def load_ivfadc(self, index, w):
        ix = index
        self.d = int(ix["d"])
        C_, m, K = int(ix["C"]), int(ix["m"]), int(ix["K"])
        self._table("coarse_quantization", self.T_COARSE, C_, ids=np.arange(C_, dtype=np.int32),
                    vec=np.asarray(ix["coarse"], np.float32))
        pos = np.repeat(np.arange(m, dtype=np.int32), K)
        code = np.tile(np.arange(K, dtype=np.int32), m)
        self._table("residual_codebook", self.T_CODEBOOK, m * K, ids=np.arange(m * K, dtype=np.int32), a=pos, b=code,
                    vec=np.asarray(ix["residual_codebook"], np.float32).reshape(m * K, -1))
        self._table("fine_quantization", self.T_FINE, int(ix["N"]), ids=np.asarray(ix["ids"], np.int32),
                    a=np.asarray(ix["coarse_ids"], np.int32), vec=np.asarray(ix["codes"], np.int16))
        self.set_config("get_vecs_name()", "google_vecs_norm")
        self.set_config("get_vecs_name_residual_codebook()", "residual_codebook")
        self.set_config("get_vecs_name_residual_quantization()", "fine_quantization")
        self.set_config("get_vecs_name_coarse_quantization()", "coarse_quantization")
        self.set_config("get_w()", w)

    def load_pq(self, index):
        ix = index
        self.d = int(ix["d"])
        m, K = int(ix["m"]), int(ix["K"])
        pos = np.repeat(np.arange(m, dtype=np.int32), K)
        code = np.tile(np.arange(K, dtype=np.int32), m)
        self._table("pq_codebook", self.T_CODEBOOK, m * K, ids=np.arange(m * K, dtype=np.int32), a=pos, b=code,
                    vec=np.asarray(ix["pq_codebook"], np.float32).reshape(m * K, -1))
        self._table("pq_quantization", self.T_PQ, int(ix["N"]), ids=np.asarray(ix["ids"], np.int32),
                    vec=np.asarray(ix["pq_codes"], np.int16))
        self.set_config("get_vecs_name()", "google_vecs_norm")
        self.set_config("get_vecs_name_codebook()", "pq_codebook")
        self.set_config("get_vecs_name_pq_quantization()", "pq_quantization")

    def load_ivpq(self, ivpq, vectors, vec_ids):
        """ivpq: dict from freddy_b200.index_build.make_ivpq_index; vectors: the normalized table"""
        self.d = int(ivpq["d"])
        m, K, Kc, N = int(ivpq["m"]), int(ivpq["K"]), int(ivpq["Kc"]), int(ivpq["N"])
        self._table("codebook_ivpq", self.T_CODEBOOK, m * K, ids=np.arange(m * K, dtype=np.int32),
                    a=np.repeat(np.arange(m, dtype=np.int32), K), b=np.tile(np.arange(K, dtype=np.int32), m),
                    vec=np.asarray(ivpq["ivpq_codebook"], np.float32).reshape(m * K, -1))
        self._table("coarse_quantization_ivpq", self.T_CODEBOOK, 2 * Kc, ids=np.arange(2 * Kc, dtype=np.int32),
                    a=np.repeat(np.arange(2, dtype=np.int32), Kc), b=np.tile(np.arange(Kc, dtype=np.int32), 2),
                    vec=np.asarray(ivpq["coarse_multi"], np.float32).reshape(2 * Kc, -1))
        self._table("fine_quantization_ivpq", self.T_FINE, N, ids=np.asarray(ivpq["ids"], np.int32),
                    a=np.asarray(ivpq["ivpq_coarse_ids"], np.int32), vec=np.asarray(ivpq["ivpq_codes"], np.int16))
        self._table("google_vecs_norm", self.T_VECS, len(vec_ids), ids=np.asarray(vec_ids, np.int32),
                    vec=np.asarray(vectors, np.float32))
        st = np.asarray(ivpq["stats"], np.float32)
        self._table("stat_table", self.T_STATS, len(st), ids=np.arange(len(st), dtype=np.int32), freq=st)
        self.set_config("get_vecs_name()", "google_vecs_norm")
        self.set_config("get_vecs_name_ivpq_codebook()", "codebook_ivpq")
        self.set_config("get_vecs_name_ivpq_quantization()", "fine_quantization_ivpq")
        self.set_config("get_vecs_name_coarse_quantization_multi()", "coarse_quantization_ivpq")
        self.set_config("get_statistics_table()", "stat_table")
        self.R.ref_ivpq_search_in.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int,
                                              C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int,
                                              C.c_void_p, C.c_void_p, C.c_void_p]

    def load_vectors_table(self, vectors, vec_ids, name="google_vecs_norm"):
        self._table(name, self.T_VECS, len(vec_ids), ids=np.asarray(vec_ids, np.int32), vec=np.asarray(vectors, np.float32))
        self.set_config("get_vecs_name()", name)

    def ivfadc_batch_search(self, query_ids, k):
        qi = np.ascontiguousarray(query_ids, np.int32)
        n = len(qi)
        self.R.ref_ivfadc_batch_search.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_int), C.c_void_p,
                                                   C.c_void_p, C.c_void_p]
        qo, ids, raw = np.empty(n * k, np.int32), np.empty(n * k, np.int32), np.empty(n * k, np.float32)
        nq = C.c_int(0)
        r = self.R.ref_ivfadc_batch_search(_p(qi), n, k, n, C.byref(nq), _p(qo), _p(ids), _p(raw))
        if r < 0:
            self._err("ivfadc_batch_search", r)
        m = nq.value
        return qo[:m * k].reshape(m, k)[:, 0].copy(), ids[:m * k].reshape(m, k), raw[:m * k].reshape(m, k)

    def ivpq_search_in(self, queries, query_ids, k, targets, alpha, pvf, method, use_targetlist, confidence, dbl_threshold):
        q = np.ascontiguousarray(queries, np.float32).reshape(-1, self.d)
        nq = len(q)
        qi = np.ascontiguousarray(query_ids, np.int32)
        t = np.ascontiguousarray(targets, np.int32)
        qo, ids, raw = np.empty(nq * k, np.int32), np.empty(nq * k, np.int32), np.empty(nq * k, np.float32)
        n = self.R.ref_ivpq_search_in(_p(q), nq, self.d, _p(qi), k, _p(t), len(t), alpha, pvf, method,
                                      1 if use_targetlist else 0, confidence, dbl_threshold, _p(qo), _p(ids), _p(raw))
        if n != nq * k:
            self._err("ivpq_search_in", n)
        return qo.reshape(nq, k), ids.reshape(nq, k), raw.reshape(nq, k)

    def grouping_pq(self, ids, group_ids):
        ii, gi = np.ascontiguousarray(ids, np.int32), np.ascontiguousarray(group_ids, np.int32)
        oi, og = np.empty(max(1, len(ii)), np.int32), np.empty(max(1, len(ii)), np.int32)
        self.R.ref_grouping_pq.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        n = self.R.ref_grouping_pq(_p(ii), len(ii), _p(gi), len(gi), _p(oi), _p(og))
        if n < 0:
            self._err("grouping_pq", n)
        return oi[:n].copy(), og[:n].copy()

    # ---- insert_batch (both libraries) and the SRFs only the shim has -----------------------------------
    def load_insert_tables(self, index, ivpq, vectors, vec_ids, w=3):
        """everything insert_batch touches: pq / residual / ivpq codebooks and row tables, coarse tables, both
        word-vector tables (original = normalized here)"""
        self.load_ivfadc(index, w)
        self.load_pq(index)
        self.load_ivpq(ivpq, vectors, vec_ids)
        self._table("google_vecs", self.T_VECS, len(vec_ids), ids=np.asarray(vec_ids, np.int32), vec=np.asarray(vectors, np.float32))
        self.set_config("get_vecs_name_original()", "google_vecs")
        self.set_config("get_vecs_name()", "google_vecs_norm")

    def insert_batch(self, terms, tokens, norm_vectors, raw_vectors):
        """terms: what the caller passes; tokens / vectors: the rows the SQL-side tokenisation returns for the new
        terms (the emulator answers the tokenize() statement with them).  Returns the DML statements issued."""
        nv = np.ascontiguousarray(norm_vectors, np.float32)
        rv = np.ascontiguousarray(raw_vectors, np.float32)
        toks = (C.c_char_p * len(tokens))(*[t.encode() for t in tokens])
        self.R.ref_set_tokenization.argtypes = [C.c_int, C.POINTER(C.c_char_p), C.c_void_p, C.c_void_p, C.c_int]
        self.R.ref_set_tokenization(len(tokens), toks, _p(nv), _p(rv), nv.shape[1])
        self.R.ref_statement_log.restype = C.c_char_p
        self.R.ref_clear_statement_log()
        tt = (C.c_char_p * len(terms))(*[t.encode() for t in terms])
        self.R.ref_insert_batch.argtypes = [C.c_int, C.POINTER(C.c_char_p)]
        r = self.R.ref_insert_batch(len(terms), tt)
        if r != 0:
            self._err("insert_batch", r)
        return self.R.ref_statement_log().decode().splitlines()

    def knn_exact_search(self, query, k, targets=None):
        q = np.ascontiguousarray(query, np.float32).ravel()
        t = None if targets is None else np.ascontiguousarray(targets, np.int32)
        ids, sims = np.full(k, -1, np.int32), np.zeros(k, np.float32)
        self.R.ref_knn_exact_search.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        n = self.R.ref_knn_exact_search(_p(q), len(q), k, None if t is None else _p(t), 0 if t is None else len(t), _p(ids), _p(sims))
        if n < 0:
            self._err("knn_exact_search", n)
        return ids[:n], sims[:n]

    def ivfadc_search_pv(self, query, k):
        q = np.ascontiguousarray(query, np.float32).ravel()
        ids, sims = np.full(k, -1, np.int32), np.zeros(k, np.float32)
        self.R.ref_ivfadc_search_pv.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        n = self.R.ref_ivfadc_search_pv(_p(q), len(q), k, _p(ids), _p(sims))
        if n < 0:
            self._err("ivfadc_search_pv", n)
        return ids[:n], sims[:n]

    def pq_search_pv(self, query, k):
        q = np.ascontiguousarray(query, np.float32).ravel()
        ids, sims = np.full(k, -1, np.int32), np.zeros(k, np.float32)
        self.R.ref_pq_search_pv.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        n = self.R.ref_pq_search_pv(_p(q), len(q), k, _p(ids), _p(sims))
        if n < 0:
            self._err("pq_search_pv", n)
        return ids[:n], sims[:n]

    def analogy_3cosadd_batch(self, ids_abc):
        t = np.ascontiguousarray(ids_abc, np.int32).reshape(-1, 3)
        ids, sc = np.empty(len(t), np.int32), np.empty(len(t), np.float32)
        self.R.ref_analogy_3cosadd_batch.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        n = self.R.ref_analogy_3cosadd_batch(_p(t), len(t), _p(ids), _p(sc))
        if n != len(t):
            self._err("analogy_3cosadd_batch", n)
        return ids, sc

    def cosine_similarity_batch(self, a, b, variant):
        a, b = np.ascontiguousarray(a, np.float32), np.ascontiguousarray(b, np.float32)
        out = np.empty(len(a), np.float64)
        self.R.ref_cosine_similarity_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]
        n = self.R.ref_cosine_similarity_batch(_p(a), _p(b), len(a), a.shape[1], variant, _p(out))
        if n != len(a):
            self._err("cosine_similarity_batch", n)
        return out

    def repin(self):
        return self.R.ref_freddy_repin()

    def sidecar_serve(self, d, max_k, seconds):
        """shim only, blocking: freddy_sidecar_serve(dims, max_k, seconds) -> queries answered"""
        return self.R.ref_freddy_sidecar_serve(int(d), int(max_k), int(seconds))

    def sidecar_stop(self):
        return self.R.ref_freddy_sidecar_stop()

    def ivfadc_search(self, queries, k):
        q = np.ascontiguousarray(queries, np.float32).reshape(-1, self.d)
        ids = np.empty((len(q), k), np.int32)
        raw = np.empty((len(q), k), np.float32)
        txt = np.empty((len(q), k), np.float32)
        for i in range(len(q)):
            n = self.R.ref_ivfadc_search(_p(q[i]), self.d, k, _p(ids[i]), _p(raw[i]), _p(txt[i]))
            if n != k:
                self._err("ivfadc_search", n)
        return ids, raw, txt

    def pq_search(self, queries, k):
        q = np.ascontiguousarray(queries, np.float32).reshape(-1, self.d)
        ids, raw = np.empty((len(q), k), np.int32), np.empty((len(q), k), np.float32)
        for i in range(len(q)):
            n = self.R.ref_pq_search(_p(q[i]), self.d, k, _p(ids[i]), _p(raw[i]))
            if n != k:
                self._err("pq_search", n)
        return ids, raw

    def pq_search_in(self, query, k, targets):
        q = np.ascontiguousarray(query, np.float32).reshape(self.d)
        t = np.ascontiguousarray(targets, np.int32)
        ids, raw = np.empty(k, np.int32), np.empty(k, np.float32)
        n = self.R.ref_pq_search_in(_p(q), self.d, k, _p(t), len(t), _p(ids), _p(raw))
        if n != k:
            self._err("pq_search_in", n)
        return ids, raw

    def pq_search_in_batch(self, queries, query_ids, k, targets, use_targetlist):
        q = np.ascontiguousarray(queries, np.float32).reshape(-1, self.d)
        nq = len(q)
        qi = np.ascontiguousarray(query_ids, np.int32)
        t = np.ascontiguousarray(targets, np.int32)
        qo, ids, raw = np.empty(nq * k, np.int32), np.empty(nq * k, np.int32), np.empty(nq * k, np.float32)
        n = self.R.ref_pq_search_in_batch(_p(q), nq, self.d, _p(qi), k, _p(t), len(t), 1 if use_targetlist else 0,
                                          _p(qo), _p(ids), _p(raw))
        if n != nq * k:
            self._err("pq_search_in_batch", n)
        return qo.reshape(nq, k), ids.reshape(nq, k), raw.reshape(nq, k)


class OracleIndex:
    """In-memory image of the index tables for the oracle (keeps arrays alive)."""

    def __init__(self, index, flat_pq=False):
        L = lib()
        self.k = None
        self.arr = {}
        ix = FoIndex()
        ix.d, ix.m, ix.K, ix.N = int(index["d"]), int(index["m"]), int(index["K"]), int(index["N"])
        if flat_pq:
            ix.C = 0
            self.arr["cb"] = np.ascontiguousarray(index["pq_codebook"], np.float32)
            self.arr["codes"] = np.ascontiguousarray(index["pq_codes"], np.int16)
            ix.coarse, ix.coarse_ids = None, None
        else:
            ix.C = int(index["C"])
            self.arr["coarse"] = np.ascontiguousarray(index["coarse"], np.float32)
            self.arr["cb"] = np.ascontiguousarray(index["residual_codebook"], np.float32)
            self.arr["codes"] = np.ascontiguousarray(index["codes"], np.int16)
            self.arr["coarse_ids"] = np.ascontiguousarray(index["coarse_ids"], np.int32)
            ix.coarse = _p(self.arr["coarse"])
            ix.coarse_ids = _p(self.arr["coarse_ids"])
        self.arr["ids"] = np.ascontiguousarray(index["ids"], np.int32)
        ix.codebook, ix.ids, ix.codes = _p(self.arr["cb"]), _p(self.arr["ids"]), _p(self.arr["codes"])
        self.ix = ix
        rc = L.fo_index_prepare(C.byref(ix))
        if rc:
            raise RuntimeError(f"fo_index_prepare failed: {rc}")

    def __del__(self):
        try:
            lib().fo_index_release(C.byref(self.ix))
        except Exception:
            pass

    @staticmethod
    def _unpack(tk, nq, k):
        a = np.frombuffer(tk, dtype=[("id", np.int32), ("distance", np.float32)]).reshape(nq, k)
        return a["id"].copy(), a["distance"].copy()

    def ivfadc_search(self, queries, k, w, threads=1):
        q = np.ascontiguousarray(queries, np.float32).reshape(-1, self.ix.d)
        nq = q.shape[0]
        tk = (TopKEntry * (nq * k))()
        rows = C.c_int64(0)
        rc = lib().fo_ivfadc_search_many(C.byref(self.ix), _p(q), nq, k, w, threads, tk, C.byref(rows))
        ids, d = self._unpack(tk, nq, k)
        return ids, d, rc, rows.value

    def pq_search(self, queries, k):
        q = np.ascontiguousarray(queries, np.float32).reshape(-1, self.ix.d)
        nq = q.shape[0]
        ids, ds = np.empty((nq, k), np.int32), np.empty((nq, k), np.float32)
        for i in range(nq):
            tk = (TopKEntry * k)()
            lib().fo_pq_search(C.byref(self.ix), _p(q[i]), k, tk)
            ids[i], ds[i] = self._unpack(tk, 1, k)
        return ids, ds

    def grouping_pq(self, vectors, vec_ids, ids, group_ids):
        """oracle: (ids, group ids, rc) of grouping_pq (freddy.c:1178-1401); flat PQ index"""
        v = np.ascontiguousarray(vectors, np.float32)
        vi = np.ascontiguousarray(vec_ids, np.int32)
        ii, gi = np.ascontiguousarray(ids, np.int32), np.ascontiguousarray(group_ids, np.int32)
        oi, og = np.empty(max(1, len(ii)), np.int32), np.empty(max(1, len(ii)), np.int32)
        n = lib().fo_grouping_pq(C.byref(self.ix), _p(v), _p(vi), len(vi), _p(ii), len(ii), _p(gi), len(gi), _p(oi), _p(og))
        return (oi[:max(n, 0)].copy(), og[:max(n, 0)].copy(), n)

    def pq_search_in_batch(self, queries, k, targets, use_target_lists=False):
        q = np.ascontiguousarray(queries, np.float32).reshape(-1, self.ix.d)
        nq = q.shape[0]
        t = np.ascontiguousarray(targets, np.int32)
        tk = (TopKEntry * (nq * k))()
        lib().fo_pq_search_in_batch(C.byref(self.ix), _p(q), nq, k, _p(t), t.shape[0], 1 if use_target_lists else 0, tk)
        return self._unpack(tk, nq, k)


# ---- exact cosine k-NN and post-verification (SQL-level functions) ----------------
def cosine_similarity_bytea_many(q, vectors):
    """core_functions.c:67-81 for one query against many rows: `scalar += v1[i] * v2[i]` in float4, product and
    sum rounded separately, left to right (vectorised over rows, sequential over dimensions)"""
    v = np.ascontiguousarray(vectors, np.float32)
    q = np.ascontiguousarray(q, np.float32)
    acc = np.zeros(v.shape[0], np.float32)
    for i in range(v.shape[1]):
        acc = acc + q[i] * v[:, i]          # float32 * float32 -> float32, + float32 -> float32
    return acc


def _order_desc(sims, rows, k):
    """ORDER BY similarity DESC FETCH FIRST k; equal similarities by table row (the engine's stated rule)"""
    order = np.lexsort((rows, -sims.astype(np.float64)))
    return order[:k]


def knn_exact(vectors, vec_ids, queries, k, targets=None):
    """k_nearest_neighbour(bytea, k) / knn_in_exact(bytea, k, int[])   freddy--0.0.1.sql:426-439, :1026-1038"""
    v = np.ascontiguousarray(vectors, np.float32)
    ids = np.asarray(vec_ids, np.int32)
    rows = np.arange(len(v))
    if targets is not None:
        rows = np.nonzero(np.isin(ids, np.asarray(targets, np.int32)))[0]       # WHERE id = ANY(...): table order, once
    out_ids = np.full((len(queries), k), -1, np.int32)
    out_s = np.zeros((len(queries), k), np.float32)
    for qi, q in enumerate(np.ascontiguousarray(queries, np.float32)):
        s = cosine_similarity_bytea_many(q, v[rows])
        sel = _order_desc(s, rows, k)
        out_ids[qi, :len(sel)] = ids[rows[sel]]
        out_s[qi, :len(sel)] = s[sel]
    return out_ids, out_s


def pq_search_pv(oracle_index, vectors, vec_ids, queries, k, pvf):
    """k_nearest_neighbour_pq_pv(bytea, k)   freddy--0.0.1.sql:624-662: candidates = pq_search(v, pvf*k) (flat PQ oracle
    index), INNER JOIN vectors ON idx = id, ORDER BY cosine_similarity_bytea DESC FETCH FIRST k.  (The reference's SQL
    names the candidate's word where its ivfadc twin, :574-591, names the vector; the twin's form is restated.)"""
    cand, _ = oracle_index.pq_search(queries, k * pvf)
    return _rerank(vectors, vec_ids, queries, cand, k)


def ivfadc_search_pv(oracle_index, vectors, vec_ids, queries, k, pvf, w, threads=4):
    """k_nearest_neighbour_ivfadc_pv(bytea, k)   freddy--0.0.1.sql:574-591: candidates = ivfadc_search(v, pvf*k),
    INNER JOIN vectors ON idx = id, ORDER BY cosine_similarity_bytea DESC FETCH FIRST k"""
    cand, _, rc, _ = oracle_index.ivfadc_search(queries, k * pvf, w, threads=threads)
    assert rc == 0
    return _rerank(vectors, vec_ids, queries, cand, k)


def _rerank(vectors, vec_ids, queries, cand, k):
    v = np.ascontiguousarray(vectors, np.float32)
    ids = np.asarray(vec_ids, np.int32)
    row_of = {int(i): r for r, i in enumerate(ids)}
    out_ids = np.full((len(queries), k), -1, np.int32)
    out_s = np.zeros((len(queries), k), np.float32)
    for qi, q in enumerate(np.ascontiguousarray(queries, np.float32)):
        rows = np.array([row_of[int(c)] for c in cand[qi] if int(c) in row_of], np.int64)
        if len(rows) == 0:
            continue
        s = cosine_similarity_bytea_many(q, v[rows])
        sel = _order_desc(s, rows, k)
        out_ids[qi, :len(sel)] = ids[rows[sel]]
        out_s[qi, :len(sel)] = s[sel]
    return out_ids, out_s


# ---- quantisation of new rows (insert_batch) -----------------------------------------------------------
def encode(vectors, codebook, coarse=None):
    """oracle: (coarse_ids or None, codes[n][m] int16, rc) as insert_batch assigns them (freddy.c:1567-1582,
    index_utils.c:923-939)"""
    v = np.ascontiguousarray(vectors, np.float32)
    cb = np.ascontiguousarray(codebook, np.float32)            # [m][K][sub]
    m, K, _ = cb.shape
    n, d = v.shape
    codes = np.empty((n, m), np.int16)
    if coarse is not None:
        cq = np.ascontiguousarray(coarse, np.float32)
        cids = np.empty(n, np.int32)
        rc = lib().fo_encode(_p(v), n, d, _p(cq), cq.shape[0], _p(cb), m, K, _p(cids), _p(codes))
        return cids, codes, rc
    rc = lib().fo_encode(_p(v), n, d, None, 0, _p(cb), m, K, None, _p(codes))
    return None, codes, rc


def reference_update_codebook_assignments(vectors, codebook):
    """nearestCentroids of the reference's own updateCodebook (index_utils.c:908-957, compiled from
    /root/reference into oracle/_ref) for raw vectors against a [m][K][sub] codebook"""
    R = ref_lib()
    v = np.ascontiguousarray(vectors, np.float32)
    cb = np.ascontiguousarray(codebook, np.float32).copy()     # updateCodebook drifts the codebook in place
    m, K, sub = cb.shape
    n = v.shape[0]

    class Entry(C.Structure):
        _fields_ = [("pos", C.c_int), ("code", C.c_int), ("vector", C.POINTER(C.c_float)), ("count", C.c_int)]

    entries = (Entry * (m * K))()
    flat = cb.reshape(m * K, sub)
    for j in range(m * K):
        entries[j].pos, entries[j].code, entries[j].count = j // K, j % K, 1
        entries[j].vector = flat[j].ctypes.data_as(C.POINTER(C.c_float))
    rows = (C.POINTER(C.c_float) * n)(*[v[i].ctypes.data_as(C.POINTER(C.c_float)) for i in range(n)])
    nearest = (C.POINTER(C.c_int) * n)()
    incs = (C.c_int * (m * K))()
    R.updateCodebook.restype = None
    R.updateCodebook.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
    R.updateCodebook(rows, n, sub, entries, m, K, nearest, incs)
    return np.array([[nearest[i][p] for p in range(m)] for i in range(n)], np.int16)


def create_statistics(table_ids, table_cells, listed_ids, n_cells):
    """create_statistics (freddy--0.0.1.sql:150-171) restated: the statistics table of the kNN-join.
    total = count(*) of `table JOIN vecs ON column = word` (:164) — a listed id counts once per occurrence and per
    index row carrying it; coarse_freq(c) = (count of those rows with coarse_id = c)::float / total stored as float4
    (:162, :166: float8 division, then the float4 column); the extra row n_cells holds the total (:168).
    listed_ids None = every row of the index once."""
    table_ids = np.asarray(table_ids, np.int64)
    table_cells = np.asarray(table_cells, np.int64)
    if listed_ids is None:
        counts = np.bincount(table_cells, minlength=n_cells).astype(np.float64)
    else:
        vals, mult = np.unique(np.asarray(listed_ids, np.int64), return_counts=True)
        pos = np.searchsorted(vals, table_ids)
        pos[pos >= len(vals)] = 0
        weight = np.where(vals[pos] == table_ids, mult[pos], 0) if len(vals) else np.zeros(len(table_ids), np.int64)
        counts = np.bincount(table_cells, weights=weight.astype(np.float64), minlength=n_cells)
    total = counts.sum()
    return np.concatenate([(counts / total).astype(np.float32), np.asarray([total], np.float32)])
