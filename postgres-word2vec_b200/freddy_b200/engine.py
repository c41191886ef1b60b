"""Host-side handle over the C-ABI engine: one per process / GPU, index pinned once.

Mirrors what a Postgres backend would hold per session: the
coarse table, codebooks and code tables are uploaded once; searches then run on
the GPU only.
"""
import ctypes as C

import numpy as np

from . import _lib


class FreddyError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"freddy_b200 error {code}: {msg}")
        self.code = code
        self.msg = msg

    def __reduce__(self):                       # crosses multiprocessing pipes
        return (FreddyError, (self.code, self.msg))


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


class Engine:
    def __init__(self, device=0):
        self._lib = _lib.load()
        h = C.c_void_p()
        rc = self._lib.fb_create(int(device), C.byref(h))
        if rc != 0:
            raise FreddyError(rc, (self._lib.fb_last_error(None) or b"").decode())
        self._h = h
        self.device = device
        self.d = None
        self._cb_m = {}
        self._cb_d = {}

    def close(self):
        if getattr(self, "_h", None):
            if getattr(self, "_sidecar", None):          # the server thread uses the engine: stop it first
                self._lib.fb_sidecar_stop(self._sidecar, None)
                self._sidecar = None
            self._lib.fb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc):
        if rc != 0:
            raise FreddyError(rc, (self._lib.fb_last_error(self._h) or b"").decode())

    # ---- index upload -------------------------------------------------
    def load_coarse(self, coarse):
        coarse = _f32(coarse)
        self.d = coarse.shape[1]
        self._check(self._lib.fb_load_coarse(self._h, _ptr(coarse), coarse.shape[0], coarse.shape[1]))

    def load_codebook(self, kind, codebook):
        cb = _f32(codebook)
        m, K, sub = cb.shape
        if self.d is None:
            self.d = m * sub
        self._check(self._lib.fb_load_codebook(self._h, kind, _ptr(cb), m, K, sub))
        self._cb_m[kind] = m
        self._cb_d[kind] = m * sub          # every index kind has its own dimension (a session may pin several)

    def load_fine(self, ids, coarse_ids, codes):
        ids, coarse_ids = _i32(ids), _i32(coarse_ids)
        codes = np.ascontiguousarray(codes, dtype=np.int16)
        self._check(self._lib.fb_load_fine(self._h, _ptr(ids), _ptr(coarse_ids), _ptr(codes), codes.shape[0], codes.shape[1]))

    def load_pq(self, ids, codes):
        ids = _i32(ids)
        codes = np.ascontiguousarray(codes, dtype=np.int16)
        self._check(self._lib.fb_load_pq(self._h, _ptr(ids), _ptr(codes), codes.shape[0], codes.shape[1]))

    def load_ivfadc_index(self, index):
        """index: dict with coarse, residual_codebook, ids, coarse_ids, codes."""
        self.load_coarse(index["coarse"])
        self.load_codebook(_lib.FB_CB_RESIDUAL, index["residual_codebook"])
        self.load_fine(index["ids"], index["coarse_ids"], index["codes"])

    def load_pq_index(self, index):
        self.load_codebook(_lib.FB_CB_PQ, index["pq_codebook"])
        self.load_pq(index["ids"], index["pq_codes"])

    # ---- searches (host buffers) --------------------------------------
    def ivfadc_search(self, queries, k, w, out_ids=None, out_dists=None):
        q = _f32(queries).reshape(-1, self.d)
        nq = q.shape[0]
        ids = out_ids if out_ids is not None else np.empty((nq, k), np.int32)
        dists = out_dists if out_dists is not None else np.empty((nq, k), np.float32)
        self._check(self._lib.fb_ivfadc_search(self._h, _ptr(q), nq, k, w, _ptr(ids), _ptr(dists)))
        return ids, dists

    def ivfadc_search_ptr(self, q_ptr, nq, k, w, ids_ptr, dists_ptr):
        """host pointers given as integers (e.g. pinned torch tensors' data_ptr())"""
        self._check(self._lib.fb_ivfadc_search(self._h, C.c_void_p(q_ptr), nq, k, w, C.c_void_p(ids_ptr), C.c_void_p(dists_ptr)))

    def ivfadc_search_dev(self, d_queries_ptr, nq, k, w, d_ids_ptr, d_dists_ptr):
        """device pointers; enqueued on the engine stream, see synchronize()"""
        self._check(self._lib.fb_ivfadc_search_dev(self._h, C.c_void_p(d_queries_ptr), nq, k, w,
                                                   C.c_void_p(d_ids_ptr), C.c_void_p(d_dists_ptr)))

    def ivfadc_batch_search(self, query_ids, k):
        qi = _i32(query_ids)
        n = qi.shape[0]
        oq, ids, dists = np.empty(n, np.int32), np.empty((n, k), np.int32), np.empty((n, k), np.float32)
        nout = C.c_int(0)
        self._check(self._lib.fb_ivfadc_batch_search(self._h, _ptr(qi), n, k, _ptr(oq), _ptr(ids), _ptr(dists), C.byref(nout)))
        return oq[:nout.value], ids[:nout.value], dists[:nout.value]

    def pq_search(self, queries, k):
        q = _f32(queries).reshape(-1, self._cb_d[_lib.FB_CB_PQ])
        nq = q.shape[0]
        ids, dists = np.empty((nq, k), np.int32), np.empty((nq, k), np.float32)
        self._check(self._lib.fb_pq_search(self._h, _ptr(q), nq, k, _ptr(ids), _ptr(dists)))
        return ids, dists

    def pq_search_in_batch(self, queries, k, targets, use_target_lists=False):
        q = _f32(queries).reshape(-1, self._cb_d[_lib.FB_CB_PQ])
        nq = q.shape[0]
        t = _i32(targets)
        ids, dists = np.empty((nq, k), np.int32), np.empty((nq, k), np.float32)
        self._check(self._lib.fb_pq_search_in_batch(self._h, _ptr(q), nq, k, _ptr(t), t.shape[0],
                                                    1 if use_target_lists else 0, _ptr(ids), _ptr(dists)))
        return ids, dists

    def set_stream(self, cuda_stream):
        """cuda_stream: integer handle (e.g. torch.cuda.current_stream().cuda_stream) or 0/None"""
        self._check(self._lib.fb_set_stream(self._h, C.c_void_p(cuda_stream or 0)))

    # ---- kNN-join ----------------------------------------------------------
    def load_ivpq_index(self, ivpq):
        """ivpq: dict from index_build.make_ivpq_index"""
        self.load_codebook(_lib.FB_CB_IVPQ, ivpq["ivpq_codebook"])
        cm = _f32(ivpq["coarse_multi"])
        ids, cids = _i32(ivpq["ids"]), _i32(ivpq["ivpq_coarse_ids"])
        codes = np.ascontiguousarray(ivpq["ivpq_codes"], dtype=np.int16)
        st = _f32(ivpq["stats"]) if ivpq.get("stats") is not None else None     # None: create_statistics over all rows, on the device
        self.d = int(ivpq["d"])
        self._check(self._lib.fb_load_ivpq(self._h, _ptr(cm), int(ivpq["Kc"]), self.d, _ptr(ids), _ptr(cids),
                                           _ptr(codes), codes.shape[0], codes.shape[1], _ptr(st) if st is not None else None))
        self._ivpq_cells = int(ivpq["Kc"]) ** 2

    def ivpq_statistics(self, ids=None, install=False):
        """create_statistics (freddy--0.0.1.sql:150-171) over the pinned IVPQ table; ids = the ids of the user's column
        (repeats count), None = every row.  Returns float32[Kc*Kc + 1]."""
        out = np.empty(self._ivpq_cells + 1, np.float32)
        if ids is None:
            self._check(self._lib.fb_ivpq_statistics(self._h, None, 0, _ptr(out), 1 if install else 0))
        else:
            t = _i32(ids)
            self._check(self._lib.fb_ivpq_statistics(self._h, _ptr(t), t.shape[0], _ptr(out), 1 if install else 0))
        return out

    # ---- sidecar (include/freddy_sidecar.h) ------------------------------
    def sidecar_start(self, name, max_k=32, slots=256, max_batch=256, linger_us=0):
        """serve the single-query requests backends post into the shared-memory segment `name` ("/..."); the engine
        must not be used until sidecar_stop()"""
        h = C.c_void_p()
        self._check(self._lib.fb_sidecar_start(self._h, name.encode(), max_k, slots, max_batch, linger_us, C.byref(h)))
        self._sidecar = h

    def sidecar_stop(self):
        """-> dict(batches, queries, largest_batch)"""
        c = (C.c_int64 * 3)()
        rc = self._lib.fb_sidecar_stop(self._sidecar, c)
        self._sidecar = None
        if rc != 0:
            raise FreddyError(rc, "sidecar loop failed")
        return {"batches": int(c[0]), "queries": int(c[1]), "largest_batch": int(c[2])}

    def table_checksum(self, table):
        """(layout checksum, code checksum) of a pinned table: 0 fine, 1 pq, 2 ivpq"""
        out = np.zeros(2, np.uint64)
        self._check(self._lib.fb_table_checksum(self._h, table, _ptr(out)))
        return int(out[0]), int(out[1])

    def ivpq_search_in(self, queries, k, targets, alpha, pvf, method, use_target_lists, confidence,
                       double_threshold=10_000_000):
        q = _f32(queries).reshape(-1, self._cb_d[_lib.FB_CB_IVPQ])
        nq = q.shape[0]
        t = _i32(targets)
        ids, dists = np.empty((nq, k), np.int32), np.empty((nq, k), np.float32)
        self._check(self._lib.fb_ivpq_search_in(self._h, _ptr(q), nq, k, _ptr(t), t.shape[0], alpha, pvf, method,
                                                1 if use_target_lists else 0, confidence, double_threshold,
                                                _ptr(ids), _ptr(dists)))
        return ids, dists

    # ---- dense word-vector UDFs -----------------------------------------
    def load_vectors(self, ids, vectors):
        ids = _i32(ids)
        v = _f32(vectors)
        self.vec_d = v.shape[1]
        self._check(self._lib.fb_load_vectors(self._h, _ptr(ids), _ptr(v), v.shape[0], v.shape[1]))

    def cosine_similarity(self, a, b, variant=0):
        a, b = _f32(a), _f32(b)
        n, d = a.shape
        out = np.empty(n, np.float64)
        self._check(self._lib.fb_cosine_similarity(self._h, variant, _ptr(a), _ptr(b), n, d, _ptr(out)))
        return out

    def vec_op(self, op, a, b=None):
        a = _f32(a)
        b = _f32(b) if b is not None else None
        out = np.empty_like(a)
        self._check(self._lib.fb_vec_op(self._h, op, _ptr(a), _ptr(b), a.shape[0], a.shape[1], _ptr(out)))
        return out

    def analogy_3cosadd(self, ids_abc):
        t = _i32(ids_abc).reshape(-1, 3)
        out_ids, out_s = np.empty(len(t), np.int32), np.empty(len(t), np.float32)
        self._check(self._lib.fb_analogy_3cosadd(self._h, _ptr(t), len(t), _ptr(out_ids), _ptr(out_s)))
        return out_ids, out_s

    def analogy_scan(self, qvecs, exclude_ids=None):
        q = _f32(qvecs)
        ex = _i32(exclude_ids).reshape(-1, 3) if exclude_ids is not None else None
        out_ids, out_s = np.empty(len(q), np.int32), np.empty(len(q), np.float32)
        self._check(self._lib.fb_analogy_scan(self._h, _ptr(q), _ptr(ex), len(q), _ptr(out_ids), _ptr(out_s)))
        return out_ids, out_s

    def knn_exact(self, queries, k, targets=None):
        """k_nearest_neighbour(bytea, k) / knn_in_exact(bytea, k, int[]) for a batch of query vectors"""
        q = _f32(queries).reshape(-1, self.vec_d)
        t = _i32(targets) if targets is not None else None
        ids, sims = np.empty((len(q), k), np.int32), np.empty((len(q), k), np.float32)
        self._check(self._lib.fb_knn_exact(self._h, _ptr(q), len(q), k, _ptr(t), 0 if t is None else len(t),
                                           _ptr(ids), _ptr(sims)))
        return ids, sims

    def pq_search_pv(self, queries, k, pvf):
        """k_nearest_neighbour_pq_pv(bytea, k) for a batch: pq_search(v, pvf*k) re-ranked exactly"""
        q = _f32(queries).reshape(-1, self._cb_d[_lib.FB_CB_PQ])
        ids, sims = np.empty((len(q), k), np.int32), np.empty((len(q), k), np.float32)
        self._check(self._lib.fb_pq_search_pv(self._h, _ptr(q), len(q), k, pvf, _ptr(ids), _ptr(sims)))
        return ids, sims

    def ivfadc_search_pv(self, queries, k, pvf, w):
        """k_nearest_neighbour_ivfadc_pv(bytea, k) for a batch: ivfadc_search(v, pvf*k) re-ranked exactly"""
        q = _f32(queries).reshape(-1, self.d)
        ids, sims = np.empty((len(q), k), np.int32), np.empty((len(q), k), np.float32)
        self._check(self._lib.fb_ivfadc_search_pv(self._h, _ptr(q), len(q), k, pvf, w, _ptr(ids), _ptr(sims)))
        return ids, sims

    def encode_ivfadc(self, vectors):
        """(coarse_ids, codes) of new rows as insert_batch quantises them (coarse table + residual codebook loaded)"""
        v = _f32(vectors)
        n = v.shape[0]
        m = self._cb_m[_lib.FB_CB_RESIDUAL]
        cids, codes = np.empty(n, np.int32), np.empty((n, m), np.int16)
        self._check(self._lib.fb_encode_ivfadc(self._h, _ptr(v), n, _ptr(cids), _ptr(codes)))
        return cids, codes

    def encode_ivfadc_dev(self, d_vectors_ptr, n, d_coarse_ids_ptr, d_codes_ptr):
        """device pointers (e.g. torch tensors' data_ptr()): [n][d] fp32 in, [n] int32 and [n][m] int16 out"""
        self._check(self._lib.fb_encode_ivfadc_dev(self._h, C.c_void_p(d_vectors_ptr), n, C.c_void_p(d_coarse_ids_ptr), C.c_void_p(d_codes_ptr)))

    def encode_pq_dev(self, d_vectors_ptr, n, d_codes_ptr, kind=None):
        kind = _lib.FB_CB_PQ if kind is None else kind
        self._check(self._lib.fb_encode_pq_dev(self._h, kind, C.c_void_p(d_vectors_ptr), n, C.c_void_p(d_codes_ptr)))

    def encode_pq(self, vectors, kind=None):
        """codes of raw rows against the pq (default) / ivpq codebook"""
        kind = _lib.FB_CB_PQ if kind is None else kind
        v = _f32(vectors)
        n = v.shape[0]
        codes = np.empty((n, self._cb_m[kind]), np.int16)
        self._check(self._lib.fb_encode_pq(self._h, kind, _ptr(v), n, _ptr(codes)))
        return codes

    def append_fine(self, ids, coarse_ids, codes):
        """rows insert_batch adds to fine_quantization, appended to the pinned table on the device"""
        ids, cids = _i32(ids), _i32(coarse_ids)
        codes = np.ascontiguousarray(codes, dtype=np.int16).reshape(len(ids), -1)
        self._check(self._lib.fb_append_fine(self._h, _ptr(ids), _ptr(cids), _ptr(codes), len(ids)))

    def append_pq(self, ids, codes, kind=None, cells=None):
        kind = _lib.FB_CB_PQ if kind is None else kind
        ids = _i32(ids)
        cells = _i32(cells) if cells is not None else None
        codes = np.ascontiguousarray(codes, dtype=np.int16).reshape(len(ids), -1)
        self._check(self._lib.fb_append_pq(self._h, kind, _ptr(ids), _ptr(cells), _ptr(codes), len(ids)))

    def append_vectors(self, ids, vectors):
        ids = _i32(ids)
        v = _f32(vectors).reshape(len(ids), -1)
        self._check(self._lib.fb_append_vectors(self._h, _ptr(ids), _ptr(v), len(ids)))

    def grouping_pq(self, ids, group_ids):
        """grouping_pq(int[], int[]) -> (ids, group ids) of the selected pq rows in table order   freddy.c:1178-1401"""
        ids, gids = _i32(ids), _i32(group_ids)
        out_i, out_g = np.empty(len(ids), np.int32), np.empty(len(ids), np.int32)
        n = C.c_int(0)
        self._check(self._lib.fb_grouping_pq(self._h, _ptr(ids), len(ids), _ptr(gids), len(gids), _ptr(out_i), _ptr(out_g),
                                             C.byref(n)))
        return out_i[:n.value].copy(), out_g[:n.value].copy()

    def synchronize(self):
        self._check(self._lib.fb_synchronize(self._h))

    # ---- knobs --------------------------------------------------------
    def set_option(self, option, value):
        self._check(self._lib.fb_set_option(self._h, option, int(value)))

    def counters(self):
        c = _lib.Counters()
        self._check(self._lib.fb_get_counters(self._h, C.byref(c)))
        return {name: getattr(c, name) for name, _ in c._fields_}

    def reset_counters(self):
        self._check(self._lib.fb_reset_counters(self._h))


def round_through_text(dists):
    """snprintf('%f') -> float4in, as the SRFs return distances (freddy.c:401-408)"""
    lib = _lib.load()
    flat = np.asarray(dists, dtype=np.float32).ravel()
    return np.array([lib.fb_round_through_text(float(x)) for x in flat], np.float32).reshape(np.shape(dists))
