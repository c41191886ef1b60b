"""Host-side mirror of the reference's search SRFs (same names, argument meaning,
row shape and error behaviour), over the C-ABI engine.

In the reference these are fmgr-V1 set-returning functions in the `freddy`
shared library (freddy--0.0.1.sql:370-392); vectors travel as `bytea` = raw
little-endian float4[d] (index_utils.c:1098-1106) and id sets as int4[].  A
`Session` plays the role of one Postgres backend: tables are pinned once, the
config-as-SQL-functions `set_w()/get_w()` (freddy--0.0.1.sql:28-33, default 3 at
:189) live on it, and every SRF returns its rows the way the C code emits them:
`k` rows per query, ascending distance, padded with id = -1, distances rounded
through "%f" text (freddy.c:401-408).
"""
import numpy as np

from . import _lib
from .engine import Engine, round_through_text


def vec_to_bytea(vec):
    """float4[] -> bytea payload (freddy.c:1758-1783 vec_to_bytea)"""
    return np.ascontiguousarray(vec, dtype="<f4").tobytes()


def bytea_to_vec(b):
    return np.frombuffer(b, dtype="<f4")


class Session:
    def __init__(self, device=0, engine=None):
        self.engine = engine if engine is not None else Engine(device)
        self._w = 3            # SELECT set_w(3)   freddy--0.0.1.sql:189
        self._pvf = 20         # SELECT set_pvf(20) freddy--0.0.1.sql:188

    # ---- config-as-functions -------------------------------------------
    def set_w(self, w):
        self._w = int(w)

    def get_w(self):
        return self._w

    def set_pvf(self, f):
        self._pvf = int(f)

    def get_pvf(self):
        return self._pvf

    # ---- index tables ("pin once per session") ---------------------------
    def load_ivfadc(self, coarse, residual_codebook, ids, coarse_ids, codes):
        self.engine.load_coarse(coarse)
        self.engine.load_codebook(_lib.FB_CB_RESIDUAL, residual_codebook)
        self.engine.load_fine(ids, coarse_ids, codes)

    def load_pq(self, pq_codebook, ids, codes):
        self.engine.load_codebook(_lib.FB_CB_PQ, pq_codebook)
        self.engine.load_pq(ids, codes)

    # ---- SRFs -------------------------------------------------------------
    @staticmethod
    def _rows(ids, dists):
        d = round_through_text(dists)
        return [(int(i), float(x)) for i, x in zip(ids.ravel(), d.ravel())]

    def ivfadc_search(self, query_bytea, k):
        """ivfadc_search(bytea, int) -> SETOF (id int4, distance float4)   freddy.c:174-410"""
        q = bytea_to_vec(query_bytea)
        ids, d = self.engine.ivfadc_search(q[None, :], int(k), self._w)
        return self._rows(ids, d)

    def ivfadc_batch_search(self, ids, k):
        """ivfadc_batch_search(int[], int) -> SETOF (query_id, id, distance)   freddy.c:677-1024"""
        oq, ri, rd = self.engine.ivfadc_batch_search(np.asarray(ids, np.int32), int(k))
        rd = round_through_text(rd)
        return [(int(q), int(i), float(x)) for q, a, b in zip(oq, ri, rd) for i, x in zip(a, b)]

    def ivpq_search_in(self, query_byteas, query_ids, k, input_ids, alpha, pvf, method, use_targetlist, confidence,
                       double_threshold):
        """ivpq_search_in(bytea[], int[], int, int[], int, int, int, bool, float4, int)
        -> SETOF (query_id, target_id, distance)   ivpq_search_in.c:61-721"""
        if len(query_byteas) != len(query_ids):
            raise ValueError(f"Number of query vectors and query vector ids differs! ( {len(query_ids)}, {len(query_byteas)})")
        q = np.stack([bytea_to_vec(b) for b in query_byteas])
        ids, d = self.engine.ivpq_search_in(q, int(k), np.asarray(input_ids, np.int32), int(alpha), int(pvf), int(method),
                                            bool(use_targetlist), float(confidence), int(double_threshold))
        d = round_through_text(d)
        return [(int(qid), int(i), float(x)) for qid, ri, rd in zip(query_ids, ids, d) for i, x in zip(ri, rd)]

    def analogy_3cosadd(self, id1, id2, id3):
        """analogy_3cosadd(w1, w2, w3) by word ids -> id of the answer   freddy--0.0.1.sql:1270-1288"""
        ids, _ = self.engine.analogy_3cosadd(np.array([[id1, id2, id3]], np.int32))
        return int(ids[0])

    # ---- SQL-level functions around the SRFs (plpgsql in the reference) ----
    def k_nearest_neighbour(self, query_bytea, k):
        """k_nearest_neighbour(bytea, int) -> TABLE (id, similarity float4)   freddy--0.0.1.sql:426-439
        (the SQL returns the word; this mirror returns the row's id)"""
        ids, s = self.engine.knn_exact(bytea_to_vec(query_bytea)[None, :], int(k))
        return [(int(i), float(x)) for i, x in zip(ids[0], s[0]) if i >= 0]

    def knn_in_exact(self, query_bytea, k, input_ids):
        """knn_in_exact(bytea, int, int[]) -> TABLE (id, similarity float4)   freddy--0.0.1.sql:1026-1038"""
        ids, s = self.engine.knn_exact(bytea_to_vec(query_bytea)[None, :], int(k), np.asarray(input_ids, np.int32))
        return [(int(i), float(x)) for i, x in zip(ids[0], s[0]) if i >= 0]

    def k_nearest_neighbour_ivfadc_pv(self, query_bytea, k):
        """k_nearest_neighbour_ivfadc_pv(bytea, int) -> TABLE (id, similarity float4)   freddy--0.0.1.sql:574-591,
        post-verification factor get_pvf(), probes get_w()"""
        ids, s = self.engine.ivfadc_search_pv(bytea_to_vec(query_bytea)[None, :], int(k), self._pvf, self._w)
        return [(int(i), float(x)) for i, x in zip(ids[0], s[0]) if i >= 0]

    def k_nearest_neighbour_pq_pv(self, query_bytea, k):
        """k_nearest_neighbour_pq_pv(bytea, int) -> TABLE (id, similarity float4)   freddy--0.0.1.sql:624-641,
        post-verification factor get_pvf()"""
        ids, s = self.engine.pq_search_pv(bytea_to_vec(query_bytea)[None, :], int(k), self._pvf)
        return [(int(i), float(x)) for i, x in zip(ids[0], s[0]) if i >= 0]

    def pq_search(self, query_bytea, k):
        """pq_search(bytea, int) -> SETOF (id, distance)   freddy.c:28-170"""
        q = bytea_to_vec(query_bytea)
        ids, d = self.engine.pq_search(q[None, :], int(k))
        return self._rows(ids, d)

    def pq_search_in(self, query_bytea, k, input_ids):
        """pq_search_in(bytea, int, int[]) -> SETOF (id, distance)   freddy.c:1028-1174"""
        q = bytea_to_vec(query_bytea)
        ids, d = self.engine.pq_search_in_batch(q[None, :], int(k), np.asarray(input_ids, np.int32))
        return self._rows(ids, d)

    def pq_search_in_batch(self, query_byteas, query_ids, k, input_ids, use_targetlist):
        """pq_search_in_batch(bytea[], int[], int, int[], bool)
        -> SETOF (query_id, target_id, distance)   freddy.c:414-675"""
        if len(query_byteas) != len(query_ids):
            # freddy.c:495
            raise ValueError("Number of query vectors and query vector ids differs!")
        q = np.stack([bytea_to_vec(b) for b in query_byteas]) if len(query_byteas) else np.zeros((0, self.engine.d), np.float32)
        ids, d = self.engine.pq_search_in_batch(q, int(k), np.asarray(input_ids, np.int32), bool(use_targetlist))
        d = round_through_text(d)
        return [(int(qid), int(i), float(x)) for qid, ri, rd in zip(query_ids, ids, d) for i, x in zip(ri, rd)]
