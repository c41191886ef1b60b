"""ctypes binding of libfreddy_b200.so (C-ABI: include/freddy_b200.h).

The library is the product; there is no Python or CPU fallback.  If the shared
object is missing or no CUDA device is present every entry point fails loudly.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "libfreddy_b200.so")

FB_OK = 0
FB_ERR_INVALID, FB_ERR_CUDA, FB_ERR_UNSUPPORTED, FB_ERR_REFERENCE_UB = -1, -2, -3, -4
FB_CB_RESIDUAL, FB_CB_PQ, FB_CB_IVPQ = 0, 1, 2
FB_OPT_FORCE_EXACT_PATH, FB_OPT_PROFILE, FB_OPT_QUERY_CHUNK, FB_OPT_QSCAN_MIN_QUERIES, FB_OPT_PACKED_FP32 = 1, 2, 3, 4, 5
FB_OPT_LUT_TILE = 6
FB_OPT_LUT_CTAS_PER_SM, FB_OPT_OVERLAP = 7, 8
FB_OPT_PIPELINE, FB_OPT_PIPE_CHUNK, FB_OPT_PIPE_DEBUG, FB_OPT_PLACEMENT_WINDOW = 9, 10, 11, 12
FB_OPT_PIPE_SHAPE, FB_OPT_PIPE_RAMP, FB_OPT_CUDA_GRAPHS, FB_OPT_ZERO_COPY_UPLOAD = 13, 14, 15, 16
FB_OPT_PREFILTER, FB_OPT_BYTE_CODES, FB_OPT_PREFILTER_LOCKSTEP, FB_OPT_DEVICE_BUILD, FB_OPT_SUBSET_PLACEMENT = 17, 18, 19, 20, 21


class Counters(C.Structure):
    _fields_ = [
        ("queries", C.c_int64), ("rows_scanned", C.c_int64), ("scan_bytes", C.c_int64),
        ("exact_path_queries", C.c_int64), ("kernel_launches", C.c_int64),
        ("ms_coarse", C.c_double), ("ms_lut", C.c_double), ("ms_scan", C.c_double),
        ("ms_finalize", C.c_double), ("ms_exact", C.c_double), ("n_scan_launches", C.c_int64),
        ("exact_coarse_tie", C.c_int64), ("exact_coarse_far", C.c_int64), ("exact_few_rows", C.c_int64),
        ("exact_scan_tie", C.c_int64), ("exact_forced", C.c_int64),
        ("ms_pipe", C.c_double), ("n_pipe_launches", C.c_int64),
        ("prefilter_queries", C.c_int64), ("prefilter_overflow_queries", C.c_int64), ("prefilter_candidates", C.c_int64),
    ]


# every symbol include/freddy_b200.h declares: (restype, argtypes)
_P = C.c_void_p
SIGNATURES = {
    "fb_create": (C.c_int, [C.c_int, C.POINTER(_P)]),
    "fb_placement_order": (C.c_int, [_P, C.c_int, C.c_int, C.c_int, C.c_int, _P]),
    "fb_destroy": (None, [_P]),
    "fb_last_error": (C.c_char_p, [_P]),
    "fb_load_coarse": (C.c_int, [_P, _P, C.c_int, C.c_int]),
    "fb_load_codebook": (C.c_int, [_P, C.c_int, _P, C.c_int, C.c_int, C.c_int]),
    "fb_load_fine": (C.c_int, [_P, _P, _P, _P, C.c_int64, C.c_int]),
    "fb_load_pq": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int]),
    "fb_ivfadc_search": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "fb_ivfadc_search_dev": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "fb_ivfadc_batch_search": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, _P, C.POINTER(C.c_int)]),
    "fb_pq_search": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P]),
    "fb_pq_search_in_batch": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, C.c_int, C.c_int, _P, _P]),
    "fb_load_ivpq": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, _P, _P, C.c_int64, C.c_int, _P]),
    "fb_ivpq_search_in": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_float,
                                    C.c_int, _P, _P]),
    "fb_ivpq_statistics": (C.c_int, [_P, _P, C.c_int64, _P, C.c_int]),
    "fb_table_checksum": (C.c_int, [_P, C.c_int, _P]),
    "fb_sidecar_start": (C.c_int, [_P, C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_void_p)]),
    "fb_sidecar_running": (C.c_int, [_P]),
    "fb_sidecar_stop": (C.c_int, [_P, _P]),
    "fb_load_vectors": (C.c_int, [_P, _P, _P, C.c_int64, C.c_int]),
    "fb_encode_ivfadc": (C.c_int, [_P, _P, C.c_int64, _P, _P]),
    "fb_encode_pq": (C.c_int, [_P, C.c_int, _P, C.c_int64, _P]),
    "fb_encode_ivfadc_dev": (C.c_int, [_P, _P, C.c_int64, _P, _P]),
    "fb_encode_pq_dev": (C.c_int, [_P, C.c_int, _P, C.c_int64, _P]),
    "fb_append_fine": (C.c_int, [_P, _P, _P, _P, C.c_int64]),
    "fb_append_pq": (C.c_int, [_P, C.c_int, _P, _P, _P, C.c_int64]),
    "fb_append_vectors": (C.c_int, [_P, _P, _P, C.c_int64]),
    "fb_grouping_pq": (C.c_int, [_P, _P, C.c_int, _P, C.c_int, _P, _P, C.POINTER(C.c_int)]),
    "fb_knn_exact": (C.c_int, [_P, _P, C.c_int, C.c_int, _P, C.c_int, _P, _P]),
    "fb_ivfadc_search_pv": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, C.c_int, _P, _P]),
    "fb_pq_search_pv": (C.c_int, [_P, _P, C.c_int, C.c_int, C.c_int, _P, _P]),
    "fb_cosine_similarity": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, C.c_int, _P]),
    "fb_vec_op": (C.c_int, [_P, C.c_int, _P, _P, C.c_int, C.c_int, _P]),
    "fb_analogy_3cosadd": (C.c_int, [_P, _P, C.c_int, _P, _P]),
    "fb_analogy_scan": (C.c_int, [_P, _P, _P, C.c_int, _P, _P]),
    "fb_synchronize": (C.c_int, [_P]),
    "fb_set_stream": (C.c_int, [_P, _P]),
    "fb_set_option": (C.c_int, [_P, C.c_int, C.c_int64]),
    "fb_get_counters": (C.c_int, [_P, C.POINTER(Counters)]),
    "fb_reset_counters": (C.c_int, [_P]),
    "fb_round_through_text": (C.c_float, [C.c_float]),
    "fb_version": (C.c_char_p, []),
}

_lib = None


def load():
    """dlopen the product library and bind every declared symbol."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python __graft_entry__.py` or "
            "`make -C postgres-word2vec_b200/csrc` (there is no fallback implementation)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
