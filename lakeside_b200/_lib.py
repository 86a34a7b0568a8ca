"""ctypes binding of liblakeside_b200.so (the same calling convention a JNA / Panama binding uses).

The library is built in-tree by ``__graft_entry__.build()`` / ``make -C lakeside_b200/csrc``.  There is no Python or
CPU fallback: if the shared object is missing, or no CUDA device is visible, every compute entry point raises.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import POINTER, c_char_p, c_double, c_int, c_int32, c_int64, c_size_t, c_uint8, c_void_p

LK_OK, LK_ERR_INVALID, LK_ERR_UNSUPPORTED, LK_ERR_IO, LK_ERR_CUDA, LK_ERR_QUERY, LK_ERR_NOMEM = range(7)
_CODE_NAMES = {1: "INVALID", 2: "UNSUPPORTED", 3: "IO", 4: "CUDA", 5: "QUERY", 6: "NOMEM"}

# LK_LIB selects another in-tree build of the same library (kernel tuning experiments); there is still no fallback
LIB_PATH = os.environ.get("LK_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "liblakeside_b200.so")


class LakesideError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"LK_ERR_{_CODE_NAMES.get(code, code)}: {message}")
        self.code = code
        self.message = message


class LakesideUnsupported(LakesideError):
    """LK_ERR_UNSUPPORTED: the query/file shape is outside the GPU path (the adapter streams nothing)."""


class LakesideQueryError(LakesideError):
    """LK_ERR_QUERY: the reference's SQL would fail to bind in DuckDB (Commons.scala:249-253 -> empty stream)."""


_SIGNATURES = {
    "lk_init": (c_int, [c_char_p]),
    "lk_shutdown": (None, []),
    "lk_last_error": (c_char_p, []),
    "lk_version": (c_char_p, []),
    "lk_device_count": (c_int, []),
    "lk_cache_stats": (c_int, [POINTER(ctypes.c_int64)]),
    "lk_cache_configure": (c_int, [ctypes.c_int64]),
    "lk_cache_clear": (None, []),
    "lk_host_alloc": (c_void_p, [c_size_t]),
    "lk_host_free": (None, [c_void_p]),
    "lk_eval": (c_int, [c_char_p, POINTER(c_char_p), c_int, POINTER(c_void_p)]),
    "lk_query_create": (c_int, [c_char_p, c_char_p, POINTER(c_void_p)]),
    "lk_query_add_segment_file": (c_int, [c_void_p, c_char_p]),
    "lk_query_add_segment_buffer": (c_int, [c_void_p, c_void_p, c_size_t]),
    "lk_query_prepare": (c_int, [c_void_p]),
    "lk_query_export_dictionaries": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_size_t)]),
    "lk_query_import_dictionaries": (c_int, [c_void_p, c_void_p, c_size_t]),
    "lk_query_execute": (c_int, [c_void_p]),
    "lk_query_sync": (c_int, [c_void_p]),
    "lk_query_partial_dense": (c_int, [c_void_p, POINTER(c_int64), POINTER(c_int), POINTER(c_void_p), POINTER(c_int)]),
    "lk_query_partial_sparse": (c_int, [c_void_p, c_int, POINTER(c_void_p), POINTER(c_int64), POINTER(c_int)]),
    "lk_query_plan": (c_int, [c_void_p]),
    "lk_query_merge_sparse": (c_int, [c_void_p, c_void_p, c_int64]),
    "lk_query_finalize_device": (c_int, [c_void_p]),
    "lk_query_finalize": (c_int, [c_void_p, POINTER(c_void_p)]),
    "lk_query_survivors": (c_int64, [c_void_p]),
    "lk_query_phase": (c_int, [c_void_p, POINTER(ctypes.c_uint32), POINTER(ctypes.c_uint32)]),
    "lk_query_set_phase": (c_int, [c_void_p, ctypes.c_uint32, ctypes.c_uint32]),
    "lk_formula_eval": (c_int, [c_void_p, c_void_p, c_char_p, c_int64, POINTER(c_int64), POINTER(c_double), POINTER(c_int32), POINTER(c_int64), POINTER(c_int64)]),
    "lk_query_eval": (c_int64, [c_void_p, c_char_p, c_char_p, c_char_p, POINTER(c_double), c_int64]),
    "lk_query_timings": (c_int, [c_void_p, POINTER(c_double)]),
    "lk_query_touched_bytes": (c_int64, [c_void_p]),
    "lk_query_total_rows": (c_int64, [c_void_p]),
    "lk_query_stream": (c_int, [c_void_p, POINTER(c_void_p)]),
    "lk_query_info_json": (c_int, [c_void_p, POINTER(c_char_p)]),
    "lk_query_destroy": (None, [c_void_p]),
    "lk_result_num_rows": (c_int64, [c_void_p]),
    "lk_result_to_sse": (c_int64, [c_void_p, c_int64, c_int64, POINTER(c_char_p), c_int, POINTER(c_char_p), c_int, c_void_p, c_int64]),
    "lk_result_num_values": (c_int, [c_void_p]),
    "lk_result_num_tags": (c_int, [c_void_p]),
    "lk_result_num_cols": (c_int, [c_void_p]),
    "lk_result_col_name": (c_char_p, [c_void_p, c_int]),
    "lk_result_ts": (POINTER(c_int64), [c_void_p]),
    "lk_result_value": (POINTER(c_double), [c_void_p, c_int]),
    "lk_result_value_null": (POINTER(c_uint8), [c_void_p, c_int]),
    "lk_result_tag_codes": (POINTER(c_int32), [c_void_p, c_int]),
    "lk_result_tag_dict": (c_int, [c_void_p, c_int, POINTER(c_int32), POINTER(POINTER(c_char_p))]),
    "lk_result_get_long": (c_int64, [c_void_p, c_int64, c_int]),
    "lk_result_get_double": (c_double, [c_void_p, c_int64, c_int]),
    "lk_result_get_string": (c_char_p, [c_void_p, c_int64, c_int]),
    "lk_result_free": (None, [c_void_p]),
    "lk_merge_streams": (c_int, [c_int, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int64), c_int,
                                 c_void_p, c_void_p, c_void_p, c_void_p]),
    "lk_merge_create": (c_int, [c_int, POINTER(c_void_p), POINTER(c_void_p), POINTER(c_void_p), POINTER(c_int64), c_int, POINTER(c_void_p)]),
    "lk_merge_run": (c_int, [c_void_p]),
    "lk_merge_sync": (c_int, [c_void_p]),
    "lk_merge_timings": (c_int, [c_void_p, POINTER(c_double)]),
    "lk_merge_download": (c_int, [c_void_p, c_void_p, c_void_p, c_void_p, c_void_p]),
    "lk_merge_reduce": (c_int, [c_void_p, c_int, POINTER(c_int64), c_void_p, c_void_p, c_void_p]),
    "lk_merge_destroy": (None, [c_void_p]),
    "lk_comm_create": (c_int, [c_int, c_int, c_int64, c_int, POINTER(c_void_p)]),
    "lk_comm_handle": (c_int, [c_void_p, POINTER(c_void_p), POINTER(c_size_t)]),
    "lk_comm_connect": (c_int, [c_void_p, c_char_p, c_size_t]),
    "lk_comm_destroy": (None, [c_void_p]),
    "lk_query_set_comm": (c_int, [c_void_p, c_void_p]),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)
_lib = None


def load() -> ctypes.CDLL:
    """Loads liblakeside_b200.so; raises (loudly) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(make -C lakeside_b200/csrc).  lakeside_b200 has no CPU fallback.")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGNATURES.items():
            fn = getattr(lib, name)
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def check(rc: int):
    if rc != LK_OK:
        msg = (load().lk_last_error() or b"").decode("utf-8", "replace")
        cls = LakesideUnsupported if rc == LK_ERR_UNSUPPORTED else LakesideQueryError if rc == LK_ERR_QUERY else LakesideError
        raise cls(rc, msg)
