"""ctypes view of include/probly_b200.h (the C ABI).  Structures mirror the header field by field.

The shared library is built in-tree by probly_search_b200/build.py.  There is no Python or CPU
implementation behind these entry points: if the library cannot be loaded the import fails, and
every query entry point returns PB_ERR_NO_DEVICE on a box without a GPU.
"""
from __future__ import annotations

import ctypes as C
import os

PB_OK = 0
PB_ERR_INVALID = -1
PB_ERR_CUDA = -2
PB_ERR_NO_DEVICE = -3
PB_ERR_CAPACITY = -4
PB_ERR_UNSUPPORTED = -5
PB_ERR_DUPLICATE_KEY = -6
PB_ERR_NOMEM = -7
PB_SCORER_BM25 = 0
PB_SCORER_ZERO_TO_ONE = 1
PB_MAX_FIELDS = 4
PB_MAX_TOP_K = 32

u8p = C.POINTER(C.c_uint8)
u32p = C.POINTER(C.c_uint32)
u64p = C.POINTER(C.c_uint64)
f64p = C.POINTER(C.c_double)


class DocTokens(C.Structure):
    _fields_ = [("tok_bytes", C.c_void_p), ("tok_off", C.c_void_p), ("value_tok_count", C.c_void_p),
                ("field_value_count", C.c_void_p)]


class BuilderInfo(C.Structure):
    _fields_ = [("num_fields", C.c_uint32), ("n_live_docs", C.c_uint64), ("n_doc_ordinals", C.c_uint64),
                ("n_removed_pending", C.c_uint64), ("n_terms", C.c_uint64), ("n_nodes", C.c_uint64),
                ("n_rows", C.c_uint64), ("n_pointers", C.c_uint64), ("field_sum", C.c_uint64 * PB_MAX_FIELDS),
                ("field_avg", C.c_double * PB_MAX_FIELDS)]


class IndexImage(C.Structure):
    _fields_ = [("version", C.c_uint32), ("num_fields", C.c_uint32),
                ("n_nodes", C.c_uint64), ("n_edges", C.c_uint64), ("n_terms", C.c_uint64),
                ("n_rows", C.c_uint64), ("n_rows_padded", C.c_uint64), ("n_docs", C.c_uint64),
                ("max_term_bytes", C.c_uint32), ("max_tf", C.c_uint32 * PB_MAX_FIELDS),
                ("max_fl", C.c_uint32 * PB_MAX_FIELDS),
                ("node_edge_begin", u32p), ("node_term_lo", u32p), ("node_term_hi", u32p),
                ("node_parent", u32p), ("node_char", u32p), ("edge_char", u32p), ("edge_child", u32p),
                ("term_row_begin", u64p), ("term_byte_len", u32p), ("term_node", u32p),
                ("post_blocks", u32p),
                ("doc_key", u64p), ("removed_bitmap", u32p), ("n_removed", C.c_uint64),
                ("n_live_docs", C.c_uint64), ("field_avg", C.c_double * PB_MAX_FIELDS)]


class DeviceLayout(C.Structure):
    _fields_ = [("narrow", C.c_uint32), ("bytes_per_row", C.c_uint32), ("fl_bits", C.c_uint32 * PB_MAX_FIELDS),
                ("posting_bytes", C.c_uint64)]


class QueryBatchDesc(C.Structure):
    _fields_ = [("n_queries", C.c_uint64), ("query_term_off", C.c_void_p), ("term_byte_off", C.c_void_p),
                ("term_bytes", C.c_void_p), ("scorer", C.c_uint32), ("bm25_k1", C.c_double),
                ("bm25_b", C.c_double), ("fields_boost", C.c_void_p), ("n_fields_boost", C.c_uint32),
                ("top_k", C.c_uint32)]


class QueryResults(C.Structure):
    _fields_ = [("n_results", C.c_void_p), ("doc_digest", C.c_void_p), ("score_digest", C.c_void_p),
                ("topk_n", C.c_void_p), ("topk_doc", C.c_void_p), ("topk_score", C.c_void_p)]


class BatchStats(C.Structure):
    _fields_ = [("n_queries", C.c_uint64), ("n_query_terms", C.c_uint64), ("n_segments", C.c_uint64),
                ("rows_streamed", C.c_uint64), ("rows_streamed_direct", C.c_uint64),
                ("rows_scored", C.c_uint64), ("rows_diverted", C.c_uint64), ("legacy_records", C.c_uint64),
                ("results_emitted", C.c_uint64), ("pointer_visits", C.c_uint64), ("gpu_launches", C.c_uint32),
                ("side_rounds", C.c_uint32), ("ms_total", C.c_float), ("ms_descend", C.c_float),
                ("ms_plan", C.c_float), ("ms_score", C.c_float), ("ms_side", C.c_float),
                ("ms_finalize", C.c_float), ("score_launches", C.c_uint32), ("ms_gather", C.c_float),
                ("ms_side_mark", C.c_float), ("ms_side_score", C.c_float), ("ms_side_fold", C.c_float),
                ("ms_union", C.c_float), ("rows_streamed_side", C.c_uint64), ("rows_streamed_union", C.c_uint64),
                ("union_queries", C.c_uint64), ("rows_streamed_compact", C.c_uint64)]

    def as_dict(self) -> dict:
        return {n: getattr(self, n) for n, _ in self._fields_}


# every symbol include/probly_b200.h declares (checked by tests/test_capi_symbols.py)
EXPORTS = [
    "pb_builder_create", "pb_builder_destroy", "pb_builder_add_document", "pb_builder_add_documents",
    "pb_builder_remove_document", "pb_builder_vacuum", "pb_builder_get_info", "pb_builder_flatten",
    "pb_device_count", "pb_index_create", "pb_index_set_live_state", "pb_index_destroy",
    "pb_index_expand_term", "pb_index_term_df_live", "pb_index_device_layout", "pb_device_read_bandwidth",
    "pb_image_save", "pb_image_load", "pb_image_file_image", "pb_image_file_free", "pb_query_batch", "pb_batch_create", "pb_batch_run",
    "pb_batch_fetch", "pb_batch_destroy", "pb_batch_device_results", "pb_batch_get_stats", "pb_index_last_stats", "pb_query_full",
    "pb_host_alloc", "pb_host_free", "pb_last_error", "pb_version",
    "pb_comm_unique_id", "pb_comm_create", "pb_comm_info", "pb_comm_destroy",
    "pb_batch_set_gather", "pb_batch_reload", "pb_batch_run_local", "pb_batch_gather", "pb_batch_sync",
    "pb_batch_fetch_gathered", "pb_batch_device_gathered",
    "pb_group_create", "pb_group_size", "pb_group_set_live_state", "pb_group_query_batch", "pb_group_member_stats",
    "pb_group_destroy",
    "pb_builder_flatten_from", "pb_builder_flatten_term_ids", "pb_index_set_df_extra", "pb_index_attach_delta", "pb_index_create_from_builder", "pb_builder_flatten_structure",
]
PB_COMM_ID_BYTES = 128

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PB_LIB_PATH") or os.path.join(_HERE, "_lib", "libprobly_b200.so")   # PB_LIB_PATH: tuning variants
_lib = None


class ProblyError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"probly_b200 error {code}: {msg}")
        self.code = code


def lib() -> C.CDLL:
    """Loads libprobly_b200.so; builds it first if the sources are newer (needs nvcc)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build as _build
        _build.build()
    L = C.CDLL(LIB_PATH)
    vp, u32, u64, i32 = C.c_void_p, C.c_uint32, C.c_uint64, C.c_int
    P = C.POINTER
    sig = {
        "pb_builder_create": (i32, [u32, P(vp)]),
        "pb_builder_destroy": (None, [vp]),
        "pb_builder_add_document": (i32, [vp, u64, P(DocTokens)]),
        "pb_builder_add_documents": (i32, [vp, u64, vp, vp, vp, vp]),
        "pb_builder_remove_document": (i32, [vp, u64]),
        "pb_builder_vacuum": (i32, [vp]),
        "pb_builder_get_info": (i32, [vp, P(BuilderInfo)]),
        "pb_builder_flatten": (i32, [vp, P(IndexImage)]),
        "pb_device_count": (i32, []),
        "pb_index_create": (i32, [P(IndexImage), i32, P(vp)]),
        "pb_index_set_live_state": (i32, [vp, vp, u64, u64, vp]),
        "pb_index_destroy": (None, [vp]),
        "pb_index_expand_term": (i32, [vp, vp, u64, vp, u64, P(u64), P(u64)]),
        "pb_index_term_df_live": (i32, [vp, vp, u64]),
        "pb_index_device_layout": (i32, [vp, vp]),
        "pb_image_save": (i32, [vp, C.c_char_p]),
        "pb_image_load": (i32, [C.c_char_p, vp]),
        "pb_image_file_image": (vp, [vp]),
        "pb_image_file_free": (None, [vp]),
        "pb_device_read_bandwidth": (i32, [i32, u64, C.c_uint32, vp]),
        "pb_query_batch": (i32, [vp, P(QueryBatchDesc), P(QueryResults)]),
        "pb_batch_create": (i32, [vp, P(QueryBatchDesc), P(vp)]),
        "pb_batch_run": (i32, [vp]),
        "pb_batch_fetch": (i32, [vp, P(QueryResults)]),
        "pb_batch_destroy": (None, [vp]),
        "pb_batch_device_results": (i32, [vp, P(QueryResults)]),
        "pb_batch_get_stats": (i32, [vp, P(BatchStats)]),
        "pb_index_last_stats": (i32, [vp, P(BatchStats)]),
        "pb_query_full": (i32, [vp, P(QueryBatchDesc), u64, vp, vp, vp, P(u64)]),
        "pb_host_alloc": (vp, [C.c_size_t]),
        "pb_host_free": (None, [vp]),
        "pb_last_error": (C.c_char_p, []),
        "pb_version": (C.c_char_p, []),
        "pb_comm_unique_id": (i32, [vp]),
        "pb_comm_create": (i32, [vp, i32, i32, i32, P(vp)]),
        "pb_comm_info": (i32, [vp, P(i32), P(i32), P(i32)]),
        "pb_comm_destroy": (None, [vp]),
        "pb_batch_set_gather": (i32, [vp, vp, u64]),
        "pb_batch_reload": (i32, [vp, P(QueryBatchDesc)]),
        "pb_batch_run_local": (i32, [vp]),
        "pb_batch_gather": (i32, [vp]),
        "pb_batch_sync": (i32, [vp]),
        "pb_batch_fetch_gathered": (i32, [vp, u64, P(QueryResults)]),
        "pb_batch_device_gathered": (i32, [vp, P(vp), P(u64), P(u64)]),
        "pb_group_create": (i32, [P(IndexImage), P(i32), i32, P(vp)]),
        "pb_group_size": (i32, [vp]),
        "pb_group_set_live_state": (i32, [vp, vp, u64, u64, vp]),
        "pb_group_query_batch": (i32, [vp, P(QueryBatchDesc), P(QueryResults)]),
        "pb_group_member_stats": (i32, [vp, i32, P(BatchStats)]),
        "pb_group_destroy": (None, [vp]),
        "pb_builder_flatten_from": (i32, [vp, u64, P(IndexImage)]),
        "pb_builder_flatten_term_ids": (i32, [vp, vp, u64, P(u64)]),
        "pb_index_set_df_extra": (i32, [vp, vp, u64]),
        "pb_index_attach_delta": (i32, [vp, vp, vp, u64, vp, u64]),
        "pb_index_create_from_builder": (i32, [vp, u64, i32, P(vp)]),
        "pb_builder_flatten_structure": (i32, [vp, u64, P(IndexImage)]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != PB_OK:
        raise ProblyError(rc, lib().pb_last_error().decode("utf-8", "replace"))


def last_error() -> str:
    return lib().pb_last_error().decode("utf-8", "replace")
