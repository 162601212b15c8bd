"""ctypes binding of libotters_b200.so (include/otters_b200.h).

The library is the product: if it is missing, importing this module fails loudly — there is no
Python/NumPy/CPU fallback anywhere in this package.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libotters_b200.so")

if not os.path.exists(LIB_PATH):
    raise ImportError(
        f"{LIB_PATH} is missing: build the CUDA library first "
        "(python -c 'import __graft_entry__ as g; g.build()' or make -C otters_b200/csrc). "
        "otters_b200 has no CPU fallback."
    )

lib = C.CDLL(LIB_PATH)

c_u64p = C.POINTER(C.c_uint64)
c_u32p = C.POINTER(C.c_uint32)
c_f32p = C.POINTER(C.c_float)
c_u8p = C.POINTER(C.c_uint8)


class ScanTuning(C.Structure):
    _fields_ = [
        ("warps_per_cta", C.c_uint32),
        ("slots_per_warp", C.c_uint32),
        ("kc_floats", C.c_uint32),
        ("ctas_per_sm", C.c_uint32),
        ("unit_rows", C.c_uint32),
        ("disable_fused_predicate", C.c_uint32),
        ("batch_mode", C.c_uint32),
        ("batch_cta_group", C.c_uint32),
        ("scan_mode", C.c_uint32),
        ("planners", C.c_uint32),
        ("timing", C.c_uint32),
        ("batch_passes", C.c_uint32),
        ("separate_select", C.c_uint32),
        ("lazy_prune", C.c_uint32),
    ]


class LastWork(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_uint64),
        ("rows_scored", C.c_uint64),
        ("scan_bytes", C.c_uint64),
        ("meta_bytes", C.c_uint64),
        ("scan_ms", C.c_float),
        ("prune_ms", C.c_float),
        ("rowmask_ms", C.c_float),
        ("select_ms", C.c_float),
        ("batch_used", C.c_uint32),
        ("batch_fallback", C.c_uint32),
        ("batch_candidates", C.c_uint64),
        ("batch_max_err", C.c_float),
        ("batch_delta", C.c_float),
        ("h2d_bytes", C.c_uint64),
        ("d2h_bytes", C.c_uint64),
        ("batch_passes", C.c_uint32),
        ("batch_attempts", C.c_uint32),
    ]


class VecQuery(C.Structure):
    _fields_ = [
        ("queries", c_f32p),
        ("nq", C.c_uint32),
        ("dim", C.c_uint32),
        ("metric", C.c_int32),
        ("take_type", C.c_int32),
        ("k", C.c_uint64),
        ("has_filter", C.c_int32),
        ("thr", C.c_float),
        ("cmp", C.c_int32),
        ("row_mask_words", c_u64p),
        ("row_mask_bits", C.c_uint64),
    ]


class Column(C.Structure):
    _fields_ = [
        ("name", C.c_char_p),
        ("dtype", C.c_int32),
        ("values", C.c_void_p),
        ("null_words", c_u64p),
        ("str_offsets", c_u64p),
        ("str_bytes", c_u8p),
    ]


class BuildParams(C.Structure):
    _fields_ = [
        ("n_rows", C.c_uint64),
        ("dim", C.c_uint32),
        ("chunk_size", C.c_uint64),
        ("bloom_mode", C.c_int32),
        ("bloom_fpr", C.c_double),
        ("bloom_bits", C.c_uint64),
        ("vectors_kind", C.c_int32),
        ("vectors", C.c_void_p),
        ("synthetic_seed", C.c_uint64),
        ("synthetic_first_row", C.c_uint64),
        ("synthetic_map", C.c_void_p),
        ("columns", C.POINTER(Column)),
        ("n_columns", C.c_uint32),
        ("vector_format", C.c_int32),
    ]


class BuildStats(C.Structure):
    _fields_ = [
        ("n_rows", C.c_uint64),
        ("dim", C.c_uint64),
        ("n_chunks", C.c_uint64),
        ("vectors_ingest_s", C.c_double),
        ("zonemap_build_s", C.c_double),
        ("build_total_s", C.c_double),
    ]


class QueryStats(C.Structure):
    _fields_ = [
        ("total_chunks", C.c_uint64),
        ("pruned_chunks", C.c_uint64),
        ("evaluated_chunks", C.c_uint64),
        ("vectors_compared", C.c_uint64),
        ("prune_s", C.c_double),
        ("score_s", C.c_double),
        ("merge_s", C.c_double),
        ("total_s", C.c_double),
    ]


class Leaf(C.Structure):
    _fields_ = [
        ("col", C.c_uint32),
        ("op", C.c_int32),
        ("kind", C.c_int32),
        ("i", C.c_int64),
        ("f", C.c_double),
        ("s", C.c_char_p),
        ("slen", C.c_uint64),
    ]


class Filter(C.Structure):
    _fields_ = [
        ("n_clauses", C.c_uint32),
        ("clause_offsets", c_u32p),
        ("leaves", C.POINTER(Leaf)),
    ]


class ShardMap(C.Structure):
    _fields_ = [("row_base", C.c_uint64), ("world", C.c_uint32), ("rank", C.c_uint32), ("block_rows", C.c_uint64)]


class TopkRecord(C.Structure):
    _fields_ = [("row", C.c_uint64), ("score", C.c_float), ("qid", C.c_uint32)]


class PeerExchange(C.Structure):
    _fields_ = [
        ("world", C.c_uint32),
        ("rank", C.c_uint32),
        ("k_max", C.c_uint64),
        ("peer_records", C.POINTER(C.c_void_p)),
        ("peer_flags", C.POINTER(C.c_void_p)),
    ]


VECTORS_HOST, VECTORS_DEVICE, VECTORS_SYNTHETIC = 0, 1, 2

_p = C.c_void_p


def _sig(name, restype, *argtypes):
    fn = getattr(lib, name)
    fn.restype = restype
    fn.argtypes = list(argtypes)
    return fn


# every symbol include/otters_b200.h declares (tests/test_abi_exports.py checks the two lists agree)
otters_ctx_create = _sig("otters_ctx_create", C.c_int, C.c_int, _p, C.POINTER(_p))
otters_ctx_destroy = _sig("otters_ctx_destroy", C.c_int, _p)
otters_ctx_synchronize = _sig("otters_ctx_synchronize", C.c_int, _p)
otters_ctx_join = _sig("otters_ctx_join", C.c_int, _p)
otters_last_error = _sig("otters_last_error", C.c_char_p)
otters_version = _sig("otters_version", C.c_char_p)
otters_ctx_set_tuning = _sig("otters_ctx_set_tuning", C.c_int, _p, C.POINTER(ScanTuning))
otters_ctx_last_work = _sig("otters_ctx_last_work", C.c_int, _p, C.POINTER(LastWork))
otters_vecstore_create = _sig("otters_vecstore_create", C.c_int, _p, C.c_uint32, C.POINTER(_p))
otters_vecstore_create_fmt = _sig("otters_vecstore_create_fmt", C.c_int, _p, C.c_uint32, C.c_int32, C.POINTER(_p))
otters_vecstore_format = _sig("otters_vecstore_format", C.c_int32, _p)
otters_vecstore_destroy = _sig("otters_vecstore_destroy", C.c_int, _p)
otters_vecstore_reserve = _sig("otters_vecstore_reserve", C.c_int, _p, C.c_uint64)
otters_vecstore_add = _sig("otters_vecstore_add", C.c_int, _p, c_f32p, C.c_uint64)
otters_vecstore_add_device = _sig("otters_vecstore_add_device", C.c_int, _p, _p, C.c_uint64)
otters_vecstore_add_synthetic = _sig("otters_vecstore_add_synthetic", C.c_int, _p, C.c_uint64, C.c_uint64, C.c_uint64)
otters_vecstore_set_rows = _sig("otters_vecstore_set_rows", C.c_int, _p, c_u64p, c_f32p, C.c_uint64)
otters_metastore_set_rows = _sig("otters_metastore_set_rows", C.c_int, _p, c_u64p, c_f32p, C.c_uint64)
otters_vecstore_len = _sig("otters_vecstore_len", C.c_uint64, _p)
otters_vecstore_dim = _sig("otters_vecstore_dim", C.c_uint32, _p)
otters_vecstore_inv_norms = _sig("otters_vecstore_inv_norms", C.c_int, _p, C.c_uint64, C.c_uint64, c_f32p)
otters_vecstore_query = _sig(
    "otters_vecstore_query", C.c_int, _p, C.POINTER(VecQuery), c_u64p, c_f32p, c_u32p, C.c_uint64, c_u64p
)
otters_metastore_build = _sig(
    "otters_metastore_build", C.c_int, _p, C.POINTER(BuildParams), C.POINTER(_p), C.POINTER(BuildStats)
)
otters_metastore_save = _sig("otters_metastore_save", C.c_int, _p, C.c_char_p, _p, C.c_uint64)
otters_metastore_load = _sig("otters_metastore_load", C.c_int, _p, C.c_char_p, C.POINTER(_p))
otters_metastore_user_blob = _sig("otters_metastore_user_blob", C.c_int, _p, C.POINTER(_p), c_u64p)
otters_metastore_n_columns = _sig("otters_metastore_n_columns", C.c_uint32, _p)
otters_metastore_column_info = _sig("otters_metastore_column_info", C.c_int, _p, C.c_uint32, C.POINTER(C.c_char_p), C.POINTER(C.c_int32))
otters_metastore_dim = _sig("otters_metastore_dim", C.c_uint32, _p)
otters_metastore_format = _sig("otters_metastore_format", C.c_int32, _p)
otters_metastore_destroy = _sig("otters_metastore_destroy", C.c_int, _p)
otters_metastore_n_chunks = _sig("otters_metastore_n_chunks", C.c_uint64, _p)
otters_metastore_chunk_size = _sig("otters_metastore_chunk_size", C.c_uint64, _p)
otters_metastore_len = _sig("otters_metastore_len", C.c_uint64, _p)
otters_metastore_query = _sig(
    "otters_metastore_query",
    C.c_int,
    _p,
    C.POINTER(VecQuery),
    C.POINTER(Filter),
    c_u64p,
    c_f32p,
    c_u32p,
    C.c_uint64,
    c_u64p,
    C.POINTER(QueryStats),
)
otters_vecstore_query_batch = _sig("otters_vecstore_query_batch", C.c_int, _p, C.POINTER(VecQuery), c_u64p, c_f32p, c_u64p)
otters_metastore_query_batch = _sig(
    "otters_metastore_query_batch", C.c_int, _p, C.POINTER(VecQuery), C.POINTER(Filter), c_u64p, c_f32p, c_u64p, C.POINTER(QueryStats)
)
otters_metastore_gather = _sig("otters_metastore_gather", C.c_int, _p, C.c_uint32, c_u64p, C.c_uint64, _p, c_u8p)
otters_metastore_dict_entry = _sig("otters_metastore_dict_entry", C.c_int, _p, C.c_uint32, C.c_uint32, C.POINTER(c_u8p), c_u64p)
otters_metastore_last_stats = _sig("otters_metastore_last_stats", C.c_int, _p, C.POINTER(QueryStats))
otters_metastore_chunk_mask = _sig("otters_metastore_chunk_mask", C.c_int, _p, C.POINTER(Filter), c_u8p)
otters_metastore_row_mask = _sig("otters_metastore_row_mask", C.c_int, _p, C.POINTER(Filter), c_u8p)
otters_metastore_zonemap_i64 = _sig(
    "otters_metastore_zonemap_i64", C.c_int, _p, C.c_uint32, C.POINTER(C.c_int64), C.POINTER(C.c_int64), c_u64p
)
otters_metastore_zonemap_f64 = _sig(
    "otters_metastore_zonemap_f64", C.c_int, _p, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double), c_u64p
)
otters_metastore_inv_norms = _sig("otters_metastore_inv_norms", C.c_int, _p, C.c_uint64, C.c_uint64, c_f32p)
otters_query_local_device = _sig(
    "otters_query_local_device", C.c_int, _p, _p, C.POINTER(VecQuery), C.POINTER(Filter), C.POINTER(ShardMap), _p, C.POINTER(QueryStats)
)
otters_vecstore_add_synthetic_sharded = _sig(
    "otters_vecstore_add_synthetic_sharded", C.c_int, _p, C.POINTER(ShardMap), C.c_uint64, C.c_uint64
)
otters_topk_merge_device = _sig(
    "otters_topk_merge_device", C.c_int, _p, _p, C.c_uint64, C.c_uint64, C.c_int32, c_u64p, c_f32p, c_u32p, C.c_uint64, c_u64p
)

otters_query_exchange = _sig(
    "otters_query_exchange", C.c_int, _p, _p, C.POINTER(VecQuery), C.POINTER(Filter), C.POINTER(ShardMap), C.POINTER(PeerExchange),
    C.c_uint64, c_u64p, c_f32p, c_u32p, C.c_uint64, c_u64p, C.POINTER(QueryStats)
)
otters_query_submit = _sig(
    "otters_query_submit", C.c_int, _p, _p, C.POINTER(VecQuery), C.POINTER(Filter), C.POINTER(ShardMap), C.POINTER(PeerExchange),
    C.c_uint64, c_u64p
)
otters_query_wait = _sig(
    "otters_query_wait", C.c_int, _p, C.c_uint64, c_u64p, c_f32p, c_u32p, C.c_uint64, c_u64p, C.POINTER(QueryStats)
)
EXCHANGE_SLOTS = 4  # OTTERS_EXCHANGE_SLOTS

BOUND_SYMBOLS = sorted(n for n in dir() if n.startswith("otters_"))


def last_error() -> str:
    msg = otters_last_error()
    return msg.decode("utf-8", "replace") if msg else ""
