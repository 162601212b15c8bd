//! Raw FFI declarations of libotters_b200.so — GENERATED from include/otters_b200.h by scripts/gen_rust_sys.py; do not edit.
//! Every item mirrors the C header one to one; see the header for the reference lines each entry point replaces.
//! tests/test_rust_sys.py keeps this file in step with the header and checks every repr(C) layout.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const OTTERS_OK: c_int = 0;
pub const OTTERS_ERR_INVALID: c_int = 1;
pub const OTTERS_ERR_CUDA: c_int = 2;
pub const OTTERS_ERR_NOMEM: c_int = 3;
pub const OTTERS_ERR_UNSUPPORTED: c_int = 4;
pub const OTTERS_METRIC_COSINE: c_int = 0;
pub const OTTERS_METRIC_EUCLIDEAN: c_int = 1;
pub const OTTERS_METRIC_DOT: c_int = 2;
pub const OTTERS_TAKE_MIN: c_int = 0;
pub const OTTERS_TAKE_MAX: c_int = 1;
pub const OTTERS_CMP_LT: c_int = 0;
pub const OTTERS_CMP_GT: c_int = 1;
pub const OTTERS_CMP_LTE: c_int = 2;
pub const OTTERS_CMP_GTE: c_int = 3;
pub const OTTERS_CMP_EQ: c_int = 4;
pub const OTTERS_OP_EQ: c_int = 0;
pub const OTTERS_OP_NEQ: c_int = 1;
pub const OTTERS_OP_LT: c_int = 2;
pub const OTTERS_OP_LTE: c_int = 3;
pub const OTTERS_OP_GT: c_int = 4;
pub const OTTERS_OP_GTE: c_int = 5;
pub const OTTERS_DTYPE_INT32: c_int = 0;
pub const OTTERS_DTYPE_INT64: c_int = 1;
pub const OTTERS_DTYPE_FLOAT32: c_int = 2;
pub const OTTERS_DTYPE_FLOAT64: c_int = 3;
pub const OTTERS_DTYPE_STRING: c_int = 4;
pub const OTTERS_DTYPE_DATETIME: c_int = 5;
pub const OTTERS_LIT_I64: c_int = 0;
pub const OTTERS_LIT_F64: c_int = 1;
pub const OTTERS_LIT_STR: c_int = 2;
pub const OTTERS_VECTORS_FMT_F32: c_int = 0;
pub const OTTERS_VECTORS_FMT_BF16: c_int = 1;
pub const OTTERS_VECTORS_HOST: c_int = 0;
pub const OTTERS_VECTORS_DEVICE: c_int = 1;
pub const OTTERS_VECTORS_SYNTHETIC: c_int = 2;
pub const OTTERS_EXCHANGE_SLOTS: c_int = 4;

#[repr(C)]
pub struct otters_ctx { _private: [u8; 0] }
#[repr(C)]
pub struct otters_vecstore { _private: [u8; 0] }
#[repr(C)]
pub struct otters_metastore { _private: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct otters_scan_tuning {
    pub warps_per_cta: u32,
    pub slots_per_warp: u32,
    pub kc_floats: u32,
    pub ctas_per_sm: u32,
    pub unit_rows: u32,
    pub disable_fused_predicate: u32,
    pub batch_mode: u32,
    pub batch_cta_group: u32,
    pub scan_mode: u32,
    pub planners: u32,
    pub timing: u32,
    pub batch_passes: u32,
    pub separate_select: u32,
    pub lazy_prune: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct otters_last_work {
    pub kernel_launches: u64,
    pub rows_scored: u64,
    pub scan_bytes: u64,
    pub meta_bytes: u64,
    pub scan_ms: f32,
    pub prune_ms: f32,
    pub rowmask_ms: f32,
    pub select_ms: f32,
    pub batch_used: u32,
    pub batch_fallback: u32,
    pub batch_candidates: u64,
    pub batch_max_err: f32,
    pub batch_delta: f32,
    pub h2d_bytes: u64,
    pub d2h_bytes: u64,
    pub batch_passes: u32,
    pub batch_attempts: u32,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct otters_vec_query {
    pub queries: *const f32,
    pub nq: u32,
    pub dim: u32,
    pub metric: i32,
    pub take_type: i32,
    pub k: u64,
    pub has_filter: i32,
    pub thr: f32,
    pub cmp: i32,
    pub row_mask_words: *const u64,
    pub row_mask_bits: u64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct otters_column {
    pub name: *const c_char,
    pub dtype: i32,
    pub values: *const c_void,
    pub null_words: *const u64,
    pub str_offsets: *const u64,
    pub str_bytes: *const u8,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct otters_build_params {
    pub n_rows: u64,
    pub dim: u32,
    pub chunk_size: u64,
    pub bloom_mode: i32,
    pub bloom_fpr: f64,
    pub bloom_bits: u64,
    pub vectors_kind: i32,
    pub vectors: *const f32,
    pub synthetic_seed: u64,
    pub synthetic_first_row: u64,
    pub synthetic_map: *const c_void,
    pub columns: *const otters_column,
    pub n_columns: u32,
    pub vector_format: i32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct otters_build_stats {
    pub n_rows: u64,
    pub dim: u64,
    pub n_chunks: u64,
    pub vectors_ingest_s: f64,
    pub zonemap_build_s: f64,
    pub build_total_s: f64,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct otters_query_stats {
    pub total_chunks: u64,
    pub pruned_chunks: u64,
    pub evaluated_chunks: u64,
    pub vectors_compared: u64,
    pub prune_s: f64,
    pub score_s: f64,
    pub merge_s: f64,
    pub total_s: f64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct otters_leaf {
    pub col: u32,
    pub op: i32,
    pub kind: i32,
    pub i: i64,
    pub f: f64,
    pub s: *const u8,
    pub slen: u64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct otters_filter {
    pub n_clauses: u32,
    pub clause_offsets: *const u32,
    pub leaves: *const otters_leaf,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct otters_topk_record {
    pub row: u64,
    pub score: f32,
    pub qid: u32,
}

#[repr(C)]
#[derive(Clone, Copy, Default)]
pub struct otters_shard_map {
    pub row_base: u64,
    pub world: u32,
    pub rank: u32,
    pub block_rows: u64,
}

#[repr(C)]
#[derive(Clone, Copy)]
pub struct otters_peer_exchange {
    pub world: u32,
    pub rank: u32,
    pub k_max: u64,
    pub peer_records: *const *mut c_void,
    pub peer_flags: *const *mut u32,
}

extern "C" {
    pub fn otters_ctx_create(device: c_int, cuda_stream: *mut c_void, out: *mut *mut otters_ctx) -> c_int;
    pub fn otters_ctx_destroy(ctx: *mut otters_ctx) -> c_int;
    pub fn otters_ctx_synchronize(ctx: *mut otters_ctx) -> c_int;
    pub fn otters_ctx_join(ctx: *mut otters_ctx) -> c_int;
    pub fn otters_last_error() -> *const c_char;
    pub fn otters_version() -> *const c_char;
    pub fn otters_ctx_set_tuning(ctx: *mut otters_ctx, t: *const otters_scan_tuning) -> c_int;
    pub fn otters_ctx_last_work(ctx: *mut otters_ctx, out: *mut otters_last_work) -> c_int;
    pub fn otters_vecstore_create(ctx: *mut otters_ctx, dim: u32, out: *mut *mut otters_vecstore) -> c_int;
    pub fn otters_vecstore_create_fmt(ctx: *mut otters_ctx, dim: u32, vector_format: i32, out: *mut *mut otters_vecstore) -> c_int;
    pub fn otters_vecstore_format(vs: *const otters_vecstore) -> i32;
    pub fn otters_vecstore_destroy(vs: *mut otters_vecstore) -> c_int;
    pub fn otters_vecstore_reserve(vs: *mut otters_vecstore, n_rows: u64) -> c_int;
    pub fn otters_vecstore_add(vs: *mut otters_vecstore, rows: *const f32, n: u64) -> c_int;
    pub fn otters_vecstore_add_device(vs: *mut otters_vecstore, d_rows: *const f32, n: u64) -> c_int;
    pub fn otters_vecstore_add_synthetic(vs: *mut otters_vecstore, first_row: u64, n: u64, seed: u64) -> c_int;
    pub fn otters_vecstore_set_rows(vs: *mut otters_vecstore, rows: *const u64, data: *const f32, n: u64) -> c_int;
    pub fn otters_vecstore_len(vs: *const otters_vecstore) -> u64;
    pub fn otters_vecstore_dim(vs: *const otters_vecstore) -> u32;
    pub fn otters_vecstore_inv_norms(vs: *const otters_vecstore, first: u64, n: u64, out: *mut f32) -> c_int;
    pub fn otters_vecstore_query(vs: *mut otters_vecstore, q: *const otters_vec_query, out_idx: *mut u64, out_score: *mut f32, out_qid: *mut u32, cap: u64, out_len: *mut u64) -> c_int;
    pub fn otters_metastore_build(ctx: *mut otters_ctx, p: *const otters_build_params, out: *mut *mut otters_metastore, stats: *mut otters_build_stats) -> c_int;
    pub fn otters_metastore_destroy(ms: *mut otters_metastore) -> c_int;
    pub fn otters_metastore_set_rows(ms: *mut otters_metastore, rows: *const u64, data: *const f32, n: u64) -> c_int;
    pub fn otters_metastore_n_chunks(ms: *const otters_metastore) -> u64;
    pub fn otters_metastore_chunk_size(ms: *const otters_metastore) -> u64;
    pub fn otters_metastore_len(ms: *const otters_metastore) -> u64;
    pub fn otters_metastore_query(ms: *mut otters_metastore, q: *const otters_vec_query, filter: *const otters_filter, out_idx: *mut u64, out_score: *mut f32, out_qid: *mut u32, cap: u64, out_len: *mut u64, stats: *mut otters_query_stats) -> c_int;
    pub fn otters_vecstore_query_batch(vs: *mut otters_vecstore, q: *const otters_vec_query, out_idx: *mut u64, out_score: *mut f32, out_len: *mut u64) -> c_int;
    pub fn otters_metastore_query_batch(ms: *mut otters_metastore, q: *const otters_vec_query, filter: *const otters_filter, out_idx: *mut u64, out_score: *mut f32, out_len: *mut u64, stats: *mut otters_query_stats) -> c_int;
    pub fn otters_metastore_gather(ms: *mut otters_metastore, col: u32, rows: *const u64, n: u64, out_values: *mut c_void, out_nulls: *mut u8) -> c_int;
    pub fn otters_metastore_dict_entry(ms: *const otters_metastore, col: u32, code: u32, bytes: *mut *const u8, len: *mut u64) -> c_int;
    pub fn otters_metastore_last_stats(ms: *const otters_metastore, out: *mut otters_query_stats) -> c_int;
    pub fn otters_metastore_chunk_mask(ms: *mut otters_metastore, filter: *const otters_filter, keep: *mut u8) -> c_int;
    pub fn otters_metastore_row_mask(ms: *mut otters_metastore, filter: *const otters_filter, keep: *mut u8) -> c_int;
    pub fn otters_metastore_zonemap_i64(ms: *const otters_metastore, col: u32, mn: *mut i64, mx: *mut i64, non_null: *mut u64) -> c_int;
    pub fn otters_metastore_zonemap_f64(ms: *const otters_metastore, col: u32, mn: *mut f64, mx: *mut f64, non_null: *mut u64) -> c_int;
    pub fn otters_metastore_inv_norms(ms: *const otters_metastore, first: u64, n: u64, out: *mut f32) -> c_int;
    pub fn otters_query_local_device(vs: *mut otters_vecstore, ms: *mut otters_metastore, q: *const otters_vec_query, filter: *const otters_filter, map: *const otters_shard_map, d_records: *mut c_void, stats: *mut otters_query_stats) -> c_int;
    pub fn otters_query_exchange(vs: *mut otters_vecstore, ms: *mut otters_metastore, q: *const otters_vec_query, filter: *const otters_filter, map: *const otters_shard_map, ex: *const otters_peer_exchange, seq: u64, out_idx: *mut u64, out_score: *mut f32, out_qid: *mut u32, cap: u64, out_len: *mut u64, stats: *mut otters_query_stats) -> c_int;
    pub fn otters_metastore_save(ms: *mut otters_metastore, path: *const c_char, user: *const c_void, user_bytes: u64) -> c_int;
    pub fn otters_metastore_load(ctx: *mut otters_ctx, path: *const c_char, out: *mut *mut otters_metastore) -> c_int;
    pub fn otters_metastore_user_blob(ms: *const otters_metastore, bytes: *mut *const c_void, len: *mut u64) -> c_int;
    pub fn otters_metastore_n_columns(ms: *const otters_metastore) -> u32;
    pub fn otters_metastore_column_info(ms: *const otters_metastore, col: u32, name: *mut *const c_char, dtype: *mut i32) -> c_int;
    pub fn otters_metastore_dim(ms: *const otters_metastore) -> u32;
    pub fn otters_metastore_format(ms: *const otters_metastore) -> i32;
    pub fn otters_query_submit(vs: *mut otters_vecstore, ms: *mut otters_metastore, q: *const otters_vec_query, filter: *const otters_filter, map: *const otters_shard_map, ex: *const otters_peer_exchange, seq: u64, ticket: *mut u64) -> c_int;
    pub fn otters_query_wait(ctx: *mut otters_ctx, ticket: u64, out_idx: *mut u64, out_score: *mut f32, out_qid: *mut u32, cap: u64, out_len: *mut u64, stats: *mut otters_query_stats) -> c_int;
    pub fn otters_vecstore_add_synthetic_sharded(vs: *mut otters_vecstore, map: *const otters_shard_map, n_local: u64, seed: u64) -> c_int;
    pub fn otters_topk_merge_device(ctx: *mut otters_ctx, d_records: *const c_void, n_records: u64, k: u64, take_type: i32, out_idx: *mut u64, out_score: *mut f32, out_qid: *mut u32, cap: u64, out_len: *mut u64) -> c_int;
}
