//! Hand-written declarations of include/otters_b200.h (bindgen is not needed: the ABI is small and stable).
//! Every item mirrors the C header one to one; see the header for the reference lines each entry point replaces.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

pub const OTTERS_OK: c_int = 0;

#[repr(C)]
pub struct otters_ctx { _private: [u8; 0] }
#[repr(C)]
pub struct otters_vecstore { _private: [u8; 0] }
#[repr(C)]
pub struct otters_metastore { _private: [u8; 0] }

#[repr(C)]
#[derive(Clone, Copy)]
pub struct otters_vec_query {
    pub queries: *const f32,
    pub nq: u32,
    pub dim: u32,
    pub metric: i32,     // 0 Cosine, 1 Euclidean, 2 DotProduct  (src/vec.rs:11-16)
    pub take_type: i32,  // 0 Min, 1 Max                          (src/vec.rs:18-22)
    pub k: u64,
    pub has_filter: i32,
    pub thr: f32,
    pub cmp: i32,        // 0 Lt, 1 Gt, 2 Lte, 3 Gte, 4 Eq         (src/vec.rs:24-31)
    pub row_mask_words: *const u64,
    pub row_mask_bits: u64,
}

#[repr(C)]
pub struct otters_column {
    pub name: *const c_char,
    pub dtype: i32,      // 0 Int32, 1 Int64, 2 Float32, 3 Float64, 4 String, 5 DateTime (src/type_utils.rs:11-19)
    pub values: *const c_void,
    pub null_words: *const u64,
    pub str_offsets: *const u64,
    pub str_bytes: *const u8,
}

#[repr(C)]
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
}

#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct otters_build_stats {
    pub n_rows: u64,
    pub dim: u64,
    pub n_chunks: u64,
    pub vectors_ingest_s: f64,
    pub zonemap_build_s: f64,
    pub build_total_s: f64,
}

#[repr(C)]
#[derive(Default, Clone, Copy)]
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
pub struct otters_leaf {
    pub col: u32,
    pub op: i32,     // 0 Eq, 1 Neq, 2 Lt, 3 Lte, 4 Gt, 5 Gte (src/expr.rs:83-91)
    pub kind: i32,   // 0 I64, 1 F64, 2 Str
    pub i: i64,
    pub f: f64,
    pub s: *const u8,
    pub slen: u64,
}

#[repr(C)]
pub struct otters_filter {
    pub n_clauses: u32,
    pub clause_offsets: *const u32,
    pub leaves: *const otters_leaf,
}

extern "C" {
    pub fn otters_ctx_create(device: c_int, cuda_stream: *mut c_void, out: *mut *mut otters_ctx) -> c_int;
    pub fn otters_ctx_destroy(ctx: *mut otters_ctx) -> c_int;
    pub fn otters_last_error() -> *const c_char;

    pub fn otters_vecstore_create(ctx: *mut otters_ctx, dim: u32, out: *mut *mut otters_vecstore) -> c_int;
    pub fn otters_vecstore_destroy(vs: *mut otters_vecstore) -> c_int;
    pub fn otters_vecstore_add(vs: *mut otters_vecstore, rows: *const f32, n: u64) -> c_int;
    pub fn otters_vecstore_len(vs: *const otters_vecstore) -> u64;
    pub fn otters_vecstore_query(vs: *mut otters_vecstore, q: *const otters_vec_query, out_idx: *mut u64, out_score: *mut f32,
                                 out_qid: *mut u32, cap: u64, out_len: *mut u64) -> c_int;

    pub fn otters_metastore_build(ctx: *mut otters_ctx, p: *const otters_build_params, out: *mut *mut otters_metastore,
                                  stats: *mut otters_build_stats) -> c_int;
    pub fn otters_metastore_destroy(ms: *mut otters_metastore) -> c_int;
    pub fn otters_metastore_n_chunks(ms: *const otters_metastore) -> u64;
    pub fn otters_metastore_query(ms: *mut otters_metastore, q: *const otters_vec_query, filter: *const otters_filter,
                                  out_idx: *mut u64, out_score: *mut f32, out_qid: *mut u32, cap: u64, out_len: *mut u64,
                                  stats: *mut otters_query_stats) -> c_int;
    pub fn otters_metastore_last_stats(ms: *const otters_metastore, out: *mut otters_query_stats) -> c_int;
}
