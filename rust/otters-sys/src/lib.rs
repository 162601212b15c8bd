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

/// otters_shard_map: how local rows of a shard map to global row ids (contiguous or block-cyclic).
#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct otters_shard_map {
    pub row_base: u64,
    pub world: u32,
    pub rank: u32,
    pub block_rows: u64,
}

/// otters_topk_record: one entry of a shard's local top-k as it travels between GPUs.
#[repr(C)]
#[derive(Clone, Copy)]
pub struct otters_topk_record {
    pub row: u64,   // global row id; u64::MAX = empty slot
    pub score: f32,
    pub qid: u32,
}

/// otters_peer_exchange: record / flag areas of every rank as mapped into this process (CUDA IPC / VMM).
#[repr(C)]
pub struct otters_peer_exchange {
    pub world: u32,
    pub rank: u32,
    pub k_max: u64,
    pub peer_records: *const *mut c_void,
    pub peer_flags: *const *mut u32,
}

/// otters_scan_tuning: profiling knobs (0 = automatic everywhere); results never depend on them.
#[repr(C)]
#[derive(Default, Clone, Copy)]
pub struct otters_scan_tuning {
    pub warps_per_cta: u32,
    pub slots_per_warp: u32,
    pub kc_floats: u32,
    pub ctas_per_sm: u32,
    pub unit_rows: u32,
    pub disable_fused_predicate: u32,
    pub batch_mode: u32,       // 0 auto, 1 always the tcgen05 kernel for batches, 2 never
    pub batch_cta_group: u32,  // 0 auto (CTA pairs), 1 single CTAs, 2 pairs
    pub scan_mode: u32,        // K1 front-end: 0 auto, 1 autonomous warps, 2 planner + worker warps
    pub planners: u32,         // planner warps per CTA (0 auto)
    pub timing: u32,           // 0 auto, 1 always record phase events, 2 never
    pub batch_passes: u32,     // 0 auto (single-pass tf32 selection, then 3xTF32), 1 single pass only, 3 3xTF32 only
}

extern "C" {
    pub fn otters_ctx_set_tuning(ctx: *mut otters_ctx, t: *const otters_scan_tuning) -> c_int;
    /// Row-sharded search, NCCL flavour: the local top-k stays in HBM as k records ...
    pub fn otters_query_local_device(vs: *mut otters_vecstore, ms: *mut otters_metastore, q: *const otters_vec_query,
                                     filter: *const otters_filter, map: *const otters_shard_map, d_records: *mut c_void,
                                     stats: *mut otters_query_stats) -> c_int;
    /// ... and after the all-gather every rank merges world * k records.
    pub fn otters_topk_merge_device(ctx: *mut otters_ctx, d_records: *const c_void, n_records: u64, k: u64, take_type: i32,
                                    out_idx: *mut u64, out_score: *mut f32, out_qid: *mut u32, cap: u64, out_len: *mut u64) -> c_int;
    /// Row-sharded search with the exchange fused into the selection kernel (peer stores over NVLink, no NCCL call).
    pub fn otters_query_exchange(vs: *mut otters_vecstore, ms: *mut otters_metastore, q: *const otters_vec_query,
                                 filter: *const otters_filter, map: *const otters_shard_map, ex: *const otters_peer_exchange,
                                 seq: u64, out_idx: *mut u64, out_score: *mut f32, out_qid: *mut u32, cap: u64,
                                 out_len: *mut u64, stats: *mut otters_query_stats) -> c_int;

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
