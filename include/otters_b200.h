/*
 * otters_b200.h — C ABI of libotters_b200.so, the B200-native (sm_100a) implementation of the
 * otters exact-search hot path.
 *
 * The reference (AtharvBhat/otters, Rust) has no FFI seam of its own: the drop-in boundary is its
 * public Rust API, and these entry points are what a thin `otters-sys` crate binds directly beneath
 * `VecQueryPlan::collect` (src/vec.rs:206-311), `MetaStoreBuilder::build` (src/meta.rs:151-305) and
 * `MetaQueryPlan::collect` (src/meta.rs:632-829).  INTEGRATION.md shows that binding.  Each entry
 * point below cites the reference interface it replaces (paths relative to /root/reference).
 *
 * Conventions
 *  - every function returns an int status (OTTERS_OK == 0); on failure a thread-local message is
 *    available from otters_last_error().  Validation messages reproduce the reference's strings
 *    (src/vec.rs:170-203, :357-363).
 *  - plain pointers and sizes only; the caller owns all in/out buffers; the library copies inputs
 *    before returning; opaque handles are released with the matching *_destroy.
 *  - enum codes follow the reference's declaration order.
 *  - there is NO CPU fallback: every query runs on the CUDA device of the context.
 */
#ifndef OTTERS_B200_H
#define OTTERS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define OTTERS_API __attribute__((visibility("default")))
#else
#define OTTERS_API
#endif

#define OTTERS_OK 0
#define OTTERS_ERR_INVALID 1     /* validation error; message mirrors the reference's Err(String) */
#define OTTERS_ERR_CUDA 2        /* CUDA runtime failure */
#define OTTERS_ERR_NOMEM 3       /* host or device allocation failure */
#define OTTERS_ERR_UNSUPPORTED 4 /* configuration outside what the kernels support */

/* src/vec.rs:11-16 */
#define OTTERS_METRIC_COSINE 0
#define OTTERS_METRIC_EUCLIDEAN 1 /* SQUARED euclidean, no sqrt (src/vec_compute.rs:35-54) */
#define OTTERS_METRIC_DOT 2
/* src/vec.rs:18-22 */
#define OTTERS_TAKE_MIN 0
#define OTTERS_TAKE_MAX 1
/* src/vec.rs:24-31 */
#define OTTERS_CMP_LT 0
#define OTTERS_CMP_GT 1
#define OTTERS_CMP_LTE 2
#define OTTERS_CMP_GTE 3
#define OTTERS_CMP_EQ 4
/* src/expr.rs:83-91 */
#define OTTERS_OP_EQ 0
#define OTTERS_OP_NEQ 1
#define OTTERS_OP_LT 2
#define OTTERS_OP_LTE 3
#define OTTERS_OP_GT 4
#define OTTERS_OP_GTE 5
/* src/type_utils.rs:11-19 */
#define OTTERS_DTYPE_INT32 0
#define OTTERS_DTYPE_INT64 1
#define OTTERS_DTYPE_FLOAT32 2
#define OTTERS_DTYPE_FLOAT64 3
#define OTTERS_DTYPE_STRING 4
#define OTTERS_DTYPE_DATETIME 5 /* i64 epoch milliseconds UTC (src/col.rs:18) */
/* src/expr.rs:192-210 (NumericLiteral::{I64,F64} / ColumnFilter::String) */
#define OTTERS_LIT_I64 0
#define OTTERS_LIT_F64 1
#define OTTERS_LIT_STR 2

typedef struct otters_ctx otters_ctx;
typedef struct otters_vecstore otters_vecstore;
typedef struct otters_metastore otters_metastore;

/* ---------------------------------------------------------------------------------------------
 * context: one CUDA device + one stream.  `cuda_stream` may be NULL (the library creates its own
 * non-blocking stream) or an existing cudaStream_t on `device` that all work is enqueued on.
 * ------------------------------------------------------------------------------------------- */
OTTERS_API int otters_ctx_create(int device, void *cuda_stream, otters_ctx **out);
OTTERS_API int otters_ctx_destroy(otters_ctx *ctx);
OTTERS_API int otters_ctx_synchronize(otters_ctx *ctx); /* waits for the context's stream and for every lane (otters_query_submit) */
/* Orders the context's own stream after everything enqueued on its lanes so far (no host wait): work or events the caller
 * enqueues on that stream afterwards run once all submitted queries have finished. */
OTTERS_API int otters_ctx_join(otters_ctx *ctx);
OTTERS_API const char *otters_last_error(void);
OTTERS_API const char *otters_version(void);

/* Tuning knobs of the scan kernel (0 = automatic).  For profiling sweeps; results never change. */
typedef struct {
    uint32_t warps_per_cta;
    uint32_t slots_per_warp;
    uint32_t kc_floats;    /* columns staged per slot (multiple of 8) */
    uint32_t ctas_per_sm;
    uint32_t unit_rows;    /* 32, 64 or 128 rows per dynamically scheduled work unit */
    uint32_t disable_fused_predicate; /* row predicate: 0 = automatic and 1 = its own kernel (K0b: surviving-row bitmask, then the
                                         scan reads mask words), 2 = evaluated per work unit inside the scan kernel */
    uint32_t batch_mode;   /* query batches: 0 = automatic, 1 = always the tensor-core kernel (when k <= 1024), 2 = never */
    uint32_t batch_cta_group; /* tensor-core kernel: 0 = automatic (single CTAs), 1 = single CTAs, 2 = CTA pairs (tcgen05 cta_group::2) */
    uint32_t scan_mode;    /* K1 front-end: 0 = automatic, 1 = autonomous warps (scan_kernel.cuh), 2 = planner + worker warps (scan_planner.cu) */
    uint32_t planners;     /* planner front-end: planner warps per CTA (0 = automatic: 2 for filtered stores up to 256-d, else 1) */
    uint32_t timing;       /* per-phase CUDA events (otters_last_work *_ms, otters_query_stats durations): 0 = automatic (only
                              for blocking MetaStore queries that ask for stats), 1 = always, 2 = never */
    uint32_t batch_passes; /* tensor-core kernel, arithmetic rung of the SELECTION: 0 = automatic (bf16 operands from a bf16 shadow
                              of the rows when device memory allows, then single-pass tf32, then 3xTF32, each tried when the
                              certificate of the one before fails), 1 = single-pass tf32 only, 2 = bf16 only, 3 = 3xTF32 only.
                              Results never change: every returned score is re-computed in the reference's arithmetic from the
                              fp32 rows and the selection is certified or redone */
    uint32_t separate_select; /* 1: run the final selection (K3) as its own kernel instead of in the last CTA of the scan kernel */
    uint32_t lazy_prune;      /* 1: evaluate the zonemap / Bloom chunk rules lazily inside the scan kernel (per work unit) instead
                                 of running K0 first: one launch per query, but measured slower — every unit pays a dependent
                                 L2 round trip and 8 units re-evaluate each 1024-row chunk (profiles/r2_prune_ab.txt) */
} otters_scan_tuning;
OTTERS_API int otters_ctx_set_tuning(otters_ctx *ctx, const otters_scan_tuning *t);

/* Counters of the work enqueued by the LAST query on this context (for bench.py's gpu_launches and
 * the roofline numerator). */
typedef struct {
    uint64_t kernel_launches;   /* kernels of this library launched by the last query */
    uint64_t rows_scored;       /* rows whose vector was streamed from HBM (x queries) */
    uint64_t scan_bytes;        /* algorithmic bytes of the scan kernel: rows_scored*(dim*4 [+4 cosine]) */
    uint64_t meta_bytes;        /* algorithmic bytes of prune + row-mask kernels */
    float scan_ms;              /* device time of the scan kernel(s) of the last query (CUDA events) */
    float prune_ms, rowmask_ms, select_ms;
    /* query batches served by the tcgen05 kernel (K2): scan_ms is then the device time of that kernel */
    uint32_t batch_used;        /* 1: the tensor-core kernel produced the result */
    uint32_t batch_fallback;    /* 1: its candidate set could not be verified and the batch was re-run query by query */
    uint64_t batch_candidates;  /* (row, query) pairs re-scored in the reference's exact arithmetic */
    float batch_max_err;        /* largest |tensor-core score - exact score| over the re-scored pairs */
    float batch_delta;          /* the error bound the selection assumed (must exceed batch_max_err) */
    uint64_t h2d_bytes;         /* host-to-device bytes of the last query (input image: control block + filter + queries; row mask) */
    uint64_t d2h_bytes;         /* device-to-host bytes of the last query (result header + candidates / stats) */
    uint32_t batch_passes;      /* rung of the accepted tensor-core run: 2 = bf16, 1 = single-pass tf32, 3 = 3xTF32 (0: K2 not used) */
    uint32_t batch_attempts;    /* tensor-core runs made for the last query (a failed certificate costs one) */
} otters_last_work;
OTTERS_API int otters_ctx_last_work(otters_ctx *ctx, otters_last_work *out);

/* ---------------------------------------------------------------------------------------------
 * VecStore — src/vec.rs:338-411.  Rows live in HBM (row-major f32, 16-byte aligned rows) with the
 * per-row inverse L2 norm precomputed exactly as add_vector does (serial f32 sum of squares,
 * 1/sqrt, 0.0 for a zero row; src/vec.rs:357-371).
 * ------------------------------------------------------------------------------------------- */
OTTERS_API int otters_vecstore_create(otters_ctx *ctx, uint32_t dim, otters_vecstore **out); /* VecStore::new */
/* Reduced-precision rows (the reference's roadmap item "Quantization for vectors", README.md:208).
 * OTTERS_VECTORS_FMT_BF16: every element is rounded to bf16 (nearest even) when it is added and the store keeps 2 bytes per
 * element — half the bytes of a scan that runs at the HBM limit.  The parity contract is the reference's arithmetic applied
 * to the ROUNDED rows: scores, inverse norms and results are bit-identical to the CPU path run on f32(bf16(x)) (queries
 * stay fp32; widening a bf16 value to fp32 is exact).  Rows are still passed in as fp32.  Query batches on a bf16 store run on the
 * tensor cores with the stored rows as the bf16 operand (one rung; exact re-scoring from the same rows). */
#define OTTERS_VECTORS_FMT_F32 0
#define OTTERS_VECTORS_FMT_BF16 1
OTTERS_API int otters_vecstore_create_fmt(otters_ctx *ctx, uint32_t dim, int32_t vector_format, otters_vecstore **out);
OTTERS_API int32_t otters_vecstore_format(const otters_vecstore *vs);
OTTERS_API int otters_vecstore_destroy(otters_vecstore *vs);
OTTERS_API int otters_vecstore_reserve(otters_vecstore *vs, uint64_t n_rows);
/* VecStore::add_vectors with contiguous row-major host rows (n * dim floats). */
OTTERS_API int otters_vecstore_add(otters_vecstore *vs, const float *rows, uint64_t n);
/* Same, rows already in device memory on the context's device. */
OTTERS_API int otters_vecstore_add_device(otters_vecstore *vs, const float *d_rows, uint64_t n);
/* Appends n rows of the counter-based synthetic generator x = (splitmix64(seed ^ (row*dim+col)) >> 40) * 2^-23 - 1
 * (row = absolute row id starting at first_row), generated on the device.  Bench/test utility. */
OTTERS_API int otters_vecstore_add_synthetic(otters_vecstore *vs, uint64_t first_row, uint64_t n, uint64_t seed);
/* Overwrites n stored rows (row ids local to this store, any order; data = n * dim floats in host memory) and recomputes
 * their inverse norms.  Bench/test utility: how near-duplicates of a query are planted into a synthetic store (SURVEY.md
 * §8d, config C5).  Waits for every query in flight first. */
OTTERS_API int otters_vecstore_set_rows(otters_vecstore *vs, const uint64_t *rows, const float *data, uint64_t n);
OTTERS_API uint64_t otters_vecstore_len(const otters_vecstore *vs); /* VecStore::len */
OTTERS_API uint32_t otters_vecstore_dim(const otters_vecstore *vs);
/* debug/parity: copies inverse norms [first, first+n) to host */
OTTERS_API int otters_vecstore_inv_norms(const otters_vecstore *vs, uint64_t first, uint64_t n, float *out);

/* The inputs of VecQueryPlan::collect (src/vec.rs:55-66, :206-219).  The host-side plan builder
 * resolves the defaults before the call: k = n_vecs when no take*() was given (:213), take_type =
 * Max when none (:214) / inferred from the metric by take() (:92-116). */
typedef struct {
    const float *queries;            /* nq * dim, row-major, host memory */
    uint32_t nq;
    uint32_t dim;                    /* must equal the store's dim, else the reference's error string */
    int32_t metric;                  /* OTTERS_METRIC_* */
    int32_t take_type;               /* OTTERS_TAKE_* */
    uint64_t k;                      /* take count */
    int32_t has_filter;              /* .filter(thr, cmp) / .vec_filter(thr, cmp) present */
    float thr;
    int32_t cmp;                     /* OTTERS_CMP_* */
    const uint64_t *row_mask_words;  /* .with_row_mask(BitVec): Lsb0 words, bit = 1 keep; NULL = none */
    uint64_t row_mask_bits;          /* mask length in bits; rows >= this are kept (src/vec.rs:234,297) */
} otters_vec_query;

/* VecQueryPlan::collect (src/vec.rs:206-311).  Writes min(*out_len, cap) results best-first;
 * *out_len is the full result count.  One merged list for a batch (src/vec.rs:217-219); out_qid
 * (nullable, extension) receives the query index of each result.  Ties: better score, then lower
 * row, then lower query index. */
OTTERS_API int otters_vecstore_query(otters_vecstore *vs, const otters_vec_query *q, uint64_t *out_idx, float *out_score,
                          uint32_t *out_qid, uint64_t cap, uint64_t *out_len);

/* ---------------------------------------------------------------------------------------------
 * MetaStore — src/meta.rs:48-60, :151-305.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    const char *name;               /* UTF-8, NUL-terminated */
    int32_t dtype;                  /* OTTERS_DTYPE_* */
    const void *values;             /* typed array of n_rows entries (src/col.rs:11-19); NULL for String */
    const uint64_t *null_words;     /* Lsb0 words, bit = 1 NULL (src/col.rs:26); NULL = no nulls */
    const uint64_t *str_offsets;    /* String: n_rows + 1 byte offsets into str_bytes */
    const uint8_t *str_bytes;       /* String: concatenated bytes */
} otters_column;

#define OTTERS_VECTORS_HOST 0
#define OTTERS_VECTORS_DEVICE 1
#define OTTERS_VECTORS_SYNTHETIC 2

typedef struct {
    uint64_t n_rows;
    uint32_t dim;
    uint64_t chunk_size;            /* with_chunk_size; 0 -> 1 (src/meta.rs:86-89); default 1024 */
    int32_t bloom_mode;             /* 0 = Fpr(bloom_fpr) (src/meta.rs:92-101), 1 = Bits(bloom_bits) (:106-110) */
    double bloom_fpr;
    uint64_t bloom_bits;
    int32_t vectors_kind;           /* OTTERS_VECTORS_* */
    const float *vectors;           /* host or device pointer, n_rows * dim floats row-major */
    uint64_t synthetic_seed;        /* OTTERS_VECTORS_SYNTHETIC */
    uint64_t synthetic_first_row;
    const void *synthetic_map;      /* optional otters_shard_map*: global row ids of the generated rows */
    const otters_column *columns;
    uint32_t n_columns;
    int32_t vector_format;          /* OTTERS_VECTORS_FMT_* (0 = fp32 rows; see otters_vecstore_create_fmt) */
} otters_build_params;

/* src/meta.rs:844-852 (durations in seconds) */
typedef struct {
    uint64_t n_rows;
    uint64_t dim;
    uint64_t n_chunks;
    double vectors_ingest_s;
    double zonemap_build_s;
    double build_total_s;
} otters_build_stats;

/* src/meta.rs:832-842 (durations in seconds) */
typedef struct {
    uint64_t total_chunks;
    uint64_t pruned_chunks;
    uint64_t evaluated_chunks;
    uint64_t vectors_compared;
    double prune_s;
    double score_s;
    double merge_s;
    double total_s;
} otters_query_stats;

/* MetaStoreBuilder::build (src/meta.rs:151-305): chunking, per-chunk zonemaps (min/max/non-null,
 * src/meta_compute.rs:32-132), per-chunk string Bloom filters, dictionary codes for strings. */
OTTERS_API int otters_metastore_build(otters_ctx *ctx, const otters_build_params *p, otters_metastore **out,
                           otters_build_stats *stats /* nullable */);
OTTERS_API int otters_metastore_destroy(otters_metastore *ms);
OTTERS_API int otters_metastore_set_rows(otters_metastore *ms, const uint64_t *rows, const float *data, uint64_t n); /* as otters_vecstore_set_rows */
OTTERS_API uint64_t otters_metastore_n_chunks(const otters_metastore *ms); /* MetaStore::n_chunks */
OTTERS_API uint64_t otters_metastore_chunk_size(const otters_metastore *ms);
OTTERS_API uint64_t otters_metastore_len(const otters_metastore *ms);

/* One leaf of the compiled filter (ColumnFilter, src/expr.rs:192-210). */
typedef struct {
    uint32_t col;       /* index into otters_build_params.columns */
    int32_t op;         /* OTTERS_OP_* */
    int32_t kind;       /* OTTERS_LIT_* */
    int64_t i;          /* NumericLiteral::I64 (DateTime literals: epoch millis) */
    double f;           /* NumericLiteral::F64 */
    const uint8_t *s;   /* string literal bytes */
    uint64_t slen;
} otters_leaf;

/* CompiledFilter (src/expr.rs:212-226): AND over clauses of OR over leaves. */
typedef struct {
    uint32_t n_clauses;
    const uint32_t *clause_offsets; /* n_clauses + 1 offsets into leaves */
    const otters_leaf *leaves;
} otters_filter;

/* MetaQueryPlan::collect (src/meta.rs:632-721): zonemap/Bloom chunk pruning (:407-544), per-row CNF
 * predicate (src/meta_compute.rs:194-318), scoring, top-k, stats.  `q->row_mask_words` must be NULL.
 * `filter` may be NULL (no meta_filter).  Result-column gathering (src/meta.rs:723-828) stays on
 * the host side of the boundary or, on the device, otters_metastore_gather. */
OTTERS_API int otters_metastore_query(otters_metastore *ms, const otters_vec_query *q, const otters_filter *filter,
                           uint64_t *out_idx, float *out_score, uint32_t *out_qid, uint64_t cap, uint64_t *out_len,
                           otters_query_stats *stats /* nullable */);
/* Per-query top-k for a batch — an extension: the reference merges a batch into ONE list (src/vec.rs:217-219), which
 * otters_vecstore_query / otters_metastore_query reproduce.  Here every query i of q (nq queries) gets its own list:
 * out_idx / out_score hold nq rows of q->k entries, out_len[i] results are valid in row i, and row i is exactly what the
 * single-query call returns for query i (bit-identical).  The queries are pipelined over the context's two lanes.  Stats
 * follow the reference's batch convention (chunks counted once, vectors_compared = sum over chunks of len * nq). */
OTTERS_API int otters_vecstore_query_batch(otters_vecstore *vs, const otters_vec_query *q, uint64_t *out_idx, float *out_score,
                                uint64_t *out_len /* [nq] */);
OTTERS_API int otters_metastore_query_batch(otters_metastore *ms, const otters_vec_query *q, const otters_filter *filter,
                                 uint64_t *out_idx, float *out_score, uint64_t *out_len /* [nq] */,
                                 otters_query_stats *stats /* nullable */);
/* MetaQueryResults.data (src/meta.rs:723-821): gathers column `col` at the n result rows on the device.  out_values is
 * typed like the column (int32 / int64 / float / double / DateTime millis); for String columns it receives uint32
 * dictionary codes, which otters_metastore_dict_entry maps back to bytes (valid while the store lives).  out_nulls[i] = 1
 * marks a NULL row (its value slot holds the column's sentinel / an undefined code). */
OTTERS_API int otters_metastore_gather(otters_metastore *ms, uint32_t col, const uint64_t *rows, uint64_t n, void *out_values,
                            uint8_t *out_nulls);
OTTERS_API int otters_metastore_dict_entry(const otters_metastore *ms, uint32_t col, uint32_t code, const uint8_t **bytes, uint64_t *len);
/* MetaStore::last_query_stats (src/meta.rs:395-397); returns OTTERS_ERR_INVALID if no query ran yet */
OTTERS_API int otters_metastore_last_stats(const otters_metastore *ms, otters_query_stats *out);

/* parity/debug exports: results of the prune and row-mask kernels for `filter` (one byte per chunk /
 * per row, 1 = keep) and the zonemap tables built on the device side of the boundary */
OTTERS_API int otters_metastore_chunk_mask(otters_metastore *ms, const otters_filter *filter, uint8_t *keep);
OTTERS_API int otters_metastore_row_mask(otters_metastore *ms, const otters_filter *filter, uint8_t *keep);
OTTERS_API int otters_metastore_zonemap_i64(const otters_metastore *ms, uint32_t col, int64_t *mn, int64_t *mx, uint64_t *non_null);
OTTERS_API int otters_metastore_zonemap_f64(const otters_metastore *ms, uint32_t col, double *mn, double *mx, uint64_t *non_null);
OTTERS_API int otters_metastore_inv_norms(const otters_metastore *ms, uint64_t first, uint64_t n, float *out);

/* ---------------------------------------------------------------------------------------------
 * Row-sharded multi-GPU search: every rank holds a contiguous row range, computes a
 * local top-k that stays in device memory as fixed 16-byte records, the ranks all-gather the
 * records (NCCL), and every rank runs the same final merge.
 * ------------------------------------------------------------------------------------------- */
typedef struct {
    uint64_t row;   /* global row id (see otters_shard_map); UINT64_MAX = empty slot */
    float score;
    uint32_t qid;
} otters_topk_record;

/* How local rows of a shard map to global row ids.  world <= 1 or block_rows == 0: contiguous,
 * global = row_base + local.  Otherwise block-cyclic: the store's blocks of block_rows rows are dealt
 * round-robin to the ranks (block b of rank r is global block b * world + r), which keeps every shard
 * balanced under range filters:  global = row_base + ((local / block_rows) * world + rank) * block_rows
 * + local % block_rows. */
typedef struct {
    uint64_t row_base;
    uint32_t world;
    uint32_t rank;
    uint64_t block_rows;
} otters_shard_map;

/* Local query whose result stays on the device: writes exactly k records to d_records (device
 * memory, padded with empty slots).  `filter`/`ms` semantics as above; pass vs for a VecStore or ms
 * for a MetaStore (exactly one non-NULL).  `map` (nullable = identity) turns local rows into global
 * row ids.  Stats counters are local to the shard; with stats == NULL the call does not synchronise. */
OTTERS_API int otters_query_local_device(otters_vecstore *vs, otters_metastore *ms, const otters_vec_query *q,
                              const otters_filter *filter, const otters_shard_map *map, void *d_records,
                              otters_query_stats *stats /* nullable */);
/* Fused exchange over peer memory (NVLink / NVSwitch).  Every rank owns a record area of OTTERS_EXCHANGE_SLOTS * world *
 * k_max otters_topk_record and a flag area of OTTERS_EXCHANGE_SLOTS * world uint32 (zero-initialised), both mapped into every
 * peer process (CUDA IPC / symmetric memory; otters_b200/sharded.py uses torch.distributed._symmetric_memory for the mapping).
 * peer_records[p] / peer_flags[p] are rank p's areas as mapped into THIS process (p == rank: the local areas).  The areas
 * are OTTERS_EXCHANGE_SLOTS deep on the query sequence number: twice the two queries a context keeps in flight, so a rank
 * that runs ahead never overwrites records a slower peer is still merging. */
#define OTTERS_EXCHANGE_SLOTS 4
typedef struct {
    uint32_t world;              /* 2..8 */
    uint32_t rank;
    uint64_t k_max;              /* records per (query parity, rank) slot; take counts must not exceed it */
    void *const *peer_records;   /* [world] */
    uint32_t *const *peer_flags; /* [world] */
} otters_peer_exchange;

/* Row-sharded query with the exchange fused into the selection: ONE kernel per query prunes, scans, selects the local
 * top-k (in its last CTA), stores its k records straight into every peer's record area, publishes them with a release
 * flag, waits for the other ranks' flags and merges world * k records — no NCCL call and no extra launch on the query
 * path.  `seq` numbers the queries of this exchange (1, 2, 3, ... identical on every rank; it selects the area slot).  Every rank must call it for every query.  With out_idx = out_score = out_qid = NULL and
 * cap = 0 the call only enqueues (no copy, no sync).  Returns OTTERS_ERR_UNSUPPORTED for plans the fused selection
 * does not serve (take counts above 1024 or above k_max, batches routed to the tensor-core kernel); callers then
 * use otters_query_local_device + an all-gather + otters_topk_merge_device.  Stats are local to the shard. */
OTTERS_API int otters_query_exchange(otters_vecstore *vs, otters_metastore *ms, const otters_vec_query *q,
                          const otters_filter *filter, const otters_shard_map *map, const otters_peer_exchange *ex,
                          uint64_t seq, uint64_t *out_idx, float *out_score, uint32_t *out_qid, uint64_t cap,
                          uint64_t *out_len, otters_query_stats *stats /* nullable */);

/* Persistence — the reference's roadmap item "Persistence (save/load MetaStore to/from disk)" (README.md:206).  The file is
 * the store's HBM image (rows as stored — fp32 or bf16 —, inverse norms, column values, null words, zonemap tables, Bloom
 * filters, dictionaries) plus `user_bytes` opaque bytes of the caller (the host mirrors keep their row-order permutation
 * there); otters_metastore_load allocates and copies, nothing is recomputed: a loaded store returns the same bytes and the
 * same statistics.  Little-endian, versioned ("OTTERSB2", 1); a file whose size does not match its header is rejected
 * before anything is allocated.  otters_metastore_user_blob / _n_columns / _column_info / _dim / _format describe a loaded
 * store to the host side (pointers valid while the store lives). */
OTTERS_API int otters_metastore_save(otters_metastore *ms, const char *path, const void *user /* nullable */, uint64_t user_bytes);
OTTERS_API int otters_metastore_load(otters_ctx *ctx, const char *path, otters_metastore **out);
OTTERS_API int otters_metastore_user_blob(const otters_metastore *ms, const void **bytes, uint64_t *len);
OTTERS_API uint32_t otters_metastore_n_columns(const otters_metastore *ms);
OTTERS_API int otters_metastore_column_info(const otters_metastore *ms, uint32_t col, const char **name, int32_t *dtype);
OTTERS_API uint32_t otters_metastore_dim(const otters_metastore *ms);
OTTERS_API int32_t otters_metastore_format(const otters_metastore *ms);

/* Non-blocking queries.  otters_query_submit enqueues ONE query — chunk pruning, row predicate, scan, selection, and the peer
 * exchange when `ex` is given — on one of the two lanes of the store's context (a lane = its own CUDA stream + scratch +
 * pinned input / result buffers) and returns at once with a ticket; otters_query_wait blocks until that query has finished
 * and copies its result and statistics out (durations are not split per phase on this path).  Two tickets may be
 * outstanding per context: submit query i+1 before waiting for query i, and its input copy, its launch and the head of its
 * scan overlap the tail, the selection and the exchange of query i — the per-call host cost no longer sits between two
 * scans.  Inputs are copied before submit returns.  A lane holds one query: submitting twice more without waiting abandons
 * the older result.  vs / ms / filter / map as in otters_query_local_device; ex == NULL: single-GPU query (seq ignored),
 * else as in otters_query_exchange (every rank submits every query with the same seq).  otters_ctx_synchronize waits for
 * all lanes.  The store must not be modified while tickets are outstanding. */
OTTERS_API int otters_query_submit(otters_vecstore *vs, otters_metastore *ms, const otters_vec_query *q, const otters_filter *filter,
                        const otters_shard_map *map, const otters_peer_exchange *ex, uint64_t seq, uint64_t *ticket);
OTTERS_API int otters_query_wait(otters_ctx *ctx, uint64_t ticket, uint64_t *out_idx, float *out_score, uint32_t *out_qid, uint64_t cap,
                      uint64_t *out_len, otters_query_stats *stats /* nullable */);

/* Appends n_local rows of the synthetic generator whose global row ids follow `map` (bench/test utility). */
OTTERS_API int otters_vecstore_add_synthetic_sharded(otters_vecstore *vs, const otters_shard_map *map, uint64_t n_local, uint64_t seed);
/* Merges n_records device records (e.g. world_size * k after the all-gather) into the global best k.  Any record
 * order is accepted; when the input is a concatenation of blocks of k records ordered best-first — exactly what the
 * ranks' otters_query_local_device calls leave — the merge is a binary search per block instead of an all-pairs count.
 * With out_idx = out_score = out_qid = NULL and cap = 0 the call only enqueues the merge (no copy, no sync). */
OTTERS_API int otters_topk_merge_device(otters_ctx *ctx, const void *d_records, uint64_t n_records, uint64_t k, int32_t take_type,
                             uint64_t *out_idx, float *out_score, uint32_t *out_qid, uint64_t cap, uint64_t *out_len);

#ifdef __cplusplus
}
#endif
#endif /* OTTERS_B200_H */
