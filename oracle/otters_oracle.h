/*
 * otters_oracle.h — CPU restatement of the otters exact-search hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / CPU baseline.
 *
 * Parity status: the reference crate (Rust) cannot be compiled in this image
 * (no cargo/rustc), so the oracle is pinned against the reference's own
 * known-answer tests (tests/golden/reference_kats.json, transcribed from
 * /root/reference/tests/ *.rs files) — see tests/test_oracle_kats.py.  Two pieces of
 * third-party arithmetic are NOT on disk and are restated from memory of the
 * published crates:
 *   - wide 0.7.33 f32x8::reduce_add — default (non-AVX) build order
 *     (((l0+l1)+l2)+l3) + (((l4+l5)+l6)+l7); the AVX order is selectable with
 *     oracle_set_reduce_order(1).  Differences are <= a few ulp.
 *   - fastbloom 0.14.0 — bit layout/hash unknown offline: Bloom-dependent
 *     prune counts are PARITY UNPINNED against the real crate (result rows are
 *     unaffected: row-level string compare is exact).  The oracle uses this
 *     repo's documented Bloom spec.
 *
 * Every function cites the reference file:line it follows (paths relative to
 * /root/reference).
 */
#ifndef OTTERS_ORACLE_H
#define OTTERS_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* enum codes follow reference declaration order */
enum { ORA_COSINE = 0, ORA_EUCLIDEAN = 1, ORA_DOT = 2 };            /* src/vec.rs:11-16 */
enum { ORA_TAKE_MIN = 0, ORA_TAKE_MAX = 1 };                         /* src/vec.rs:18-22 */
enum { ORA_LT = 0, ORA_GT = 1, ORA_LTE = 2, ORA_GTE = 3, ORA_EQ = 4 }; /* src/vec.rs:24-31 */
enum { ORA_OP_EQ = 0, ORA_OP_NEQ = 1, ORA_OP_LT = 2, ORA_OP_LTE = 3, ORA_OP_GT = 4, ORA_OP_GTE = 5 }; /* src/expr.rs:83-91 */
enum { ORA_I32 = 0, ORA_I64 = 1, ORA_F32 = 2, ORA_F64 = 3, ORA_STR = 4, ORA_DT = 5 }; /* src/type_utils.rs:11-19 */
enum { ORA_LIT_I64 = 0, ORA_LIT_F64 = 1, ORA_LIT_STR = 2 };          /* src/expr.rs:192-210 */
enum { ORA_MODE_FAITHFUL = 0, ORA_MODE_CANONICAL = 1 };

/* 0 = wide's non-AVX order (reference default build), 1 = wide's AVX order */
void oracle_set_reduce_order(int order);

/* src/vec_compute.rs:9-22 / 25-32 / 35-54 ; src/vec.rs:365-368 */
float oracle_dot(const float *a, const float *b, size_t n);
float oracle_cosine(const float *a, const float *b, float a_inv, float b_inv, size_t n);
float oracle_l2(const float *a, const float *b, size_t n);
float oracle_inv_norm(const float *v, size_t n);
void oracle_inv_norms(const float *rows, size_t n_rows, size_t dim, float *out);

/* src/vec_compute.rs:56-74 */
uint8_t oracle_filter_mask_bits(const float *scores8, float thr, int cmp);

typedef struct {
    const float *queries; /* nq * dim, row-major */
    uint32_t nq;
    uint32_t dim;
    int32_t metric;
    int32_t take_type;
    uint64_t k;
    int32_t has_filter;
    float thr;
    int32_t cmp;
    const uint64_t *row_mask_words; /* Lsb0, bit=1 keep; NULL = none (src/vec.rs:231-237) */
    uint64_t row_mask_bits;         /* rows >= this are kept */
} ora_vec_query;

/* src/vec.rs:206-311 (collect).  Returns number of results written (<= cap).
 * mode FAITHFUL restates TopKCollector (src/vec_compute.rs:77-294) with stable
 * arrival-order ties; CANONICAL selects the best k of all candidates ordered by
 * (score, row, query id) — the tie rule this repo imposes. */
uint64_t oracle_vecstore_query(const float *vectors, const float *inv_norms, uint64_t n_vecs,
                               const ora_vec_query *q, int mode, uint64_t *out_idx, float *out_score,
                               uint32_t *out_qid, uint64_t cap);

/* ---- metadata ---- */
typedef struct {
    int32_t dtype;
    const void *values;          /* typed array, n_rows entries; String: NULL */
    const uint64_t *null_words;  /* Lsb0, bit=1 null (src/col.rs:21-28); NULL = no nulls */
    const uint64_t *str_offsets; /* String: n_rows+1 byte offsets */
    const uint8_t *str_bytes;    /* String: concatenated UTF-8 */
} ora_column;

typedef struct {
    uint32_t col;
    int32_t op;
    int32_t kind;
    int64_t i;
    double f;
    const uint8_t *s;
    uint64_t slen;
} ora_leaf;

typedef struct {
    uint32_t n_clauses;
    const uint32_t *clause_offsets; /* n_clauses + 1 */
    const ora_leaf *leaves;
} ora_filter;

typedef struct {
    uint64_t total_chunks, pruned_chunks, evaluated_chunks, vectors_compared;
    double prune_s, score_s, merge_s, total_s;
} ora_query_stats;

typedef struct ora_metastore ora_metastore;

/* src/meta.rs:151-305 (build); bloom_mode 0 = Fpr(bloom_fpr), 1 = Bits(bloom_bits).
 * The store keeps pointers to the caller's arrays (they must outlive it). */
ora_metastore *oracle_meta_build(const float *vectors, uint64_t n_rows, uint32_t dim, uint64_t chunk_size,
                                 const ora_column *cols, uint32_t n_cols, int bloom_mode, double bloom_fpr,
                                 uint64_t bloom_bits);
void oracle_meta_free(ora_metastore *);
uint64_t oracle_meta_n_chunks(const ora_metastore *);
/* zonemap export for parity tests: numeric columns -> min/max as double or int64 per dtype */
int oracle_meta_zonemap_i64(const ora_metastore *, uint32_t col, int64_t *mn, int64_t *mx, uint64_t *non_null);
int oracle_meta_zonemap_f64(const ora_metastore *, uint32_t col, double *mn, double *mx, uint64_t *non_null);
/* Bloom export: bits per chunk and packed words (stride = words of chunk 0) */
uint64_t oracle_meta_bloom_words_stride(const ora_metastore *, uint32_t col);
int oracle_meta_bloom_export(const ora_metastore *, uint32_t col, uint64_t *words, uint64_t *m_bits, uint32_t *k_hashes,
                             uint64_t *non_null);

/* src/meta.rs:407-544: chunk keep mask (one byte per chunk) */
void oracle_meta_chunk_mask(const ora_metastore *, const ora_filter *, uint8_t *keep);
/* src/meta_compute.rs:194-318: row keep mask (one byte per row; rows of pruned chunks = 0) */
void oracle_meta_row_mask(const ora_metastore *, const ora_filter *, uint8_t *keep);

/* src/meta.rs:632-721 (+ meta_compute.rs:153-192).  n_threads <= 0 -> all cores. */
uint64_t oracle_meta_query(const ora_metastore *, const ora_vec_query *q, const ora_filter *f, int mode, int n_threads,
                           uint64_t *out_idx, float *out_score, uint32_t *out_qid, uint64_t cap,
                           ora_query_stats *stats);

/* Bloom spec helpers (this repo's spec) */
void oracle_bloom_params(uint64_t n_items, int mode, double fpr, uint64_t bits, uint64_t *m_bits, uint32_t *k_hashes);
void oracle_bloom_hash(const uint8_t *s, uint64_t len, uint64_t *h1, uint64_t *h2);

/* counter-based synthetic generator: x = (splitmix64(seed ^ (row*dim+col)) >> 40) * 2^-23 - 1 */
void oracle_synth_fill(float *out, uint64_t row0, uint64_t n_rows, uint32_t dim, uint64_t seed);

int oracle_num_threads(void);

#ifdef __cplusplus
}
#endif
#endif
