/*
 * otters_oracle.c — CPU restatement of the otters exact-search hot path (plain C).
 *
 * TEST INFRASTRUCTURE ONLY — see otters_oracle.h for the rules and the parity
 * status.  Build with -ffp-contract=off: the reference multiplies then adds
 * (wide::f32x8 has no fused multiply-add in these kernels), so no FMA may be
 * formed here either.
 *
 * Citations are file:line under /root/reference.
 */
#define _GNU_SOURCE
#include "otters_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------- */
/* scoring kernels                                                            */
/* ------------------------------------------------------------------------- */

static int g_reduce_order = 0;
void oracle_set_reduce_order(int order) { g_reduce_order = order; }

/* wide 0.7.33 f32x8::reduce_add.  Non-AVX build: a.reduce_add() + b.reduce_add()
 * with f32x4::reduce_add = sequential array sum.  AVX build: (lo+hi) quads,
 * then (q0+q2),(q1+q3), then their sum. */
static inline float reduce_add8(const float l[8]) {
    if (g_reduce_order == 0) {
        float a = ((l[0] + l[1]) + l[2]) + l[3];
        float b = ((l[4] + l[5]) + l[6]) + l[7];
        return a + b;
    } else {
        float s0 = l[0] + l[4], s1 = l[1] + l[5], s2 = l[2] + l[6], s3 = l[3] + l[7];
        float d0 = s0 + s2, d1 = s1 + s3;
        return d0 + d1;
    }
}

/* src/vec_compute.rs:9-22 — 8 lane accumulators over chunks_exact(8) (multiply,
 * then add), reduce_add, plus a serial scalar sum of the remainder.  Rust's
 * f32 Sum starts from -0.0 (identity for +). */
float oracle_dot(const float *a, const float *b, size_t n) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    size_t nb = n / 8;
    for (size_t j = 0; j < nb; ++j) {
        const float *pa = a + 8 * j, *pb = b + 8 * j;
        for (int l = 0; l < 8; ++l) {
            float p = pa[l] * pb[l];
            acc[l] = acc[l] + p;
        }
    }
    float tail = -0.0f;
    for (size_t i = nb * 8; i < n; ++i) {
        float p = a[i] * b[i];
        tail = tail + p;
    }
    return reduce_add8(acc) + tail;
}

/* src/vec_compute.rs:25-32 — dot * inv1 * inv2, left to right */
float oracle_cosine(const float *a, const float *b, float a_inv, float b_inv, size_t n) {
    float d = oracle_dot(a, b, n);
    float t = d * a_inv;
    return t * b_inv;
}

/* src/vec_compute.rs:35-54 — sum of (a-b)^2, same lane structure, no sqrt */
float oracle_l2(const float *a, const float *b, size_t n) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    size_t nb = n / 8;
    for (size_t j = 0; j < nb; ++j) {
        const float *pa = a + 8 * j, *pb = b + 8 * j;
        for (int l = 0; l < 8; ++l) {
            float d = pa[l] - pb[l];
            float p = d * d;
            acc[l] = acc[l] + p;
        }
    }
    float tail = -0.0f;
    for (size_t i = nb * 8; i < n; ++i) {
        float d = a[i] - b[i];
        float p = d * d;
        tail = tail + p;
    }
    return reduce_add8(acc) + tail;
}

/* src/vec.rs:365-368 and :390-397 — serial f32 sum of squares, sqrt, 1/norm or 0 */
float oracle_inv_norm(const float *v, size_t n) {
    float s = -0.0f;
    for (size_t i = 0; i < n; ++i) {
        float p = v[i] * v[i];
        s = s + p;
    }
    float norm = sqrtf(s);
    return norm != 0.0f ? 1.0f / norm : 0.0f;
}

void oracle_inv_norms(const float *rows, size_t n_rows, size_t dim, float *out) {
#pragma omp parallel for schedule(static)
    for (long long r = 0; r < (long long)n_rows; ++r) out[r] = oracle_inv_norm(rows + (size_t)r * dim, dim);
}

static inline int cmp_score(float s, float thr, int cmp) {
    switch (cmp) {
    case ORA_LT: return s < thr;
    case ORA_GT: return s > thr;
    case ORA_LTE: return s <= thr;
    case ORA_GTE: return s >= thr;
    case ORA_EQ: return s == thr;
    }
    return 0;
}

/* src/vec_compute.rs:56-74 */
uint8_t oracle_filter_mask_bits(const float *scores8, float thr, int cmp) {
    uint8_t bits = 0;
    for (int i = 0; i < 8; ++i)
        if (cmp_score(scores8[i], thr, cmp)) bits |= (uint8_t)(1u << i);
    return bits;
}

/* ------------------------------------------------------------------------- */
/* TopKCollector — src/vec_compute.rs:77-294                                  */
/* ------------------------------------------------------------------------- */

typedef struct {
    uint64_t idx;
    float score;
    uint32_t qid;
    uint64_t seq; /* arrival order; makes the reference's unstable sorts deterministic */
} cand_t;

/* f32::total_cmp */
static inline int32_t total_key(float f) {
    int32_t b;
    memcpy(&b, &f, 4);
    b ^= (int32_t)(((uint32_t)(b >> 31)) >> 1);
    return b;
}
static inline int total_cmp(float a, float b) {
    int32_t ka = total_key(a), kb = total_key(b);
    return (ka > kb) - (ka < kb);
}

typedef struct {
    cand_t *buf;
    uint64_t len, cap, k;
    int take_type;
    int has_filter;
    float f_thr;
    int f_cmp;
    int is_sorted;
    float threshold;
    int has_eff;
    float eff_thr;
    int eff_cmp;
    uint64_t seq;
} collector_t;

/* :89-125 */
static void collector_init(collector_t *c, uint64_t k, int take_type, int has_filter, float thr, int cmp) {
    memset(c, 0, sizeof(*c));
    c->k = k;
    c->take_type = take_type;
    c->has_filter = has_filter;
    c->f_thr = thr;
    c->f_cmp = cmp;
    c->is_sorted = 1;
    c->threshold = take_type == ORA_TAKE_MIN ? INFINITY : -INFINITY;
    if (has_filter) {
        float combined;
        if (take_type == ORA_TAKE_MIN && (cmp == ORA_LT || cmp == ORA_LTE))
            combined = fminf(thr, c->threshold);
        else if (take_type == ORA_TAKE_MAX && (cmp == ORA_GT || cmp == ORA_GTE))
            combined = fmaxf(thr, c->threshold);
        else
            combined = thr;
        c->has_eff = 1;
        c->eff_thr = combined;
        c->eff_cmp = cmp;
    }
    c->cap = k < 1024 ? (k ? k : 1) : 1024;
    c->buf = (cand_t *)malloc(c->cap * sizeof(cand_t));
}

static void collector_free(collector_t *c) { free(c->buf); }

/* :127-141 */
static int collector_effective(const collector_t *c, float *thr, int *cmp) {
    if (c->has_eff) {
        *thr = c->eff_thr;
        *cmp = c->eff_cmp;
        return 1;
    }
    if (c->len == c->k) {
        *thr = c->threshold;
        *cmp = c->take_type == ORA_TAKE_MIN ? ORA_LT : ORA_GT;
        return 1;
    }
    return 0;
}

/* :143-165 */
static void collector_update_effective(collector_t *c) {
    if (c->len != c->k) return;
    if (c->has_eff) {
        if (c->take_type == ORA_TAKE_MIN && (c->eff_cmp == ORA_LT || c->eff_cmp == ORA_LTE))
            c->eff_thr = fminf(c->eff_thr, c->threshold);
        else if (c->take_type == ORA_TAKE_MAX && (c->eff_cmp == ORA_GT || c->eff_cmp == ORA_GTE))
            c->eff_thr = fmaxf(c->eff_thr, c->threshold);
    } else {
        c->has_eff = 1;
        c->eff_thr = c->threshold;
        c->eff_cmp = c->take_type == ORA_TAKE_MIN ? ORA_LT : ORA_GT;
    }
}

static int g_sort_take; /* comparator context (single-threaded use per sort call is guarded) */
static int cand_cmp_ctx(const void *pa, const void *pb, int take) {
    const cand_t *a = (const cand_t *)pa, *b = (const cand_t *)pb;
    int c = take == ORA_TAKE_MIN ? total_cmp(a->score, b->score) : total_cmp(b->score, a->score);
    if (c) return c;
    return (a->seq > b->seq) - (a->seq < b->seq);
}
static int cand_cmp_min(const void *a, const void *b) { return cand_cmp_ctx(a, b, ORA_TAKE_MIN); }
static int cand_cmp_max(const void *a, const void *b) { return cand_cmp_ctx(a, b, ORA_TAKE_MAX); }

/* :270-288 — sort_unstable_by(total_cmp); ties made deterministic by arrival order */
static void collector_sort(collector_t *c) {
    (void)g_sort_take;
    if (!c->is_sorted) {
        qsort(c->buf, c->len, sizeof(cand_t), c->take_type == ORA_TAKE_MIN ? cand_cmp_min : cand_cmp_max);
        c->is_sorted = 1;
    }
}

/* :236-268 */
static void collector_push_single(collector_t *c, uint64_t idx, float score, uint32_t qid) {
    if (isnan(score)) return;
    if (c->len == c->k) {
        int should = c->take_type == ORA_TAKE_MIN ? (score < c->threshold) : (score > c->threshold);
        if (should) {
            /* :260-268 binary_search_by(total_cmp); insert after equal scores (arrival-stable) */
            uint64_t lo = 0, hi = c->len;
            while (lo < hi) {
                uint64_t mid = (lo + hi) / 2;
                int o = c->take_type == ORA_TAKE_MIN ? total_cmp(c->buf[mid].score, score)
                                                     : total_cmp(score, c->buf[mid].score);
                if (o <= 0) lo = mid + 1; else hi = mid;
            }
            /* insert at lo, pop last */
            if (lo < c->len) {
                memmove(&c->buf[lo + 1], &c->buf[lo], (c->len - 1 - lo) * sizeof(cand_t));
                c->buf[lo].idx = idx;
                c->buf[lo].score = score;
                c->buf[lo].qid = qid;
                c->buf[lo].seq = c->seq++;
            }
            c->threshold = c->buf[c->k - 1].score;
            collector_update_effective(c);
        }
    } else {
        if (c->len == c->cap) {
            uint64_t nc = c->cap * 2;
            if (nc > c->k) nc = c->k;
            c->buf = (cand_t *)realloc(c->buf, nc * sizeof(cand_t));
            c->cap = nc;
        }
        cand_t *e = &c->buf[c->len++];
        e->idx = idx;
        e->score = score;
        e->qid = qid;
        e->seq = c->seq++;
        c->is_sorted = 0;
        if (c->len == c->k) {
            collector_sort(c);
            c->threshold = c->buf[c->k - 1].score;
            collector_update_effective(c);
        }
    }
}

/* :168-208 */
static void collector_push_chunk_masked(collector_t *c, uint64_t chunk_idx, const float *scores8, const uint8_t *rowmask8,
                                        uint32_t qid) {
    if (c->k == 0) return;
    float thr;
    int cmp;
    uint8_t tbits = 0xFF;
    if (collector_effective(c, &thr, &cmp)) tbits = oracle_filter_mask_bits(scores8, thr, cmp);
    uint8_t sbits = 0xFF;
    if (rowmask8) {
        sbits = 0;
        for (int i = 0; i < 8; ++i)
            if (rowmask8[i]) sbits |= (uint8_t)(1u << i);
    }
    uint8_t bits = tbits & sbits;
    if (!bits) return;
    for (int i = 0; i < 8; ++i)
        if ((bits >> i) & 1) collector_push_single(c, chunk_idx * 8 + (uint64_t)i, scores8[i], qid);
}

/* :210-234 — remainder rows use the RAW filter, not the effective threshold */
static void collector_push_scalar(collector_t *c, uint64_t idx, float score, uint32_t qid) {
    if (c->k == 0) return;
    if (c->has_filter && !cmp_score(score, c->f_thr, c->f_cmp)) return;
    collector_push_single(c, idx, score, qid);
}

/* ------------------------------------------------------------------------- */
/* VecQueryPlan::collect — src/vec.rs:206-311                                 */
/* ------------------------------------------------------------------------- */

static inline int mask_keep(const uint64_t *words, uint64_t nbits, uint64_t row) {
    /* src/vec.rs:234,297 — rm.get(row).unwrap_or(true) */
    if (!words || row >= nbits) return 1;
    return (int)((words[row >> 6] >> (row & 63)) & 1);
}

static inline float score_one(int metric, const float *q, const float *v, float q_inv, float v_inv, size_t dim) {
    switch (metric) {
    case ORA_COSINE: return oracle_cosine(q, v, q_inv, v_inv, dim);
    case ORA_EUCLIDEAN: return oracle_l2(q, v, dim);
    default: return oracle_dot(q, v, dim);
    }
}

/* canonical ordering: better score first (IEEE compare, so -0.0 == +0.0), then lower row, then lower query */
static int g_canon_take;
static int canon_cmp(const void *pa, const void *pb) {
    const cand_t *a = (const cand_t *)pa, *b = (const cand_t *)pb;
    if (a->score != b->score) {
        if (g_canon_take == ORA_TAKE_MIN) return a->score < b->score ? -1 : 1;
        return a->score > b->score ? -1 : 1;
    }
    if (a->idx != b->idx) return a->idx < b->idx ? -1 : 1;
    return (a->qid > b->qid) - (a->qid < b->qid);
}

static uint64_t emit(const cand_t *buf, uint64_t n, uint64_t base, uint64_t *out_idx, float *out_score, uint32_t *out_qid,
                     uint64_t cap) {
    uint64_t m = n < cap ? n : cap;
    for (uint64_t i = 0; i < m; ++i) {
        if (out_idx) out_idx[i] = base + buf[i].idx;
        if (out_score) out_score[i] = buf[i].score;
        if (out_qid) out_qid[i] = buf[i].qid;
    }
    return m;
}

/* Runs the scan into a collector (faithful) or a flat candidate list (canonical). */
static void scan_faithful(const float *vectors, const float *inv_norms, uint64_t n_vecs, const ora_vec_query *q,
                          const float *q_inv, collector_t *col) {
    const size_t dim = q->dim;
    uint64_t full = n_vecs / 8;
    for (uint64_t b = 0; b < full; ++b) { /* :222-267 */
        uint64_t base_row = b * 8;
        uint8_t bm[8];
        int has_bm = q->row_mask_words != NULL;
        if (has_bm)
            for (int i = 0; i < 8; ++i) bm[i] = (uint8_t)mask_keep(q->row_mask_words, q->row_mask_bits, base_row + i);
        float scratch[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (uint32_t qi = 0; qi < q->nq; ++qi) {
            const float *qv = q->queries + (size_t)qi * dim;
            for (int i = 0; i < 8; ++i) {
                if (has_bm && !bm[i]) continue; /* :248-252 masked rows are not scored */
                uint64_t row = base_row + i;
                scratch[i] = score_one(q->metric, qv, vectors + row * dim, q_inv[qi], inv_norms[row], dim);
            }
            collector_push_chunk_masked(col, b, scratch, has_bm ? bm : NULL, qi);
        }
    }
    uint64_t rem0 = full * 8; /* :270-303 */
    if (rem0 < n_vecs) {
        for (uint32_t qi = 0; qi < q->nq; ++qi) {
            const float *qv = q->queries + (size_t)qi * dim;
            for (uint64_t row = rem0; row < n_vecs; ++row) {
                float s = score_one(q->metric, qv, vectors + row * dim, q_inv[qi], inv_norms[row], dim);
                if (!mask_keep(q->row_mask_words, q->row_mask_bits, row)) continue;
                collector_push_scalar(col, row, s, qi);
            }
        }
    }
}

typedef struct {
    cand_t *v;
    uint64_t n, cap;
} candvec_t;
static void cv_push(candvec_t *c, uint64_t idx, float s, uint32_t qid) {
    if (c->n == c->cap) {
        c->cap = c->cap ? c->cap * 2 : 1024;
        c->v = (cand_t *)realloc(c->v, c->cap * sizeof(cand_t));
    }
    c->v[c->n].idx = idx;
    c->v[c->n].score = s;
    c->v[c->n].qid = qid;
    c->v[c->n].seq = c->n;
    c->n++;
}

static void scan_canonical(const float *vectors, const float *inv_norms, uint64_t n_vecs, uint64_t idx_base,
                           const ora_vec_query *q, const float *q_inv, const uint8_t *row_keep_bytes, candvec_t *out) {
    const size_t dim = q->dim;
    for (uint64_t row = 0; row < n_vecs; ++row) {
        if (row_keep_bytes ? !row_keep_bytes[row] : !mask_keep(q->row_mask_words, q->row_mask_bits, row)) continue;
        for (uint32_t qi = 0; qi < q->nq; ++qi) {
            float s = score_one(q->metric, q->queries + (size_t)qi * dim, vectors + row * dim, q_inv[qi], inv_norms[row],
                                dim);
            if (isnan(s)) continue;
            if (q->has_filter && !cmp_score(s, q->thr, q->cmp)) continue;
            cv_push(out, idx_base + row, s, qi);
        }
    }
}

static uint64_t finish_canonical(candvec_t *cv, int take, uint64_t k, uint64_t *out_idx, float *out_score, uint32_t *out_qid,
                                 uint64_t cap) {
    g_canon_take = take;
    qsort(cv->v, cv->n, sizeof(cand_t), canon_cmp);
    uint64_t n = cv->n < k ? cv->n : k;
    return emit(cv->v, n, 0, out_idx, out_score, out_qid, cap);
}

uint64_t oracle_vecstore_query(const float *vectors, const float *inv_norms, uint64_t n_vecs, const ora_vec_query *q,
                               int mode, uint64_t *out_idx, float *out_score, uint32_t *out_qid, uint64_t cap) {
    if (q->nq == 0) return 0;
    float *q_inv = (float *)malloc(sizeof(float) * q->nq);
    for (uint32_t i = 0; i < q->nq; ++i) q_inv[i] = oracle_inv_norm(q->queries + (size_t)i * q->dim, q->dim);
    uint64_t n = 0;
    if (mode == ORA_MODE_FAITHFUL) {
        collector_t col;
        collector_init(&col, q->k, q->take_type, q->has_filter, q->thr, q->cmp);
        scan_faithful(vectors, inv_norms, n_vecs, q, q_inv, &col);
        collector_sort(&col); /* :290-293 into_sorted_vec */
        n = emit(col.buf, col.len, 0, out_idx, out_score, out_qid, cap);
        collector_free(&col);
    } else {
        candvec_t cv = {0, 0, 0};
        if (q->k) scan_canonical(vectors, inv_norms, n_vecs, 0, q, q_inv, NULL, &cv);
        n = finish_canonical(&cv, q->take_type, q->k, out_idx, out_score, out_qid, cap);
        free(cv.v);
    }
    free(q_inv);
    return n;
}

/* ------------------------------------------------------------------------- */
/* Bloom filter — this repo's spec (fastbloom is not on disk: parity unpinned) */
/* ------------------------------------------------------------------------- */

static inline uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
}

void oracle_bloom_hash(const uint8_t *s, uint64_t len, uint64_t *h1, uint64_t *h2) {
    uint64_t h = 0xCBF29CE484222325ULL; /* FNV-1a 64 */
    for (uint64_t i = 0; i < len; ++i) {
        h ^= s[i];
        h *= 0x100000001B3ULL;
    }
    *h1 = mix64(h);
    *h2 = mix64(h ^ 0x9E3779B97F4A7C15ULL) | 1ULL;
}

void oracle_bloom_params(uint64_t n_items, int mode, double fpr, uint64_t bits, uint64_t *m_bits, uint32_t *k_hashes) {
    uint64_t n = n_items ? n_items : 1;
    uint64_t m;
    if (mode == 0) {
        double ln2 = 0.6931471805599453;
        double mm = ceil(-(double)n * log(fpr) / (ln2 * ln2));
        m = mm < 64.0 ? 64 : (uint64_t)mm;
    } else {
        m = bits < 64 ? 64 : bits;
    }
    m = (m + 63) / 64 * 64;
    double kk = floor((double)m / (double)n * 0.6931471805599453 + 0.5);
    uint32_t k = kk < 1.0 ? 1u : (kk > 16.0 ? 16u : (uint32_t)kk);
    *m_bits = m;
    *k_hashes = k;
}

/* probe i = (a + i*b) mod m with a = h1 mod m, b = h2 mod m (1 if that is 0): double hashing in Z_m, stepped without
 * a division per probe */
static inline void bloom_insert(uint64_t *words, uint64_t m, uint32_t k, const uint8_t *s, uint64_t len) {
    uint64_t h1, h2;
    oracle_bloom_hash(s, len, &h1, &h2);
    uint64_t bit = h1 % m, step = h2 % m;
    if (step == 0) step = 1;
    for (uint32_t i = 0; i < k; ++i) {
        words[bit >> 6] |= 1ULL << (bit & 63);
        bit += step;
        if (bit >= m) bit -= m;
    }
}
static inline int bloom_contains(const uint64_t *words, uint64_t m, uint32_t k, const uint8_t *s, uint64_t len) {
    uint64_t h1, h2;
    oracle_bloom_hash(s, len, &h1, &h2);
    uint64_t bit = h1 % m, step = h2 % m;
    if (step == 0) step = 1;
    for (uint32_t i = 0; i < k; ++i) {
        if (!((words[bit >> 6] >> (bit & 63)) & 1)) return 0;
        bit += step;
        if (bit >= m) bit -= m;
    }
    return 1;
}

/* ------------------------------------------------------------------------- */
/* MetaStore build — src/meta.rs:151-305, src/meta_compute.rs:32-132           */
/* ------------------------------------------------------------------------- */

typedef struct {
    /* numeric zonemaps (ZoneStat, meta_compute.rs:18-24), packed as meta.rs:237-271 */
    int64_t *imin, *imax; /* Int32 (already truncated to i32 range semantics), Int64, DateTime */
    double *fmin, *fmax;  /* Float32 (rounded to f32), Float64 */
    uint64_t *non_null;
    /* string Bloom */
    uint64_t *bloom;      /* n_chunks * stride words */
    uint64_t stride;
    uint64_t *m_bits;
    uint32_t *k_hashes;
} zone_t;

struct ora_metastore {
    const float *vectors;
    uint64_t n_rows;
    uint32_t dim;
    uint64_t chunk_size;
    uint64_t n_chunks;
    ora_column *cols;
    uint32_t n_cols;
    float *inv_norms;
    zone_t *zones;
};

static inline int is_null(const ora_column *c, uint64_t row) {
    return c->null_words ? (int)((c->null_words[row >> 6] >> (row & 63)) & 1) : 0;
}

ora_metastore *oracle_meta_build(const float *vectors, uint64_t n_rows, uint32_t dim, uint64_t chunk_size,
                                 const ora_column *cols, uint32_t n_cols, int bloom_mode, double bloom_fpr,
                                 uint64_t bloom_bits) {
    ora_metastore *m = (ora_metastore *)calloc(1, sizeof(*m));
    if (chunk_size < 1) chunk_size = 1; /* meta.rs:86-89 */
    if (bloom_mode == 0) {              /* meta.rs:92-101 */
        if (!isfinite(bloom_fpr)) bloom_fpr = 0.01;
        if (bloom_fpr < 1e-2) bloom_fpr = 1e-2;
        if (bloom_fpr > 0.5) bloom_fpr = 0.5;
    } else if (bloom_bits < 64) bloom_bits = 64; /* meta.rs:106-110 */
    m->vectors = vectors;
    m->n_rows = n_rows;
    m->dim = dim;
    m->chunk_size = chunk_size;
    m->n_chunks = (n_rows + chunk_size - 1) / chunk_size;
    m->n_cols = n_cols;
    m->cols = (ora_column *)malloc(sizeof(ora_column) * (n_cols ? n_cols : 1));
    memcpy(m->cols, cols, sizeof(ora_column) * n_cols);
    m->inv_norms = (float *)malloc(sizeof(float) * (n_rows ? n_rows : 1));
    oracle_inv_norms(vectors, n_rows, dim, m->inv_norms); /* vec.rs:357-371 via meta.rs:209-212 */
    m->zones = (zone_t *)calloc(n_cols ? n_cols : 1, sizeof(zone_t));
    uint64_t nc = m->n_chunks;
    for (uint32_t ci = 0; ci < n_cols; ++ci) {
        const ora_column *c = &m->cols[ci];
        zone_t *z = &m->zones[ci];
        z->non_null = (uint64_t *)calloc(nc ? nc : 1, sizeof(uint64_t));
        if (c->dtype == ORA_STR) {
            uint64_t m0;
            uint32_t k0;
            oracle_bloom_params(chunk_size < n_rows ? chunk_size : (n_rows ? n_rows : 1), bloom_mode, bloom_fpr, bloom_bits,
                                &m0, &k0);
            z->stride = m0 / 64;
            z->bloom = (uint64_t *)calloc((nc ? nc : 1) * z->stride, sizeof(uint64_t));
            z->m_bits = (uint64_t *)calloc(nc ? nc : 1, sizeof(uint64_t));
            z->k_hashes = (uint32_t *)calloc(nc ? nc : 1, sizeof(uint32_t));
        } else if (c->dtype == ORA_F32 || c->dtype == ORA_F64) {
            z->fmin = (double *)malloc(sizeof(double) * (nc ? nc : 1));
            z->fmax = (double *)malloc(sizeof(double) * (nc ? nc : 1));
        } else {
            z->imin = (int64_t *)malloc(sizeof(int64_t) * (nc ? nc : 1));
            z->imax = (int64_t *)malloc(sizeof(int64_t) * (nc ? nc : 1));
        }
#pragma omp parallel for schedule(static)
        for (long long ch = 0; ch < (long long)nc; ++ch) {
            uint64_t start = (uint64_t)ch * chunk_size, end = start + chunk_size;
            if (end > n_rows) end = n_rows;
            uint64_t cnt = 0;
            switch (c->dtype) {
            case ORA_I32: { /* meta_compute.rs:41-54, packed meta.rs:252-257 (`as i32`) */
                int64_t mn = INT64_MAX, mx = INT64_MIN;
                const int32_t *v = (const int32_t *)c->values;
                for (uint64_t i = start; i < end; ++i)
                    if (!is_null(c, i)) {
                        int64_t x = v[i];
                        if (x < mn) mn = x;
                        if (x > mx) mx = x;
                        ++cnt;
                    }
                z->imin[ch] = (int64_t)(int32_t)(uint32_t)(uint64_t)mn;
                z->imax[ch] = (int64_t)(int32_t)(uint32_t)(uint64_t)mx;
                break;
            }
            case ORA_I64:
            case ORA_DT: { /* meta_compute.rs:55-68, :117-130 */
                int64_t mn = INT64_MAX, mx = INT64_MIN;
                const int64_t *v = (const int64_t *)c->values;
                for (uint64_t i = start; i < end; ++i)
                    if (!is_null(c, i)) {
                        int64_t x = v[i];
                        if (x < mn) mn = x;
                        if (x > mx) mx = x;
                        ++cnt;
                    }
                z->imin[ch] = mn;
                z->imax[ch] = mx;
                break;
            }
            case ORA_F32: { /* meta_compute.rs:69-83 (f64::min/max ignore NaN), packed `as f32` meta.rs:240-245 */
                double mn = INFINITY, mx = -INFINITY;
                const float *v = (const float *)c->values;
                for (uint64_t i = start; i < end; ++i)
                    if (!is_null(c, i)) {
                        double x = v[i];
                        mn = fmin(mn, x);
                        mx = fmax(mx, x);
                        ++cnt;
                    }
                z->fmin[ch] = (double)(float)mn;
                z->fmax[ch] = (double)(float)mx;
                break;
            }
            case ORA_F64: { /* meta_compute.rs:84-98 */
                double mn = INFINITY, mx = -INFINITY;
                const double *v = (const double *)c->values;
                for (uint64_t i = start; i < end; ++i)
                    if (!is_null(c, i)) {
                        double x = v[i];
                        mn = fmin(mn, x);
                        mx = fmax(mx, x);
                        ++cnt;
                    }
                z->fmin[ch] = mn;
                z->fmax[ch] = mx;
                break;
            }
            case ORA_STR: { /* meta_compute.rs:99-116: sized for end-start items; only non-null inserted */
                uint64_t mb;
                uint32_t kh;
                oracle_bloom_params(end - start, bloom_mode, bloom_fpr, bloom_bits, &mb, &kh);
                z->m_bits[ch] = mb;
                z->k_hashes[ch] = kh;
                uint64_t *w = z->bloom + (uint64_t)ch * z->stride;
                for (uint64_t i = start; i < end; ++i)
                    if (!is_null(c, i)) {
                        bloom_insert(w, mb, kh, c->str_bytes + c->str_offsets[i], c->str_offsets[i + 1] - c->str_offsets[i]);
                        ++cnt;
                    }
                break;
            }
            }
            z->non_null[ch] = cnt;
        }
    }
    return m;
}

void oracle_meta_free(ora_metastore *m) {
    if (!m) return;
    for (uint32_t i = 0; i < m->n_cols; ++i) {
        zone_t *z = &m->zones[i];
        free(z->imin); free(z->imax); free(z->fmin); free(z->fmax); free(z->non_null);
        free(z->bloom); free(z->m_bits); free(z->k_hashes);
    }
    free(m->zones);
    free(m->cols);
    free(m->inv_norms);
    free(m);
}

uint64_t oracle_meta_n_chunks(const ora_metastore *m) { return m->n_chunks; }

int oracle_meta_zonemap_i64(const ora_metastore *m, uint32_t col, int64_t *mn, int64_t *mx, uint64_t *non_null) {
    if (col >= m->n_cols || !m->zones[col].imin) return -1;
    memcpy(mn, m->zones[col].imin, m->n_chunks * 8);
    memcpy(mx, m->zones[col].imax, m->n_chunks * 8);
    memcpy(non_null, m->zones[col].non_null, m->n_chunks * 8);
    return 0;
}
int oracle_meta_zonemap_f64(const ora_metastore *m, uint32_t col, double *mn, double *mx, uint64_t *non_null) {
    if (col >= m->n_cols || !m->zones[col].fmin) return -1;
    memcpy(mn, m->zones[col].fmin, m->n_chunks * 8);
    memcpy(mx, m->zones[col].fmax, m->n_chunks * 8);
    memcpy(non_null, m->zones[col].non_null, m->n_chunks * 8);
    return 0;
}
uint64_t oracle_meta_bloom_words_stride(const ora_metastore *m, uint32_t col) {
    return col < m->n_cols ? m->zones[col].stride : 0;
}
int oracle_meta_bloom_export(const ora_metastore *m, uint32_t col, uint64_t *words, uint64_t *m_bits, uint32_t *k_hashes,
                             uint64_t *non_null) {
    if (col >= m->n_cols || !m->zones[col].bloom) return -1;
    const zone_t *z = &m->zones[col];
    memcpy(words, z->bloom, m->n_chunks * z->stride * 8);
    memcpy(m_bits, z->m_bits, m->n_chunks * 8);
    memcpy(k_hashes, z->k_hashes, m->n_chunks * 4);
    memcpy(non_null, z->non_null, m->n_chunks * 8);
    return 0;
}

/* ------------------------------------------------------------------------- */
/* chunk pruning — src/meta.rs:407-544, src/type_utils.rs:446-584,739-889      */
/* ------------------------------------------------------------------------- */

/* Rust `f64 as i64` / `f64 as i32`: saturating, NaN -> 0 */
static inline int64_t f64_as_i64(double v) {
    if (isnan(v)) return 0;
    if (v >= 9223372036854775808.0) return INT64_MAX;
    if (v <= -9223372036854775808.0) return INT64_MIN;
    return (int64_t)v;
}
static inline int32_t f64_as_i32(double v) {
    if (isnan(v)) return 0;
    if (v >= 2147483647.0) return INT32_MAX;
    if (v <= -2147483648.0) return INT32_MIN;
    return (int32_t)v;
}
static inline int32_t i64_as_i32(int64_t v) { return (int32_t)(uint32_t)(uint64_t)v; }

#define RANGE_SAT(op, mn, mx, t) \
    ((op) == ORA_OP_EQ ? ((mn) <= (t) && (t) <= (mx)) : (op) == ORA_OP_LT ? ((mn) < (t)) : (op) == ORA_OP_LTE ? ((mn) <= (t)) \
     : (op) == ORA_OP_GT ? ((mx) > (t)) : (op) == ORA_OP_GTE ? ((mx) >= (t)) : 1)

static void chunk_leaf(const ora_metastore *m, const ora_leaf *lf, uint8_t *clause) {
    uint64_t nc = m->n_chunks;
    if (lf->col >= m->n_cols) {
        /* unknown column: numeric -> no table -> nothing set; string -> "conservatively keep" (meta.rs:540-542) */
        if (lf->kind == ORA_LIT_STR) memset(clause, 1, nc);
        return;
    }
    const ora_column *c = &m->cols[lf->col];
    const zone_t *z = &m->zones[lf->col];
    if (lf->kind == ORA_LIT_STR) { /* meta.rs:523-544 */
        if (c->dtype != ORA_STR) {
            memset(clause, 1, nc);
            return;
        }
        for (uint64_t ch = 0; ch < nc; ++ch) {
            if (z->non_null[ch] == 0) continue;
            if (lf->op == ORA_OP_EQ) {
                if (bloom_contains(z->bloom + ch * z->stride, z->m_bits[ch], z->k_hashes[ch], lf->s, lf->slen)) clause[ch] = 1;
            } else if (lf->op == ORA_OP_NEQ) {
                clause[ch] = 1;
            }
        }
        return;
    }
    /* meta.rs:431-521 */
    if (lf->kind == ORA_LIT_F64) {
        if (c->dtype == ORA_F32) {
            float t = (float)lf->f;
            for (uint64_t ch = 0; ch < nc; ++ch) {
                float mn = (float)z->fmin[ch], mx = (float)z->fmax[ch];
                if (RANGE_SAT(lf->op, mn, mx, t) && z->non_null[ch] > 0) clause[ch] = 1;
            }
        } else if (c->dtype == ORA_F64) {
            double t = lf->f;
            for (uint64_t ch = 0; ch < nc; ++ch)
                if (RANGE_SAT(lf->op, z->fmin[ch], z->fmax[ch], t) && z->non_null[ch] > 0) clause[ch] = 1;
        } else if (c->dtype == ORA_I64 || c->dtype == ORA_DT) { /* `_` arm: packed_ranges_i64 holds Int64/DateTime only */
            int64_t t = f64_as_i64(lf->f);
            for (uint64_t ch = 0; ch < nc; ++ch)
                if (RANGE_SAT(lf->op, z->imin[ch], z->imax[ch], t) && z->non_null[ch] > 0) clause[ch] = 1;
        }
    } else { /* I64 literal */
        if (c->dtype == ORA_I32) {
            int32_t t = i64_as_i32(lf->i);
            for (uint64_t ch = 0; ch < nc; ++ch) {
                int32_t mn = (int32_t)z->imin[ch], mx = (int32_t)z->imax[ch];
                if (RANGE_SAT(lf->op, mn, mx, t) && z->non_null[ch] > 0) clause[ch] = 1;
            }
        } else if (c->dtype == ORA_I64 || c->dtype == ORA_DT) {
            int64_t t = lf->i;
            for (uint64_t ch = 0; ch < nc; ++ch)
                if (RANGE_SAT(lf->op, z->imin[ch], z->imax[ch], t) && z->non_null[ch] > 0) clause[ch] = 1;
        }
        /* Float columns with an I64 literal: `_ => {}` (meta.rs:518) */
    }
}

void oracle_meta_chunk_mask(const ora_metastore *m, const ora_filter *f, uint8_t *keep) {
    uint64_t nc = m->n_chunks;
    memset(keep, 1, nc);
    if (!f) return;
    uint8_t *clause = (uint8_t *)malloc(nc ? nc : 1);
    for (uint32_t ci = 0; ci < f->n_clauses; ++ci) { /* meta.rs:412-427 */
        memset(clause, 0, nc);
        for (uint32_t li = f->clause_offsets[ci]; li < f->clause_offsets[ci + 1]; ++li) chunk_leaf(m, &f->leaves[li], clause);
        for (uint64_t ch = 0; ch < nc; ++ch) keep[ch] &= clause[ch];
    }
    free(clause);
}

/* ------------------------------------------------------------------------- */
/* row masks — src/meta_compute.rs:194-318, src/type_utils.rs:306-444,586-736  */
/* ------------------------------------------------------------------------- */

#define ROW_SAT(op, v, t) \
    ((op) == ORA_OP_EQ ? ((v) == (t)) : (op) == ORA_OP_NEQ ? ((v) != (t)) : (op) == ORA_OP_LT ? ((v) < (t)) \
     : (op) == ORA_OP_LTE ? ((v) <= (t)) : (op) == ORA_OP_GT ? ((v) > (t)) : ((v) >= (t)))

static void row_leaf(const ora_metastore *m, const ora_leaf *lf, uint64_t base, uint64_t len, uint8_t *clause) {
    if (lf->col >= m->n_cols) return; /* columns.get(column) == None */
    const ora_column *c = &m->cols[lf->col];
    if (lf->kind == ORA_LIT_STR) { /* meta_compute.rs:291-318 */
        if (c->dtype != ORA_STR) return;
        for (uint64_t off = 0; off < len; ++off) {
            uint64_t r = base + off;
            if (is_null(c, r)) continue;
            uint64_t l = c->str_offsets[r + 1] - c->str_offsets[r];
            int eq = (l == lf->slen) && (l == 0 || memcmp(c->str_bytes + c->str_offsets[r], lf->s, l) == 0);
            int sat = lf->op == ORA_OP_EQ ? eq : lf->op == ORA_OP_NEQ ? !eq : 0;
            if (sat) clause[off] = 1;
        }
        return;
    }
    switch (c->dtype) { /* meta_compute.rs:244-288 */
    case ORA_F32: {
        float t = lf->kind == ORA_LIT_F64 ? (float)lf->f : (float)lf->i;
        const float *v = (const float *)c->values;
        for (uint64_t off = 0; off < len; ++off)
            if (ROW_SAT(lf->op, v[base + off], t) && !is_null(c, base + off)) clause[off] = 1;
        break;
    }
    case ORA_I32: {
        int32_t t = lf->kind == ORA_LIT_I64 ? i64_as_i32(lf->i) : f64_as_i32(lf->f);
        const int32_t *v = (const int32_t *)c->values;
        for (uint64_t off = 0; off < len; ++off)
            if (ROW_SAT(lf->op, v[base + off], t) && !is_null(c, base + off)) clause[off] = 1;
        break;
    }
    case ORA_F64: {
        double t = lf->kind == ORA_LIT_F64 ? lf->f : (double)lf->i;
        const double *v = (const double *)c->values;
        for (uint64_t off = 0; off < len; ++off)
            if (ROW_SAT(lf->op, v[base + off], t) && !is_null(c, base + off)) clause[off] = 1;
        break;
    }
    case ORA_I64:
    case ORA_DT: {
        int64_t t = lf->kind == ORA_LIT_I64 ? lf->i : f64_as_i64(lf->f);
        const int64_t *v = (const int64_t *)c->values;
        for (uint64_t off = 0; off < len; ++off)
            if (ROW_SAT(lf->op, v[base + off], t) && !is_null(c, base + off)) clause[off] = 1;
        break;
    }
    default: break; /* String column with numeric leaf: nothing */
    }
}

/* meta_compute.rs:194-232 for one chunk; keep/clause are len bytes */
static void row_mask_chunk(const ora_metastore *m, const ora_filter *f, uint64_t base, uint64_t len, uint8_t *keep,
                           uint8_t *clause) {
    memset(keep, 1, len);
    for (uint32_t ci = 0; ci < f->n_clauses; ++ci) {
        memset(clause, 0, len);
        for (uint32_t li = f->clause_offsets[ci]; li < f->clause_offsets[ci + 1]; ++li)
            row_leaf(m, &f->leaves[li], base, len, clause);
        for (uint64_t i = 0; i < len; ++i) keep[i] &= clause[i];
    }
}

void oracle_meta_row_mask(const ora_metastore *m, const ora_filter *f, uint8_t *keep) {
    uint64_t nc = m->n_chunks;
    uint8_t *ck = (uint8_t *)malloc(nc ? nc : 1);
    oracle_meta_chunk_mask(m, f, ck);
#pragma omp parallel
    {
        uint8_t *clause = (uint8_t *)malloc(m->chunk_size);
#pragma omp for schedule(dynamic, 16)
        for (long long ch = 0; ch < (long long)nc; ++ch) {
            uint64_t base = (uint64_t)ch * m->chunk_size;
            uint64_t len = base + m->chunk_size <= m->n_rows ? m->chunk_size : m->n_rows - base;
            if (!ck[ch]) memset(keep + base, 0, len);
            else if (!f) memset(keep + base, 1, len);
            else row_mask_chunk(m, f, base, len, keep + base, clause);
        }
        free(clause);
    }
    free(ck);
}

/* ------------------------------------------------------------------------- */
/* MetaQueryPlan::collect — src/meta.rs:632-721, src/meta_compute.rs:153-192   */
/* ------------------------------------------------------------------------- */

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int oracle_num_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

static int merge_cmp_min(const void *pa, const void *pb) {
    const cand_t *a = (const cand_t *)pa, *b = (const cand_t *)pb;
    if (a->score < b->score) return -1;
    if (a->score > b->score) return 1;
    return (a->seq > b->seq) - (a->seq < b->seq);
}
static int merge_cmp_max(const void *pa, const void *pb) {
    const cand_t *a = (const cand_t *)pa, *b = (const cand_t *)pb;
    if (a->score > b->score) return -1;
    if (a->score < b->score) return 1;
    return (a->seq > b->seq) - (a->seq < b->seq);
}

uint64_t oracle_meta_query(const ora_metastore *m, const ora_vec_query *q, const ora_filter *f, int mode, int n_threads,
                           uint64_t *out_idx, float *out_score, uint32_t *out_qid, uint64_t cap, ora_query_stats *stats) {
    double t_total = now_s();
    uint64_t nc = m->n_chunks;
#ifdef _OPENMP
    int nt = n_threads > 0 ? n_threads : omp_get_max_threads();
#else
    int nt = 1;
    (void)n_threads;
#endif
    /* (1) prune — meta.rs:646-660 */
    double t0 = now_s();
    uint8_t *ck = (uint8_t *)malloc(nc ? nc : 1);
    oracle_meta_chunk_mask(m, f, ck);
    uint64_t *cand = (uint64_t *)malloc(sizeof(uint64_t) * (nc ? nc : 1));
    uint64_t n_cand = 0;
    for (uint64_t ch = 0; ch < nc; ++ch)
        if (ck[ch]) cand[n_cand++] = ch;
    double prune_s = now_s() - t0;

    /* the per-chunk VecStore::collect swallows validation errors (meta_compute.rs:182):
       wrong-dimension or empty query batches yield no rows but stats are still counted */
    int chunk_err = (q->nq == 0) || (q->dim != m->dim);

    /* (2) score — meta.rs:671-697 */
    t0 = now_s();
    uint64_t vectors_compared = 0;
    float *q_inv = (float *)malloc(sizeof(float) * (q->nq ? q->nq : 1));
    if (!chunk_err)
        for (uint32_t i = 0; i < q->nq; ++i) q_inv[i] = oracle_inv_norm(q->queries + (size_t)i * q->dim, q->dim);
    uint64_t n_out = 0;

    if (mode == ORA_MODE_FAITHFUL) {
        cand_t **res = (cand_t **)calloc(n_cand ? n_cand : 1, sizeof(cand_t *));
        uint64_t *res_n = (uint64_t *)calloc(n_cand ? n_cand : 1, sizeof(uint64_t));
#pragma omp parallel num_threads(nt)
        {
            uint8_t *keep = (uint8_t *)malloc(m->chunk_size);
            uint8_t *clause = (uint8_t *)malloc(m->chunk_size);
            uint64_t *mw = (uint64_t *)malloc(((m->chunk_size + 63) / 64) * 8);
#pragma omp for schedule(dynamic, 1)
            for (long long ci = 0; ci < (long long)n_cand; ++ci) {
                uint64_t ch = cand[ci];
                uint64_t base = ch * m->chunk_size;
                uint64_t len = base + m->chunk_size <= m->n_rows ? m->chunk_size : m->n_rows - base;
                if (chunk_err) continue;
                ora_vec_query cq = *q;
                cq.row_mask_words = NULL;
                cq.row_mask_bits = 0;
                if (f) { /* meta_compute.rs:169-170 */
                    row_mask_chunk(m, f, base, len, keep, clause);
                    memset(mw, 0, ((len + 63) / 64) * 8);
                    for (uint64_t i = 0; i < len; ++i)
                        if (keep[i]) mw[i >> 6] |= 1ULL << (i & 63);
                    cq.row_mask_words = mw;
                    cq.row_mask_bits = len;
                }
                collector_t col; /* meta_compute.rs:172-182: chunk.vec_store.query(..).filter(..).with_row_mask(..).take(k) */
                collector_init(&col, q->k, q->take_type, q->has_filter, q->thr, q->cmp);
                scan_faithful(m->vectors + base * m->dim, m->inv_norms + base, len, &cq, q_inv, &col);
                collector_sort(&col);
                for (uint64_t i = 0; i < col.len; ++i) col.buf[i].idx += base; /* :185 */
                res[ci] = col.buf;
                res_n[ci] = col.len;
            }
            free(keep);
            free(clause);
            free(mw);
        }
        uint64_t total = 0;
        for (uint64_t ci = 0; ci < n_cand; ++ci) {
            uint64_t ch = cand[ci];
            uint64_t base = ch * m->chunk_size;
            uint64_t len = base + m->chunk_size <= m->n_rows ? m->chunk_size : m->n_rows - base;
            vectors_compared += len * q->nq; /* meta_compute.rs:166 */
            total += res_n[ci];
        }
        cand_t *agg = (cand_t *)malloc(sizeof(cand_t) * (total ? total : 1));
        uint64_t p = 0;
        for (uint64_t ci = 0; ci < n_cand; ++ci) {
            for (uint64_t i = 0; i < res_n[ci]; ++i) {
                agg[p] = res[ci][i];
                agg[p].seq = p;
                ++p;
            }
            free(res[ci]);
        }
        double score_s = now_s() - t0;
        /* (3) merge — meta.rs:699-709 */
        t0 = now_s();
        qsort(agg, total, sizeof(cand_t), q->take_type == ORA_TAKE_MIN ? merge_cmp_min : merge_cmp_max);
        if (total > q->k) total = q->k;
        double merge_s = now_s() - t0;
        n_out = emit(agg, total, 0, out_idx, out_score, out_qid, cap);
        free(agg);
        free(res);
        free(res_n);
        if (stats) {
            stats->score_s = score_s;
            stats->merge_s = merge_s;
        }
    } else {
        /* canonical: global selection over surviving rows (set-equal to per-chunk top-k + merge) */
        candvec_t *parts = (candvec_t *)calloc(n_cand ? n_cand : 1, sizeof(candvec_t));
#pragma omp parallel num_threads(nt)
        {
            uint8_t *keep = (uint8_t *)malloc(m->chunk_size);
            uint8_t *clause = (uint8_t *)malloc(m->chunk_size);
#pragma omp for schedule(dynamic, 1)
            for (long long ci = 0; ci < (long long)n_cand; ++ci) {
                uint64_t ch = cand[ci];
                uint64_t base = ch * m->chunk_size;
                uint64_t len = base + m->chunk_size <= m->n_rows ? m->chunk_size : m->n_rows - base;
                if (chunk_err || q->k == 0) continue;
                if (f) row_mask_chunk(m, f, base, len, keep, clause);
                else memset(keep, 1, len);
                ora_vec_query cq = *q;
                cq.row_mask_words = NULL;
                scan_canonical(m->vectors + base * m->dim, m->inv_norms + base, len, base, &cq, q_inv, keep, &parts[ci]);
            }
            free(keep);
            free(clause);
        }
        candvec_t all = {0, 0, 0};
        for (uint64_t ci = 0; ci < n_cand; ++ci) {
            uint64_t ch = cand[ci];
            uint64_t base = ch * m->chunk_size;
            uint64_t len = base + m->chunk_size <= m->n_rows ? m->chunk_size : m->n_rows - base;
            vectors_compared += len * q->nq;
            for (uint64_t i = 0; i < parts[ci].n; ++i) cv_push(&all, parts[ci].v[i].idx, parts[ci].v[i].score, parts[ci].v[i].qid);
            free(parts[ci].v);
        }
        free(parts);
        double score_s = now_s() - t0;
        t0 = now_s();
        n_out = finish_canonical(&all, q->take_type, q->k, out_idx, out_score, out_qid, cap);
        free(all.v);
        if (stats) {
            stats->score_s = score_s;
            stats->merge_s = now_s() - t0;
        }
    }
    free(q_inv);
    free(cand);
    free(ck);
    if (stats) { /* meta.rs:711-721 */
        stats->total_chunks = nc;
        stats->evaluated_chunks = n_cand;
        stats->pruned_chunks = nc - n_cand;
        stats->vectors_compared = vectors_compared;
        stats->prune_s = prune_s;
        stats->total_s = now_s() - t_total;
    }
    return n_out;
}

/* ------------------------------------------------------------------------- */
/* synthetic generator                                        */
/* ------------------------------------------------------------------------- */

static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ULL;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ULL;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBULL;
    return x ^ (x >> 31);
}

void oracle_synth_fill(float *out, uint64_t row0, uint64_t n_rows, uint32_t dim, uint64_t seed) {
#pragma omp parallel for schedule(static)
    for (long long r = 0; r < (long long)n_rows; ++r) {
        uint64_t row = row0 + (uint64_t)r;
        float *o = out + (size_t)r * dim;
        for (uint32_t c = 0; c < dim; ++c) {
            uint64_t u = splitmix64(seed ^ (row * (uint64_t)dim + c));
            o[c] = (float)(u >> 40) * (1.0f / 8388608.0f) - 1.0f;
        }
    }
}
