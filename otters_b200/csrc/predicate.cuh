// predicate.cuh — per-row evaluation of one lowered filter leaf, shared by the row-mask kernel (meta.cu)
// and the fused predicate stage of the scan kernel (scan_kernel.cuh).
//
// Semantics follow the reference's row kernels (src/type_utils.rs:306-444,586-736) and leaf helpers
// (src/meta_compute.rs:235-318): IEEE compares (NaN satisfies only Neq), a NULL row fails every leaf
// including Neq, strings compare by byte equality (here: dictionary-code equality).
#pragma once
#include "internal.h"

namespace otters {

// A compare is reduced to one of four STATES of (value, literal) — 0: value > literal, 1: value < literal, 2: equal,
// 3: unordered (a NaN on either side) — and the leaf carries a 4-bit truth table indexed by the state (DevLeaf::tt, set when
// the filter is lowered: Eq 0100, Neq 1011, Lt 0010, Lte 0110, Gt 0001, Gte 0101).  That is a handful of straight-line
// instructions per (row, leaf) where a switch over the operator cost ~60 with its branches: the predicate kernels were
// bound by instruction issue, not by memory.  IEEE semantics are kept: NaN satisfies only Neq (src/type_utils.rs:306-444).
template <typename T>
__device__ __forceinline__ uint32_t cmp_state(T v, T t) {
    const uint32_t lt = v < t ? 1u : 0u, eq = v == t ? 2u : 0u, gt = v > t ? 4u : 0u;
    return (lt | eq | gt) ? (lt | eq) : 3u;  // integers are never unordered
}
__device__ __forceinline__ bool tt_sat(uint32_t tt, uint32_t state) { return ((tt >> state) & 1u) != 0; }

template <typename T>
__device__ __forceinline__ bool row_sat(int op, T v, T t) {  // reference form, kept for the chunk rules' documentation
    switch (op) {
    case OTTERS_OP_EQ: return v == t;
    case OTTERS_OP_NEQ: return v != t;
    case OTTERS_OP_LT: return v < t;
    case OTTERS_OP_LTE: return v <= t;
    case OTTERS_OP_GT: return v > t;
    default: return v >= t;
    }
}

// state of the 64-bit raw value `raw` of a row under leaf lf (dictionary strings: code equality, src/meta_compute.rs:291-318)
__device__ __forceinline__ uint32_t leaf_state_raw(const DevLeaf& lf, uint64_t raw) {
    switch (lf.exec) {
    case LEAF_I32: return cmp_state<int32_t>((int32_t)(uint32_t)raw, lf.i32);
    case LEAF_I64: return cmp_state<int64_t>((int64_t)raw, lf.i64);
    case LEAF_F32: return cmp_state<float>(__uint_as_float((uint32_t)raw), lf.f32);
    case LEAF_F64: return cmp_state<double>(__longlong_as_double((long long)raw), lf.f64);
    default: return (lf.code_valid && (uint32_t)raw == lf.code) ? 2u : 0u;
    }
}

__device__ __forceinline__ bool row_leaf_sat(const DevLeaf& lf, uint32_t row) {
    const bool is_null = lf.null_words && ((lf.null_words[row >> 5] >> (row & 31)) & 1u);
    const bool wide = lf.exec == LEAF_I64 || lf.exec == LEAF_F64;
    const uint64_t raw = wide ? reinterpret_cast<const unsigned long long*>(lf.values)[row]
                              : (uint64_t)reinterpret_cast<const uint32_t*>(lf.values)[row];
    return tt_sat(lf.tt, leaf_state_raw(lf, raw)) && !is_null;
}

// CNF over one row: AND over clauses of OR over leaves; every leaf is evaluated so that its loads overlap
__device__ __forceinline__ bool row_passes(const DevLeaf* leaves, const uint32_t* clause_off, uint32_t n_clauses, uint32_t row) {
    bool keep = true;
    for (uint32_t ci = 0; ci < n_clauses; ++ci) {
        bool any = false;
        for (uint32_t li = clause_off[ci]; li < clause_off[ci + 1]; ++li) any |= row_leaf_sat(leaves[li], row);
        keep &= any;
    }
    return keep;
}

// --- chunk-level rules: zonemap ranges + Bloom filters (reference src/meta.rs:407-544, src/type_utils.rs:446-584,739-889) ---------
// Shared by the stand-alone prune kernels (meta.cu) and the lazy per-unit pruning inside the scan kernels.
__device__ __forceinline__ bool chunk_leaf_rule(const DevLeaf& lf, uint32_t ch);

template <typename T>
__device__ __forceinline__ bool range_sat(int op, T mn, T mx, T t) {
    switch (op) {
    case OTTERS_OP_EQ: return mn <= t && t <= mx;
    case OTTERS_OP_LT: return mn < t;
    case OTTERS_OP_LTE: return mn <= t;
    case OTTERS_OP_GT: return mx > t;
    case OTTERS_OP_GTE: return mx >= t;
    default: return true;  // Neq
    }
}

__device__ __forceinline__ bool chunk_leaf_sat(const DevLeaf& lf, uint32_t ch) {
    const bool has_rows = __ldg(lf.non_null + ch) != 0;  // every rule is ANDed with non_null > 0
    return chunk_leaf_rule(lf, ch) && has_rows;
}

__device__ __forceinline__ bool chunk_leaf_rule(const DevLeaf& lf, uint32_t ch) {
    switch (lf.exec) {
    case LEAF_I32: return range_sat<int32_t>(lf.op, ((const int32_t*)lf.zmin)[ch], ((const int32_t*)lf.zmax)[ch], lf.i32);
    case LEAF_I64: return range_sat<int64_t>(lf.op, ((const int64_t*)lf.zmin)[ch], ((const int64_t*)lf.zmax)[ch], lf.i64);
    case LEAF_F32: return range_sat<float>(lf.op, ((const float*)lf.zmin)[ch], ((const float*)lf.zmax)[ch], lf.f32);
    case LEAF_F64: return range_sat<double>(lf.op, ((const double*)lf.zmin)[ch], ((const double*)lf.zmax)[ch], lf.f64);
    default: {  // LEAF_STR — src/meta.rs:523-544
        if (lf.op == OTTERS_OP_NEQ) return true;
        if (lf.op != OTTERS_OP_EQ) return false;
        const uint64_t* w = lf.bloom + (size_t)ch * lf.bloom_stride;
        // every full chunk has the same filter geometry: only the (shorter) last chunk loads its own
        const bool full = ch < lf.bloom_full_chunks;
        const uint64_t m = full ? lf.bloom_m0 : lf.bloom_mbits[ch];
        const uint32_t kh = full ? lf.bloom_k0 : lf.bloom_k[ch];
        // probe i = (a + i*b) mod m, a = h1 mod m, b = h2 mod m (1 if 0), stepped without a division per probe; a and b
        // come precomputed for the filter size of a full chunk
        uint64_t bit, step;
        if (m == lf.bloom_m0) {
            bit = lf.bloom_a0;
            step = lf.bloom_b0;
        } else {
            bit = lf.h1 % m;
            step = lf.h2 % m;
            if (step == 0) step = 1;
        }
        bool all = true;  // every probe is issued (no early exit) so that the word loads overlap
        for (uint32_t i = 0; i < kh; ++i) {
            all &= ((__ldg(w + (bit >> 6)) >> (bit & 63)) & 1ull) != 0;
            bit += step;
            if (bit >= m) bit -= m;
        }
        return all;
    }
    }
}

// CNF over one chunk: AND over clauses of OR over leaves; every leaf is evaluated so that its zonemap / Bloom loads overlap
__device__ __forceinline__ bool chunk_passes(const DevLeaf* leaves, const uint32_t* clause_off, uint32_t n_clauses, uint32_t ch) {
    bool keep = true;
    for (uint32_t ci = 0; ci < n_clauses; ++ci) {
        bool any = false;
        for (uint32_t li = clause_off[ci]; li < clause_off[ci + 1]; ++li) any |= chunk_leaf_sat(leaves[li], ch);
        keep &= any;
    }
    return keep;
}

// --- memory-level-parallel form for the planner warps of K1 (scan_planner.cu) ------------------------------------
// A lane owns up to 4 consecutive rows r .. r+rpl-1 (all inside one 32-bit null word; bit j of `bits` = row r+j is
// still a candidate).  Phase 1 issues EVERY value and null-word load of every (leaf, row) pair, phase 2 compares, so
// a unit's metadata arrives in one memory round trip instead of one per leaf.  Same semantics as row_passes().
constexpr uint32_t kMlpLeaves = 6;

__device__ __forceinline__ bool leaf_sat_raw(const DevLeaf& lf, uint64_t raw) { return tt_sat(lf.tt, leaf_state_raw(lf, raw)); }

__device__ __forceinline__ uint32_t rows_pass_mlp(const DevLeaf* leaves, const uint32_t* clause_off, uint32_t n_clauses,
                                                  uint32_t n_leaves, uint32_t r, uint32_t bits, uint32_t rpl) {
    if (n_clauses == 0 || bits == 0) return bits;
    if (n_leaves > kMlpLeaves) {  // large CNFs: row by row
        uint32_t out = 0;
        for (uint32_t j = 0; j < rpl; ++j)
            if (((bits >> j) & 1u) && row_passes(leaves, clause_off, n_clauses, r + j)) out |= 1u << j;
        return out;
    }
    uint64_t raw[kMlpLeaves][4];
    uint32_t nul[kMlpLeaves];
#pragma unroll
    for (uint32_t li = 0; li < kMlpLeaves; ++li) {
        nul[li] = 0;
#pragma unroll
        for (uint32_t j = 0; j < 4; ++j) raw[li][j] = 0;
        if (li < n_leaves) {
            const DevLeaf& lf = leaves[li];
            if (lf.null_words) nul[li] = __ldg(lf.null_words + (r >> 5)) >> (r & 31);
            const bool wide = lf.exec == LEAF_I64 || lf.exec == LEAF_F64;
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j)
                if (j < rpl && ((bits >> j) & 1u))
                    raw[li][j] = wide ? __ldg(reinterpret_cast<const unsigned long long*>(lf.values) + r + j)
                                      : (uint64_t)__ldg(reinterpret_cast<const uint32_t*>(lf.values) + r + j);
        }
    }
    uint32_t keep = bits, any = 0, ci = 0;
    uint32_t next = clause_off[1];  // first leaf of the next clause
#pragma unroll
    for (uint32_t li = 0; li < kMlpLeaves; ++li) {
        if (li < n_leaves) {
            while (li == next && ci + 1 < n_clauses) {  // clause ci is complete (also steps over empty clauses)
                keep &= any;
                any = 0;
                ++ci;
                next = clause_off[ci + 1];
            }
            const DevLeaf& lf = leaves[li];
            uint32_t m = 0;
#pragma unroll
            for (uint32_t j = 0; j < 4; ++j)
                if (j < rpl && leaf_sat_raw(lf, raw[li][j]) && !((nul[li] >> j) & 1u)) m |= 1u << j;
            any |= m;
        }
    }
    keep &= any;  // the clause the last leaf belongs to
    for (++ci; ci < n_clauses; ++ci) keep = 0;  // trailing clauses without leaves can never be satisfied
    return keep;
}


// --- leaf-major form for the row-mask kernel (meta.cu) -------------------------------------------------------------
// A lane owns row `lane` of four consecutive 32-row words starting at row `r0` (a multiple of 128): for every leaf the
// four values are loaded together (coalesced: 32 lanes read 32 consecutive values; the null words are warp-uniform), the
// type dispatch happens once per leaf, and the four rows' clause state advances side by side.  live = 4-bit mask of the
// lane's rows that are still candidates (chunk kept, inside the store); returns the 4-bit mask of rows passing the CNF.
__device__ __forceinline__ uint32_t rows_pass_x4(const DevLeaf* leaves, const uint32_t* clause_off, uint32_t n_clauses, uint32_t r0,
                                                 uint32_t lane, uint32_t live) {
    uint32_t keep = live;
    for (uint32_t ci = 0; ci < n_clauses; ++ci) {
        uint32_t any = 0;
        for (uint32_t li = clause_off[ci]; li < clause_off[ci + 1]; ++li) {
            const DevLeaf& lf = leaves[li];
            const bool wide = lf.exec == LEAF_I64 || lf.exec == LEAF_F64;
            uint64_t raw[4];
            uint32_t nul[4];
#pragma unroll
            for (uint32_t u = 0; u < 4; ++u) {
                const uint32_t row = r0 + 32u * u + lane;
                raw[u] = 0;
                nul[u] = 0;
                if ((live >> u) & 1u) {
                    raw[u] = wide ? __ldg(reinterpret_cast<const unsigned long long*>(lf.values) + row)
                                  : (uint64_t)__ldg(reinterpret_cast<const uint32_t*>(lf.values) + row);
                    if (lf.null_words) nul[u] = __ldg(lf.null_words + (row >> 5));
                }
            }
            uint32_t m = 0;
            switch (lf.exec) {  // one dispatch per leaf, four rows each
#define OTTERS_X4(expr)                                                          \
    _Pragma("unroll") for (uint32_t u = 0; u < 4; ++u) {                        \
        const uint32_t st = (expr);                                              \
        m |= (((lf.tt >> st) & 1u) & ~(nul[u] >> lane) & 1u) << u;               \
    }
            case LEAF_I32: OTTERS_X4(cmp_state<int32_t>((int32_t)(uint32_t)raw[u], lf.i32)) break;
            case LEAF_I64: OTTERS_X4(cmp_state<int64_t>((int64_t)raw[u], lf.i64)) break;
            case LEAF_F32: OTTERS_X4(cmp_state<float>(__uint_as_float((uint32_t)raw[u]), lf.f32)) break;
            case LEAF_F64: OTTERS_X4(cmp_state<double>(__longlong_as_double((long long)raw[u]), lf.f64)) break;
            default: OTTERS_X4((lf.code_valid && (uint32_t)raw[u] == lf.code) ? 2u : 0u) break;
#undef OTTERS_X4
            }
            any |= m;
        }
        keep &= any;
    }
    return keep;
}

}  // namespace otters
