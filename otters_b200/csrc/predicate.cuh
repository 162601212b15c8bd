// predicate.cuh — per-row evaluation of one lowered filter leaf, shared by the row-mask kernel (meta.cu)
// and the fused predicate stage of the scan kernel (scan.cu).
//
// Semantics follow the reference's row kernels (src/type_utils.rs:306-444,586-736) and leaf helpers
// (src/meta_compute.rs:235-318): IEEE compares (NaN satisfies only Neq), a NULL row fails every leaf
// including Neq, strings compare by byte equality (here: dictionary-code equality).
#pragma once
#include "internal.h"

namespace otters {

template <typename T>
__device__ __forceinline__ bool row_sat(int op, T v, T t) {
    switch (op) {
    case OTTERS_OP_EQ: return v == t;
    case OTTERS_OP_NEQ: return v != t;
    case OTTERS_OP_LT: return v < t;
    case OTTERS_OP_LTE: return v <= t;
    case OTTERS_OP_GT: return v > t;
    default: return v >= t;
    }
}

__device__ __forceinline__ bool row_leaf_sat(const DevLeaf& lf, uint32_t row) {
    const bool is_null = lf.null_words && ((lf.null_words[row >> 5] >> (row & 31)) & 1u);
    bool sat;
    switch (lf.exec) {
    case LEAF_I32: sat = row_sat<int32_t>(lf.op, ((const int32_t*)lf.values)[row], lf.i32); break;
    case LEAF_I64: sat = row_sat<int64_t>(lf.op, ((const int64_t*)lf.values)[row], lf.i64); break;
    case LEAF_F32: sat = row_sat<float>(lf.op, ((const float*)lf.values)[row], lf.f32); break;
    case LEAF_F64: sat = row_sat<double>(lf.op, ((const double*)lf.values)[row], lf.f64); break;
    default: {  // dictionary-coded string equality (src/meta_compute.rs:291-318)
        const bool eq = lf.code_valid && ((const uint32_t*)lf.values)[row] == lf.code;
        sat = lf.op == OTTERS_OP_EQ ? eq : (lf.op == OTTERS_OP_NEQ ? !eq : false);
    }
    }
    return sat && !is_null;
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

}  // namespace otters
