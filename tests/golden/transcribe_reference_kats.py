#!/usr/bin/env python3
"""Writes tests/golden/reference_kats.json: the known-answer tests of the reference's own test-suite
for the exact-search hot path, transcribed by hand (the Rust crate cannot be compiled in this image,
so these are the only golden vectors the reference offers — SURVEY.md §8c).

Every entry cites the reference test it was transcribed from (path:lines under /root/reference).
Run:  python tests/golden/transcribe_reference_kats.py
"""
import json
import os

V5 = [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [1.0, 1.0, 0.0], [0.5, 0.5, 0.5]]  # create_test_vectors :5-13

VEC = []


def vec(name, cite, dim, vectors, queries, metric, calls, expect, plan_new=False):
    VEC.append(dict(name=name, cite=f"tests/vec_store_tests.rs:{cite}", dim=dim, vectors=vectors, queries=queries,
                    metric=metric, calls=calls, expect=expect, plan_new=plan_new))


DIM_ERR_23 = "Query vector length 2 does not match expected dimension 3"

vec("test_query_plan_creation/single", "36-45", 3, [], [1.0, 0.0, 0.0], "Cosine", [], {"len": 0})
vec("test_query_plan_creation/multi", "36-45", 3, [], [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]], "Cosine", [], {"len": 0})
vec("test_dimension_mismatch_error_handling", "51-63", 3, [[1.0, 0.0, 0.0]], [1.0, 0.0], "Cosine", [["take", 5]],
    {"error_contains": DIM_ERR_23})
vec("test_empty_query_batch_error_handling", "65-76", 3, [], [], "Cosine", [["take", 5]], {"error_equals": "No queries provided"})
vec("test_error_propagation_through_chain", "78-94", 3, [], [1.0, 0.0], "Cosine",
    [["filter", 0.5, "Gt"], ["take", 5], ["take_min", 3]], {"error_contains": DIM_ERR_23})
vec("test_successful_chain_after_valid_query", "96-119", 2, [[1.0, 0.0], [0.8, 0.6], [0.0, 1.0]], [1.0, 0.0], "Cosine",
    [["filter", 0.5, "Gt"], ["take", 5]], {"all_scores": ["Gt", 0.5], "indices_set": [0, 1]})
vec("test_mixed_dimension_batch_error", "121-139", 3, [[1.0, 0.0, 0.0]], [[1.0, 0.0, 0.0], [1.0, 0.0], [1.0, 0.0, 0.0]],
    "Cosine", [["take", 5]], {"error_contains": DIM_ERR_23})
vec("test_cosine_similarity_basic", "145-161", 3, V5, [1.0, 0.0, 0.0], "Cosine", [["take", 5]],
    {"len": 5, "score_by_index": {"0": 1.0}, "tol": 1e-6})
vec("test_cosine_orthogonal_vectors", "163-181", 2, [[1.0, 0.0], [0.0, 1.0]], [1.0, 0.0], "Cosine", [["take", 2]],
    {"len": 2, "score_by_index": {"0": 1.0, "1": 0.0}, "tol": 1e-6})
vec("test_euclidean_distance_basic", "187-201", 3, V5, [1.0, 0.0, 0.0], "Euclidean", [["take_min", 5]],
    {"score_by_index": {"0": 0.0}, "tol": 1e-6})
vec("test_dot_product_basic", "207-221", 3, V5, [1.0, 0.0, 0.0], "DotProduct", [["take", 5]],
    {"score_by_index": {"0": 1.0}, "tol": 1e-6})
vec("test_dot_product_orthogonal_vectors", "223-249", 2, [[1.0, 0.0], [0.0, 1.0], [2.0, 0.0], [-1.0, 0.0]], [1.0, 0.0],
    "DotProduct", [["take", 4]], {"len": 4, "score_by_index": {"0": 1.0, "1": 0.0, "2": 2.0, "3": -1.0}, "tol": 1e-6})
vec("test_dot_product_ranking", "251-276", 2, [[3.0, 4.0], [1.0, 1.0], [0.0, 1.0], [-1.0, 0.0]], [3.0, 4.0], "DotProduct",
    [["take", 4]], {"len": 4, "sorted": "desc", "scores": [25.0, 7.0, 4.0, -3.0], "tol": 1e-6})
vec("test_dot_product_filtering", "278-298", 2, [[2.0, 0.0], [1.0, 0.0], [0.5, 0.0], [-1.0, 0.0]], [1.0, 0.0], "DotProduct",
    [["filter", 1.0, "Gt"], ["take", 10]], {"len": 1, "scores": [2.0], "tol": 1e-6})
vec("test_dot_product_take_max", "300-321", 2, [[1.0, 0.0], [2.0, 0.0], [0.5, 0.0], [-1.0, 0.0]], [1.0, 0.0], "DotProduct",
    [["take_max", 2]], {"len": 2, "scores": [2.0, 1.0], "tol": 1e-6})
vec("test_dot_product_take_min", "323-343", 2, [[1.0, 0.0], [2.0, 0.0], [0.5, 0.0], [-1.0, 0.0]], [1.0, 0.0], "DotProduct",
    [["take_min", 2]], {"len": 2, "scores": [-1.0, 0.5], "tol": 1e-6})
vec("test_dot_product_batch_queries", "345-359", 2, [[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]], [[1.0, 0.0], [0.0, 1.0]],
    "DotProduct", [["take", 3]], {"len": 3})
vec("test_top_k_cosine", "365-386", 2, [[1.0, 0.0], [0.8, 0.6], [0.0, 1.0], [-1.0, 0.0]], [1.0, 0.0], "Cosine", [["take", 2]],
    {"len": 2, "sorted": "desc"})
vec("test_top_k_euclidean", "388-409", 2, [[1.0, 0.0], [1.1, 0.0], [0.0, 1.0], [-1.0, 0.0]], [1.0, 0.0], "Euclidean",
    [["take_min", 2]], {"len": 2, "sorted": "asc"})
vec("test_take_more_than_available", "411-428", 2, [[1.0, 0.0], [0.0, 1.0]], [1.0, 0.0], "Cosine", [["take", 10]], {"len": 2})
vec("test_take_zero_results", "430-445", 2, [[1.0, 0.0], [0.0, 1.0]], [1.0, 0.0], "Cosine", [["take", 0]], {"len": 0})
vec("test_filtering", "451-472", 2, [[1.0, 0.0], [0.8, 0.6], [0.0, 1.0], [-1.0, 0.0]], [1.0, 0.0], "Cosine",
    [["filter", 0.5, "Gt"], ["take", 10]], {"all_scores": ["Gt", 0.5], "indices_set": [0, 1]})
vec("test_empty_store", "488-499", 3, [], [1.0, 0.0, 0.0], "Cosine", [["take", 5]], {"len": 0})
vec("test_cosine_similarity_correctness", "544-608", 2, [[1.0, 0.0], [-1.0, 0.0], [0.0, 1.0], [1.0, 1.0]], [1.0, 0.0], "Cosine",
    [["take", 4]], {"len": 4, "score_by_index": {"0": 1.0, "1": -1.0, "2": 0.0, "3": 0.7071067811865475}, "tol": 1e-5})
vec("test_euclidean_distance_correctness", "610-656", 2, [[0.0, 0.0], [3.0, 4.0], [1.0, 1.0], [0.0, 5.0], [-3.0, -4.0]],
    [0.0, 0.0], "Euclidean", [["take_min", 5]],
    {"score_by_index": {"0": 0.0, "1": 25.0, "2": 2.0, "3": 25.0, "4": 25.0}, "tol": 1e-6})
vec("test_dot_product_correctness", "658-745", 3,
    [[2.0, 3.0, 1.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [-1.0, 0.0, 0.0], [1.0, 1.0, 1.0]], [2.0, 3.0, 1.0],
    "DotProduct", [["take", 6]],
    {"len": 6, "sorted": "desc", "score_by_index": {"0": 14.0, "1": 2.0, "2": 3.0, "3": 1.0, "4": -2.0, "5": 6.0}, "tol": 1e-6})
vec("test_top_k_ranking_correctness", "747-798", 2, [[1.0, 0.0], [0.8, 0.6], [0.6, 0.8], [0.0, 1.0]], [1.0, 0.0], "Cosine",
    [["take", 4]], {"scores": [1.0, 0.8, 0.6, 0.0], "sorted": "desc", "tol": 1e-6})
vec("test_euclidean_ranking_correctness", "800-851", 2, [[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0], [2.0, 0.0], [3.0, 4.0]],
    [0.0, 0.0], "Euclidean", [["take_min", 6]], {"scores": [0.0, 1.0, 1.0, 2.0, 4.0, 25.0], "sorted": "asc", "tol": 1e-6})
FT = [[1.0, 0.0], [0.8, 0.6], [0.6, 0.8], [0.0, 1.0], [-0.6, 0.8]]
vec("test_filter_threshold_correctness/gt0.7", "853-896", 2, FT, [1.0, 0.0], "Cosine", [["filter", 0.7, "Gt"], ["take", 10]],
    {"all_scores": ["Gt", 0.7], "indices_set": [0, 1]})
vec("test_filter_threshold_correctness/gte0.6", "853-896", 2, FT, [1.0, 0.0], "Cosine", [["filter", 0.6, "Gte"], ["take", 10]],
    {"all_scores": ["Gte", 0.6]})
vec("test_filter_threshold_correctness/lt0.5", "853-896", 2, FT, [1.0, 0.0], "Cosine", [["filter", 0.5, "Lt"], ["take", 10]],
    {"all_scores": ["Lt", 0.5], "indices_set": [3, 4]})
vec("test_batch_query_correctness", "898-924", 2, [[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0]], [[1.0, 0.0], [0.0, 1.0]], "Cosine",
    [["take", 2]], {"count_score": [1.0, 2], "tol": 1e-6})
vec("test_api_design_showcase", "930-958", 3, [[i / 100.0, (i * 2) / 100.0, (i * 3) / 100.0] for i in range(100)],
    [0.5, 0.5, 0.5], "Cosine", [["filter", 0.8, "Gt"], ["take_min", 10]], {"len": 10, "all_scores": ["Gt", 0.8], "sorted": "asc"})
vec("test_error_in_chain_stops_execution", "960-980", 3, [], [1.0, 0.0], "Cosine",
    [["filter", 0.5, "Gt"], ["take", 10], ["take_min", 5]], {"error_contains": DIM_ERR_23})
vec("test_vec_query_plan_new", "986-996", 0, [], None, None, [], {"error_contains": "Query vectors or their norms are not set"},
    plan_new=True)
vec("test_error_propagation_in_filter", "998-1008", 0, [], None, None, [["filter", 0.5, "Gt"]], {"error": True}, plan_new=True)
vec("test_error_propagation_in_take_methods/take", "1010-1019", 0, [], None, None, [["take", 5]], {"error": True}, plan_new=True)
vec("test_error_propagation_in_take_methods/take_min", "1010-1019", 0, [], None, None, [["take_min", 5]], {"error": True},
    plan_new=True)
vec("test_error_propagation_in_take_methods/take_max", "1010-1019", 0, [], None, None, [["take_max", 5]], {"error": True},
    plan_new=True)
vec("test_empty_query_vectors_in_batch", "1021-1029", 3, [], [], "Cosine", [], {"error_contains": "No queries provided"})
OPS = [[1.0, 0.0], [0.0, 1.0], [0.5, 0.5], [0.8, 0.6]]
for thr, op in [(0.9, "Lt"), (0.1, "Gt"), (1.0, "Lte"), (0.0, "Gte"), (1.0, "Eq")]:
    vec(f"test_filter_with_all_comparison_operators/{op}", "1031-1090", 2, OPS, [1.0, 0.0], "Cosine",
        [["filter", thr, op], ["take", 10]], {"nonempty": True, "all_scores": [op, thr]})
vec("test_add_vector_with_zero_norm", "1092-1109", 3, [[0.0, 0.0, 0.0]], [1.0, 0.0, 0.0], "Cosine", [["take", 1]],
    {"len": 1, "scores": [0.0], "tol": 0.0})
vec("test_query_with_zero_norm_query_vector", "1111-1124", 3, [[1.0, 0.0, 0.0]], [0.0, 0.0, 0.0], "Cosine", [["take", 1]],
    {"len": 1, "scores": [0.0], "tol": 0.0})
vec("test_filter_and_merge_with_no_filtering", "1128-1142", 2, [[1.0, 0.0], [0.0, 1.0], [0.5, 0.5]], [1.0, 0.0], "Cosine",
    [["take", 2]], {"len": 2})
CF = [[1.0, 0.0], [0.0, 1.0], [0.9, 0.1]]
vec("test_take_closest_and_farthest_methods/min2", "1162-1206", 2, CF, [1.0, 0.0], "Euclidean", [["take_min", 2]],
    {"len": 2, "indices": [0, 2]})
vec("test_take_closest_and_farthest_methods/max2", "1162-1206", 2, CF, [1.0, 0.0], "Euclidean", [["take_max", 2]],
    {"len": 2, "indices": [1, 2]})
vec("test_take_closest_and_farthest_methods/batch_min1", "1162-1206", 2, CF, [[1.0, 0.0], [0.0, 1.0]], "Euclidean",
    [["take_min", 1]], {"len": 1, "scores": [0.0], "tol": 0.0})
vec("test_take_closest_and_farthest_methods/batch_max1", "1162-1206", 2, CF, [[1.0, 0.0], [0.0, 1.0]], "Euclidean",
    [["take_max", 1]], {"len": 1, "scores": [2.0], "tol": 1e-6})
vec("test_query_batch_conversions/single", "1208-1231", 3, [[1.0, 0.0, 0.0]], [1.0, 0.0, 0.0], "Cosine", [["take", 1]], {"len": 1})
vec("test_query_batch_conversions/multi", "1208-1231", 3, [[1.0, 0.0, 0.0]], [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0]], "Cosine",
    [["take", 2]], {"max_len": 2})
vec("test_error_states_in_chained_operations", "1233-1252", 3, [[1.0, 0.0, 0.0]], [1.0, 0.0], "Cosine",
    [["filter", 0.5, "Gt"], ["take", 5], ["take_min", 2], ["take_max", 1]], {"error_contains": "does not match expected dimension"})
EC = [[1.0, 0.0], [0.0, 1.0], [-1.0, 0.0]]
vec("test_filtering_edge_cases/empty", "1254-1284", 2, EC, [1.0, 0.0], "Cosine", [["filter", 1.5, "Gt"], ["take", 10]], {"len": 0})
vec("test_filtering_edge_cases/eq", "1254-1284", 2, EC, [1.0, 0.0], "Cosine", [["filter", 1.0, "Eq"], ["take", 10]],
    {"len": 1, "indices": [0]})

# scoring functions (tests/vec_store_tests.rs:505-538)
FUNCS = [
    dict(name="test_dot_product", cite="tests/vec_store_tests.rs:505-515", fn="dot", a=[1.0, 2.0, 3.0, 4.0], b=[2.0, 3.0, 4.0, 5.0],
         expect=40.0, tol=0.0),
    dict(name="test_euclidean_distance_squared", cite="tests/vec_store_tests.rs:517-527", fn="l2", a=[1.0, 2.0], b=[4.0, 6.0],
         expect=25.0, tol=0.0),
    dict(name="test_cosine_similarity", cite="tests/vec_store_tests.rs:529-538", fn="cosine", a=[1.0, 0.0], b=[1.0, 0.0],
         a_inv=1.0, b_inv=1.0, expect=1.0, tol=1e-6),
]

ADD_ERRORS = [
    dict(name="test_vecstore_creation", cite="tests/vec_store_tests.rs:19-26", dim=3, rows=[[1.0, 2.0, 3.0], [1.0, 2.0]],
         error_contains="Input vector length 2 does not match expected dimension 3", ok_rows=1),
    dict(name="test_dimension_mismatch_during_add_vectors", cite="tests/vec_store_tests.rs:1144-1160", dim=3,
         rows=[[1.0, 0.0, 0.0], [1.0, 0.0]], error_contains="Input vector length 2 does not match expected dimension 3", ok_rows=1),
]

# ---- MetaStore KATs (tests/meta_tests.rs, tests/meta_zonemap_tests.rs) ----------------------------------
META = []


def meta(name, cite, vectors, columns, chunk_size, queries, metric, expr, vec_filter, take, expect):
    META.append(dict(name=name, cite=cite, vectors=vectors, columns=columns, chunk_size=chunk_size, queries=queries,
                     metric=metric, expr=expr, vec_filter=vec_filter, take=take, expect=expect))


def cmp_(c, op, v):
    return ["cmp", c, op, v]


meta("meta_basic_pruning_and_stats", "tests/meta_tests.rs:4-39",
     [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.5, 0.5, 0.0], [0.0, 0.0, 1.0]],
     [["age", "Int32", [10, 20, 30, None]], ["grade", "String", ["A", "B", "A", "C"]]], 2, [1.0, 0.0, 0.0], "Cosine",
     ["and", cmp_("age", "gt", 15), cmp_("grade", "eq", "A")], None, 4,
     {"indices_set": [2], "stats": {"total_chunks": 2}, "stats_ge": {"evaluated_chunks": 1}})
meta("meta_string_eq_prunes_chunks", "tests/meta_tests.rs:41-88",
     [[1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [1.0, 1.0, 0.0], [0.0, 0.0, 1.0], [1.0, 0.0, 1.0], [0.5, 0.5, 0.0]],
     [["age", "Int32", [10, 11, 12, 20, 21, 22]], ["grade", "String", ["B", "C", "B+", "A", "A", "C"]]], 3, [1.0, 0.0, 0.0],
     "Cosine", cmp_("grade", "eq", "A"), None, 6,
     {"indices_set": [3, 4], "stats": {"total_chunks": 2}, "stats_ge": {"pruned_chunks": 1}})
meta("meta_datetime_range_filter", "tests/meta_tests.rs:90-119", [[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]],
     [["ts", "DateTime", ["2023-01-01T00:00:00Z", "2023-06-01T00:00:00Z", "2024-01-01T00:00:00Z"]]], 2, [1.0, 0.0], "DotProduct",
     ["and", cmp_("ts", "gte", "2023-01-01T00:00:00Z"), cmp_("ts", "lt", "2024-01-01T00:00:00Z")], None, 3,
     {"indices_set": [0, 1]})
meta("meta_global_scope_merge_and_vec_threshold", "tests/meta_tests.rs:121-153",
     [[1.0, 0.0], [0.0, 1.0], [1.0, 1.0], [2.0, 0.0]], [["grade", "String", ["A", "B", "A", "A"]]], 2,
     [[1.0, 0.0], [0.0, 1.0]], "DotProduct", cmp_("grade", "eq", "A"), [0.5, "Gt"], 2,
     {"max_len": 2, "scores": [2.0, 1.0], "tol": 1e-6, "stats_le_total": True})
meta("meta_stats_without_meta_filter", "tests/meta_tests.rs:168-184", [[1.0, 0.0], [0.0, 1.0], [1.0, 1.0]], [], 2, [1.0, 0.0],
     "Cosine", None, None, 3, {"len": 3, "stats": {"total_chunks": 2, "evaluated_chunks": 2, "pruned_chunks": 0, "vectors_compared": 3}})

# build_store() of tests/meta_zonemap_tests.rs:17-67
ZV = [[1.0, 0.0] for _ in range(9)]
ZC = [
    ["val", "Int32", [1, 2, None, 10, 11, 12, None, None, None]],
    ["ts", "DateTime", ["2024-01-01T00:00:00Z", None, "2024-06-01T00:00:00Z", "2026-01-01T00:00:00Z", "2026-06-01T00:00:00Z",
                        "2024-12-31T23:59:59Z", None, None, None]],
    ["grade", "String", ["A", "B", None, "C", "A", "A", None, None, None]],
]
meta("zonemap_prunes_numeric_with_nulls", "tests/meta_zonemap_tests.rs:69-89", ZV, ZC, 3, [1.0, 0.0], "DotProduct",
     cmp_("val", "gt", 5), None, 9,
     {"indices_set": [3, 4, 5], "stats": {"total_chunks": 3, "evaluated_chunks": 1, "pruned_chunks": 2}})
meta("zonemap_boundary_conditions/gte2", "tests/meta_zonemap_tests.rs:91-116", ZV, ZC, 3, [1.0, 0.0], "Cosine",
     cmp_("val", "gte", 2), None, 9, {"indices_set": [1, 3, 4, 5], "stats": {"total_chunks": 3, "pruned_chunks": 1}})
meta("zonemap_boundary_conditions/gt2", "tests/meta_zonemap_tests.rs:91-116", ZV, ZC, 3, [1.0, 0.0], "Cosine",
     cmp_("val", "gt", 2), None, 9, {"indices_set": [3, 4, 5], "stats": {"evaluated_chunks": 1, "pruned_chunks": 2}})
meta("zonemap_all_null_chunk_pruned_for_equality", "tests/meta_zonemap_tests.rs:118-131", ZV, ZC, 3, [1.0, 0.0], "Cosine",
     cmp_("grade", "eq", "A"), None, 9, {"indices_set": [0, 4, 5], "stats": {"total_chunks": 3}, "stats_ge": {"pruned_chunks": 1}})
meta("zonemap_and_clause_numeric_datetime", "tests/meta_zonemap_tests.rs:133-156", ZV, ZC, 3, [1.0, 0.0], "DotProduct",
     ["and", cmp_("val", "gt", 5), cmp_("ts", "lt", "2025-01-01T00:00:00Z")], None, 9,
     {"len": 1, "indices": [5], "stats": {"total_chunks": 3, "evaluated_chunks": 1, "pruned_chunks": 2}})
meta("zonemap_ne_comparator_with_null_only_chunk", "tests/meta_zonemap_tests.rs:158-174", ZV, ZC, 3, [1.0, 0.0], "Cosine",
     cmp_("val", "neq", 1), None, 9, {"indices_set": [1, 3, 4, 5], "stats": {"total_chunks": 3}, "stats_ge": {"pruned_chunks": 1}})

META_BUILD_ERRORS = [
    dict(name="meta_build_mismatched_column_len_errors", cite="tests/meta_tests.rs:155-166", vectors=[[1.0], [2.0]],
         columns=[["age", "Int32", [1]]], chunk_size=2, error=True),
]

# ---- expression compiler KATs (tests/expr_tests.rs) -------------------------------------------------------
SCHEMA = {"age": "Int64", "score": "Float64", "name": "String", "ts": "DateTime"}  # :8-16
EXPR = [
    dict(name="numeric_gt_simple", cite="tests/expr_tests.rs:18-30", expr=cmp_("age", "gt", 25),
         clauses=[[["age", "Gt", "i64", 25]]]),
    dict(name="literal_on_left_is_invalid", cite="tests/expr_tests.rs:32-42", expr=["cmp_raw", ["lit", 25], "lt", ["col", "age"]],
         error="InvalidComparison"),
    dict(name="string_eq_allowed", cite="tests/expr_tests.rs:44-56", expr=cmp_("name", "eq", "alice"),
         clauses=[[["name", "Eq", "str", "alice"]]]),
    dict(name="string_or_multiple_equalities", cite="tests/expr_tests.rs:58-78",
         expr=["or", cmp_("name", "eq", "Alice"), cmp_("name", "eq", "Bob")],
         clauses=[[["name", "Eq", "str", "Alice"], ["name", "Eq", "str", "Bob"]]]),
    dict(name="string_unsupported_op_err", cite="tests/expr_tests.rs:80-89", expr=cmp_("name", "gt", "bob"),
         error="UnsupportedStringOp", error_column="name"),
    dict(name="type_mismatch_errs/string_on_int", cite="tests/expr_tests.rs:91-102", expr=cmp_("age", "eq", "x"),
         error="TypeMismatch", error_column="age", error_got="string"),
    dict(name="type_mismatch_errs/float_on_int", cite="tests/expr_tests.rs:91-102", expr=cmp_("age", "gt", 25.5),
         error="TypeMismatch", error_column="age", error_got="float"),
    dict(name="float_column_widen_int_literal", cite="tests/expr_tests.rs:104-116", expr=cmp_("score", "gte", 80),
         clauses=[[["score", "Gte", "f64", 80.0]]]),
    dict(name="float_column_float_literal", cite="tests/expr_tests.rs:118-130", expr=cmp_("score", "gt", 80.5),
         clauses=[[["score", "Gt", "f64", 80.5]]]),
    dict(name="and_yields_two_clauses", cite="tests/expr_tests.rs:132-139", expr=["and", cmp_("age", "gt", 25), cmp_("score", "gte", 80.0)],
         clause_sizes=[1, 1]),
    dict(name="or_yields_one_clause_with_two_leaves", cite="tests/expr_tests.rs:141-147",
         expr=["or", cmp_("age", "gt", 25), cmp_("age", "lt", 18)], clause_sizes=[2]),
    dict(name="complex_cnf_distribution", cite="tests/expr_tests.rs:149-161",
         expr=["and", cmp_("age", "gt", 25), ["or", cmp_("score", "gte", 80.0), cmp_("age", "lt", 18)]], clause_sizes=[1, 2]),
    dict(name="unknown_column_error", cite="tests/expr_tests.rs:163-168", expr=cmp_("missing", "eq", 1), error="UnknownColumn",
         error_column="missing"),
    dict(name="datetime_string_literal_compiles", cite="tests/expr_tests.rs:170-190", expr=cmp_("ts", "gte", "2023-01-02T03:04:05Z"),
         clauses=[[["ts", "Gte", "i64", 1672628645000]]]),
    dict(name="datetime_non_string_literal_err", cite="tests/expr_tests.rs:192-199", expr=cmp_("ts", "eq", 1700000000000),
         error="TypeMismatch", error_column="ts", error_got="datetime string"),
    dict(name="tautology_in_or_clause_is_removed", cite="tests/expr_tests.rs:201-208",
         expr=["and", ["or", cmp_("name", "eq", "bob"), cmp_("name", "neq", "bob")], cmp_("age", "gt", 5)],
         clauses=[[["age", "Gt", "i64", 5]]]),
]

# ---- compare-mask KATs (tests/simd_types_tests.rs): lane-wise i64x8 / f64x8 compares -> u8 (bit i = lane i),
# the primitives the row and zonemap predicates are built from (src/type_utils.rs:306-584).
# required = bits the reference asserts are set; forbidden = bits it asserts are clear -----------------------------
DESC = [5, 4, 3, 2, 1, 0, -1, -2]
ASC = [1, 2, 3, 4, 5, 6, 7, 8]
MASKS = []
for ty, cites in (("i64", ["34-44", "46-54", "56-65", "67-75", "77-86"]), ("f64", ["137-145", "147-155", "157-166", "168-176", "178-187"])):
    f = (lambda x: x) if ty == "i64" else float
    conv = lambda xs: [f(x) for x in xs]
    MASKS += [
        dict(cite=f"tests/simd_types_tests.rs:{cites[0]}", ty=ty, op="eq", a=conv(ASC), b=conv([1, 2, 3, 4, 9, 10, 11, 12]), required=0x0F, forbidden=0xF0),
        dict(cite=f"tests/simd_types_tests.rs:{cites[1]}", ty=ty, op="gt", a=conv(DESC), b=conv(ASC), required=0x03, forbidden=0),
        dict(cite=f"tests/simd_types_tests.rs:{cites[2]}", ty=ty, op="gte", a=conv(DESC), b=conv([5, 3, 3, 3, 1, 1, 0, 0]), required=0b00010111, forbidden=0),
        dict(cite=f"tests/simd_types_tests.rs:{cites[3]}", ty=ty, op="lt", a=conv(ASC), b=conv(DESC), required=0x03, forbidden=0),
        dict(cite=f"tests/simd_types_tests.rs:{cites[4]}", ty=ty, op="lte", a=conv([1, 3, 3, 4, 1, 0, -1, -2]), b=conv([5, 3, 3, 2, 1, 0, 0, 0]), required=0b00110111, forbidden=0),
    ]

if __name__ == "__main__":
    out = dict(
        note="Transcribed from the reference's own tests (see each 'cite'); generated by tests/golden/transcribe_reference_kats.py",
        vec=VEC, funcs=FUNCS, add_errors=ADD_ERRORS, meta=META, meta_build_errors=META_BUILD_ERRORS, expr_schema=SCHEMA, expr=EXPR,
        masks=MASKS,
    )
    path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_kats.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1)
    print(f"wrote {path}: {len(VEC)} vec, {len(META)} meta, {len(EXPR)} expr, {len(FUNCS)} fn KATs")
