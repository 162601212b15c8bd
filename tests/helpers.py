"""Shared test helpers: KAT replay on the oracle and on the product, comparison utilities."""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import otters_b200 as ob  # noqa: E402  (host mirror + ctypes binding; no device work at import)
from oracle import oracle as ora  # noqa: E402

METRIC = {"Cosine": ob.Metric.Cosine, "Euclidean": ob.Metric.Euclidean, "DotProduct": ob.Metric.DotProduct}
CMP = {"Lt": ob.Cmp.Lt, "Gt": ob.Cmp.Gt, "Lte": ob.Cmp.Lte, "Gte": ob.Cmp.Gte, "Eq": ob.Cmp.Eq}
DTYPE = {n: getattr(ob.DataType, n) for n in ("Int32", "Int64", "Float32", "Float64", "String", "DateTime")}
PYCMP = {
    "Lt": lambda s, t: s < t, "Gt": lambda s, t: s > t, "Lte": lambda s, t: s <= t, "Gte": lambda s, t: s >= t,
    "Eq": lambda s, t: s == t,
}


def load_kats():
    with open(os.path.join(ROOT, "tests", "golden", "reference_kats.json")) as f:
        return json.load(f)


# ---- plan resolution exactly as the reference's builder does it (src/vec.rs:92-116, :213-214) ----------
def resolve_vec_calls(calls, metric, n_vecs):
    take_count, take_type, flt = None, None, None
    for c in calls:
        if c[0] == "filter":
            flt = (float(c[1]), CMP[c[2]])
        elif c[0] == "take":
            take_count = c[1]
            if take_type is None and metric is not None:
                take_type = ob.TakeType.Min if metric == ob.Metric.Euclidean else ob.TakeType.Max
        elif c[0] == "take_min":
            take_count, take_type = c[1], ob.TakeType.Min
        elif c[0] == "take_max":
            take_count, take_type = c[1], ob.TakeType.Max
    k = take_count if take_count is not None else n_vecs
    tt = take_type if take_type is not None else ob.TakeType.Max
    return k, tt, flt


def as_batch(queries):
    if queries is None:
        return None
    if len(queries) == 0:
        return []
    if isinstance(queries[0], (list, tuple)):
        return [list(q) for q in queries]
    return [list(queries)]


def oracle_vec_kat(kat, mode):
    """Replays a VecStore KAT on the oracle.  Returns ("ok", idx, scores) or ("err", message)."""
    if kat["plan_new"]:
        return ("err", "Query vectors or their norms are not set")  # src/vec.rs:174-176: validate() before anything else
    dim = kat["dim"]
    metric = METRIC[kat["metric"]]
    batch = as_batch(kat["queries"])
    # VecQueryPlan::validate (src/vec.rs:170-203) — host-side checks, restated
    if len(batch) == 0:
        return ("err", "No queries provided")
    for q in batch:
        if len(q) != dim:
            return ("err", f"Query vector length {len(q)} does not match expected dimension {dim}")
    vectors = np.asarray(kat["vectors"], dtype=np.float32).reshape(-1, dim)
    k, tt, flt = resolve_vec_calls(kat["calls"], metric, vectors.shape[0])
    idx, score, _ = ora.vecstore_query(vectors, np.asarray(batch, np.float32), metric, tt, k, flt, None, mode)
    return ("ok", idx, score)


def product_vec_kat(kat):
    """Replays a VecStore KAT through the public API of otters_b200."""
    try:
        if kat["plan_new"]:
            plan = ob.VecQueryPlan.new()
        else:
            store = ob.VecStore(kat["dim"])
            if kat["vectors"]:
                store.add_vectors(kat["vectors"])
            plan = store.query(kat["queries"], METRIC[kat["metric"]])
        for c in kat["calls"]:
            if c[0] == "filter":
                plan = plan.filter(c[1], CMP[c[2]])
            else:
                plan = getattr(plan, c[0])(c[1])
        res = plan.collect()
    except ob.OttersError as e:
        return ("err", str(e))
    return ("ok", np.array([r.index for r in res], np.uint64), np.array([r.score for r in res], np.float32))


def check_expect(kat, result):
    exp = kat["expect"]
    name = kat["name"]
    if any(key in exp for key in ("error", "error_contains", "error_equals")):
        assert result[0] == "err", f"{name}: expected an error, got {result}"
        if "error_contains" in exp:
            assert exp["error_contains"] in result[1], f"{name}: {result[1]!r}"
        if "error_equals" in exp:
            assert exp["error_equals"] == result[1], f"{name}: {result[1]!r}"
        return
    assert result[0] == "ok", f"{name}: unexpected error {result[1]!r}"
    idx, score = [int(i) for i in result[1]], [float(s) for s in result[2]]
    tol = exp.get("tol", 1e-6)
    if "len" in exp:
        assert len(idx) == exp["len"], f"{name}: len {len(idx)}"
    if "max_len" in exp:
        assert len(idx) <= exp["max_len"], name
    if exp.get("nonempty"):
        assert len(idx) > 0, name
    if "indices" in exp:
        assert idx == exp["indices"], f"{name}: {idx}"
    if "indices_set" in exp:
        assert sorted(set(idx)) == sorted(exp["indices_set"]), f"{name}: {idx}"
    if "scores" in exp:
        assert len(score) == len(exp["scores"]), f"{name}: {score}"
        for a, b in zip(score, exp["scores"]):
            assert abs(a - b) <= tol, f"{name}: {score}"
    if "score_by_index" in exp:
        got = dict(zip(idx, score))
        for i, s in exp["score_by_index"].items():
            assert int(i) in got, f"{name}: index {i} missing from {idx}"
            assert abs(got[int(i)] - s) <= tol, f"{name}: idx {i} score {got[int(i)]}"
    if "all_scores" in exp:
        op, thr = exp["all_scores"]
        assert all(PYCMP[op](s, np.float32(thr)) for s in score), f"{name}: {score}"
    if "sorted" in exp:
        for a, b in zip(score, score[1:]):
            assert (a >= b) if exp["sorted"] == "desc" else (a <= b), f"{name}: {score}"
    if "count_score" in exp:
        s0, cnt = exp["count_score"]
        assert sum(1 for s in score if abs(s - s0) <= tol) == cnt, f"{name}: {score}"


# ---- expressions from their JSON form -------------------------------------------------------------------
def build_expr(e):
    if e is None:
        return None
    tag = e[0]
    if tag == "cmp":
        return getattr(ob.col(e[1]), e[2])(e[3])
    if tag == "cmp_raw":  # arbitrary operands (literal on the left etc.)
        def side(x):
            return ob.lit(x[1]) if x[0] == "lit" else ob.col(x[1])
        op = {"eq": ob.CmpOp.Eq, "neq": ob.CmpOp.Neq, "lt": ob.CmpOp.Lt, "lte": ob.CmpOp.Lte, "gt": ob.CmpOp.Gt, "gte": ob.CmpOp.Gte}[e[2]]
        return ob.Expr("cmp", side(e[1]), side(e[3]), op)
    if tag == "and":
        return build_expr(e[1]) & build_expr(e[2])
    if tag == "or":
        return build_expr(e[1]) | build_expr(e[2])
    raise ValueError(tag)


def build_columns(spec):
    cols = []
    for name, dt, values in spec:
        cols.append(ob.Column(name, DTYPE[dt]).from_values(values))
    return cols


def oracle_meta_kat(kat, mode):
    cols = build_columns(kat["columns"])
    vectors = np.asarray(kat["vectors"], np.float32)
    store = ora.MetaStore(vectors, cols, kat["chunk_size"])
    schema = {c.name(): c.dtype() for c in cols}
    col_index = {c.name(): i for i, c in enumerate(cols)}
    metric = METRIC[kat["metric"]]
    tt = ob.TakeType.Min if metric == ob.Metric.Euclidean else ob.TakeType.Max
    expr = build_expr(kat["expr"])
    fp = ora.FilterPack.from_compiled(expr.compile(schema), col_index) if expr is not None else None
    vf = (kat["vec_filter"][0], CMP[kat["vec_filter"][1]]) if kat["vec_filter"] else None
    batch = np.asarray(as_batch(kat["queries"]), np.float32)
    idx, score, qid, stats = store.query(batch, metric, tt, kat["take"], vf, fp, mode)
    return ("ok", idx, score), stats


def product_meta_kat(kat):
    cols = build_columns(kat["columns"])
    store = ob.MetaStore.from_columns(cols).with_vectors(kat["vectors"]).with_chunk_size(kat["chunk_size"]).build()
    metric = METRIC[kat["metric"]]
    batch = as_batch(kat["queries"])
    plan = store.query(batch[0], metric) if len(batch) == 1 else store.query_batch(batch, metric)
    expr = build_expr(kat["expr"])
    if expr is not None:
        plan = plan.meta_filter(expr)
    if kat["vec_filter"]:
        plan = plan.vec_filter(kat["vec_filter"][0], CMP[kat["vec_filter"][1]])
    res = plan.take(kat["take"]).collect()
    st = store.last_query_stats()
    stats = dict(total_chunks=st.total_chunks, pruned_chunks=st.pruned_chunks, evaluated_chunks=st.evaluated_chunks,
                 vectors_compared=st.vectors_compared)
    return ("ok", np.array(res.indices, np.uint64), np.array(res.scores, np.float32)), stats


def check_meta_expect(kat, result, stats):
    check_expect(kat, result)
    exp = kat["expect"]
    for key, v in exp.get("stats", {}).items():
        assert stats[key] == v, f"{kat['name']}: stats.{key} = {stats[key]}, expected {v}"
    for key, v in exp.get("stats_ge", {}).items():
        assert stats[key] >= v, f"{kat['name']}: stats.{key} = {stats[key]}, expected >= {v}"
    if exp.get("stats_le_total"):
        assert stats["evaluated_chunks"] <= stats["total_chunks"]
    assert stats["pruned_chunks"] == stats["total_chunks"] - stats["evaluated_chunks"]


# ---- parity comparison ----------------------------------------------------------------------------------
def assert_same_results(got, want, what="", exact_scores=True, rtol=1e-5):
    gi, gs = np.asarray(got[0]), np.asarray(got[1], np.float32)
    wi, ws = np.asarray(want[0]), np.asarray(want[1], np.float32)
    assert len(gi) == len(wi), f"{what}: result count {len(gi)} != {len(wi)}"
    if not np.array_equal(gi.astype(np.uint64), wi.astype(np.uint64)):
        bad = np.nonzero(gi.astype(np.uint64) != wi.astype(np.uint64))[0][:5]
        raise AssertionError(f"{what}: indices differ at {bad}: got {gi[bad]} ({gs[bad]}) want {wi[bad]} ({ws[bad]})")
    if exact_scores:
        # +0.0 / -0.0 are the same score (documented canonicalisation of the sign of zero)
        same = (gs.view(np.uint32) == ws.view(np.uint32)) | ((gs == 0) & (ws == 0))
        assert same.all(), f"{what}: scores not bit-identical at {np.nonzero(~same)[0][:5]}: {gs[~same][:5]} vs {ws[~same][:5]}"
    else:
        tol = rtol * np.maximum(np.abs(ws), 1e-30)
        assert (np.abs(gs - ws) <= tol).all(), f"{what}: scores differ beyond {rtol} relative"
    if len(got) > 2 and len(want) > 2 and got[2] is not None and want[2] is not None:
        assert np.array_equal(np.asarray(got[2], np.uint32), np.asarray(want[2], np.uint32)), f"{what}: query ids differ"
