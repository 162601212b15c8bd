"""Parity at BASELINE.json's FULL sizes, through properties that do not need the oracle to scan the whole store
(SURVEY.md §8c/§8d): the stores are generated on the device by the counter-based generator, and
  * every returned (row, score) is re-derived by the oracle from that row alone (rows are regenerated on the host) — bit-identical;
  * a planted row (the query itself) must come back first with the exact score;
  * the result over the full store equals the merge of the results over two disjoint row-mask halves (partition property);
  * results are ordered, rows are unique, a repeated query returns the same bytes (idempotence);
  * MetaStore: every returned row satisfies the predicate, and the chunk statistics equal the oracle's on the same
    metadata (pruning does not depend on the vectors, so the oracle runs it with 1-d stand-in vectors);
  * batches: the tensor-core path and the per-query exact path return the same merged list.
They cost a few seconds each on a B200 (the 10M x 768 stores are 30.7 GB of HBM)."""
import gc

import numpy as np
import pytest

from helpers import assert_same_results, ob, ora

pytestmark = pytest.mark.gpu

SEED = 0x07735


def oracle_rows(rows, dim, q, metric, bf16=False):
    """Exact scores of individual store rows (regenerated on the host; rounded like the store's for a bf16 store) against one query."""
    v = np.concatenate([ora.synth_fill(int(r), 1, dim, SEED) for r in rows], axis=0)
    if bf16:
        v = ora.round_bf16(v)
    idx, score, _ = ora.vecstore_query(v, q[None, :], metric, ob.TakeType.Max, len(rows), None, None, ora.CANONICAL)
    out = np.zeros(len(rows), np.float32)
    out[np.asarray(idx, np.int64)] = score
    return out


def check_against_row_oracle(got, dim, q, metric, bf16=False):
    idx, score = np.asarray(got[0], np.int64), np.asarray(got[1], np.float32)
    assert len(set(idx.tolist())) == len(idx), "rows must be unique in a single-query result"
    want = oracle_rows(idx, dim, q, metric, bf16)
    same = (want.view(np.uint32) == score.view(np.uint32)) | ((want == 0) & (score == 0))
    assert same.all(), f"scores differ from the per-row oracle at {np.nonzero(~same)[0][:5]}"


def merge_host(parts, k, take_max):
    idx = np.concatenate([p[0] for p in parts]).astype(np.int64)
    sc = np.concatenate([p[1] for p in parts]).astype(np.float32)
    order = np.lexsort((idx, -sc if take_max else sc))[:k]
    return idx[order].astype(np.uint64), sc[order]


@pytest.fixture(scope="module")
def big_vecstore(ctx):
    s = ob.VecStore(768, ctx)
    s.add_synthetic(0, 10_000_000, SEED)  # BASELINE config 4: 10M x 768 fp32, 30.72 GB in HBM
    yield s
    s.close()
    gc.collect()


@pytest.mark.parametrize("metric", [ob.Metric.Euclidean, ob.Metric.Cosine], ids=lambda m: m.name)
def test_fullsize_vecstore_properties(metric, big_vecstore, ctx):
    n, dim, k = 10_000_000, 768, 100
    take_max = metric != ob.Metric.Euclidean
    planted = 7_654_321
    q = ora.synth_fill(planted, 1, dim, SEED)[0]  # the query is a row of the store
    got = big_vecstore.query(q, metric).take(k).collect_arrays()
    assert len(got[0]) == k
    s = got[1]
    assert np.all(s[:-1] >= s[1:]) if take_max else np.all(s[:-1] <= s[1:]), "best-first order"
    assert int(got[0][0]) == planted
    if metric == ob.Metric.Euclidean:
        assert float(s[0]) == 0.0
    check_against_row_oracle(got, dim, q, metric)
    # idempotence
    again = big_vecstore.query(q, metric).take(k).collect_arrays()
    assert_same_results(again, got, "repeat")
    # nothing outside the result beats its last entry (a random sample of other rows, scored by the oracle)
    rng = np.random.default_rng(5)
    sample = np.setdiff1d(rng.integers(0, n, 400), np.asarray(got[0], np.int64))
    others = oracle_rows(sample, dim, q, metric)
    assert np.all(others <= s[-1]) if take_max else np.all(others >= s[-1])
    # partition property: even rows + odd rows, merged on the host, give the same list
    mask = np.zeros(n, bool)
    mask[0::2] = True
    even = big_vecstore.query(q, metric).with_row_mask(mask).take(k).collect_arrays()
    odd = big_vecstore.query(q, metric).with_row_mask(~mask).take(k).collect_arrays()
    assert np.all(np.asarray(even[0]) % 2 == 0) and np.all(np.asarray(odd[0]) % 2 == 1)
    assert_same_results(merge_host([even, odd], k, take_max), got[:2], "partition")
    # both front-ends of the scan kernel
    for mode in (1, 2):
        ctx.set_tuning(scan_mode=mode)
        assert_same_results(big_vecstore.query(q, metric).take(k).collect_arrays(), got, f"scan_mode {mode}")
    ctx.set_tuning()


def test_fullsize_bf16_vecstore_properties(ctx):
    """10M x 768 kept as bf16 rows (15.4 GB): the same properties, the oracle working on the rounded rows."""
    n, dim, k = 10_000_000, 768, 100
    s = ob.VecStore(dim, ctx, ob.VectorFormat.Bf16)
    s.add_synthetic(0, n, SEED)
    planted = 7_654_321
    q = ora.round_bf16(ora.synth_fill(planted, 1, dim, SEED))[0]  # the query is a (rounded) row of the store
    for metric in (ob.Metric.Cosine, ob.Metric.Euclidean):
        take_max = metric != ob.Metric.Euclidean
        got = s.query(q, metric).take(k).collect_arrays()
        sc = got[1]
        assert len(got[0]) == k and (np.all(sc[:-1] >= sc[1:]) if take_max else np.all(sc[:-1] <= sc[1:]))
        assert int(got[0][0]) == planted and (metric != ob.Metric.Euclidean or float(sc[0]) == 0.0)
        check_against_row_oracle(got, dim, q, metric, bf16=True)
        assert ctx.last_work()["scan_bytes"] == n * (dim * 2 + (4 if metric == ob.Metric.Cosine else 0))
        rng = np.random.default_rng(6)
        sample = np.setdiff1d(rng.integers(0, n, 400), np.asarray(got[0], np.int64))
        others = oracle_rows(sample, dim, q, metric, bf16=True)
        assert np.all(others <= sc[-1]) if take_max else np.all(others >= sc[-1])
        mask = np.zeros(n, bool)
        mask[0::2] = True
        even = s.query(q, metric).with_row_mask(mask).take(k).collect_arrays()
        odd = s.query(q, metric).with_row_mask(~mask).take(k).collect_arrays()
        assert_same_results(merge_host([even, odd], k, take_max), got[:2], "partition")
        for mode in (1, 2):
            ctx.set_tuning(scan_mode=mode)
            assert_same_results(s.query(q, metric).take(k).collect_arrays(), got, f"scan_mode {mode}")
        ctx.set_tuning()
    s.close()
    gc.collect()


def test_fullsize_metastore_target(ctx):
    """North-star target: MetaStore 10M x 768 Cosine top-100 with price.gt & item.eq & ts.gte (bench.py's generators)."""
    import bench_workloads as bw

    n, dim, chunk, k = 10_000_000, 768, 1024, 100
    wl = bw.Workload("target")
    cols = [c.to_ob(ob) for c in wl.columns(np.arange(n))]
    store = ob.MetaStore.from_columns(cols).with_synthetic_vectors(n, dim, SEED).with_chunk_size(chunk).with_context(ctx).build()
    expr = wl.expr(ob)
    q = ora.synth_fill(123_456, 1, dim, SEED)[0] + np.float32(0.25) * ora.synth_fill(0, 1, dim, 99)[0]
    res = store.query(q, ob.Metric.Cosine).meta_filter(expr).take(k).collect()
    st = store.last_query_stats()
    idx, score = np.array(res.indices, np.int64), np.array(res.scores, np.float32)
    assert len(idx) == k and np.all(score[:-1] >= score[1:])
    check_against_row_oracle((idx, score), dim, q, ob.Metric.Cosine)
    # every returned row satisfies the predicate (evaluated on the host columns)
    fp = ora.FilterPack.from_compiled(expr.compile(store.schema()), store.column_index())
    stand_in = np.ones((n, 1), np.float32)  # pruning and row masks do not depend on the vectors
    ost = ora.MetaStore(stand_in, cols, chunk)
    keep_rows = ost.row_mask(fp)
    assert keep_rows[idx].all()
    # chunk statistics: exactly the oracle's
    _, _, _, ostats = ost.query(np.ones((1, 1), np.float32), ob.Metric.DotProduct, ob.TakeType.Max, 1, None, fp, ora.CANONICAL)
    assert (st.total_chunks, st.pruned_chunks, st.evaluated_chunks, st.vectors_compared) == (
        ostats["total_chunks"], ostats["pruned_chunks"], ostats["evaluated_chunks"], ostats["vectors_compared"])
    assert st.total_chunks == 9766 and 0 < st.evaluated_chunks < st.total_chunks
    # rows scored on the device == rows the oracle's mask keeps
    assert ctx.last_work()["rows_scored"] == int(keep_rows.sum())
    # a sample of kept rows outside the result cannot beat its last entry
    rng = np.random.default_rng(7)
    kept = np.nonzero(keep_rows)[0]
    sample = np.setdiff1d(rng.choice(kept, 400, replace=False), idx)
    assert np.all(oracle_rows(sample, dim, q, ob.Metric.Cosine) <= score[-1])
    del store
    gc.collect()


def _fullsize_meta_workload(ctx, name):
    """Builds bench.py's full-size MetaStore workload `name` on the device; returns (workload, store, host columns, oracle filter)."""
    import bench_workloads as bw

    wl = bw.Workload(name)
    spec = wl.columns(np.arange(wl.rows))
    store = (ob.MetaStore.from_columns([c.to_ob(ob) for c in spec]).with_synthetic_vectors(wl.rows, wl.dim, SEED)
             .with_chunk_size(wl.chunk).with_context(ctx).build())
    prow, pvec = wl.planted()
    if len(prow):
        store.set_rows(prow, pvec)
    idx = {c.name(): i for i, c in enumerate(spec)}
    fp = ora.FilterPack([[(idx[n], op, kind, val) for n, op, kind, val in cl] for cl in wl.clauses()])
    return wl, store, spec, fp, (prow, pvec)


def _check_stats_and_mask(ctx, wl, store, spec, fp, idx):
    """Chunk statistics equal the oracle's on the same metadata; returned rows pass the predicate; rows scored == rows kept."""
    st = store.last_query_stats()
    stand_in = np.ones((wl.rows, 1), np.float32)  # pruning and row masks do not depend on the vectors
    ost = ora.MetaStore(stand_in, spec, wl.chunk)
    keep_rows = ost.row_mask(fp)
    assert keep_rows[idx].all(), "a returned row fails the metadata filter"
    assert np.array_equal(keep_rows.astype(bool), wl.row_mask(spec)), "NumPy restatement of the CNF disagrees with the oracle"
    _, _, _, ostats = ost.query(np.ones((1, 1), np.float32), ob.Metric.DotProduct, ob.TakeType.Max, 1, None, fp, ora.CANONICAL)
    assert (st.total_chunks, st.pruned_chunks, st.evaluated_chunks, st.vectors_compared) == (
        ostats["total_chunks"], ostats["pruned_chunks"], ostats["evaluated_chunks"], ostats["vectors_compared"])
    assert 0 < st.evaluated_chunks < st.total_chunks
    assert ctx.last_work()["rows_scored"] == int(keep_rows.sum())
    return keep_rows


def test_fullsize_metastore_c3(ctx):
    """BASELINE config 3: MetaStore 10M x 128 Cosine, chunk 1024, price.gt & item.eq & ts.gte (zonemap + Bloom prune), top-100."""
    wl, store, spec, fp, _ = _fullsize_meta_workload(ctx, "c3")
    dim, k = wl.dim, wl.k
    q = ora.synth_fill(4_321_987, 1, dim, SEED)[0] + np.float32(0.25) * ora.synth_fill(0, 1, dim, 98)[0]
    res = store.query(q, ob.Metric.Cosine).meta_filter(wl.expr(ob)).take(k).collect()
    idx, score = np.array(res.indices, np.int64), np.array(res.scores, np.float32)
    assert len(idx) == k and np.all(score[:-1] >= score[1:])
    check_against_row_oracle((idx, score), dim, q, ob.Metric.Cosine)
    keep_rows = _check_stats_and_mask(ctx, wl, store, spec, fp, idx)
    assert store.last_query_stats().total_chunks == 9766
    # a sample of kept rows outside the result cannot beat its last entry
    rng = np.random.default_rng(9)
    sample = np.setdiff1d(rng.choice(np.nonzero(keep_rows)[0], 2000, replace=False), idx)
    assert np.all(oracle_rows(sample, dim, q, ob.Metric.Cosine) <= score[-1])
    # the non-blocking API, both predicate placements and both front-ends return the same bytes
    again = store.query(q, ob.Metric.Cosine).meta_filter(wl.expr(ob)).take(k).submit().wait()
    assert_same_results((again.indices, again.scores), (idx, score), "submit/wait")
    for tune in (dict(disable_fused_predicate=2), dict(disable_fused_predicate=2, lazy_prune=1), dict(scan_mode=2), dict(scan_mode=1)):
        ctx.set_tuning(**tune)
        r2 = store.query(q, ob.Metric.Cosine).meta_filter(wl.expr(ob)).take(k).collect()
        assert_same_results((r2.indices, r2.scores), (idx, score), str(tune))
    ctx.set_tuning()
    del store
    gc.collect()


def test_fullsize_metastore_c5(ctx):
    """BASELINE config 5 as specified: MetaStore 5M x 1536 Cosine, Int32 / Float64 / String predicates
    qty.gte & (price.lt | item.eq) & brand.neq, ~5,000 planted near-duplicates of the query, vec_filter(0.8, Gt), take(1000).
    The answer is known by construction: the best 1000 planted rows that pass the metadata filter, all with cosine > 0.8."""
    wl, store, spec, fp, (prow, pvec) = _fullsize_meta_workload(ctx, "c5")
    dim, k = wl.dim, wl.k
    q = wl.queries()[0, 0]
    res = store.query(q, ob.Metric.Cosine).meta_filter(wl.expr(ob)).vec_filter(0.8, ob.Cmp.Gt).take(k).collect()
    idx, score = np.array(res.indices, np.int64), np.array(res.scores, np.float32)
    assert len(idx) == k and np.all(score[:-1] >= score[1:]) and np.all(score > np.float32(0.8))
    keep_rows = _check_stats_and_mask(ctx, wl, store, spec, fp, idx)
    # expected: oracle scores of the planted rows alone, filtered by the predicate and the threshold, best 1000
    pidx, psc, _ = ora.vecstore_query(pvec, q[None, :], ob.Metric.Cosine, ob.TakeType.Max, len(prow), (0.8, ob.Cmp.Gt), None, ora.CANONICAL)
    pidx = np.asarray(pidx, np.int64)
    ok = keep_rows[prow[pidx]].astype(bool)
    assert ok.sum() > k, "the workload must plant more passing near-duplicates than it asks for"
    assert_same_results((idx, score), (prow[pidx][ok][:k], np.asarray(psc)[ok][:k]), "planted set's top-1000")
    # without the threshold the same 1000 rows lead the list (nothing else comes near 0.8); take(k) > 1024 takes the sort path
    res2 = store.query(q, ob.Metric.Cosine).meta_filter(wl.expr(ob)).take(1500).collect()
    assert_same_results((res2.indices[:k], res2.scores[:k]), (idx, score), "no vec_filter, sort path")
    n_ok = int(ok.sum())
    assert n_ok >= 1500 or (res2.scores[n_ok] < 0.8 <= res2.scores[n_ok - 1])
    assert np.isin(np.array(res2.indices[: min(n_ok, 1500)], np.int64), prow).all()
    del store
    gc.collect()


def test_fullsize_batched_config2(ctx):
    """BASELINE config 2: VecStore 1M x 768 DotProduct, 1024 queries, one merged top-100."""
    n, dim, nq, k = 1_000_000, 768, 1024, 100
    s = ob.VecStore(dim, ctx)
    s.add_synthetic(0, n, SEED)
    q = ora.synth_fill(0, nq, dim, 0xBEEF)
    ctx.set_tuning(batch_mode=1)
    idx, score, qid = s.query(q, ob.Metric.DotProduct).take(k).collect_arrays()
    w = ctx.last_work()
    assert w["batch_used"] == 1 and w["batch_fallback"] == 0 and w["batch_max_err"] <= 0.5 * w["batch_delta"], w
    assert w["batch_passes"] in (1, 2, 3) and w["rows_scored"] == n * nq, w
    # 3xTF32 only: same bytes
    ctx.set_tuning(batch_mode=1, batch_passes=3)
    three = s.query(q, ob.Metric.DotProduct).take(k).collect_arrays()
    assert ctx.last_work()["batch_passes"] == 3
    assert_same_results(three, (idx, score, qid), "3xTF32 vs automatic ladder")
    # bf16 rung only (bf16 shadow of the rows, 2^-7 bound): certified on this workload, same bytes
    ctx.set_tuning(batch_mode=1, batch_passes=2)
    half = s.query(q, ob.Metric.DotProduct).take(k).collect_arrays()
    w2 = ctx.last_work()
    assert w2["batch_used"] == 1 and w2["batch_passes"] == 2 and w2["batch_attempts"] == 1 and w2["batch_max_err"] <= w2["batch_delta"], w2
    assert_same_results(half, (idx, score, qid), "bf16 rung vs automatic ladder")
    assert len(idx) == k and np.all(score[:-1] >= score[1:])
    # every returned (row, query) pair re-derived by the oracle from that row and that query alone
    for i in range(k):
        want = oracle_rows([int(idx[i])], dim, q[int(qid[i])], ob.Metric.DotProduct)[0]
        assert want.view(np.uint32) == score[i : i + 1].view(np.uint32)[0], (i, want, score[i])
    # the exact per-query path (1024 streaming scans) returns the same merged list
    ctx.set_tuning(batch_mode=2)
    ref = s.query(q, ob.Metric.DotProduct).take(k).collect_arrays()
    ctx.set_tuning()
    assert_same_results((idx, score, qid), ref, "tensor-core path vs per-query path")
    s.close()
