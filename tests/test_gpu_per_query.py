"""Per-query top-k for batches (otters_*_query_batch: an extension beyond the reference's merged list, src/vec.rs:217-219) and
the device-side result-column gather (otters_metastore_gather: MetaQueryResults.data, src/meta.rs:723-821).
Parity bar: list i of a batch is bit-identical to the oracle's single-query answer for query i; gathered columns equal
the host columns at the result rows, NULLs preserved."""
import numpy as np
import pytest

from helpers import assert_same_results, ob, ora
from test_gpu_parity import FILTERS, meta_columns

pytestmark = pytest.mark.gpu

METRICS = [ob.Metric.Cosine, ob.Metric.Euclidean, ob.Metric.DotProduct]


@pytest.mark.parametrize("metric", METRICS, ids=lambda m: m.name)
def test_vecstore_per_query_lists(metric, ctx):
    n, dim, nq = 30000, 72, 7
    v = ora.synth_fill(0, n, dim, 101)
    q = ora.synth_fill(0, nq, dim, 102)
    q[4] = q[1]  # a repeated query gets the same list twice
    store = ob.VecStore(dim)
    store.add_vectors(v)
    rng = np.random.default_rng(4)
    mask = rng.random(n) < 0.5
    for k, flt, m in ((1, None, None), (50, None, None), (50, (0.01, ob.Cmp.Gt), mask), (2000, None, None), (0, None, None)):
        for tt, call in ((ob.TakeType.Max, "take_max"), (ob.TakeType.Min, "take_min")):
            plan = store.query(q, metric)
            if m is not None:
                plan = plan.with_row_mask(m)
            if flt:
                plan = plan.filter(*flt)
            lists = getattr(plan, call)(k).collect_per_query()
            assert len(lists) == nq
            for i in range(nq):
                want = ora.vecstore_query(v, q[i:i + 1], metric, tt, k, flt, m, ora.CANONICAL)
                assert_same_results(lists[i], want[:2], f"{metric.name} k={k} {call} query {i}")
    assert np.array_equal(lists[4][0], lists[1][0])


def test_vecstore_per_query_validation(ctx):
    store = ob.VecStore(8)
    store.add_vectors(ora.synth_fill(0, 10, 8, 1))
    with pytest.raises(ob.OttersError) as ei:
        store.query(np.ones((2, 5), np.float32), ob.Metric.Cosine).take(3).collect_per_query()
    assert str(ei.value) == "Query vector length 5 does not match expected dimension 8"


@pytest.mark.parametrize("cs", [96, 1024])
def test_metastore_per_query_lists_and_device_gather(cs, ctx):
    n, dim, nq = 12000, 40, 5
    vectors = ora.synth_fill(0, n, dim, 103)
    cols = meta_columns(n, cs, 104)
    store = ob.MetaStore.from_columns(cols).with_vectors(vectors).with_chunk_size(cs).build()
    ost = ora.MetaStore(vectors, cols, cs)
    q = ora.synth_fill(0, nq, dim, 105)
    for fi in (None, 1, 4, 8, 11):
        expr = FILTERS[fi]() if fi is not None else None
        fp = ora.FilterPack.from_compiled(expr.compile(store.schema()), store.column_index()) if expr is not None else None
        for metric, k, vf in ((ob.Metric.Cosine, 20, None), (ob.Metric.Euclidean, 300, None), (ob.Metric.DotProduct, 64, (0.5, ob.Cmp.Gt)), (ob.Metric.Cosine, 1500, None)):
            plan = store.query_batch(q, metric)
            if expr is not None:
                plan = plan.meta_filter(expr)
            if vf:
                plan = plan.vec_filter(*vf)
            results = plan.take(k).collect_per_query()
            st = store.last_query_stats()
            tt = ob.TakeType.Min if metric == ob.Metric.Euclidean else ob.TakeType.Max
            assert len(results) == nq
            for i, res in enumerate(results):
                oi, os_, _, ostats1 = ost.query(q[i:i + 1], metric, tt, k, vf, fp, ora.CANONICAL)
                assert_same_results((res.indices, res.scores), (oi, os_), f"cs={cs} filter={fi} {metric.name} k={k} query {i}")
                assert all(x == i for x in res.query_ids)
                # result columns were gathered on the device: same values and NULLs as the host columns at those rows
                for name in res.columns:
                    host = store.columns()[name].gather(res.indices)
                    dev = res.data[name]
                    for j in range(len(res.indices)):
                        a, b = host.get(j), dev.get(j)
                        assert (a is None and b is None) or a == b or (isinstance(a, float) and np.isnan(a) and np.isnan(b)), (name, j, a, b)
            # batch statistics follow the reference's convention: chunks once, vectors_compared = sum(len) * Q
            _, _, _, ostats = ost.query(q, metric, tt, k, vf, fp, ora.CANONICAL)
            for key in ("total_chunks", "pruned_chunks", "evaluated_chunks", "vectors_compared"):
                assert getattr(st, key) == ostats[key], f"stats.{key}: {getattr(st, key)} != {ostats[key]}"


def test_gather_rejects_bad_rows(ctx):
    n = 100
    cols = meta_columns(n, 16, 106)
    store = ob.MetaStore.from_columns(cols).with_vectors(ora.synth_fill(0, n, 8, 107)).with_chunk_size(16).build()
    name = sorted(store.schema())[0]
    with pytest.raises(ob.OttersError):
        store.gather(name, [5, 100])
    assert len(store.gather(name, [])) == 0
