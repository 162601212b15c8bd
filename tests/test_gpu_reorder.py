"""MetaStoreBuilder.with_row_order on the device path: a store clustered on its filter columns answers with the caller's row
ids, the same rows and bit-identical scores as the store in input order (and as the oracle), prunes more chunks, and its
result columns (gathered on the device at store positions) belong to the reported rows."""
import numpy as np
import pytest

from helpers import assert_same_results, ob, ora

pytestmark = pytest.mark.gpu


def make_columns(n, seed):
    rng = np.random.default_rng(seed)
    price = ob.Column.from_numpy("price", ob.DataType.Float64, rng.uniform(0, 100, n), rng.random(n) < 0.02)
    qty = ob.Column.from_numpy("qty", ob.DataType.Int32, rng.integers(0, 1000, n).astype(np.int32))
    item = ob.Column.from_categories("item", [f"item{i:02d}" for i in range(40)], rng.integers(0, 40, n))
    return [price, qty, item]


@pytest.mark.parametrize("method", ["sort", "zorder"])
@pytest.mark.parametrize("fmt", [ob.VectorFormat.F32, ob.VectorFormat.Bf16], ids=["f32", "bf16"])
def test_row_order_same_results_more_pruning(method, fmt, ctx):
    n, dim, cs = 30000, 64, 256
    v = ora.synth_fill(0, n, dim, 0x7735)
    cols = make_columns(n, 9)
    q = ora.synth_fill(0, 1, dim, 0xBEEF)[0]
    expr = ob.col("price").lt(12.0) & ob.col("qty").gte(600) & ob.col("item").neq("item03")
    plain = ob.MetaStore.from_columns(cols).with_vectors(v).with_chunk_size(cs).with_vector_format(fmt).with_context(ctx).build()
    clustered = (ob.MetaStore.from_columns(cols).with_vectors(v).with_chunk_size(cs).with_vector_format(fmt)
                 .with_row_order(["price", "qty"], method).with_context(ctx).build())
    assert plain.row_order() is None and sorted(clustered.row_order().tolist()) == list(range(n))
    for metric in (ob.Metric.Cosine, ob.Metric.Euclidean):
        a = plain.query(q, metric).meta_filter(expr).take(50).collect()
        sa = plain.last_query_stats()
        b = clustered.query(q, metric).meta_filter(expr).take(50).collect()
        sb = clustered.last_query_stats()
        assert len(a.indices) == 50
        assert_same_results((b.indices, b.scores), (a.indices, a.scores), f"{method} {metric.name}")
        assert sa.pruned_chunks == 0 and sb.pruned_chunks >= sb.total_chunks // 2, (sa, sb)
        assert sb.vectors_compared < sa.vectors_compared // 2
        # result columns: gathered at store positions, reported against the caller's row ids
        for name, colobj in zip(("price", "qty", "item"), cols):
            for j, row in enumerate(b.indices[:10]):
                assert b.data[name].get(j) == colobj.get(row), (name, j, row)
    # the oracle on the caller's data agrees as well
    vv = ora.round_bf16(v) if fmt == ob.VectorFormat.Bf16 else v
    ost = ora.MetaStore(vv, cols, cs)
    fp = ora.FilterPack.from_compiled(expr.compile(plain.schema()), plain.column_index())
    oi, os_, _, _ = ost.query(q[None, :], ob.Metric.Cosine, ob.TakeType.Max, 50, None, fp)
    b = clustered.query(q, ob.Metric.Cosine).meta_filter(expr).take(50).collect()
    assert_same_results((b.indices, b.scores), (oi, os_), f"{method} vs oracle")
    # submitted queries and per-query lists report the caller's row ids too
    w = clustered.query(q, ob.Metric.Cosine).meta_filter(expr).take(50).submit().wait()
    assert w.indices == b.indices
    per = clustered.query_batch(np.stack([q, -q]), ob.Metric.Cosine).meta_filter(expr).take(50).collect_per_query()
    assert per[0].indices == b.indices


def test_row_order_needs_host_vectors(ctx):
    cols = make_columns(100, 1)
    with pytest.raises(ob.OttersError):
        ob.MetaStore.from_columns(cols).with_synthetic_vectors(100, 16, 1).with_row_order("price").with_context(ctx).build()
