"""MetaStore.save / MetaStore.load (otters_metastore_save / _load; the reference's roadmap item "Persistence", README.md:206):
a loaded store returns the same bytes as the store that was saved — rows, scores, statistics, zonemap tables, masks,
gathered result columns — for fp32 and bf16 rows, with and without a row order; damaged files are rejected."""
import os

import numpy as np
import pytest

from helpers import assert_same_results, ob, ora

pytestmark = pytest.mark.gpu


def make_columns(n, seed):
    rng = np.random.default_rng(seed)
    price = ob.Column.from_numpy("price", ob.DataType.Float64, rng.uniform(0, 100, n), rng.random(n) < 0.02)
    qty = ob.Column.from_numpy("qty", ob.DataType.Int32, rng.integers(0, 1000, n).astype(np.int32))
    ts = ob.Column.from_numpy("ts", ob.DataType.DateTime, 1_700_000_000_000 + np.arange(n, dtype=np.int64) * 1000)
    item = ob.Column.from_categories("item", [f"item{i:02d}" for i in range(40)], rng.integers(0, 40, n), rng.random(n) < 0.01)
    score = ob.Column.from_numpy("score", ob.DataType.Float32, rng.standard_normal(n).astype(np.float32))
    big = ob.Column.from_numpy("big", ob.DataType.Int64, rng.integers(-2**40, 2**40, n))
    return [price, qty, ts, item, score, big]


@pytest.mark.parametrize("fmt,order", [(ob.VectorFormat.F32, None), (ob.VectorFormat.Bf16, None), (ob.VectorFormat.F32, "zorder")],
                         ids=["f32", "bf16", "f32_zorder"])
def test_save_load_round_trip(fmt, order, ctx, tmp_path):
    n, dim, cs = 12000, 72, 200
    v = ora.synth_fill(0, n, dim, 0x7735)
    cols = make_columns(n, 4)
    b = ob.MetaStore.from_columns(cols).with_vectors(v).with_chunk_size(cs).with_vector_format(fmt).with_context(ctx)
    if order:
        b = b.with_row_order(["price", "qty"], order)
    store = b.build()
    path = str(tmp_path / "store.otters")
    store.save(path)
    assert os.path.getsize(path) > n * dim * (2 if fmt == ob.VectorFormat.Bf16 else 4)
    loaded = ob.MetaStore.load(path, ctx)
    assert (loaded.len(), loaded.dim(), loaded.chunk_size(), loaded.n_chunks()) == (n, dim, cs, store.n_chunks())
    assert loaded.vector_format() == fmt and list(loaded.schema().items()) == list(store.schema().items())
    assert (loaded.row_order() is None) == (order is None)
    if order:
        assert np.array_equal(loaded.row_order(), store.row_order())
    q = ora.synth_fill(0, 2, dim, 0xBEEF)
    exprs = [ob.col("price").lt(30.0) & ob.col("qty").gte(500), ob.col("item").eq("item07") | ob.col("big").gt(0),
             ob.col("ts").gte("2023-11-14 23:00:00") & ob.col("score").lt(0.5) & ob.col("item").neq("item01")]
    for expr in exprs:
        for metric in (ob.Metric.Cosine, ob.Metric.Euclidean, ob.Metric.DotProduct):
            a = store.query(q[0], metric).meta_filter(expr).take(40).collect()
            sa = store.last_query_stats()
            r = loaded.query(q[0], metric).meta_filter(expr).take(40).collect()
            sr = loaded.last_query_stats()
            assert_same_results((r.indices, r.scores), (a.indices, a.scores), f"{metric.name}")
            assert (sa.total_chunks, sa.pruned_chunks, sa.evaluated_chunks, sa.vectors_compared) == (
                sr.total_chunks, sr.pruned_chunks, sr.evaluated_chunks, sr.vectors_compared)
            for name in a.columns:
                assert [a.data[name].get(j) for j in range(len(a.indices))] == [r.data[name].get(j) for j in range(len(r.indices))], name
        assert np.array_equal(store.chunk_mask(expr), loaded.chunk_mask(expr)) and np.array_equal(store.row_mask(expr), loaded.row_mask(expr))
    for name in ("price", "qty", "ts", "score", "big"):
        for x, y in zip(store.zonemap(name), loaded.zonemap(name)):
            assert np.array_equal(x, y, equal_nan=True), name
    assert np.array_equal(store.inv_norms().view(np.uint32), loaded.inv_norms().view(np.uint32))
    # a batch through the tensor-core path of the loaded store (fp32 rows only)
    qb = ora.synth_fill(0, 16, dim, 77)
    a = store.query_batch(qb, ob.Metric.DotProduct).take(64).collect()
    r = loaded.query_batch(qb, ob.Metric.DotProduct).take(64).collect()
    assert_same_results((r.indices, r.scores, r.query_ids), (a.indices, a.scores, a.query_ids), "batch")
    store.close()
    r2 = loaded.query(q[1], ob.Metric.Cosine).take(5).collect()  # the loaded store does not depend on the saved one
    assert len(r2.indices) == 5
    loaded.close()


def test_damaged_files_are_rejected(ctx, tmp_path):
    n, dim = 500, 8
    cols = [ob.Column.from_numpy("a", ob.DataType.Int32, np.arange(n, dtype=np.int32))]
    store = ob.MetaStore.from_columns(cols).with_vectors(ora.synth_fill(0, n, dim, 1)).with_chunk_size(64).with_context(ctx).build()
    path = str(tmp_path / "s.otters")
    store.save(path)
    raw = open(path, "rb").read()
    for name, data, msg in (("short", raw[:-10], "truncated"), ("long", raw + b"x", "truncated"), ("magic", b"NOTOTTER" + raw[8:], "not an otters"),
                            ("version", raw[:8] + (99).to_bytes(4, "little") + raw[12:], "version"), ("empty", b"", "truncated")):
        p = str(tmp_path / name)
        open(p, "wb").write(data)
        with pytest.raises(ob.OttersError, match=msg):
            ob.MetaStore.load(p, ctx)
    with pytest.raises(ob.OttersError, match="cannot open"):
        ob.MetaStore.load(str(tmp_path / "missing"), ctx)
    with pytest.raises(ob.OttersError, match="cannot open"):
        store.save(str(tmp_path / "no_such_dir" / "x"))
    assert ob.MetaStore.load(path, ctx).len() == n
