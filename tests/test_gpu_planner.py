"""K1's planner front-end (otters_b200/csrc/scan_planner.cu: planner warps publish 16-row tiles of surviving rows, worker
warps stream them) must give exactly what the autonomous-warp front-end and the oracle give.  The whole GPU suite can
also be run against it with OTTERS_SCAN_MODE=2; these cases keep it covered in the default run."""
import numpy as np
import pytest

from helpers import assert_same_results, ob, ora

pytestmark = pytest.mark.gpu

METRICS = [ob.Metric.Cosine, ob.Metric.Euclidean, ob.Metric.DotProduct]


@pytest.fixture(params=[1, 4], ids=["one_planner", "four_planners"])
def pctx(ctx, request):
    ctx.set_tuning(scan_mode=2, planners=request.param, batch_mode=2)
    yield ctx
    ctx.set_tuning()


def make_store(vectors):
    s = ob.VecStore(vectors.shape[1])
    s.add_vectors(vectors)
    return s


@pytest.mark.parametrize("metric", METRICS, ids=lambda m: m.name)
@pytest.mark.parametrize("n,dim", [(1, 4), (9, 5), (100, 128), (4097, 768), (20000, 128), (3000, 1536), (700, 2052), (60000, 32)])
def test_planner_vecstore_parity(n, dim, metric, pctx):
    v = ora.synth_fill(0, n, dim, 0x7735 + n)
    q = ora.synth_fill(0, 1, dim, 0xBEEF)
    store = make_store(v)
    for k in sorted({1, min(100, n), min(1024, n)}):
        for tt, call in ((ob.TakeType.Max, "take_max"), (ob.TakeType.Min, "take_min")):
            got = getattr(store.query(q[0], metric), call)(k).collect_arrays()
            want = ora.vecstore_query(v, q, metric, tt, k, None, None, ora.CANONICAL)
            assert_same_results(got, want, f"n={n} dim={dim} {metric.name} {call}({k})")


def test_planner_row_mask_filter_and_batch(pctx):
    n = 5000
    v = ora.synth_fill(0, n, 40, 31)
    q = ora.synth_fill(0, 3, 40, 32)
    store = make_store(v)
    rng = np.random.default_rng(1)
    for mlen in (0, 63, 1000, n, n + 77):
        mask = rng.random(mlen) < 0.3
        got = store.query(q[0], ob.Metric.DotProduct).with_row_mask(mask).take(50).collect_arrays()
        want = ora.vecstore_query(v, q[:1], ob.Metric.DotProduct, ob.TakeType.Max, 50, None, mask, ora.CANONICAL)
        assert_same_results(got, want, f"mask len {mlen}")
    got = store.query(q[0], ob.Metric.Cosine).filter(0.1, ob.Cmp.Gt).take_min(40).collect_arrays()
    assert_same_results(got, ora.vecstore_query(v, q[:1], ob.Metric.Cosine, ob.TakeType.Min, 40, (0.1, ob.Cmp.Gt), None, ora.CANONICAL), "filter")
    got = store.query(q, ob.Metric.Euclidean).take(70).collect_arrays()  # batch on the per-query path (running threshold)
    assert_same_results(got, ora.vecstore_query(v, q, ob.Metric.Euclidean, ob.TakeType.Min, 70, None, None, ora.CANONICAL), "batch")


@pytest.mark.parametrize("chunk", [4, 100, 256, 1024])
def test_planner_metastore_parity(chunk, pctx):
    n, dim = 9000, 24
    rng = np.random.default_rng(chunk)
    v = ora.synth_fill(0, n, dim, 77)
    price = ob.Column.from_numpy("price", ob.DataType.Float64, np.where((np.arange(n) // max(chunk, 50)) % 2 == 0, 80.0, 10.0) + rng.random(n) * 20,
                                 rng.random(n) < 0.02)
    qty = ob.Column.from_numpy("qty", ob.DataType.Int32, rng.integers(0, 50, n).astype(np.int32), rng.random(n) < 0.02)
    ts = ob.Column.from_numpy("ts", ob.DataType.DateTime, (1_700_000_000_000 + np.arange(n) * 1000).astype(np.int64))
    w = ob.Column.from_numpy("w", ob.DataType.Float32, rng.random(n).astype(np.float32))
    big = ob.Column.from_numpy("big", ob.DataType.Int64, rng.integers(-10**12, 10**12, n).astype(np.int64))
    item = ob.Column.from_categories("item", [f"i{i}" for i in range(20)], rng.integers(0, 20, n), rng.random(n) < 0.02)
    cols = [price, qty, ts, w, big, item]
    store = ob.MetaStore.from_columns(cols).with_vectors(v).with_chunk_size(chunk).build()
    ost = ora.MetaStore(v, cols, chunk)
    q = ora.synth_fill(0, 1, dim, 78)
    exprs = [
        ob.col("price").lt(50.0) & ob.col("qty").gte(10),
        (ob.col("price").gt(85.0) | ob.col("item").eq("i3")) & ob.col("ts").gte("2023-11-14 23:00:00"),
        ob.col("w").lt(0.5) & ob.col("big").gt(0) & ob.col("item").neq("i7") & ob.col("qty").lt(40),
        # seven leaves: more than the planner's load-everything-first form takes, so it evaluates row by row
        (ob.col("price").lt(95.0) | ob.col("w").gt(0.9)) & (ob.col("qty").gte(1) | ob.col("big").lt(0)) & ob.col("item").neq("i1") & ob.col("ts").lt("2023-11-15 02:00:00") & ob.col("w").lte(0.99),
    ]
    for ei, expr in enumerate(exprs):
        res = store.query(q[0], ob.Metric.Cosine).meta_filter(expr).take(60).collect()
        st = store.last_query_stats()
        fp = ora.FilterPack.from_compiled(expr.compile(store.schema()), store.column_index())
        oi, os_, _, ostats = ost.query(q, ob.Metric.Cosine, ob.TakeType.Max, 60, None, fp, ora.CANONICAL)
        assert_same_results((np.array(res.indices), np.array(res.scores, np.float32)), (oi, os_), f"chunk {chunk} expr {ei}")
        assert (st.total_chunks, st.pruned_chunks, st.evaluated_chunks, st.vectors_compared) == (
            ostats["total_chunks"], ostats["pruned_chunks"], ostats["evaluated_chunks"], ostats["vectors_compared"])
