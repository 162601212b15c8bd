"""Stores that keep their rows as bf16 (OTTERS_VECTORS_FMT_BF16; the reference's roadmap item "Quantization for vectors",
README.md:208).  Parity contract: the reference's arithmetic applied to the ROUNDED rows — the CPU
oracle run on f32(bf16_rn(x)) must give identical rows and bit-identical scores, inverse norms, masks and statistics.
Queries stay fp32.  Every placement of the row predicate and both K1 front-ends are held to the same bytes."""
import numpy as np
import pytest

from helpers import assert_same_results, ob, ora

pytestmark = pytest.mark.gpu

METRICS = [ob.Metric.Cosine, ob.Metric.Euclidean, ob.Metric.DotProduct]
BF16 = ob.VectorFormat.Bf16


def make_store(vectors, ctx=None):
    s = ob.VecStore(vectors.shape[1], ctx, BF16)
    s.add_vectors(vectors)
    return s


def run_oracle(vectors, q, metric, tt, k, flt=None, mask=None):
    return ora.vecstore_query(ora.round_bf16(vectors), q, metric, tt, k, flt, mask, ora.CANONICAL)


def test_round_to_bf16_is_nearest_even():
    x = np.array([1.0, 1.00390625, 1.01171875, -1.00390625, 3.0e38, 0.0, -0.0, np.inf], np.float32)
    # 1 + 2^-8 is a tie between 1 and 1 + 2^-7: even mantissa (1.0) wins; 1 + 3 * 2^-8 ties upwards to 1 + 2^-6
    want = np.array([1.0, 1.0, 1.015625, -1.0, 3.0e38, 0.0, -0.0, np.inf], np.float32)
    got = ob.round_to_bf16(x)
    assert np.array_equal(got[[0, 1, 2, 3, 5, 6, 7]], want[[0, 1, 2, 3, 5, 6, 7]]) and abs(float(got[4]) - 3.0e38) <= 3.0e38 * 2.0 ** -8
    assert (got.view(np.uint32) & 0xFFFF == 0).all()
    assert np.isnan(ob.round_to_bf16(np.array([np.nan], np.float32))[0])


def test_inv_norms_are_those_of_the_rounded_rows(ctx):
    for n, dim in [(1000, 128), (333, 7), (50, 770), (17, 1)]:
        v = ora.synth_fill(0, n, dim, 11 + dim)
        v[3 % n] = 0.0
        got = make_store(v).inv_norms()
        assert np.array_equal(got.view(np.uint32), ora.inv_norms(ora.round_bf16(v)).view(np.uint32))


@pytest.mark.parametrize("metric", METRICS, ids=lambda m: m.name)
@pytest.mark.parametrize("n,dim", [(1, 4), (7, 3), (9, 5), (100, 128), (1000, 100), (4097, 768), (20000, 128), (3000, 1536), (700, 2052)])
def test_bf16_vecstore_single_query_parity(n, dim, metric, ctx):
    v = ora.synth_fill(0, n, dim, 0x7735 + n)
    q = ora.synth_fill(0, 1, dim, 0xBEEF)
    store = make_store(v)
    assert store.vector_format == BF16
    for k in sorted({1, min(10, n), min(100, n), n}):  # k = n > 1024 takes the emit-all + sort path
        for tt, call in ((ob.TakeType.Max, "take_max"), (ob.TakeType.Min, "take_min")):
            got = getattr(store.query(q[0], metric), call)(k).collect_arrays()
            assert_same_results(got, run_oracle(v, q, metric, tt, k), f"bf16 n={n} dim={dim} {metric.name} {call}({k})")


@pytest.mark.parametrize("mode", [1, 2], ids=["autonomous", "planner"])
def test_bf16_both_front_ends_filters_masks_and_batches(mode, ctx):
    n, dim = 6000, 200
    v = ora.synth_fill(0, n, dim, 77)
    q = ora.synth_fill(0, 5, dim, 78)
    store = make_store(v)
    ctx.set_tuning(scan_mode=mode)
    try:
        rng = np.random.default_rng(2)
        mask = rng.random(n) < 0.4
        for metric in METRICS:
            tt = ob.TakeType.Min if metric == ob.Metric.Euclidean else ob.TakeType.Max
            got = store.query(q[0], metric).with_row_mask(mask).take(40).collect_arrays()
            assert_same_results(got, run_oracle(v, q[:1], metric, tt, 40, None, mask), f"mask {metric.name}")
        for cmp in ob.Cmp:
            thr = 0.02
            got = store.query(q[1], ob.Metric.Cosine).filter(thr, cmp).take(64).collect_arrays()
            assert_same_results(got, run_oracle(v, q[1:2], ob.Metric.Cosine, ob.TakeType.Max, 64, (thr, cmp)), f"filter {cmp.name}")
        # five queries are too few for the tensor-core kernel in automatic mode: query by query, one merged list, ties to the lower query
        got = store.query(q, ob.Metric.DotProduct).take(100).collect_arrays()
        assert ctx.last_work()["batch_used"] == 0
        assert_same_results(got, run_oracle(v, q, ob.Metric.DotProduct, ob.TakeType.Max, 100), "merged batch")
        per = store.query(q, ob.Metric.Cosine).take(15).collect_per_query()
        for i in range(len(q)):
            want = run_oracle(v, q[i : i + 1], ob.Metric.Cosine, ob.TakeType.Max, 15)
            assert_same_results((per[i][0], per[i][1]), want, f"per-query list {i}")
    finally:
        ctx.set_tuning()


def test_bf16_synthetic_rows_and_set_rows(ctx):
    dim = 96
    s = ob.VecStore(dim, ctx, BF16)
    s.add_synthetic(100, 3000, 0x7735)
    s.add_synthetic(3100, 500, 0x7735)  # appended: the generator continues at the absolute row id
    v = ora.synth_fill(100, 3500, dim, 0x7735)
    q = ora.synth_fill(0, 1, dim, 5)
    assert np.array_equal(s.inv_norms().view(np.uint32), ora.inv_norms(ora.round_bf16(v)).view(np.uint32))
    got = s.query(q[0], ob.Metric.Cosine).take(50).collect_arrays()
    assert_same_results(got, run_oracle(v, q, ob.Metric.Cosine, ob.TakeType.Max, 50), "synthetic bf16")
    # planted rows are rounded like every other row
    plant = (q[0] * np.float32(1.0009765625)).astype(np.float32)
    s.set_rows([17, 2999], np.stack([plant, -plant]))
    v[17], v[2999] = plant, -plant
    got = s.query(q[0], ob.Metric.Cosine).take(3).collect_arrays()
    assert got[0][0] == 17
    assert_same_results(got, run_oracle(v, q, ob.Metric.Cosine, ob.TakeType.Max, 3), "set_rows bf16")
    assert np.array_equal(s.inv_norms().view(np.uint32), ora.inv_norms(ora.round_bf16(v)).view(np.uint32))


@pytest.mark.parametrize("pred", [0, 2], ids=["rowmask_kernel", "predicate_in_scan"])
def test_bf16_metastore_filtered_query(pred, ctx):
    n, dim, cs = 20000, 128, 256
    v = ora.synth_fill(0, n, dim, 0x7735)
    price = ob.Column.from_numpy("price", ob.DataType.Float64, np.where((np.arange(n) // cs) % 2 == 0, 80.0, 10.0) + (np.arange(n) % 20))
    version = ob.Column.from_numpy("version", ob.DataType.Int32, np.where((np.arange(n) // cs) % 3 == 0, 1, 3).astype(np.int32))
    store = (ob.MetaStore.from_columns([price, version]).with_vectors(v).with_chunk_size(cs).with_vector_format(BF16)
             .with_context(ctx).build())
    assert store.vector_format() == BF16
    q = ora.synth_fill(0, 3, dim, 0xBEEF)
    expr = ob.col("price").lt(50.0) & ob.col("version").gte(2)
    ost = ora.MetaStore(ora.round_bf16(v), [price, version], cs)
    fp = ora.FilterPack.from_compiled(expr.compile(store.schema()), store.column_index())
    ctx.set_tuning(disable_fused_predicate=pred)
    try:
        for metric in METRICS:
            tt = ob.TakeType.Min if metric == ob.Metric.Euclidean else ob.TakeType.Max
            res = store.query(q[0], metric).meta_filter(expr).take(25).collect()
            st = store.last_query_stats()
            oi, os_, _, ostats = ost.query(q[:1], metric, tt, 25, None, fp)
            assert_same_results((res.indices, res.scores), (oi, os_), f"bf16 metastore {metric.name}")
            assert (st.total_chunks, st.pruned_chunks, st.evaluated_chunks, st.vectors_compared) == (
                ostats["total_chunks"], ostats["pruned_chunks"], ostats["evaluated_chunks"], ostats["vectors_compared"])
        res = store.query_batch(q, ob.Metric.Cosine).meta_filter(expr).vec_filter(0.0, ob.Cmp.Gt).take(60).collect()
        oi, os_, oq, _ = ost.query(q, ob.Metric.Cosine, ob.TakeType.Max, 60, (0.0, ob.Cmp.Gt), fp)
        assert_same_results((res.indices, res.scores, res.query_ids), (oi, os_, oq), "bf16 metastore batch")
    finally:
        ctx.set_tuning()
    assert np.array_equal(store.inv_norms().view(np.uint32), ora.inv_norms(ora.round_bf16(v)).view(np.uint32))


def test_bf16_store_batches_on_the_tensor_cores(ctx):
    """K2 on a bf16 store: its rows ARE the bf16 operand of the kind::f16 contraction (no shadow copy) and the exact re-scoring
    reads the same rows; one rung only — a declined certificate goes to the streaming kernel."""
    v = ora.synth_fill(0, 3000, 768, 0x7735)
    q = ora.synth_fill(0, 1024, 768, 0xBEEF)
    store = make_store(v, ctx)
    ctx.set_tuning(batch_mode=1)
    try:
        for metric in METRICS:
            got = store.query(q, metric).take_max(100).collect_arrays()
            w = ctx.last_work()
            assert w["batch_used"] == 1 and w["batch_passes"] == 2 and w["batch_attempts"] == 1 and w["batch_max_err"] <= w["batch_delta"], w
            assert_same_results(got, run_oracle(v, q, metric, ob.TakeType.Max, 100), f"bf16 store K2 {metric.name}")
        # small, crowded stores: whatever path answers, the result is the oracle's (dims with a ragged last k-block too)
        for n, dim, nq in [(300, 24, 2), (1000, 7, 9), (4097, 100, 33), (700, 1536, 16)]:
            v2 = ora.synth_fill(0, n, dim, 11 + n)
            q2 = ora.synth_fill(0, nq, dim, 12 + nq)
            s2 = make_store(v2, ctx)
            for tt, call in ((ob.TakeType.Max, "take_max"), (ob.TakeType.Min, "take_min")):
                got = getattr(s2.query(q2, ob.Metric.Cosine), call)(30).collect_arrays()
                w = ctx.last_work()
                assert w["batch_used"] == 1 or w["batch_fallback"] == 1, w
                assert_same_results(got, run_oracle(v2, q2, ob.Metric.Cosine, tt, 30), f"bf16 store K2 n={n} dim={dim} {call}")
    finally:
        ctx.set_tuning()


def test_bf16_scan_streams_half_the_bytes(ctx):
    n, dim = 50000, 256
    v = ora.synth_fill(0, n, dim, 3)
    q = ora.synth_fill(0, 1, dim, 4)
    full, half = ob.VecStore(dim, ctx), make_store(v, ctx)
    full.add_vectors(v)
    full.query(q[0], ob.Metric.DotProduct).take(10).collect_arrays()
    b_full = ctx.last_work()["scan_bytes"]
    half.query(q[0], ob.Metric.DotProduct).take(10).collect_arrays()
    b_half = ctx.last_work()["scan_bytes"]
    assert b_full == n * dim * 4 and b_half == n * dim * 2
