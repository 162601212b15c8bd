"""Parity of the CUDA path against the CPU oracle on seeded inputs: identical row indices, bit-identical
scores (the kernel reproduces the reference's accumulation order), bit-exact prune/row masks, exact stats.
Everything goes through the C ABI (ctypes) via the public Python mirror."""
import ctypes as C

import numpy as np
import pytest

from helpers import assert_same_results, ob, ora

pytestmark = pytest.mark.gpu

METRICS = [ob.Metric.Cosine, ob.Metric.Euclidean, ob.Metric.DotProduct]


def make_store(vectors):
    s = ob.VecStore(vectors.shape[1])
    s.add_vectors(vectors)
    return s


def run_product(store, q, metric, calls=(), mask=None):
    plan = store.query(q, metric)
    if mask is not None:
        plan = plan.with_row_mask(mask)
    for c in calls:
        plan = getattr(plan, c[0])(*c[1:])
    return plan.collect_arrays()


def run_oracle(vectors, q, metric, tt, k, flt=None, mask=None):
    return ora.vecstore_query(vectors, q, metric, tt, k, flt, mask, ora.CANONICAL)


def test_inv_norms_bit_exact(ctx):
    for n, dim in [(1000, 128), (333, 7), (50, 770), (17, 1)]:
        v = ora.synth_fill(0, n, dim, 11 + dim)
        v[3 % n] = 0.0  # zero row -> 0.0 (src/vec.rs:367)
        got = make_store(v).inv_norms()
        want = ora.inv_norms(v)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))


def test_synthetic_generator_bit_exact(ctx):
    s = ob.VecStore(48)
    s.add_synthetic(100, 500, 0x7735)
    v = ora.synth_fill(100, 500, 48, 0x7735)
    assert np.array_equal(s.inv_norms().view(np.uint32), ora.inv_norms(v).view(np.uint32))
    q = ora.synth_fill(0, 1, 48, 1)
    got = s.query(q[0], ob.Metric.DotProduct).take(500).collect_arrays()
    assert_same_results(got, run_oracle(v, q, ob.Metric.DotProduct, ob.TakeType.Max, 500))


@pytest.mark.parametrize("metric", METRICS, ids=lambda m: m.name)
@pytest.mark.parametrize("n,dim", [(1, 4), (7, 3), (8, 8), (9, 5), (100, 128), (1000, 100), (4097, 768), (20000, 128), (3000, 1536), (700, 2052)])
def test_vecstore_single_query_parity(n, dim, metric, ctx):
    v = ora.synth_fill(0, n, dim, 0x7735 + n)
    q = ora.synth_fill(0, 1, dim, 0xBEEF)
    store = make_store(v)
    for k in sorted({1, min(10, n), min(100, n), n}):
        for tt, call in ((ob.TakeType.Max, "take_max"), (ob.TakeType.Min, "take_min")):
            got = run_product(store, q[0], metric, [(call, k)])
            assert_same_results(got, run_oracle(v, q, metric, tt, k), f"n={n} dim={dim} {metric.name} {call}({k})")


@pytest.mark.parametrize("metric", METRICS, ids=lambda m: m.name)
def test_vecstore_default_take_semantics(metric, ctx):
    """take() infers Min for Euclidean; no take at all => k = n and Max even for Euclidean (src/vec.rs:213-214)."""
    v = ora.synth_fill(0, 300, 24, 5)
    q = ora.synth_fill(0, 1, 24, 6)
    store = make_store(v)
    tt = ob.TakeType.Min if metric == ob.Metric.Euclidean else ob.TakeType.Max
    assert_same_results(run_product(store, q[0], metric, [("take", 17)]), run_oracle(v, q, metric, tt, 17))
    assert_same_results(run_product(store, q[0], metric), run_oracle(v, q, metric, ob.TakeType.Max, 300))
    # the last take*() wins (tests/vec_store_tests.rs:960-980)
    assert_same_results(run_product(store, q[0], metric, [("take", 50), ("take_min", 5)]), run_oracle(v, q, metric, ob.TakeType.Min, 5))


@pytest.mark.parametrize("cmp", list(ob.Cmp), ids=lambda c: c.name)
def test_vec_filter_all_comparators(cmp, ctx):
    v = ora.synth_fill(0, 5000, 64, 21)
    q = ora.synth_fill(0, 1, 64, 22)
    store = make_store(v)
    _, s_all, _ = run_oracle(v, q, ob.Metric.Cosine, ob.TakeType.Max, 5000)
    thr = float(s_all[40]) if cmp == ob.Cmp.Eq else 0.05
    for tt, call in ((ob.TakeType.Max, "take_max"), (ob.TakeType.Min, "take_min")):
        for k in (5, 5000):
            got = run_product(store, q[0], ob.Metric.Cosine, [("filter", thr, cmp), (call, k)])
            want = run_oracle(v, q, ob.Metric.Cosine, tt, k, (thr, cmp))
            assert_same_results(got, want, f"{cmp.name} {call}({k})")
            if cmp == ob.Cmp.Eq:
                assert len(got[0]) >= 1


def test_row_mask_parity(ctx):
    n = 3000
    v = ora.synth_fill(0, n, 40, 31)
    q = ora.synth_fill(0, 1, 40, 32)
    store = make_store(v)
    rng = np.random.default_rng(1)
    for mlen in (0, 1, 63, 64, 65, 1000, n, n + 77):
        mask = rng.random(mlen) < 0.3
        got = run_product(store, q[0], ob.Metric.DotProduct, [("take", 50)], mask)
        want = run_oracle(v, q, ob.Metric.DotProduct, ob.TakeType.Max, 50, None, mask)
        assert_same_results(got, want, f"mask len {mlen}")
        assert all((i >= mlen) or mask[int(i)] for i in got[0])
    none = np.zeros(n, bool)
    assert len(run_product(store, q[0], ob.Metric.DotProduct, [("take", 5)], none)[0]) == 0


@pytest.mark.parametrize("metric", METRICS, ids=lambda m: m.name)
def test_batch_merged_semantics(metric, ctx):
    """One global list over all (row, query) pairs; rows may repeat (src/vec.rs:217-219)."""
    v = ora.synth_fill(0, 2000, 72, 41)
    q = ora.synth_fill(0, 5, 72, 42)
    q[3] = q[1]  # duplicate query: exact (score,row) ties broken by query index
    store = make_store(v)
    for tt, call in ((ob.TakeType.Max, "take_max"), (ob.TakeType.Min, "take_min")):
        for k in (1, 30, 700):
            got = run_product(store, q, metric, [(call, k)])
            want = run_oracle(v, q, metric, tt, k)
            assert_same_results(got, want, f"batch {metric.name} {call}({k})")


def test_large_k_paths(ctx):
    """k > 1024 leaves the fused per-CTA top-k and takes the emit-all + full-sort path."""
    v = ora.synth_fill(0, 6000, 32, 51)
    q = ora.synth_fill(0, 2, 32, 52)
    store = make_store(v)
    for k in (1025, 3000, 6000):
        assert_same_results(run_product(store, q[0], ob.Metric.Cosine, [("take", k)]), run_oracle(v, q[:1], ob.Metric.Cosine, ob.TakeType.Max, k), f"k={k}")
    assert_same_results(run_product(store, q[0], ob.Metric.Euclidean), run_oracle(v, q[:1], ob.Metric.Euclidean, ob.TakeType.Max, 6000), "no take")
    assert_same_results(run_product(store, q, ob.Metric.DotProduct, [("take", 9000)]), run_oracle(v, q, ob.Metric.DotProduct, ob.TakeType.Max, 9000), "batch k=9000")
    assert_same_results(run_product(store, q, ob.Metric.DotProduct, [("filter", 0.0, ob.Cmp.Gt), ("take_min", 2000)]),
                        run_oracle(v, q, ob.Metric.DotProduct, ob.TakeType.Min, 2000, (0.0, ob.Cmp.Gt)), "batch filter k=2000")


def test_ties_and_special_values(ctx):
    v = np.zeros((40, 8), np.float32)
    v[:, 0] = 1.0
    v[5, 0] = np.nan  # NaN score: never returned
    v[9] = 0.0        # zero row: cosine 0
    q = np.zeros((1, 8), np.float32)
    q[0, 0] = 1.0
    store = make_store(v)
    for metric in METRICS:
        got = run_product(store, q[0], metric, [("take_max", 7)])
        want = run_oracle(v, q, metric, ob.TakeType.Max, 7)
        assert_same_results(got, want, f"ties {metric.name}")
        assert 5 not in [int(i) for i in got[0]]
    zq = np.zeros(8, np.float32)  # zero query: every cosine score is 0 -> first k rows
    got = run_product(store, zq, ob.Metric.Cosine, [("take", 4)])
    assert [int(i) for i in got[0]] == [0, 1, 2, 3] and all(s == 0 for s in got[1])


def test_abi_error_strings(ctx):
    """The C ABI itself reproduces the reference's validation messages (src/vec.rs:170-203)."""
    from otters_b200 import _ffi

    store = make_store(ora.synth_fill(0, 10, 3, 1))
    q = np.ones(2, np.float32)
    vq = _ffi.VecQuery()
    vq.queries, vq.nq, vq.dim, vq.metric, vq.take_type, vq.k = q.ctypes.data_as(_ffi.c_f32p), 1, 2, 0, 1, 5
    n = C.c_uint64()
    rc = _ffi.otters_vecstore_query(store._handle(), C.byref(vq), None, None, None, 0, C.byref(n))
    assert rc == 1 and "Query vector length 2 does not match expected dimension 3" in _ffi.last_error()
    vq.nq = 0
    rc = _ffi.otters_vecstore_query(store._handle(), C.byref(vq), None, None, None, 0, C.byref(n))
    assert rc == 1 and _ffi.last_error() == "No queries provided"


# ---------------------------------------------------------------------------------------------------------
# MetaStore
# ---------------------------------------------------------------------------------------------------------
def meta_columns(n, cs, seed, null_frac=0.03):
    rng = np.random.default_rng(seed)
    chunk = np.arange(n) // cs
    def nulls():
        m = rng.random(n) < null_frac
        return m
    price = np.where((chunk // 2) % 2 == 0, 80.0, 10.0) + rng.random(n) * 25.0
    price[rng.random(n) < 0.01] = np.nan  # non-null NaN values: only Neq is true
    f32 = (rng.standard_normal(n) * 3 + (chunk % 5)).astype(np.float32)
    i32 = (chunk % 7 * 10 + rng.integers(0, 10, n)).astype(np.int32)
    i64 = (rng.integers(-5, 5, n) + (chunk % 3) * 1_000_000_000_000).astype(np.int64)
    ts = (1_700_000_000_000 + np.arange(n) * 1000 + rng.integers(-5000, 5000, n)).astype(np.int64)
    item = [f"item_{(c // 3) % 11}" if r > 0.1 else f"rare_{i % 97}" for i, (c, r) in enumerate(zip(chunk, rng.random(n)))]
    cols = [
        ob.Column.from_numpy("price", ob.DataType.Float64, price, nulls()),
        ob.Column.from_numpy("f32", ob.DataType.Float32, f32, nulls()),
        ob.Column.from_numpy("i32", ob.DataType.Int32, i32, nulls()),
        ob.Column.from_numpy("i64", ob.DataType.Int64, i64, nulls()),
        ob.Column.from_numpy("ts", ob.DataType.DateTime, ts, nulls()),
    ]
    inull = nulls()
    inull[cs * 2: cs * 3] = True  # an all-null chunk
    scol = ob.Column.from_numpy("item", ob.DataType.String, ["" if nl else s for s, nl in zip(item, inull)], inull)
    cols.append(scol)
    return cols


FILTERS = [
    lambda: ob.col("price").gt(50.0),
    lambda: ob.col("price").lte(20) & ob.col("i32").gte(30),
    lambda: ob.col("item").eq("item_3"),
    lambda: ob.col("item").neq("item_3") & ob.col("f32").lt(1.5),
    lambda: (ob.col("i32").lt(15) | ob.col("i32").gt(55)) & ob.col("item").neq("rare_5"),
    lambda: ob.col("ts").gte("2023-11-14 22:20:00") & ob.col("ts").lt("2023-11-14 22:40:00"),
    lambda: ob.col("i64").eq(1_000_000_000_003) | ob.col("item").eq("absent-string"),
    lambda: ob.col("price").neq(85.5) & ob.col("f32").neq(0.0),
    lambda: (ob.col("price").gt(50.0) & ob.col("item").eq("item_0")) | (ob.col("i32").eq(42) & ob.col("f32").gte(2.0)),
    lambda: ob.col("i64").lt(0) & ob.col("i32").lte(5),
    lambda: ob.col("item").eq("item_1") | ob.col("item").eq("item_2") | ob.col("item").eq("rare_13"),
    lambda: (ob.col("item").eq("x") | ob.col("item").neq("x")) & ob.col("i32").gt(100000),  # tautology dropped, then nothing passes
]


@pytest.fixture(scope="module")
def meta_pair(ctx):
    n, dim, cs = 10000, 64, 96  # chunk size not a multiple of 32; last chunk short
    vectors = ora.synth_fill(0, n, dim, 61)
    cols = meta_columns(n, cs, 62)
    store = ob.MetaStore.from_columns(cols).with_vectors(vectors).with_chunk_size(cs).build()
    ost = ora.MetaStore(vectors, cols, cs)
    return store, ost, vectors, cols


def test_meta_zonemaps_and_norms_bit_exact(meta_pair):
    store, ost, vectors, cols = meta_pair
    assert store.n_chunks() == ost.n_chunks() == (10000 + 95) // 96
    for i, c in enumerate(cols):
        if c.dtype() == ob.DataType.String:
            continue
        is_f = c.dtype() in (ob.DataType.Float32, ob.DataType.Float64)
        gmn, gmx, gnn = store.zonemap(c.name())
        omn, omx, onn = ost.zonemap(i, is_f)
        assert np.array_equal(gnn, onn)
        live = onn > 0
        assert np.array_equal(gmn[live].view(np.uint64), omn[live].view(np.uint64)) and np.array_equal(gmx[live].view(np.uint64), omx[live].view(np.uint64))
    assert np.array_equal(store.inv_norms().view(np.uint32), ora.inv_norms(vectors).view(np.uint32))


@pytest.mark.parametrize("fi", range(len(FILTERS)))
def test_meta_masks_bit_exact(meta_pair, fi):
    store, ost, _, _ = meta_pair
    expr = FILTERS[fi]()
    fp = ora.FilterPack.from_compiled(expr.compile(store.schema()), store.column_index())
    assert np.array_equal(store.chunk_mask(expr), ost.chunk_mask(fp)), "prune mask differs"
    assert np.array_equal(store.row_mask(expr), ost.row_mask(fp)), "row mask differs"


@pytest.mark.parametrize("metric", METRICS, ids=lambda m: m.name)
@pytest.mark.parametrize("fi", [None] + list(range(len(FILTERS))))
def test_meta_query_parity(meta_pair, fi, metric):
    store, ost, vectors, _ = meta_pair
    q = ora.synth_fill(0, 3, 64, 63)
    expr = FILTERS[fi]() if fi is not None else None
    fp = ora.FilterPack.from_compiled(expr.compile(store.schema()), store.column_index()) if expr is not None else None
    tt = ob.TakeType.Min if metric == ob.Metric.Euclidean else ob.TakeType.Max
    for nq, k, vf in ((1, 10, None), (1, 2000, None), (3, 50, None), (1, 100, (0.02, ob.Cmp.Gt)), (1, None, None)):
        plan = store.query(q[0], metric) if nq == 1 else store.query_batch(q[:nq], metric)
        if expr is not None:
            plan = plan.meta_filter(expr)
        if vf:
            plan = plan.vec_filter(*vf)
        if k is not None:
            plan = plan.take(k)
        res = plan.collect()
        kk = k if k is not None else store.len()
        oi, os_, oq, ostats = ost.query(q[:nq], metric, tt, kk, vf, fp, ora.CANONICAL)
        assert_same_results((res.indices, res.scores, res.query_ids), (oi, os_, oq), f"filter {fi} {metric.name} nq={nq} k={k}")
        st = store.last_query_stats()
        for key in ("total_chunks", "pruned_chunks", "evaluated_chunks", "vectors_compared"):
            assert getattr(st, key) == ostats[key], f"stats.{key}: {getattr(st, key)} != {ostats[key]}"
        # result columns are gathered with NULLs preserved (src/meta.rs:723-821)
        assert res.columns == sorted(store.schema().keys())
        for name in res.columns:
            src = store.columns()[name]
            assert [res.data[name].get(j) is None for j in range(len(res.indices))] == [src.get(i) is None for i in res.indices]


def test_meta_faithful_per_chunk_path_agrees(meta_pair):
    """The reference selects top-k per chunk and merges (src/meta_compute.rs:180, src/meta.rs:702-708); the
    CUDA path selects globally.  Without score ties the two are identical."""
    store, ost, _, _ = meta_pair
    q = ora.synth_fill(0, 1, 64, 64)
    expr = FILTERS[1]()
    fp = ora.FilterPack.from_compiled(expr.compile(store.schema()), store.column_index())
    res = store.query(q[0], ob.Metric.Cosine).meta_filter(expr).take(25).collect()
    oi, os_, _, _ = ost.query(q, ob.Metric.Cosine, ob.TakeType.Max, 25, None, fp, ora.FAITHFUL)
    assert_same_results((res.indices, res.scores), (oi, os_))


def test_meta_error_and_quirk_semantics(meta_pair):
    store, _, _, _ = meta_pair
    q = ora.synth_fill(0, 1, 64, 65)[0]
    with pytest.raises(ob.OttersError) as ei:
        store.query(q, ob.Metric.Cosine).meta_filter(ob.col("i32").gt(1.5)).take(3).collect()
    assert str(ei.value).startswith("meta_filter compile error: Type mismatch for column 'i32'")
    with pytest.raises(ob.OttersError) as ei:
        store.query(q, ob.Metric.Cosine).meta_filter(ob.col("nope").eq(1)).take(3).collect()
    assert "Unknown column 'nope'" in str(ei.value)
    # wrong-dimension query: the per-chunk errors are swallowed -> Ok(empty) with stats (src/meta_compute.rs:182)
    res = store.query(np.ones(5, np.float32), ob.Metric.Cosine).take(3).collect()
    assert res.is_empty()
    st = store.last_query_stats()
    assert st.evaluated_chunks == st.total_chunks == store.n_chunks() and st.vectors_compared == store.len()
    assert store.query(q, ob.Metric.Cosine).take(0).collect().is_empty()


def test_meta_small_chunks_and_bloom_knobs(ctx):
    n, dim = 777, 20
    vectors = ora.synth_fill(0, n, dim, 71)
    q = ora.synth_fill(0, 1, dim, 72)
    for cs, bloom in ((1, ("fpr", 0.01)), (5, ("bits", 64)), (8, ("fpr", 0.5)), (1024, ("bits", 4096)), (100000, ("fpr", 0.2))):
        cols = meta_columns(n, cs, 73)
        b = ob.MetaStore.from_columns(cols).with_vectors(vectors).with_chunk_size(cs)
        b = b.with_bloom_fpr(bloom[1]) if bloom[0] == "fpr" else b.with_bloom_bits(bloom[1])
        store = b.build()
        ost = ora.MetaStore(vectors, cols, cs, bloom)
        for fi in (2, 4, 6, 10):
            expr = FILTERS[fi]()
            fp = ora.FilterPack.from_compiled(expr.compile(store.schema()), store.column_index())
            assert np.array_equal(store.chunk_mask(expr), ost.chunk_mask(fp)), f"cs={cs} filter {fi}"
            res = store.query(q[0], ob.Metric.DotProduct).meta_filter(expr).take(20).collect()
            oi, os_, _, ostats = ost.query(q, ob.Metric.DotProduct, ob.TakeType.Max, 20, None, fp)
            assert_same_results((res.indices, res.scores), (oi, os_), f"cs={cs} filter {fi}")
            assert store.last_query_stats().evaluated_chunks == ostats["evaluated_chunks"]


def test_meta_empty_store(ctx):
    store = ob.MetaStore.from_columns([]).with_vectors(np.zeros((0, 4), np.float32)).build()
    assert store.query([1, 0, 0, 0], ob.Metric.Cosine).take(3).collect().is_empty()
    st = store.last_query_stats()
    assert st.total_chunks == 0 and st.vectors_compared == 0


def test_fused_and_unfused_predicate_paths_agree(meta_pair):
    """The row predicate normally runs as its own kernel (K0b row bitmask); evaluated per work unit inside the scan kernel
    (fused K0b) it must give the same answer."""
    store, ost, _, _ = meta_pair
    q = ora.synth_fill(0, 1, 64, 66)
    try:
        for fi in range(len(FILTERS)):
            expr = FILTERS[fi]()
            got = []
            for disable in (1, 2):
                store.ctx.set_tuning(disable_fused_predicate=disable)
                res = store.query(q[0], ob.Metric.Cosine).meta_filter(expr).take(40).collect()
                got.append((res.indices, res.scores))
            assert_same_results(got[0], got[1], f"filter {fi}")
    finally:
        store.ctx.set_tuning()


def test_many_leaf_filter_falls_back_to_row_mask_kernel(meta_pair):
    """A CNF too large for the scan kernel's shared-memory staging uses the stand-alone row-mask kernel."""
    store, ost, _, _ = meta_pair
    q = ora.synth_fill(0, 1, 64, 67)
    expr = ob.col("i32").gte(0)
    for v in range(140):
        expr = expr & (ob.col("i32").neq(1000 + v) | ob.col("price").gt(1e9))
    fp = ora.FilterPack.from_compiled(expr.compile(store.schema()), store.column_index())
    res = store.query(q[0], ob.Metric.DotProduct).meta_filter(expr).take(30).collect()
    oi, os_, _, ostats = ost.query(q, ob.Metric.DotProduct, ob.TakeType.Max, 30, None, fp)
    assert_same_results((res.indices, res.scores), (oi, os_))
    assert store.last_query_stats().evaluated_chunks == ostats["evaluated_chunks"]
