"""Parity of the batched tensor-core path (K2: TMA-fed tcgen05 selection — bf16 shadow rows, single-pass tf32, then 3xTF32 — + exact re-scoring,
otters_b200/csrc/batched.cu) against the CPU oracle: one merged list over all (row, query) pairs
(reference src/vec.rs:217-219, :243-266), identical rows and query ids, bit-identical scores.
The tests force the tensor-core kernel (batch_mode=1) and check that it — not the per-query fallback —
produced the result, except where the test is about the fallback itself."""
import numpy as np
import pytest

from helpers import assert_same_results, ob, ora

pytestmark = pytest.mark.gpu

METRICS = [ob.Metric.Cosine, ob.Metric.Euclidean, ob.Metric.DotProduct]


@pytest.fixture(params=[(1, 3), (2, 3), (1, 0), (2, 0), (1, 2), (2, 2)],
                ids=["single_cta_3x", "cta_pair_3x", "single_cta_auto", "cta_pair_auto", "single_cta_bf16", "cta_pair_bf16"])
def bctx(ctx, request):
    """Forces the tensor-core kernel: single CTAs or CTA pairs (tcgen05 cta_group::2), each with the 3xTF32 contraction only,
    with the automatic ladder (bf16 selection first, then single-pass tf32, then 3xTF32, each when the certificate of the one
    before fails) and with the bf16 rung only (whose wide error band often declines on these small stores: the exact per-query
    path then answers, and the result must be the oracle's either way)."""
    cg, passes = request.param
    ctx.set_tuning(batch_mode=1, batch_cta_group=cg, batch_passes=passes)
    ctx.batch_passes_forced = passes
    yield ctx
    ctx.set_tuning()


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


def used_tensor_path(ctx, allow_fallback=False):
    w = ctx.last_work()
    forced = getattr(ctx, "batch_passes_forced", 0)
    if forced == 2:  # bf16 only: one attempt, accepted or answered exactly
        assert w["batch_attempts"] == 1 and (w["batch_used"] == 1 or w["batch_fallback"] == 1), w
        if w["batch_used"]:
            assert w["batch_passes"] == 2 and w["batch_max_err"] <= w["batch_delta"], w
        return w
    if allow_fallback:
        assert w["batch_used"] == 1 or w["batch_fallback"] == 1, w
    else:
        assert w["batch_used"] == 1 and w["batch_fallback"] == 0, w
        assert w["batch_passes"] == 3 if forced == 3 else w["batch_passes"] in (1, 2, 3), w
        # automatic ladder: every declined rung costs one run, and the store then skips it for its next batches
        assert 1 <= w["batch_attempts"] <= {2: 1, 1: 2, 3: 1 if forced == 3 else 3}[w["batch_passes"]], w
        # measured error against the rigorous bound: 3xTF32 stays far below it; single-pass truncation of the stored rows
        # is one-sided, so aligned vectors (a query that is a row) come to about a fifth of the bound; bf16 rounding is
        # two-sided but its bound carries no slack on the operand term
        margin = {3: 0.25, 1: 0.5, 2: 1.0}[w["batch_passes"]]
        assert w["batch_max_err"] <= margin * w["batch_delta"], f"tensor-core error {w['batch_max_err']} too close to the bound {w['batch_delta']}"
    return w


@pytest.mark.parametrize("metric", METRICS, ids=lambda m: m.name)
@pytest.mark.parametrize("n,dim,nq", [(300, 24, 2), (2000, 72, 5), (1000, 7, 9), (4097, 100, 33), (5000, 768, 64), (20000, 128, 300),
                                       (129, 32, 256), (128, 33, 257), (700, 1536, 16)])
def test_batched_parity(n, dim, nq, metric, bctx):
    v = ora.synth_fill(0, n, dim, 0x7735 + n)
    q = ora.synth_fill(0, nq, dim, 0xBEEF + nq)
    store = make_store(v)
    for k in sorted({1, 30, min(700, n), min(1024, n * nq)}):
        for tt, call in ((ob.TakeType.Max, "take_max"), (ob.TakeType.Min, "take_min")):
            got = run_product(store, q, metric, [(call, k)])
            # 33k pairs of 33-d vectors put many scores inside the (rigorous, hence conservative) error band around
            # the k-th one: there the library may decline to certify the tensor-core selection and answer exactly
            used_tensor_path(bctx, allow_fallback=(n, dim, nq) == (128, 33, 257))
            want = run_oracle(v, q, metric, tt, k)
            assert_same_results(got, want, f"n={n} dim={dim} nq={nq} {metric.name} {call}({k})")


def test_batched_many_queries(bctx):
    """Four query tiles (1024 queries), dot product, top-100: BASELINE config 2 at reduced row count."""
    v = ora.synth_fill(0, 3000, 768, 0x7735)
    q = ora.synth_fill(0, 1024, 768, 0xBEEF)
    store = make_store(v)
    got = run_product(store, q, ob.Metric.DotProduct, [("take", 100)])
    used_tensor_path(bctx)
    assert_same_results(got, run_oracle(v, q, ob.Metric.DotProduct, ob.TakeType.Max, 100), "1024 queries")


def test_single_pass_selection_is_accepted_on_separated_scores(ctx):
    """3000 x 768 x 1024 queries, top-100: every CTA's best excluded pair lies far below the 100th score, so the first rung
    of the ladder (bf16) certifies on its first attempt, and so does every forced rung; all return the same bytes."""
    v = ora.synth_fill(0, 3000, 768, 0x7735)
    q = ora.synth_fill(0, 1024, 768, 0xBEEF)
    store = make_store(v)
    out = {}
    for passes in (0, 1, 2, 3):
        ctx.set_tuning(batch_mode=1, batch_passes=passes)
        for metric in METRICS:
            out[passes, metric] = run_product(store, q, metric, [("take_max", 100)])
            w = ctx.last_work()
            assert w["batch_used"] == 1 and w["batch_attempts"] == 1 and w["batch_passes"] == {0: 2, 1: 1, 2: 2, 3: 3}[passes], (passes, metric, w)
            assert w["batch_max_err"] <= w["batch_delta"]
            if passes == 2:  # the bf16 bound is not vacuous either: random data comes within two orders of magnitude of it
                assert w["batch_max_err"] >= 1e-3 * w["batch_delta"], w
    ctx.set_tuning()
    for metric in METRICS:
        want = run_oracle(v, q, metric, ob.TakeType.Max, 100)
        for passes in (0, 1, 2, 3):
            assert_same_results(out[passes, metric], want, f"passes={passes} {metric.name}")


def test_single_pass_backs_off_after_a_failed_certificate(ctx):
    """Near-ties inside the bf16 and single-pass error bands: the ladder redoes the batch with 3xTF32 and the store then skips
    the rungs that declined for its next batches.  Results stay the oracle's either way."""
    rng = np.random.default_rng(3)
    base = rng.standard_normal((1, 64)).astype(np.float32)
    v = (base + np.float32(1e-4) * rng.standard_normal((600, 64)).astype(np.float32)).astype(np.float32)  # 600 near-copies
    q = rng.standard_normal((8, 64)).astype(np.float32)
    store = make_store(v)
    ctx.set_tuning(batch_mode=1)
    want = run_oracle(v, q, ob.Metric.DotProduct, ob.TakeType.Max, 10)
    got = run_product(store, q, ob.Metric.DotProduct, [("take", 10)])
    first = ctx.last_work()
    assert_same_results(got, want, "first batch")
    got = run_product(store, q, ob.Metric.DotProduct, [("take", 10)])
    second = ctx.last_work()
    assert_same_results(got, want, "second batch")
    ctx.set_tuning()
    if first["batch_passes"] not in (1, 2):  # both wide rungs declined (expected for this data): the next batch must not try them again
        assert first["batch_attempts"] == 3, first
        assert second["batch_attempts"] == 1 and second["batch_passes"] not in (1, 2), second


def test_batched_duplicate_queries_tie_on_query_index(bctx):
    v = ora.synth_fill(0, 2000, 72, 41)
    q = ora.synth_fill(0, 12, 72, 42)
    q[7] = q[1]
    q[11] = q[1]
    store = make_store(v)
    for metric in METRICS:
        got = run_product(store, q, metric, [("take", 90)])
        used_tensor_path(bctx)
        tt = ob.TakeType.Min if metric == ob.Metric.Euclidean else ob.TakeType.Max
        assert_same_results(got, run_oracle(v, q, metric, tt, 90), f"dup queries {metric.name}")


@pytest.mark.parametrize("cmp", list(ob.Cmp), ids=lambda c: c.name)
def test_batched_vec_filter(cmp, bctx):
    v = ora.synth_fill(0, 5000, 64, 21)
    q = ora.synth_fill(0, 10, 64, 22)
    store = make_store(v)
    _, s_all, _ = run_oracle(v, q, ob.Metric.Cosine, ob.TakeType.Max, 50000)
    thr = float(s_all[40]) if cmp == ob.Cmp.Eq else 0.25
    for tt, call in ((ob.TakeType.Max, "take_max"), (ob.TakeType.Min, "take_min")):
        for k in (5, 1000):
            got = run_product(store, q, ob.Metric.Cosine, [("filter", thr, cmp), (call, k)])
            used_tensor_path(bctx, allow_fallback=True)
            want = run_oracle(v, q, ob.Metric.Cosine, tt, k, (thr, cmp))
            assert_same_results(got, want, f"{cmp.name} {call}({k})")
            if cmp == ob.Cmp.Eq:
                assert len(got[0]) >= 1


def test_batched_filter_with_few_survivors_stays_on_tensor_path(bctx):
    """vec_filter leaves fewer than k pairs: nothing is ever excluded by the top-k cut, so the result is verified."""
    v = ora.synth_fill(0, 4000, 96, 61)
    q = ora.synth_fill(0, 20, 96, 62)
    store = make_store(v)
    got = run_product(store, q, ob.Metric.Cosine, [("filter", 0.3, ob.Cmp.Gt), ("take", 1000)])
    used_tensor_path(bctx)
    want = run_oracle(v, q, ob.Metric.Cosine, ob.TakeType.Max, 1000, (0.3, ob.Cmp.Gt))
    assert 0 < len(want[0]) < 1000
    assert_same_results(got, want, "few survivors")


def test_batched_row_mask(bctx):
    n = 3000
    v = ora.synth_fill(0, n, 40, 31)
    q = ora.synth_fill(0, 8, 40, 32)
    store = make_store(v)
    rng = np.random.default_rng(1)
    for mlen in (0, 63, 1000, n, n + 77):
        mask = rng.random(mlen) < 0.3
        got = run_product(store, q, ob.Metric.DotProduct, [("take", 50)], mask)
        used_tensor_path(bctx)
        want = run_oracle(v, q, ob.Metric.DotProduct, ob.TakeType.Max, 50, None, mask)
        assert_same_results(got, want, f"mask len {mlen}")
    # whole 128-row tiles masked out are skipped by every warp role
    mask = np.zeros(n, bool)
    mask[1500:1510] = True
    got = run_product(store, q, ob.Metric.Cosine, [("take", 50)], mask)
    used_tensor_path(bctx)
    assert_same_results(got, run_oracle(v, q, ob.Metric.Cosine, ob.TakeType.Max, 50, None, mask), "sparse mask")
    assert len(run_product(store, q, ob.Metric.Cosine, [("take", 5)], np.zeros(n, bool))[0]) == 0


def test_batched_near_duplicate_rows_fall_back_or_verify(bctx):
    """Rows that differ in the last bits put many exact scores inside the tensor-core error band around the
    k-th score; whatever path answers, the result must be the oracle's."""
    base = ora.synth_fill(0, 1, 64, 71)[0]
    v = np.tile(base, (600, 1)).astype(np.float32)
    v[:, 0] += np.arange(600, dtype=np.float32) * np.float32(1e-7)
    q = ora.synth_fill(0, 6, 64, 72)
    q[0] = base
    store = make_store(v)
    for metric in METRICS:
        tt = ob.TakeType.Min if metric == ob.Metric.Euclidean else ob.TakeType.Max
        got = run_product(store, q, metric, [("take", 40)])
        used_tensor_path(bctx, allow_fallback=True)
        assert_same_results(got, run_oracle(v, q, metric, tt, 40), f"near duplicates {metric.name}")


def test_batched_special_values(bctx):
    v = ora.synth_fill(0, 500, 16, 81)
    v[3] = 0.0            # zero row: cosine 0.0
    v[7, 2] = np.inf      # non-finite scores: the exact path decides
    v[9, 1] = np.nan      # NaN scores are never returned
    q = ora.synth_fill(0, 4, 16, 82)
    q[2] = 0.0            # zero query
    store = make_store(v)
    for metric in METRICS:
        tt = ob.TakeType.Min if metric == ob.Metric.Euclidean else ob.TakeType.Max
        got = run_product(store, q, metric, [("take", 60)])
        used_tensor_path(bctx, allow_fallback=True)
        assert_same_results(got, run_oracle(v, q, metric, tt, 60), f"special values {metric.name}")


def test_batched_metastore_query_batch(bctx):
    n, dim, cs = 6000, 48, 256
    v = ora.synth_fill(0, n, dim, 91)
    price = ob.Column.from_numpy("price", ob.DataType.Float64, np.where((np.arange(n) // cs) % 2 == 0, 80.0, 10.0) + (np.arange(n) % 20))
    version = ob.Column.from_numpy("version", ob.DataType.Int32, (np.arange(n) % 5).astype(np.int32))
    store = ob.MetaStore.from_columns([price, version]).with_vectors(v).with_chunk_size(cs).build()
    q = ora.synth_fill(0, 24, dim, 92)
    expr = ob.col("price").lt(50.0) & ob.col("version").gte(2)
    res = store.query_batch(q, ob.Metric.Cosine).meta_filter(expr).vec_filter(0.0, ob.Cmp.Gt).take(300).collect()
    used_tensor_path(bctx)
    st = store.last_query_stats()
    ost = ora.MetaStore(v, [price, version], cs)
    fp = ora.FilterPack.from_compiled(expr.compile(store.schema()), store.column_index())
    oi, os_, oq, ostats = ost.query(q, ob.Metric.Cosine, ob.TakeType.Max, 300, (0.0, ob.Cmp.Gt), fp, ora.CANONICAL)
    assert_same_results((np.array(res.indices), np.array(res.scores, np.float32)), (oi, os_), "metastore batch")
    assert (st.total_chunks, st.pruned_chunks, st.evaluated_chunks, st.vectors_compared) == (
        ostats["total_chunks"], ostats["pruned_chunks"], ostats["evaluated_chunks"], ostats["vectors_compared"])


def test_batch_mode_never_matches_tensor_path(ctx):
    v = ora.synth_fill(0, 5000, 128, 101)
    q = ora.synth_fill(0, 16, 128, 102)
    store = make_store(v)
    ctx.set_tuning(batch_mode=2)
    a = run_product(store, q, ob.Metric.DotProduct, [("take", 64)])
    assert ctx.last_work()["batch_used"] == 0
    ctx.set_tuning(batch_mode=0)  # automatic: 16 queries x 5000 rows goes to the tensor cores
    b = run_product(store, q, ob.Metric.DotProduct, [("take", 64)])
    assert ctx.last_work()["batch_used"] == 1
    ctx.set_tuning()
    assert_same_results(a, b, "per-query path vs tensor path")
