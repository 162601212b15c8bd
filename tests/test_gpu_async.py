"""The one-launch query path and the non-blocking API.

* K3 (selection) normally runs INSIDE the scan kernel (in its last CTA) and K0 (chunk pruning) can (`lazy_prune`: per work
  unit); `separate_select` / the default run them as their own kernels.  Every combination must return the same bytes and
  the same statistics as the oracle.
* otters_query_submit / otters_query_wait keep two queries in flight per context (two lanes); results must equal the
  blocking calls', in any interleaving, and a lane's stale ticket must be refused.
Everything goes through the C ABI (ctypes)."""
import ctypes as C

import numpy as np
import pytest

from helpers import assert_same_results, ob, ora
from otters_b200 import _ffi
from test_gpu_parity import FILTERS, meta_columns

pytestmark = pytest.mark.gpu

STAT_KEYS = ("total_chunks", "pruned_chunks", "evaluated_chunks", "vectors_compared")


@pytest.mark.parametrize("cs", [1, 5, 16, 96, 1024, 4096])
def test_fused_prune_and_select_match_standalone_kernels(cs, ctx):
    n, dim = 13000, 48
    vectors = ora.synth_fill(0, n, dim, 81)
    cols = meta_columns(n, cs, 82)
    store = ob.MetaStore.from_columns(cols).with_vectors(vectors).with_chunk_size(cs).build()
    ost = ora.MetaStore(vectors, cols, cs)
    q = ora.synth_fill(0, 1, dim, 83)
    try:
        for fi in (None, 0, 2, 4, 5, 8, 10, 11):
            expr = FILTERS[fi]() if fi is not None else None
            fp = ora.FilterPack.from_compiled(expr.compile(store.schema()), store.column_index()) if expr is not None else None
            for k in (1, 100, 1024):
                oi, os_, _, ostats = ost.query(q, ob.Metric.Cosine, ob.TakeType.Max, k, None, fp, ora.CANONICAL)
                for sel, lazy, mode, fused in ((0, 0, 0, 0), (1, 0, 0, 0), (0, 1, 0, 0), (1, 1, 0, 0), (0, 1, 2, 0), (1, 0, 2, 0), (0, 0, 1, 2), (0, 0, 2, 2)):
                    ctx.set_tuning(separate_select=sel, lazy_prune=lazy, scan_mode=mode, disable_fused_predicate=fused)
                    plan = store.query(q[0], ob.Metric.Cosine)
                    if expr is not None:
                        plan = plan.meta_filter(expr)
                    res = plan.take(k).collect()
                    what = f"cs={cs} filter={fi} k={k} separate_select={sel} lazy_prune={lazy} scan_mode={mode} fused={fused}"
                    assert_same_results((res.indices, res.scores), (oi, os_), what)
                    st = store.last_query_stats()
                    for key in STAT_KEYS:
                        assert getattr(st, key) == ostats[key], f"{what}: stats.{key} {getattr(st, key)} != {ostats[key]}"
    finally:
        ctx.set_tuning()


def test_one_launch_per_query(ctx):
    """Single queries with k <= 1024: prune kernel + row-mask kernel + ONE scan kernel (scan + select); the row predicate
    (disable_fused_predicate=2) and the chunk pruning (lazy_prune) can be folded into the scan kernel, down to one launch."""
    n, dim, cs = 20000, 64, 256
    vectors = ora.synth_fill(0, n, dim, 84)
    cols = meta_columns(n, cs, 85)
    store = ob.MetaStore.from_columns(cols).with_vectors(vectors).with_chunk_size(cs).build()
    q = ora.synth_fill(0, 1, dim, 86)[0]
    expr = FILTERS[0]()
    try:
        for sel, lazy, fused, want in ((0, 1, 0, 1), (1, 1, 0, 2), (0, 0, 2, 2), (1, 0, 2, 3), (0, 0, 0, 3), (1, 0, 0, 4)):
            ctx.set_tuning(separate_select=sel, lazy_prune=lazy, disable_fused_predicate=fused)
            store.query(q, ob.Metric.Cosine).meta_filter(expr).take(10).collect()
            assert ctx.last_work()["kernel_launches"] == want
        ctx.set_tuning()
        store.query(q, ob.Metric.Cosine).take(10).collect()  # no meta_filter: statistics travel in the input image
        assert ctx.last_work()["kernel_launches"] == 1
        vs = ob.VecStore(dim)
        vs.add_vectors(vectors)
        vs.query(q, ob.Metric.DotProduct).take(7).collect()
        assert ctx.last_work()["kernel_launches"] == 1
    finally:
        ctx.set_tuning()


def test_submit_wait_matches_blocking_calls(ctx):
    n, dim, cs = 30000, 96, 128
    vectors = ora.synth_fill(0, n, dim, 87)
    cols = meta_columns(n, cs, 88)
    store = ob.MetaStore.from_columns(cols).with_vectors(vectors).with_chunk_size(cs).build()
    ost = ora.MetaStore(vectors, cols, cs)
    qs = ora.synth_fill(0, 12, dim, 89)
    plans = []
    for i in range(12):
        fi = [None, 0, 1, 2, 4, 5][i % 6]
        expr = FILTERS[fi]() if fi is not None else None
        metric = [ob.Metric.Cosine, ob.Metric.Euclidean, ob.Metric.DotProduct][i % 3]
        k = [10, 100, 1000, 3000][i % 4]  # 3000 > 1024: the sort path, served through the same ticket interface
        plans.append((qs[i], metric, expr, k))

    def plan_of(i):
        q, metric, expr, k = plans[i]
        p = store.query(q, metric)
        if expr is not None:
            p = p.meta_filter(expr)
        return p.take(k)

    def check(i, res, st):
        q, metric, expr, k = plans[i]
        fp = ora.FilterPack.from_compiled(expr.compile(store.schema()), store.column_index()) if expr is not None else None
        tt = ob.TakeType.Min if metric == ob.Metric.Euclidean else ob.TakeType.Max
        oi, os_, _, ostats = ost.query(q[None, :], metric, tt, k, None, fp, ora.CANONICAL)
        assert_same_results((res.indices, res.scores), (oi, os_), f"query {i}")
        for key in STAT_KEYS:
            assert getattr(st, key) == ostats[key], f"query {i}: stats.{key}"

    # two in flight: submit i+1 before waiting for i
    pend = plan_of(0).submit()
    for i in range(12):
        nxt = plan_of(i + 1).submit() if i + 1 < 12 else None
        res = pend.wait()
        check(i, res, store.last_query_stats())
        pend = nxt
    # blocking calls still work in between and give the same answers
    for i in (3, 7):
        check(i, plan_of(i).collect(), store.last_query_stats())
    # waiting in the opposite order
    a, b = plan_of(1).submit(), plan_of(2).submit()
    rb = b.wait()
    check(2, rb, store.last_query_stats())
    ra = a.wait()
    check(1, ra, store.last_query_stats())
    # a lane holds one query: after two more submissions the first ticket is gone
    t0 = plan_of(4).submit()
    t1, t2 = plan_of(5).submit(), plan_of(6).submit()
    with pytest.raises(ob.OttersError) as ei:
        t0.wait()
    assert "ticket" in str(ei.value)
    check(5, t1.wait(), store.last_query_stats())
    check(6, t2.wait(), store.last_query_stats())
    ctx.synchronize()


def test_submit_wait_vecstore(ctx):
    n, dim = 50000, 40
    v = ora.synth_fill(0, n, dim, 90)
    qs = ora.synth_fill(0, 6, dim, 91)
    store = ob.VecStore(dim)
    store.add_vectors(v)
    rng = np.random.default_rng(3)
    mask = rng.random(n) < 0.4
    pend = []
    for i in range(6):
        plan = store.query(qs[i], ob.Metric.Cosine)
        if i % 2:
            plan = plan.with_row_mask(mask)
        if i % 3 == 0:
            plan = plan.filter(0.05, ob.Cmp.Gt)
        pend.append((i, plan.take(25 + i).submit()))
        if len(pend) == 2:
            j, t = pend.pop(0)
            got = t.wait()
            want = ora.vecstore_query(v, qs[j:j + 1], ob.Metric.Cosine, ob.TakeType.Max, 25 + j, (0.05, ob.Cmp.Gt) if j % 3 == 0 else None,
                                      mask if j % 2 else None, ora.CANONICAL)
            assert_same_results(got, want, f"query {j}")
    j, t = pend.pop(0)
    got = t.wait()
    want = ora.vecstore_query(v, qs[j:j + 1], ob.Metric.Cosine, ob.TakeType.Max, 25 + j, None, mask, ora.CANONICAL)
    assert_same_results(got, want, f"query {j}")
    # take(0) and a batch go through the same interface
    assert len(store.query(qs[0], ob.Metric.Cosine).take(0).submit().wait()[0]) == 0
    got = store.query(qs[:3], ob.Metric.DotProduct).take(40).submit().wait()
    assert_same_results(got, ora.vecstore_query(v, qs[:3], ob.Metric.DotProduct, ob.TakeType.Max, 40, None, None, ora.CANONICAL), "batch")


def test_wait_rejects_garbage_ticket(ctx):
    out_len = C.c_uint64(0)
    rc = _ffi.otters_query_wait(ctx.handle, 0xDEAD00, None, None, None, 0, C.byref(out_len), None)
    assert rc != 0 and "ticket" in _ffi.last_error()
