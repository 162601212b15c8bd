"""Self-consistency of the oracle on seeded random inputs (CPU only): the literal TopKCollector restatement
and the canonical selection agree whenever no score tie exists; both reduce orders agree to a few ulp; the
scores agree with float64 NumPy to 1e-5 relative; Bloom filters have no false negatives."""
import numpy as np
import pytest

from helpers import assert_same_results, ob, ora


def rand_store(n, dim, seed):
    return ora.synth_fill(0, n, dim, seed)


@pytest.mark.parametrize("n,dim,nq,k", [(1000, 128, 1, 10), (777, 100, 3, 25), (64, 3, 2, 200), (5000, 33, 1, 1500), (9, 8, 1, 9)])
@pytest.mark.parametrize("metric", [ob.Metric.Cosine, ob.Metric.Euclidean, ob.Metric.DotProduct])
def test_faithful_equals_canonical_without_ties(n, dim, nq, k, metric):
    v = rand_store(n, dim, 1234 + n)
    q = ora.synth_fill(0, nq, dim, 0xBEEF)
    for tt in (ob.TakeType.Max, ob.TakeType.Min):
        for flt in (None, (0.0, ob.Cmp.Gt), (0.05, ob.Cmp.Lte)):
            a = ora.vecstore_query(v, q, metric, tt, k, flt, None, ora.FAITHFUL)
            b = ora.vecstore_query(v, q, metric, tt, k, flt, None, ora.CANONICAL)
            assert_same_results(a, b, f"n={n} dim={dim} {metric.name} {tt.name} {flt}")


def test_row_mask_semantics():
    """bits past the mask length keep the row (src/vec.rs:234,297); masked rows never appear."""
    v = rand_store(100, 16, 7)
    q = ora.synth_fill(0, 1, 16, 9)
    mask = np.zeros(40, bool)
    mask[::3] = True
    for mode in (ora.FAITHFUL, ora.CANONICAL):
        idx, _, _ = ora.vecstore_query(v, q, ob.Metric.DotProduct, ob.TakeType.Max, 100, None, mask, mode)
        want = set(np.nonzero(mask)[0]) | set(range(40, 100))
        assert set(int(i) for i in idx) == want


def test_scores_match_float64_reference():
    v = rand_store(500, 768, 3)
    q = ora.synth_fill(0, 1, 768, 4)[0]
    d64 = v.astype(np.float64) @ q.astype(np.float64)
    for i in range(0, 500, 37):
        assert abs(ora.dot(q, v[i]) - d64[i]) <= 1e-5 * max(abs(d64[i]), 1e-3)
        l64 = float(((q.astype(np.float64) - v[i]) ** 2).sum())
        assert abs(ora.l2(q, v[i]) - l64) <= 1e-5 * l64
    inv = ora.inv_norms(v)
    assert np.allclose(inv, 1.0 / np.linalg.norm(v.astype(np.float64), axis=1), rtol=1e-6)


def test_reduce_orders_agree_to_ulps():
    v = rand_store(200, 768, 5)
    q = ora.synth_fill(0, 1, 768, 6)[0]
    a = np.array([ora.dot(q, r) for r in v], np.float32)
    ora.set_reduce_order(1)
    try:
        b = np.array([ora.dot(q, r) for r in v], np.float32)
    finally:
        ora.set_reduce_order(0)
    assert np.all(np.abs(a - b) <= 8 * np.spacing(np.maximum(np.abs(a), np.abs(b)).astype(np.float32)) + 1e-6)


def test_tail_and_small_dims():
    for dim in (1, 2, 3, 7, 8, 9, 15, 17):
        a = np.arange(1, dim + 1, dtype=np.float32)
        b = np.arange(2, dim + 2, dtype=np.float32)
        assert ora.dot(a, b) == float((a.astype(np.float64) * b).sum())
        assert ora.l2(a, b) == float(dim)
    assert ora.inv_norm(np.zeros(5, np.float32)) == 0.0


def test_nan_scores_are_dropped():
    v = np.array([[1.0, 0.0], [np.nan, 0.0], [0.5, 0.5]], np.float32)
    for mode in (ora.FAITHFUL, ora.CANONICAL):
        idx, score, _ = ora.vecstore_query(v, np.array([[1.0, 0.0]], np.float32), ob.Metric.DotProduct, ob.TakeType.Max, 3, None, None, mode)
        assert list(idx) == [0, 2] and not np.isnan(score).any()


def test_canonical_tie_order():
    v = np.array([[1.0, 0.0]] * 6, np.float32)
    idx, _, qid = ora.vecstore_query(v, np.array([[1.0, 0.0], [1.0, 0.0]], np.float32), ob.Metric.DotProduct, ob.TakeType.Max, 5, None, None, ora.CANONICAL)
    assert list(idx) == [0, 0, 1, 1, 2] and list(qid) == [0, 1, 0, 1, 0]


def test_generator_is_counter_based():
    a = ora.synth_fill(0, 64, 24, 0x7735)
    b = ora.synth_fill(32, 32, 24, 0x7735)
    assert np.array_equal(a[32:], b)
    assert a.min() >= -1.0 and a.max() < 1.0
    assert len(np.unique(a)) > 1500


def test_bloom_no_false_negatives_and_all_null_chunks():
    n, cs = 4000, 256
    rng = np.random.default_rng(0)
    names = [f"item_{(i // cs) % 7}_{rng.integers(0, 5)}" for i in range(n)]
    nulls = rng.random(n) < 0.05
    nulls[cs * 3: cs * 4] = True  # one all-null chunk
    col = ob.Column.from_numpy("item", ob.DataType.String, names, nulls)
    col._vals = ["" if nl else s for s, nl in zip(names, nulls)]
    st = ora.MetaStore(np.ones((n, 2), np.float32), [col], cs)
    for probe in ("item_0_0", "item_3_4", "absent"):
        keep = st.chunk_mask(ora.FilterPack([[(0, int(ob.CmpOp.Eq), "str", probe)]]))
        rows = st.row_mask(ora.FilterPack([[(0, int(ob.CmpOp.Eq), "str", probe)]]))
        truth = np.array([(not nl) and s == probe for s, nl in zip(names, nulls)])
        assert np.array_equal(rows.astype(bool), truth)
        for ch in range(st.n_chunks()):
            if truth[ch * cs:(ch + 1) * cs].any():
                assert keep[ch] == 1, "Bloom false negative"
        assert keep[3] == 0, "all-null chunk must be pruned"
    assert st.chunk_mask(ora.FilterPack([[(0, int(ob.CmpOp.Neq), "str", "x")]]))[3] == 0
    assert st.chunk_mask(ora.FilterPack([[(0, int(ob.CmpOp.Eq), "str", "absent")]])).sum() <= 2  # fpr 1%
