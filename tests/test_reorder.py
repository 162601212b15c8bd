"""Row ordering for better pruning (otters_b200/reorder.py; the reference's roadmap, README.md:154,212).  Host logic only:
the permutation is checked for its defining properties, and its effect is shown on the ORACLE — the reference's own zonemap
rules (src/meta.rs:407-544) prune more chunks of a clustered store while the result set stays the same."""
import numpy as np
import pytest

from helpers import ob, ora
from otters_b200 import reorder


def columns(n, seed=0):
    rng = np.random.default_rng(seed)
    price = rng.uniform(0, 100, n)
    qty = rng.integers(0, 1000, n).astype(np.int32)
    item = rng.integers(0, 50, n)
    nulls = rng.random(n) < 0.02
    return {
        "price": ob.Column.from_numpy("price", ob.DataType.Float64, price, nulls),
        "qty": ob.Column.from_numpy("qty", ob.DataType.Int32, qty),
        "item": ob.Column.from_categories("item", [f"item{i:02d}" for i in range(50)], item),
    }


@pytest.mark.parametrize("method", reorder.METHODS)
def test_row_order_is_a_stable_permutation(method):
    cols = columns(5000)
    perm = reorder.compute_row_order(cols, ["price", "qty"], method)
    assert perm.dtype == np.uint64 and sorted(perm.tolist()) == list(range(5000))
    if method == "sort":
        p = cols["price"].numpy()[perm.astype(np.int64)]
        nulls = cols["price"].null_mask()[perm.astype(np.int64)]
        live = p[~nulls]
        assert (np.diff(live) >= 0).all() and not nulls[: len(live)].any()  # ascending, NULLs last
    # stable: rows with equal keys keep their input order
    same = {"k": ob.Column.from_numpy("k", ob.DataType.Int32, np.repeat(np.arange(10, dtype=np.int32)[::-1], 7))}
    perm = reorder.compute_row_order(same, ["k"], method).astype(np.int64)
    k = same["k"].numpy()[perm]
    assert (np.diff(k) >= 0).all()
    for v in range(10):
        assert (np.diff(perm[k == v]) > 0).all()


def test_string_columns_sort_lexicographically():
    c = {"s": ob.Column("s", ob.DataType.String).from_values(["pear", "apple", None, "fig", "apple"])}
    assert reorder.compute_row_order(c, ["s"]).tolist() == [1, 4, 3, 0, 2]


def test_unknown_column_and_method_are_rejected():
    cols = columns(10)
    with pytest.raises(ob.OttersError):
        reorder.compute_row_order(cols, ["nope"])
    with pytest.raises(ob.OttersError):
        reorder.compute_row_order(cols, ["price"], "hilbert")
    with pytest.raises(ob.OttersError):
        reorder.compute_row_order(cols, [])


def test_morton_code_interleaves_most_significant_bits_first():
    a, b = np.array([0xFFFF, 0, 0x8000], np.uint64), np.array([0, 0xFFFF, 0x8000], np.uint64)
    code = reorder.morton_codes([a, b])
    assert code[0] == 0xAAAAAAAA and code[1] == 0x55555555 and code[2] == 0xC0000000


def test_zorder_narrows_every_column_and_sort_only_the_first():
    n, cs = 40000, 500
    cols = columns(n, 3)
    price, qty = cols["price"].numpy(), cols["qty"].numpy().astype(np.float64)
    base = (reorder.chunk_ranges_overlapping(price, cs, 20, 30), reorder.chunk_ranges_overlapping(qty, cs, 100, 200))
    assert base == (n // cs, n // cs)  # unordered: every chunk spans the whole range, nothing can be pruned
    out = {}
    for method in reorder.METHODS:
        perm = reorder.compute_row_order(cols, ["price", "qty"], method).astype(np.int64)
        out[method] = (reorder.chunk_ranges_overlapping(price[perm], cs, 20, 30), reorder.chunk_ranges_overlapping(qty[perm], cs, 100, 200))
    assert out["sort"][0] <= n // cs // 10 + 2 and out["sort"][1] >= n // cs - 2  # perfect on price, (almost) nothing on qty
    assert out["zorder"][0] <= n // cs // 2 and out["zorder"][1] <= n // cs // 2  # both columns prune


@pytest.mark.parametrize("method", reorder.METHODS)
def test_reference_rules_prune_more_chunks_and_return_the_same_rows(method):
    """The oracle (the reference's CPU path restated) on the clustered store: more chunks pruned, identical result set once
    the store positions are mapped back through the permutation (scores are tie-free here)."""
    n, dim, cs = 6000, 24, 128
    v = ora.synth_fill(0, n, dim, 99)
    cols = columns(n, 5)
    order = ["price", "qty", "item"]
    q = ora.synth_fill(0, 1, dim, 100)
    expr = ob.col("price").lt(15.0) & ob.col("qty").gte(700)
    schema = {name: cols[name].dtype() for name in order}
    fp = ora.FilterPack.from_compiled(expr.compile(schema), {name: i for i, name in enumerate(order)})
    plain = ora.MetaStore(v, [cols[name] for name in order], cs)
    i0, s0, _, st0 = plain.query(q, ob.Metric.Cosine, ob.TakeType.Max, 40, None, fp)
    perm = reorder.compute_row_order(cols, ["price", "qty"], method)
    pi = perm.astype(np.int64)
    clustered = ora.MetaStore(v[pi], [cols[name].gather(pi) for name in order], cs)
    i1, s1, _, st1 = clustered.query(q, ob.Metric.Cosine, ob.TakeType.Max, 40, None, fp)
    assert len(i0) == 40 and np.array_equal(perm[i1.astype(np.int64)], i0.astype(np.uint64))
    assert np.array_equal(s1.view(np.uint32), s0.view(np.uint32))
    assert st0["pruned_chunks"] == 0 and st1["pruned_chunks"] >= st1["total_chunks"] // 2
    assert st1["vectors_compared"] < st0["vectors_compared"] // 2


def test_builder_validates_row_order_without_a_device():
    b = ob.MetaStore.from_columns(list(columns(10).values()))
    with pytest.raises(ob.OttersError):
        b.with_row_order("price", "hilbert")
    assert b.with_row_order("price")._row_order == (["price"], "sort")
    assert b.with_row_order(["price", "qty"], "zorder")._row_order == (["price", "qty"], "zorder")
    assert b.with_vector_format(ob.VectorFormat.Bf16)._vector_format == ob.VectorFormat.Bf16
