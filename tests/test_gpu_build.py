"""Device-side MetaStore build (otters_b200/csrc/build.cu: zonemap, string-hash, Bloom-insert and dictionary-encode kernels
over the uploaded columns; reference src/meta.rs:151-305, src/meta_compute.rs:32-132).  The tables must be bit-identical to
the oracle's and to the host build path (OTTERS_BUILD_HOST=1), and queries over both stores must return the same bytes."""
import os

import numpy as np
import pytest

from helpers import assert_same_results, ob, ora
from test_gpu_parity import FILTERS, meta_columns

pytestmark = pytest.mark.gpu


def build(cols, vectors, cs, host, bloom=None):
    old = os.environ.get("OTTERS_BUILD_HOST")
    os.environ["OTTERS_BUILD_HOST"] = "1" if host else "0"
    try:
        b = ob.MetaStore.from_columns(cols).with_vectors(vectors).with_chunk_size(cs)
        if bloom:
            b = b.with_bloom_fpr(bloom[1]) if bloom[0] == "fpr" else b.with_bloom_bits(bloom[1])
        return b.build()
    finally:
        if old is None:
            os.environ.pop("OTTERS_BUILD_HOST", None)
        else:
            os.environ["OTTERS_BUILD_HOST"] = old


@pytest.mark.parametrize("cs,bloom", [(1, None), (7, ("bits", 64)), (96, None), (1024, ("fpr", 0.2)), (5000, ("bits", 100000)), (100000, None)])
def test_device_build_matches_host_build_and_oracle(cs, bloom, ctx):
    n, dim = 20011, 16
    vectors = ora.synth_fill(0, n, dim, 111)
    cols = meta_columns(n, cs, 112, null_frac=0.05)
    dev, host = build(cols, vectors, cs, False, bloom), build(cols, vectors, cs, True, bloom)
    ost = ora.MetaStore(vectors, cols, cs, bloom or ("fpr", 0.01))
    assert dev.n_chunks() == host.n_chunks() == ost.n_chunks()
    for i, c in enumerate(cols):
        if c.dtype() == ob.DataType.String:
            continue
        is_f = c.dtype() in (ob.DataType.Float32, ob.DataType.Float64)
        dmn, dmx, dnn = dev.zonemap(c.name())
        hmn, hmx, hnn = host.zonemap(c.name())
        omn, omx, onn = ost.zonemap(i, is_f)
        assert np.array_equal(dnn, hnn) and np.array_equal(dnn, onn), c.name()
        # every entry, the all-NULL chunks' start values included, equals the host path's; live entries equal the oracle's
        assert np.array_equal(dmn.view(np.uint64), hmn.view(np.uint64)) and np.array_equal(dmx.view(np.uint64), hmx.view(np.uint64)), c.name()
        live = onn > 0
        assert np.array_equal(dmn[live].view(np.uint64), omn[live].view(np.uint64)) and np.array_equal(dmx[live].view(np.uint64), omx[live].view(np.uint64))
    q = ora.synth_fill(0, 1, dim, 113)
    for fi in range(len(FILTERS)):
        expr = FILTERS[fi]()
        fp = ora.FilterPack.from_compiled(expr.compile(dev.schema()), dev.column_index())
        want_chunks, want_rows = ost.chunk_mask(fp), ost.row_mask(fp)
        assert np.array_equal(dev.chunk_mask(expr), want_chunks), f"filter {fi}: prune mask (Bloom filters, zonemaps)"
        assert np.array_equal(dev.row_mask(expr), want_rows), f"filter {fi}: row mask (dictionary codes)"
        assert np.array_equal(host.chunk_mask(expr), want_chunks) and np.array_equal(host.row_mask(expr), want_rows)
        rd = dev.query(q[0], ob.Metric.Cosine).meta_filter(expr).take(30).collect()
        rh = host.query(q[0], ob.Metric.Cosine).meta_filter(expr).take(30).collect()
        assert_same_results((rd.indices, rd.scores), (rh.indices, rh.scores), f"filter {fi}")
        # result strings come back through the device-built dictionary
        for name in ("item",):
            assert [rd.data[name].get(j) for j in range(len(rd.indices))] == [cols[-1].get(i) for i in rd.indices]


def test_device_build_many_distinct_and_empty_strings(ctx):
    n, cs = 30000, 512
    rng = np.random.default_rng(5)
    vals = [f"s{int(x)}" if x % 7 else "" for x in rng.integers(0, 25000, n)]  # ~17k distinct values, empty strings too
    nulls = rng.random(n) < 0.1
    col = ob.Column.from_numpy("name", ob.DataType.String, ["" if nl else s for s, nl in zip(vals, nulls)], nulls)
    vectors = ora.synth_fill(0, n, 8, 114)
    dev, host = build([col], vectors, cs, False), build([col], vectors, cs, True)
    ost = ora.MetaStore(vectors, [col], cs)
    for lit in ("s1", "s24999", "", "absent", vals[123]):
        for e in (ob.col("name").eq(lit), ob.col("name").neq(lit)):
            fp = ora.FilterPack.from_compiled(e.compile(dev.schema()), dev.column_index())
            assert np.array_equal(dev.chunk_mask(e), ost.chunk_mask(fp)) and np.array_equal(dev.row_mask(e), ost.row_mask(fp)), lit
            assert np.array_equal(host.row_mask(e), ost.row_mask(fp))
    got = dev.gather("name", list(range(0, n, 97)))
    assert [got.get(j) for j in range(len(got))] == [col.get(i) for i in range(0, n, 97)]
