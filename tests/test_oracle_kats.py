"""Pins the CPU oracle (and the host-side mirror of expr.rs / col.rs) against the reference's own
known-answer tests, transcribed in tests/golden/reference_kats.json.  No GPU needed."""
import numpy as np
import pytest

from helpers import (DTYPE, build_expr, check_expect, check_meta_expect, load_kats, ob, ora, oracle_meta_kat,
                     oracle_vec_kat)

KATS = load_kats()


@pytest.mark.parametrize("mode", [ora.FAITHFUL, ora.CANONICAL], ids=["faithful", "canonical"])
@pytest.mark.parametrize("kat", KATS["vec"], ids=[k["name"] for k in KATS["vec"]])
def test_oracle_vecstore_kats(kat, mode):
    check_expect(kat, oracle_vec_kat(kat, mode))


@pytest.mark.parametrize("kat", KATS["funcs"], ids=[k["name"] for k in KATS["funcs"]])
def test_oracle_scoring_functions(kat):
    if kat["fn"] == "dot":
        got = ora.dot(kat["a"], kat["b"])
    elif kat["fn"] == "l2":
        got = ora.l2(kat["a"], kat["b"])
    else:
        got = ora.cosine(kat["a"], kat["b"], kat["a_inv"], kat["b_inv"])
    assert abs(got - kat["expect"]) <= kat["tol"]


@pytest.mark.parametrize("mode", [ora.FAITHFUL, ora.CANONICAL], ids=["faithful", "canonical"])
@pytest.mark.parametrize("kat", KATS["meta"], ids=[k["name"] for k in KATS["meta"]])
def test_oracle_metastore_kats(kat, mode):
    result, stats = oracle_meta_kat(kat, mode)
    check_meta_expect(kat, result, stats)


@pytest.mark.parametrize("kat", KATS["expr"], ids=[k["name"] for k in KATS["expr"]])
def test_expr_compile_kats(kat):
    schema = {n: DTYPE[t] for n, t in KATS["expr_schema"].items()}
    expr = build_expr(kat["expr"])
    if "error" in kat:
        with pytest.raises(ob.ExprError) as ei:
            expr.compile(schema)
        assert type(ei.value).__name__ == kat["error"]
        if "error_column" in kat:
            assert ei.value.column == kat["error_column"]
        if "error_got" in kat:
            assert ei.value.got == kat["error_got"]
        return
    cf = expr.compile(schema)
    if "clauses" in kat:
        got = [[[lf.column, lf.cmp.name, lf.kind, lf.rhs] for lf in cl] for cl in cf.clauses]
        assert got == kat["clauses"]
    if "clause_sizes" in kat:
        assert sorted(len(c) for c in cf.clauses) == sorted(kat["clause_sizes"])


@pytest.mark.parametrize("kat", KATS["masks"], ids=[f"{k['ty']}-{k['op']}" for k in KATS["masks"]])
def test_compare_mask_kats(kat):
    """Lane compares of tests/simd_types_tests.rs, replayed through the oracle's row predicate: lane i is a
    one-row column holding a[i], compared against the literal b[i]."""
    dt = ob.DataType.Int64 if kat["ty"] == "i64" else ob.DataType.Float64
    kind = "i64" if kat["ty"] == "i64" else "f64"
    op = {"eq": ob.CmpOp.Eq, "gt": ob.CmpOp.Gt, "gte": ob.CmpOp.Gte, "lt": ob.CmpOp.Lt, "lte": ob.CmpOp.Lte}[kat["op"]]
    bits = 0
    for i, (a, b) in enumerate(zip(kat["a"], kat["b"])):
        col = ob.Column("v", dt).from_values([a])
        st = ora.MetaStore(np.ones((1, 2), np.float32), [col], 8)
        keep = st.row_mask(ora.FilterPack([[(0, int(op), kind, b)]]))
        bits |= int(keep[0]) << i
    assert bits & kat["required"] == kat["required"]
    assert bits & kat["forbidden"] == 0
    want = sum(1 << i for i, (a, b) in enumerate(zip(kat["a"], kat["b"]))
               if {"eq": a == b, "gt": a > b, "gte": a >= b, "lt": a < b, "lte": a <= b}[kat["op"]])
    assert bits == want


def test_add_vector_errors():
    for kat in KATS["add_errors"]:
        store = ob.VecStore(kat["dim"])
        ok = 0
        with pytest.raises(ob.OttersError) as ei:
            for r in kat["rows"]:
                store.add_vector(r)
                ok += 1
        assert kat["error_contains"] in str(ei.value)
        assert ok == kat["ok_rows"] and store.len() == kat["ok_rows"]


def test_meta_build_errors():
    for kat in KATS["meta_build_errors"]:
        cols = [ob.Column(n, DTYPE[t]).from_values(v) for n, t, v in kat["columns"]]
        with pytest.raises(ob.OttersError):
            ob.MetaStore.from_columns(cols).with_vectors(kat["vectors"]).with_chunk_size(kat["chunk_size"]).build()


def test_null_mask_polarity():
    """tests/column_tests.rs:18-37,159-164: mask bit true == NULL; sentinels of src/col.rs:238-326."""
    c = ob.Column("integers", ob.DataType.Int32)
    c.push(42)
    c.push(100)
    c.push(None)
    assert list(c.null_mask()) == [False, False, True]
    assert c.len() == 3 and not c.is_empty()
    assert c.numpy()[2] == -(2**31)
    words = c.null_words()
    assert words is not None and int(words[0]) == 0b100
    f = ob.Column("f", ob.DataType.Float64).from_values([1.5, None])
    assert np.isnan(f.numpy()[1])
    s = ob.Column("s", ob.DataType.String).from_values(["a", None])
    assert s.string_values() == ["a", ""]
    with pytest.raises(ob.OttersError):
        ob.Column("i", ob.DataType.Int32).push("x")


def test_datetime_parsing():
    """src/col.rs:506-529 / tests/column_tests.rs:195-221: RFC3339, date, date-time; everything else fails."""
    p = ob.parse_datetime_millis
    assert p("1970-01-01T00:00:00Z") == 0
    assert p("2023-01-02T03:04:05Z") == 1672628645000
    assert p("2023-01-02T03:04:05+01:00") == 1672628645000 - 3600_000
    assert p("2023-01-02T03:04:05.250Z") == 1672628645250
    assert p("2024-01-01") == 1704067200000
    assert p("2024-12-31 23:59:59") == 1735689599000
    assert p("not a date") is None and p("2024-13-01") is None and p("2024-01-01T00:00:00") is None
    c = ob.Column("ts", ob.DataType.DateTime).from_values(["2024-01-01", None, 5])
    assert list(c.numpy()) == [1704067200000, -(2**63), 5]
    with pytest.raises(ob.OttersError):
        ob.Column("ts", ob.DataType.DateTime).push("garbage")
    d = ob.Column("d", ob.DataType.DateTime).with_datetime_fmt("%d/%m/%Y")
    d.push("02/01/2023")
    assert d.numpy()[0] == 1672617600000
