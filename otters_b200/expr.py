"""Expression DSL for metadata filters — host-side mirror of the reference's ``expr`` module.

``col("price").lt(50.0) & col("version").gte(2)`` builds an :class:`Expr`; ``Expr.compile(schema)``
type-checks it and lowers it to conjunctive normal form (src/expr.rs:285-511).  The compiled filter
(AND over clauses of OR over typed leaves, src/expr.rs:192-226) is the predicate format the CUDA
kernels consume through ``otters_filter`` (include/otters_b200.h).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Dict, List, Union

import numpy as np

from .column import parse_datetime_millis
from .types import CmpOp, DataType, OttersError


class ExprError(OttersError):
    pass


class UnknownColumn(ExprError):
    def __init__(self, col):
        self.column = col
        super().__init__(f"Unknown column '{col}'")


class TypeMismatch(ExprError):
    def __init__(self, col, dtype, got):
        self.column, self.dtype, self.got = col, dtype, got
        super().__init__(f"Type mismatch for column '{col}': expected {dtype.name}, got literal {got}")


class UnsupportedStringOp(ExprError):
    def __init__(self, col):
        self.column = col
        super().__init__(f"Unsupported comparator for string column '{col}'")


class InvalidComparison(ExprError):
    def __init__(self):
        super().__init__("Invalid expression shape for comparison (expect column vs literal)")


class InvalidExpression(ExprError):
    def __init__(self):
        super().__init__("Invalid expression (unexpected literal or column without comparator)")


@dataclass(frozen=True)
class Literal:
    """src/expr.rs:44-80: I64 | F64 | Str."""

    kind: str  # "i64" | "f64" | "str"
    value: Union[int, float, str]

    @staticmethod
    def of(v) -> "Literal":
        if isinstance(v, Literal):
            return v
        if isinstance(v, bool):
            return Literal("i64", int(v))
        if isinstance(v, (int, np.integer)):
            return Literal("i64", int(v))
        if isinstance(v, (float, np.floating)):
            return Literal("f64", float(v))
        if isinstance(v, str):
            return Literal("str", v)
        raise TypeError(f"unsupported literal {v!r}")


class Expr:
    """src/expr.rs:93-106."""

    __slots__ = ("kind", "a", "b", "op")

    def __init__(self, kind, a=None, b=None, op=None):
        self.kind, self.a, self.b, self.op = kind, a, b, op

    # comparison builders (src/expr.rs:118-166)
    def _cmp(self, v, op):
        return Expr("cmp", self, v if isinstance(v, Expr) else lit(v), op)

    def eq(self, v):
        return self._cmp(v, CmpOp.Eq)

    def neq(self, v):
        return self._cmp(v, CmpOp.Neq)

    def lt(self, v):
        return self._cmp(v, CmpOp.Lt)

    def lte(self, v):
        return self._cmp(v, CmpOp.Lte)

    def gt(self, v):
        return self._cmp(v, CmpOp.Gt)

    def gte(self, v):
        return self._cmp(v, CmpOp.Gte)

    def and_(self, other: "Expr") -> "Expr":
        return Expr("and", self, other)

    def or_(self, other: "Expr") -> "Expr":
        return Expr("or", self, other)

    __and__ = and_
    __or__ = or_

    def compile(self, schema: Dict[str, DataType]) -> "CompiledFilter":
        """src/expr.rs:285-297."""
        return CompiledFilter(_normalize(_lower(self, schema)))

    def __repr__(self):
        if self.kind == "col":
            return f"col({self.a!r})"
        if self.kind == "lit":
            return f"lit({self.a.value!r})"
        if self.kind == "cmp":
            return f"({self.a!r} {self.op.name} {self.b!r})"
        return f"({self.a!r} {self.kind.upper()} {self.b!r})"


def col(name: str) -> Expr:
    return Expr("col", name)


def lit(v) -> Expr:
    return Expr("lit", Literal.of(v))


@dataclass(frozen=True)
class ColumnFilter:
    """One typed leaf (src/expr.rs:192-210).  ``kind`` is "i64" | "f64" (Numeric) or "str" (String)."""

    column: str
    cmp: CmpOp
    kind: str
    rhs: Union[int, float, str]

    @staticmethod
    def numeric_i64(column, cmp, v):
        return ColumnFilter(column, CmpOp(cmp), "i64", int(v))

    @staticmethod
    def numeric_f64(column, cmp, v):
        return ColumnFilter(column, CmpOp(cmp), "f64", float(v))

    @staticmethod
    def string(column, cmp, v):
        return ColumnFilter(column, CmpOp(cmp), "str", str(v))


@dataclass
class CompiledFilter:
    """AND over ``clauses`` of OR over leaves (src/expr.rs:212-226)."""

    clauses: List[List[ColumnFilter]]


def _compile_leaf(left: Expr, right: Expr, op: CmpOp, schema) -> ColumnFilter:
    """src/expr.rs:385-466."""
    if left.kind != "col" or right.kind != "lit":
        raise InvalidComparison()
    name, l = left.a, right.a
    if name not in schema:
        raise UnknownColumn(name)
    dtype = DataType(schema[name])
    if dtype == DataType.String:
        if op not in (CmpOp.Eq, CmpOp.Neq):
            raise UnsupportedStringOp(name)
        if l.kind != "str":
            raise TypeMismatch(name, dtype, "string")
        return ColumnFilter.string(name, op, l.value)
    if dtype in (DataType.Int32, DataType.Int64):
        if l.kind == "f64":
            raise TypeMismatch(name, dtype, "float")
        if l.kind == "str":
            raise TypeMismatch(name, dtype, "string")
        return ColumnFilter.numeric_i64(name, op, l.value)
    if dtype == DataType.DateTime:
        if l.kind != "str":
            raise TypeMismatch(name, dtype, "datetime string")
        ms = parse_datetime_millis(l.value)
        if ms is None:
            raise TypeMismatch(name, dtype, "datetime string")
        return ColumnFilter.numeric_i64(name, op, ms)
    # Float32 / Float64: ints are widened
    if l.kind == "str":
        raise TypeMismatch(name, dtype, "string")
    return ColumnFilter.numeric_f64(name, op, float(l.value))


def _lower(e: Expr, schema) -> List[List[ColumnFilter]]:
    """src/expr.rs:355-372, :474-511."""
    if e.kind == "and":
        a, b = _lower(e.a, schema), _lower(e.b, schema)
        if not a:
            return b
        if not b:
            return a
        return a + b
    if e.kind == "or":
        a, b = _lower(e.a, schema), _lower(e.b, schema)
        if not a:
            return b
        if not b:
            return a
        return [ca + cb for ca in a for cb in b]
    if e.kind == "cmp":
        return [[_compile_leaf(e.a, e.b, e.op, schema)]]
    raise InvalidExpression()


def _normalize(plan):
    """Drop tautology clauses ``(c == v) OR (c != v)`` (src/expr.rs:302-343)."""
    out = []
    for clause in plan:
        taut = False
        for lf in clause:
            if lf.cmp == CmpOp.Eq and any(
                x.cmp == CmpOp.Neq and x.column == lf.column and x.kind == lf.kind and _same(x.rhs, lf.rhs) for x in clause
            ):
                taut = True
                break
        if not taut:
            out.append(clause)
    return out


def _same(a, b):
    # NumericLiteral PartialEq: F64(NaN) != F64(NaN)
    return a == b
