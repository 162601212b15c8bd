"""Enums of the otters public API (codes follow the reference's declaration order).

Reference: src/vec.rs:11-31 (Metric, TakeType, Cmp), src/expr.rs:83-91 (CmpOp),
src/type_utils.rs:11-19 (DataType).
"""
from __future__ import annotations

import enum


class OttersError(Exception):
    """The reference returns ``Err(String)`` from ``collect()``/``build()``; the message is ``str(exc)``."""


class Metric(enum.IntEnum):
    Cosine = 0
    Euclidean = 1  # SQUARED euclidean distance, no sqrt (src/vec_compute.rs:35-54)
    DotProduct = 2


class TakeType(enum.IntEnum):
    Min = 0
    Max = 1


class Cmp(enum.IntEnum):
    Lt = 0
    Gt = 1
    Lte = 2
    Gte = 3
    Eq = 4


class CmpOp(enum.IntEnum):
    Eq = 0
    Neq = 1
    Lt = 2
    Lte = 3
    Gt = 4
    Gte = 5


class DataType(enum.IntEnum):
    Int32 = 0
    Int64 = 1
    Float32 = 2
    Float64 = 3
    String = 4
    DateTime = 5


def infer_default_take_type(metric: Metric) -> TakeType:
    """src/vec.rs:92-98."""
    return TakeType.Min if metric == Metric.Euclidean else TakeType.Max
