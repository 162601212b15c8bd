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


class VectorFormat(enum.IntEnum):
    """How a store keeps its rows in HBM (include/otters_b200.h OTTERS_VECTORS_FMT_*; the reference's roadmap item
    "Quantization for vectors", README.md:208).  Bf16 rounds every element to nearest even when it is added; scores are the
    reference's arithmetic applied to the rounded rows."""

    F32 = 0
    Bf16 = 1


def round_to_bf16(x):
    """f32(bf16_rn(x)) on a numpy array: the values a Bf16 store holds (round to nearest even; NaN stays NaN)."""
    import numpy as np

    a = np.ascontiguousarray(x, dtype=np.float32)
    b = a.view(np.uint32).astype(np.uint64)
    r = ((b + 0x7FFF + ((b >> 16) & 1)) & 0xFFFF0000).astype(np.uint32)
    r = np.where(np.isnan(a), (a.view(np.uint32) & np.uint32(0x80000000)) | np.uint32(0x7FFF0000), r).astype(np.uint32)
    return r.view(np.float32).reshape(a.shape)


def infer_default_take_type(metric: Metric) -> TakeType:
    """src/vec.rs:92-98."""
    return TakeType.Min if metric == Metric.Euclidean else TakeType.Max
