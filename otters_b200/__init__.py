"""otters_b200 — B200-native (sm_100a) implementation of the otters exact-search hot path.

Public surface mirrors the reference's prelude (src/prelude.rs:7-23):
``VecStore``/``VecQueryPlan``, ``MetaStore``/``MetaStoreBuilder``/``MetaQueryPlan``, ``Column``,
``DataType``, ``Metric``, ``Cmp``, ``TakeType``, ``col``/``lit``.  All scoring, filtering and top-k
selection runs in hand-written CUDA behind the C ABI in ``include/otters_b200.h``
(``otters_b200/libotters_b200.so``); importing this package without that library fails.
"""
from .types import Cmp, CmpOp, DataType, Metric, OttersError, TakeType, VectorFormat, round_to_bf16  # noqa: F401
from .column import Column, ColumnError, parse_datetime_millis  # noqa: F401
from .expr import (  # noqa: F401
    ColumnFilter,
    CompiledFilter,
    Expr,
    ExprError,
    InvalidComparison,
    InvalidExpression,
    TypeMismatch,
    UnknownColumn,
    UnsupportedStringOp,
    col,
    lit,
)
from .context import Context, default_context  # noqa: F401
from .vec import SearchResult, VecQueryPlan, VecStore  # noqa: F401
from .meta import (  # noqa: F401
    MetaBuildStats,
    MetaQueryPlan,
    MetaQueryResults,
    MetaQueryStats,
    MetaStore,
    MetaStoreBuilder,
)

__all__ = [n for n in dir() if not n.startswith("_")]
