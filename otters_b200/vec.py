"""VecStore / VecQueryPlan — host-side mirror of the reference's ``vec`` module (src/vec.rs).

The plan builder keeps the reference's semantics (deferred errors, take-type inference, "last take*
wins"); ``collect()`` is one call into the CUDA library (``otters_vecstore_query``).
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass
from typing import List, Optional, Sequence

import numpy as np

from . import _ffi
from .context import Context, check, default_context
from .types import Cmp, Metric, OttersError, TakeType, VectorFormat, infer_default_take_type


@dataclass(frozen=True)
class SearchResult:
    """src/vec.rs:33-53."""

    index: int
    score: float

    def __str__(self):
        return f"#{self.index} score={self.score:.6f}"


def _as_query_batch(queries):
    """QueryBatch: From<Vec<f32>> + From<Vec<Vec<f32>>> (src/vec.rs:320-336).  Returns a list of 1-d arrays."""
    if isinstance(queries, np.ndarray):
        if queries.ndim == 1:
            return [np.ascontiguousarray(queries, dtype=np.float32)]
        return [np.ascontiguousarray(r, dtype=np.float32) for r in queries]
    queries = list(queries)
    if len(queries) == 0:
        return []
    if isinstance(queries[0], (list, tuple, np.ndarray)):
        return [np.asarray(r, dtype=np.float32) for r in queries]
    return [np.asarray(queries, dtype=np.float32)]


def pack_mask_words(mask) -> np.ndarray:
    """bool sequence -> Lsb0 u64 words (bitvec's default layout)."""
    m = np.asarray(mask, dtype=bool)
    n = len(m)
    padded = np.zeros((n + 63) // 64 * 64, dtype=np.uint8)
    padded[:n] = m
    words = np.packbits(padded, bitorder="little").view(np.uint64).copy()
    return words if len(words) else np.zeros(1, dtype=np.uint64)


class VecStore:
    """Flat row-major f32 vectors resident in HBM (src/vec.rs:338-411).  `vector_format=VectorFormat.Bf16` keeps the rows
    as bf16 (half the bytes per scan; results are the reference's on the rounded rows, see include/otters_b200.h)."""

    def __init__(self, dim: int, ctx: Optional[Context] = None, vector_format: VectorFormat = VectorFormat.F32):
        self.dim = int(dim)
        self.vector_format = VectorFormat(vector_format)
        self._ctx = ctx
        self._h = None
        self._pending: List[np.ndarray] = []
        self._n = 0

    # -- device handle is created lazily so that plans over an empty store need no GPU work
    def _handle(self):
        if self._h is None:
            if self._ctx is None:
                self._ctx = default_context()
            h = C.c_void_p()
            check(_ffi.otters_vecstore_create_fmt(self._ctx.handle, self.dim, int(self.vector_format), C.byref(h)))
            self._h = h
        return self._h

    @property
    def ctx(self) -> Context:
        self._handle()
        return self._ctx

    def add_vector(self, vector: Sequence[float]) -> None:
        """src/vec.rs:357-371 (raises instead of returning Err)."""
        v = np.asarray(vector, dtype=np.float32)
        if v.ndim != 1 or v.shape[0] != self.dim:
            raise OttersError(f"Input vector length {v.size if v.ndim == 1 else len(vector)} does not match expected dimension {self.dim}")
        self._pending.append(v.reshape(1, -1))
        self._n += 1

    def add_vectors(self, vectors) -> None:
        """src/vec.rs:374-376: rows before a bad row are kept, as with try_for_each."""
        if isinstance(vectors, np.ndarray) and vectors.ndim == 2:
            if vectors.shape[1] != self.dim:
                raise OttersError(f"Input vector length {vectors.shape[1]} does not match expected dimension {self.dim}")
            self._pending.append(np.ascontiguousarray(vectors, dtype=np.float32))
            self._n += vectors.shape[0]
            return
        for v in vectors:
            self.add_vector(v)

    def add_synthetic(self, first_row: int, n: int, seed: int) -> None:
        """Appends device-generated synthetic rows (bench/test utility)."""
        self._flush()
        check(_ffi.otters_vecstore_add_synthetic(self._handle(), first_row, n, seed))
        self._n += n

    def add_synthetic_sharded(self, world: int, rank: int, block_rows: int, n_local: int, seed: int, row_base: int = 0) -> None:
        """Appends the rows a rank holds under block-cyclic sharding (global ids as in otters_shard_map)."""
        self._flush()
        m = _ffi.ShardMap(row_base, world, rank, block_rows)
        check(_ffi.otters_vecstore_add_synthetic_sharded(self._handle(), C.byref(m), n_local, seed))
        self._n += n_local

    def reserve(self, n: int) -> None:
        check(_ffi.otters_vecstore_reserve(self._handle(), n))

    def _flush(self) -> None:
        if not self._pending:
            return
        rows = np.ascontiguousarray(np.concatenate(self._pending, axis=0), dtype=np.float32)
        self._pending = []
        check(_ffi.otters_vecstore_add(self._handle(), rows.ctypes.data_as(_ffi.c_f32p), rows.shape[0]))

    def len(self) -> int:
        return self._n

    __len__ = len

    def is_empty(self) -> bool:
        return self._n == 0

    def set_rows(self, rows, data) -> None:
        """Overwrites stored rows (bench/test utility: planting near-duplicates into a synthetic store)."""
        self._flush()
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        data = np.ascontiguousarray(data, dtype=np.float32).reshape(len(rows), self.dim)
        check(_ffi.otters_vecstore_set_rows(self._handle(), rows.ctypes.data_as(_ffi.c_u64p), data.ctypes.data_as(_ffi.c_f32p), len(rows)))

    def inv_norms(self) -> np.ndarray:
        self._flush()
        out = np.zeros(self._n, dtype=np.float32)
        if self._n:
            check(_ffi.otters_vecstore_inv_norms(self._handle(), 0, self._n, out.ctypes.data_as(_ffi.c_f32p)))
        return out

    def query(self, queries, metric: Metric) -> "VecQueryPlan":
        """src/vec.rs:387-411."""
        return VecQueryPlan().with_vector_store(self).with_query_vectors(queries).with_metric(metric)

    def close(self):
        if self._h is not None:
            _ffi.otters_vecstore_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class VecQueryPlan:
    """Builder-style plan (src/vec.rs:55-311)."""

    def __init__(self):
        self._queries = None
        self._metric: Optional[Metric] = None
        self._filter = None
        self._take_type: Optional[TakeType] = None
        self._take_count: Optional[int] = None
        self._store: Optional[VecStore] = None
        self._row_mask = None

    @staticmethod
    def new() -> "VecQueryPlan":
        return VecQueryPlan()

    def with_vector_store(self, store: VecStore) -> "VecQueryPlan":
        self._store = store
        return self

    def with_query_vectors(self, queries) -> "VecQueryPlan":
        self._queries = _as_query_batch(queries)
        return self

    def with_metric(self, metric: Metric) -> "VecQueryPlan":
        self._metric = Metric(metric)
        return self

    def with_row_mask(self, mask) -> "VecQueryPlan":
        self._row_mask = np.asarray(mask, dtype=bool)
        return self

    def filter(self, score: float, cmp: Cmp) -> "VecQueryPlan":
        self._filter = (float(score), Cmp(cmp))
        return self

    def _take(self, count: int, tt: Optional[TakeType]) -> "VecQueryPlan":
        # src/vec.rs:103-116: the last take*() wins; take() only infers when no type is set yet
        self._take_count = int(count)
        if tt is not None:
            self._take_type = tt
        elif self._take_type is None and self._metric is not None:
            self._take_type = infer_default_take_type(self._metric)
        return self

    def take(self, count: int) -> "VecQueryPlan":
        return self._take(count, None)

    def take_min(self, count: int) -> "VecQueryPlan":
        return self._take(count, TakeType.Min)

    def take_max(self, count: int) -> "VecQueryPlan":
        return self._take(count, TakeType.Max)

    def _validate(self):
        """src/vec.rs:170-203 — the checks that do not need the device."""
        if self._queries is None:
            raise OttersError("Query vectors or their norms are not set")
        if self._metric is None:
            raise OttersError("Search metric is not set")
        if self._store is None:
            raise OttersError("Vector store is not set")
        if len(self._queries) == 0:
            raise OttersError("No queries provided")
        for q in self._queries:
            if q.shape[0] != self._store.dim:
                raise OttersError(
                    f"Query vector length {q.shape[0]} does not match expected dimension {self._store.dim}"
                )

    def _build_query(self):
        self._validate()
        store = self._store
        store._flush()
        n = store.len()
        k = self._take_count if self._take_count is not None else n  # src/vec.rs:213
        tt = self._take_type if self._take_type is not None else TakeType.Max  # src/vec.rs:214
        nq = len(self._queries)
        q = np.ascontiguousarray(np.stack(self._queries), dtype=np.float32)
        vq = _ffi.VecQuery()
        vq.queries = q.ctypes.data_as(_ffi.c_f32p)
        vq.nq, vq.dim = nq, store.dim
        vq.metric, vq.take_type, vq.k = int(self._metric), int(tt), k
        if self._filter is not None:
            vq.has_filter, vq.thr, vq.cmp = 1, self._filter[0], int(self._filter[1])
        words = None
        if self._row_mask is not None:
            words = pack_mask_words(self._row_mask)
            vq.row_mask_words = words.ctypes.data_as(_ffi.c_u64p)
            vq.row_mask_bits = len(self._row_mask)
        return vq, (q, words), min(k, n * nq)

    def collect_arrays(self):
        """Returns (indices u64, scores f32, query ids u32) best-first."""
        vq, keep, cap = self._build_query()
        store = self._store
        if cap == 0:
            return np.zeros(0, np.uint64), np.zeros(0, np.float32), np.zeros(0, np.uint32)
        idx = np.zeros(cap, np.uint64)
        score = np.zeros(cap, np.float32)
        qid = np.zeros(cap, np.uint32)
        out_len = C.c_uint64(0)
        check(
            _ffi.otters_vecstore_query(
                store._handle(),
                C.byref(vq),
                idx.ctypes.data_as(_ffi.c_u64p),
                score.ctypes.data_as(_ffi.c_f32p),
                qid.ctypes.data_as(_ffi.c_u32p),
                cap,
                C.byref(out_len),
            )
        )
        m = min(out_len.value, cap)
        return idx[:m], score[:m], qid[:m]

    def collect_per_query(self):
        """Extension (``otters_vecstore_query_batch``): a list with one (indices, scores) pair per query of the batch instead
        of the reference's single merged list; entry i is exactly what the single-query plan returns for query i."""
        vq, keep, cap = self._build_query()
        nq = len(self._queries)
        k_cap = max(min(vq.k, self._store.len()), 1)
        if vq.k > 0:
            vq.k = k_cap
        idx = np.zeros((nq, k_cap), np.uint64)
        score = np.zeros((nq, k_cap), np.float32)
        lens = np.zeros(nq, np.uint64)
        check(_ffi.otters_vecstore_query_batch(self._store._handle(), C.byref(vq), idx.ctypes.data_as(_ffi.c_u64p),
                                               score.ctypes.data_as(_ffi.c_f32p), lens.ctypes.data_as(_ffi.c_u64p)))
        return [(idx[i, : int(lens[i])].copy(), score[i, : int(lens[i])].copy()) for i in range(nq)]

    def submit(self) -> "PendingVecQuery":
        """Non-blocking form of collect_arrays() (``otters_query_submit``); at most two outstanding per context."""
        vq, keep, cap = self._build_query()
        ticket = C.c_uint64(0)
        check(_ffi.otters_query_submit(self._store._handle(), None, C.byref(vq), None, None, None, 0, C.byref(ticket)))
        return PendingVecQuery(self._store, ticket.value, cap)

    def collect(self) -> List[SearchResult]:
        """src/vec.rs:206-311."""
        idx, score, _ = self.collect_arrays()
        return [SearchResult(int(i), float(s)) for i, s in zip(idx, score)]


class PendingVecQuery:
    def __init__(self, store: VecStore, ticket: int, cap: int):
        self._store, self.ticket, self._cap = store, ticket, cap

    def wait(self):
        """(indices u64, scores f32, query ids u32) best-first, as collect_arrays() returns them."""
        cap = max(self._cap, 1)
        idx, score, qid = np.zeros(cap, np.uint64), np.zeros(cap, np.float32), np.zeros(cap, np.uint32)
        out_len = C.c_uint64(0)
        check(_ffi.otters_query_wait(self._store.ctx.handle, self.ticket, idx.ctypes.data_as(_ffi.c_u64p),
                                     score.ctypes.data_as(_ffi.c_f32p), qid.ctypes.data_as(_ffi.c_u32p), cap, C.byref(out_len), None))
        m = min(out_len.value, self._cap)
        return idx[:m], score[:m], qid[:m]
