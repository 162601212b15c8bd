"""MetaStore / MetaStoreBuilder / MetaQueryPlan — host-side mirror of the reference's ``meta`` module.

``build()`` uploads vectors and columnar metadata to HBM and builds per-chunk zonemaps / Bloom filters
(``otters_metastore_build``); ``collect()`` is one call into the CUDA library
(``otters_metastore_query``): chunk pruning, per-row predicate bitmask, scan, top-k, stats.  Gathering
the result columns for the <= k returned rows stays on the host (src/meta.rs:723-828).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Dict, List, Optional

import numpy as np

from . import _ffi
from .column import Column
from .context import Context, check, default_context
from .expr import CompiledFilter, Expr, ExprError
from .types import Cmp, DataType, Metric, OttersError, TakeType, VectorFormat, infer_default_take_type
from .vec import _as_query_batch


@dataclass
class MetaQueryStats:
    """src/meta.rs:832-842 (durations in seconds)."""

    total_chunks: int
    pruned_chunks: int
    evaluated_chunks: int
    vectors_compared: int
    prune_duration: float
    score_duration: float
    merge_duration: float
    total_duration: float


@dataclass
class MetaBuildStats:
    """src/meta.rs:844-852 (durations in seconds)."""

    n_rows: int
    dim: int
    n_chunks: int
    vectors_ingest_duration: float
    zonemap_build_duration: float
    build_total_duration: float


class MetaQueryResults:
    """src/meta.rs:23-40."""

    def __init__(self, columns, data, indices, scores, query_ids=None):
        self.columns: List[str] = columns
        self.data: Dict[str, Column] = data
        self.indices: List[int] = indices
        self.scores: List[float] = scores
        self.query_ids = query_ids

    def len(self) -> int:
        return len(self.indices)

    __len__ = len

    def is_empty(self) -> bool:
        return not self.indices

    def column(self, name: str) -> Optional[Column]:
        return self.data.get(name)


class FilterPack:
    """Keeps the ctypes buffers of an ``otters_filter`` alive."""

    def __init__(self, compiled: CompiledFilter, col_index: Dict[str, int]):
        leaves = [lf for clause in compiled.clauses for lf in clause]
        offs = [0]
        for clause in compiled.clauses:
            offs.append(offs[-1] + len(clause))
        self._offs = (C.c_uint32 * len(offs))(*offs)
        self._leaves = (_ffi.Leaf * max(len(leaves), 1))()
        self._strs = []
        for i, lf in enumerate(leaves):
            L = self._leaves[i]
            if lf.column not in col_index:
                raise OttersError(f"Unknown column '{lf.column}'")
            L.col = col_index[lf.column]
            L.op = int(lf.cmp)
            if lf.kind == "i64":
                L.kind, L.i = 0, lf.rhs
            elif lf.kind == "f64":
                L.kind, L.f = 1, lf.rhs
            else:
                b = lf.rhs.encode("utf-8")
                self._strs.append(b)
                L.kind, L.s, L.slen = 2, b, len(b)
        self.c = _ffi.Filter(len(compiled.clauses), self._offs, self._leaves)

    def byref(self):
        return C.byref(self.c)


class MetaStoreBuilder:
    """src/meta.rs:62-306."""

    def __init__(self, schema: Dict[str, DataType], columns: Dict[str, Column], order: List[str]):
        self._schema = schema
        self._columns = columns
        self._order = order
        self._vectors = None
        self._synthetic = None
        self._chunk_size = 1024
        self._bloom = ("fpr", 0.01)
        self._ctx: Optional[Context] = None
        self._vector_format = VectorFormat.F32
        self._row_order = None  # (columns, method) — see with_row_order

    def with_vector_format(self, vector_format: VectorFormat) -> "MetaStoreBuilder":
        """Extension (the reference's roadmap item "Quantization for vectors", README.md:208): ``VectorFormat.Bf16`` keeps the
        rows as bf16 — half the bytes every scan streams; scores are the reference's arithmetic on the rounded rows."""
        self._vector_format = VectorFormat(vector_format)
        return self

    def with_row_order(self, by, method: str = "sort") -> "MetaStoreBuilder":
        """Extension (the reference's roadmap item "reorder metadata for better pruning (Something like Z-ordering)",
        README.md:154,212): the store keeps its rows clustered on the filter columns ``by`` (``method`` "sort" or "zorder",
        otters_b200/reorder.py), so that zonemaps and Bloom filters prune more chunks.  Results keep reporting the caller's
        row ids; among rows with EQUAL scores the order follows the store position (the reference leaves it undefined)."""
        from .reorder import METHODS

        by = [by] if isinstance(by, str) else list(by)
        if method not in METHODS:
            raise OttersError(f"unknown row order method '{method}' (expected one of {', '.join(METHODS)})")
        self._row_order = (by, method)
        return self

    def with_vectors(self, vectors) -> "MetaStoreBuilder":
        self._vectors = vectors
        return self

    def with_synthetic_vectors(self, n_rows: int, dim: int, seed: int, first_row: int = 0, shard=None) -> "MetaStoreBuilder":
        """Vectors from the device-side synthetic generator (bench/test utility).  `shard` = (world, rank,
        block_rows) generates the rows a rank holds under block-cyclic sharding."""
        self._synthetic = (int(n_rows), int(dim), int(seed), int(first_row), shard)
        return self

    def with_context(self, ctx: Context) -> "MetaStoreBuilder":
        self._ctx = ctx
        return self

    def with_chunk_size(self, chunk_size: int) -> "MetaStoreBuilder":
        self._chunk_size = max(int(chunk_size), 1)  # src/meta.rs:86-89
        return self

    def with_bloom_fpr(self, fpr: float) -> "MetaStoreBuilder":
        f = min(max(fpr, 1e-2), 0.5) if math.isfinite(fpr) else 0.01  # src/meta.rs:92-101
        self._bloom = ("fpr", f)
        return self

    def with_bloom_bits(self, bits: int) -> "MetaStoreBuilder":
        self._bloom = ("bits", max(int(bits), 64))  # src/meta.rs:106-110
        return self

    def with_column(self, name: str, column: Column) -> "MetaStoreBuilder":
        if name not in self._schema:
            raise OttersError(f"unknown column '{name}' not present in schema")
        if self._schema[name] != column.dtype():
            raise OttersError(
                f"dtype mismatch for column '{name}': schema {self._schema[name].name}, got {column.dtype().name}"
            )
        self._columns[name] = column
        return self

    def with_columns(self, columns) -> "MetaStoreBuilder":
        for name, c in columns:
            self.with_column(name, c)
        return self

    def build(self) -> "MetaStore":
        """src/meta.rs:151-305."""
        if self._vectors is None and self._synthetic is None:
            raise OttersError("vectors must be provided to build MetaStore")
        if self._synthetic is not None:
            n_rows, dim = self._synthetic[0], self._synthetic[1]
            vec_arr = None
        else:
            v = self._vectors
            if isinstance(v, np.ndarray) and v.ndim == 2:
                vec_arr = np.ascontiguousarray(v, dtype=np.float32)
                n_rows, dim = vec_arr.shape
            else:
                rows = [np.asarray(r, dtype=np.float32) for r in v]
                n_rows = len(rows)
                dim = rows[0].shape[0] if n_rows else 0
                if dim == 0 and n_rows > 0:
                    raise OttersError("vector dimension cannot be zero")
                for i, r in enumerate(rows):
                    if r.shape[0] != dim:
                        raise OttersError(f"vector at index {i} has dim {r.shape[0]}, expected {dim}")
                vec_arr = np.ascontiguousarray(np.stack(rows), dtype=np.float32) if n_rows else np.zeros((0, 0), np.float32)
        for name in self._schema:
            colobj = self._columns.get(name)
            if colobj is None:
                raise OttersError(f"missing column '{name}' in builder columns")
            if colobj.len() != n_rows:
                raise OttersError(f"column '{name}' length {colobj.len()} does not match vectors length {n_rows}")

        # row order (extension): cluster the rows on the filter columns before anything is chunked
        perm = None
        columns = self._columns
        if self._row_order is not None:
            from .reorder import compute_row_order

            if vec_arr is None:
                raise OttersError("with_row_order needs host vectors (synthetic vectors are generated in row order)")
            perm = compute_row_order(columns, self._row_order[0], self._row_order[1])
            columns = {name: c.gather(perm) for name, c in columns.items()}
            vec_arr = np.ascontiguousarray(vec_arr[perm.astype(np.int64)]) if n_rows else vec_arr

        ctx = self._ctx or default_context()
        keep = []  # keeps numpy buffers alive during the call
        ccols = (_ffi.Column * max(len(self._order), 1))()
        for i, name in enumerate(self._order):
            colobj = columns[name]
            cc = ccols[i]
            nb = name.encode("utf-8")
            keep.append(nb)
            cc.name = nb
            cc.dtype = int(colobj.dtype())
            nw = colobj.null_words()
            if nw is not None:
                keep.append(nw)
                cc.null_words = nw.ctypes.data_as(_ffi.c_u64p)
            if colobj.dtype() == DataType.String:
                offs, data = colobj.string_buffers()
                keep += [offs, data]
                cc.str_offsets = offs.ctypes.data_as(_ffi.c_u64p)
                cc.str_bytes = data.ctypes.data_as(_ffi.c_u8p)
            else:
                arr = colobj.numpy()
                keep.append(arr)
                cc.values = arr.ctypes.data if arr.size else None
        bp = _ffi.BuildParams()
        bp.n_rows, bp.dim, bp.chunk_size = n_rows, dim, self._chunk_size
        if self._bloom[0] == "fpr":
            bp.bloom_mode, bp.bloom_fpr = 0, self._bloom[1]
        else:
            bp.bloom_mode, bp.bloom_bits = 1, self._bloom[1]
        if self._synthetic is not None:
            bp.vectors_kind = _ffi.VECTORS_SYNTHETIC
            bp.synthetic_seed, bp.synthetic_first_row = self._synthetic[2], self._synthetic[3]
            if self._synthetic[4] is not None:
                w, r, b = self._synthetic[4]
                smap = _ffi.ShardMap(self._synthetic[3], w, r, b)
                keep.append(smap)
                bp.synthetic_map = C.cast(C.pointer(smap), C.c_void_p)
        else:
            bp.vectors_kind = _ffi.VECTORS_HOST
            bp.vectors = vec_arr.ctypes.data if vec_arr.size else None
        bp.columns, bp.n_columns = ccols, len(self._order)
        bp.vector_format = int(self._vector_format)
        h = C.c_void_p()
        bs = _ffi.BuildStats()
        check(_ffi.otters_metastore_build(ctx.handle, C.byref(bp), C.byref(h), C.byref(bs)))
        stats = MetaBuildStats(bs.n_rows, bs.dim, bs.n_chunks, bs.vectors_ingest_s, bs.zonemap_build_s, bs.build_total_s)
        return MetaStore(ctx, h, dict(self._schema), dict(columns), list(self._order), self._chunk_size, dim, n_rows, stats, perm,
                         self._vector_format)


class MetaStore:
    """src/meta.rs:48-60, :308-577."""

    def __init__(self, ctx, handle, schema, columns, order, chunk_size, dim, n_rows, build_stats, perm=None,
                 vector_format=VectorFormat.F32):
        self._perm = perm  # store position -> caller's row id (with_row_order); None = identity
        self._vector_format = vector_format
        self._ctx = ctx
        self._h = handle
        self._schema = schema
        self._columns = columns
        self._order = order
        self._col_index = {n: i for i, n in enumerate(order)}
        self._chunk_size = chunk_size
        self._dim = dim
        self._n_rows = n_rows
        self._build_stats = build_stats
        self._last_stats: Optional[MetaQueryStats] = None

    @staticmethod
    def from_columns(columns: List[Column]) -> MetaStoreBuilder:
        schema, cols, order = {}, {}, []
        for c in columns:
            if c.name() not in schema:
                order.append(c.name())
            schema[c.name()] = c.dtype()
            cols[c.name()] = c
        return MetaStoreBuilder(schema, cols, order)

    @staticmethod
    def from_schema(schema) -> MetaStoreBuilder:
        sch, cols, order = {}, {}, []
        for name, dt in schema:
            if name not in sch:
                order.append(name)
            sch[name] = DataType(dt)
            cols[name] = Column(name, DataType(dt))
        return MetaStoreBuilder(sch, cols, order)

    # ---- persistence (extension: the reference's roadmap item "save/load MetaStore to/from disk", README.md:206) ----------
    def save(self, path: str) -> None:
        """Writes the store's HBM image to ``path`` (``otters_metastore_save``); a row order (with_row_order) travels in the
        file's caller blob, so a loaded store keeps reporting the caller's row ids."""
        blob = np.ascontiguousarray(self._perm, dtype=np.uint64) if self._perm is not None else None
        check(_ffi.otters_metastore_save(self._h, str(path).encode("utf-8"), C.c_void_p(blob.ctypes.data) if blob is not None else None,
                                         blob.nbytes if blob is not None else 0))

    @staticmethod
    def load(path: str, ctx: Optional[Context] = None) -> "MetaStore":
        """Loads a store written by ``save`` (``otters_metastore_load``): device arrays are copied back as they were, nothing is
        rebuilt.  The host-side Column objects are not part of the file: ``columns()`` is empty, result columns come from the
        device gather as always."""
        ctx = ctx or default_context()
        h = C.c_void_p()
        check(_ffi.otters_metastore_load(ctx.handle, str(path).encode("utf-8"), C.byref(h)))
        schema, order = {}, []
        for i in range(_ffi.otters_metastore_n_columns(h)):
            name, dt = C.c_char_p(), C.c_int32(0)
            check(_ffi.otters_metastore_column_info(h, i, C.byref(name), C.byref(dt)))
            schema[name.value.decode("utf-8")] = DataType(dt.value)
            order.append(name.value.decode("utf-8"))
        p, ln = C.c_void_p(), C.c_uint64(0)
        check(_ffi.otters_metastore_user_blob(h, C.byref(p), C.byref(ln)))
        perm = np.frombuffer(C.string_at(p, ln.value), dtype=np.uint64).copy() if ln.value else None
        return MetaStore(ctx, h, schema, {}, order, int(_ffi.otters_metastore_chunk_size(h)), int(_ffi.otters_metastore_dim(h)),
                         int(_ffi.otters_metastore_len(h)), None, perm, VectorFormat(_ffi.otters_metastore_format(h)))

    @property
    def ctx(self) -> Context:
        return self._ctx

    @property
    def handle(self):
        return self._h

    def schema(self) -> Dict[str, DataType]:
        return self._schema

    def columns(self) -> Dict[str, Column]:
        """The columns in STORE order (the input order unless the store was built with_row_order)."""
        return self._columns

    def row_order(self) -> Optional[np.ndarray]:
        """perm[i] = caller's row id of store position i, or None when the store keeps the input order."""
        return self._perm

    def vector_format(self) -> VectorFormat:
        return self._vector_format

    def column_index(self) -> Dict[str, int]:
        return self._col_index

    def n_chunks(self) -> int:
        return int(_ffi.otters_metastore_n_chunks(self._h))

    def gather(self, name: str, indices) -> Column:
        """Column ``name`` at the result rows ``indices``, gathered on the DEVICE (``otters_metastore_gather``;
        MetaQueryResults.data of the reference, src/meta.rs:723-821).  NULLs are preserved; strings come back through the
        store's dictionary."""
        col = self._col_index[name]
        dt = self._schema[name]
        rows = np.ascontiguousarray(indices, dtype=np.uint64)
        n = len(rows)
        np_t = {DataType.Int32: np.int32, DataType.Int64: np.int64, DataType.Float32: np.float32, DataType.Float64: np.float64,
                DataType.DateTime: np.int64, DataType.String: np.uint32}[dt]
        vals = np.zeros(max(n, 1), np_t)
        nulls = np.zeros(max(n, 1), np.uint8)
        check(_ffi.otters_metastore_gather(self._h, col, rows.ctypes.data_as(_ffi.c_u64p), n, C.c_void_p(vals.ctypes.data),
                                           nulls.ctypes.data_as(_ffi.c_u8p)))
        vals, nulls = vals[:n], nulls[:n].astype(bool)
        if dt != DataType.String:
            return Column.from_numpy(name, dt, vals, nulls)
        vocab: Dict[int, str] = {}
        for code in np.unique(vals[~nulls]):
            p, ln = _ffi.c_u8p(), C.c_uint64(0)
            check(_ffi.otters_metastore_dict_entry(self._h, col, int(code), C.byref(p), C.byref(ln)))
            vocab[int(code)] = C.string_at(p, ln.value).decode("utf-8")
        out = Column(name, DataType.String)
        for v, isnull in zip(vals, nulls):
            out.push(None if isnull else vocab[int(v)])
        return out

    def chunk_size(self) -> int:
        return self._chunk_size

    def len(self) -> int:
        return self._n_rows

    def dim(self) -> int:
        return self._dim

    def last_query_stats(self) -> Optional[MetaQueryStats]:
        return self._last_stats

    def build_stats(self) -> Optional[MetaBuildStats]:
        return self._build_stats

    def query(self, query, metric: Metric) -> "MetaQueryPlan":
        return MetaQueryPlan(self, [np.asarray(query, dtype=np.float32)], metric)

    def query_batch(self, queries, metric: Metric) -> "MetaQueryPlan":
        return MetaQueryPlan(self, _as_query_batch(queries) if len(queries) else [], metric)

    # ---- parity/debug exports -----------------------------------------------------------------
    def _pack(self, flt):
        if flt is None:
            return None
        compiled = flt.compile(self._schema) if isinstance(flt, Expr) else flt
        return FilterPack(compiled, self._col_index)

    def chunk_mask(self, flt) -> np.ndarray:
        out = np.zeros(max(self.n_chunks(), 1), np.uint8)
        fp = self._pack(flt)
        check(_ffi.otters_metastore_chunk_mask(self._h, fp.byref() if fp else None, out.ctypes.data_as(_ffi.c_u8p)))
        return out[: self.n_chunks()]

    def row_mask(self, flt) -> np.ndarray:
        out = np.zeros(max(self._n_rows, 1), np.uint8)
        fp = self._pack(flt)
        check(_ffi.otters_metastore_row_mask(self._h, fp.byref() if fp else None, out.ctypes.data_as(_ffi.c_u8p)))
        return out[: self._n_rows]

    def zonemap(self, name: str):
        i = self._col_index[name]
        nc = self.n_chunks()
        nn = np.zeros(max(nc, 1), np.uint64)
        if self._schema[name] in (DataType.Float32, DataType.Float64):
            mn, mx = np.zeros(max(nc, 1), np.float64), np.zeros(max(nc, 1), np.float64)
            check(
                _ffi.otters_metastore_zonemap_f64(
                    self._h, i, mn.ctypes.data_as(C.POINTER(C.c_double)), mx.ctypes.data_as(C.POINTER(C.c_double)), nn.ctypes.data_as(_ffi.c_u64p)
                )
            )
        else:
            mn, mx = np.zeros(max(nc, 1), np.int64), np.zeros(max(nc, 1), np.int64)
            check(
                _ffi.otters_metastore_zonemap_i64(
                    self._h, i, mn.ctypes.data_as(C.POINTER(C.c_int64)), mx.ctypes.data_as(C.POINTER(C.c_int64)), nn.ctypes.data_as(_ffi.c_u64p)
                )
            )
        return mn[:nc], mx[:nc], nn[:nc]

    def set_rows(self, rows, data) -> None:
        """Overwrites stored vectors (bench/test utility: planting near-duplicates into a synthetic store)."""
        rows = np.ascontiguousarray(rows, dtype=np.uint64)
        data = np.ascontiguousarray(data, dtype=np.float32).reshape(len(rows), -1)
        check(_ffi.otters_metastore_set_rows(self._h, rows.ctypes.data_as(_ffi.c_u64p), data.ctypes.data_as(_ffi.c_f32p), len(rows)))

    def inv_norms(self) -> np.ndarray:
        out = np.zeros(max(self._n_rows, 1), np.float32)
        check(_ffi.otters_metastore_inv_norms(self._h, 0, self._n_rows, out.ctypes.data_as(_ffi.c_f32p)))
        return out[: self._n_rows]

    def close(self):
        if self._h is not None:
            _ffi.otters_metastore_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class MetaQueryPlan:
    """src/meta.rs:579-830."""

    def __init__(self, store: MetaStore, queries, metric: Metric):
        self._store = store
        self._queries = queries
        self._metric = Metric(metric)
        self._meta_filter: Optional[CompiledFilter] = None
        self._meta_error: Optional[str] = None
        self._vec_filter = None
        self._take_type: Optional[TakeType] = None
        self._take_count: Optional[int] = None

    def meta_filter(self, expr) -> "MetaQueryPlan":
        """src/meta.rs:605-616: compile now, surface the error at collect()."""
        try:
            self._meta_filter = expr.compile(self._store.schema()) if isinstance(expr, Expr) else expr
            self._meta_error = None
        except ExprError as e:
            self._meta_error = f"meta_filter compile error: {e}"
        return self

    def vec_filter(self, score: float, cmp: Cmp) -> "MetaQueryPlan":
        self._vec_filter = (float(score), Cmp(cmp))
        return self

    def take(self, k: int) -> "MetaQueryPlan":
        self._take_count = int(k)
        self._take_type = infer_default_take_type(self._metric)  # src/meta.rs:623-630
        return self

    def build_query(self):
        """(VecQuery struct, FilterPack|None, keepalive) with the defaults of src/meta.rs:638-644 resolved."""
        store = self._store
        k = self._take_count if self._take_count is not None else store.len()
        tt = self._take_type if self._take_type is not None else infer_default_take_type(self._metric)
        nq = len(self._queries)
        dims = {q.shape[0] for q in self._queries}
        # mixed or wrong dimensions are swallowed per chunk by the reference (src/meta_compute.rs:182)
        bad = nq == 0 or len(dims) != 1
        vq = _ffi.VecQuery()
        q = None
        if not bad:
            q = np.ascontiguousarray(np.stack(self._queries), dtype=np.float32)
            vq.queries = q.ctypes.data_as(_ffi.c_f32p)
            vq.nq, vq.dim = nq, q.shape[1]
        else:
            vq.nq, vq.dim = nq, 0
        vq.metric, vq.take_type, vq.k = int(self._metric), int(tt), k
        if self._vec_filter is not None:
            vq.has_filter, vq.thr, vq.cmp = 1, self._vec_filter[0], int(self._vec_filter[1])
        fp = FilterPack(self._meta_filter, store.column_index()) if self._meta_filter is not None else None
        return vq, fp, q, k, nq

    def collect(self) -> MetaQueryResults:
        """src/meta.rs:632-829."""
        if self._meta_error is not None:
            raise OttersError(self._meta_error)
        store = self._store
        vq, fp, q, k, nq = self.build_query()
        cap = max(min(k, store.len() * max(nq, 1)), 1)
        idx = np.zeros(cap, np.uint64)
        score = np.zeros(cap, np.float32)
        qid = np.zeros(cap, np.uint32)
        out_len = C.c_uint64(0)
        st = _ffi.QueryStats()
        check(
            _ffi.otters_metastore_query(
                store.handle,
                C.byref(vq),
                fp.byref() if fp else None,
                idx.ctypes.data_as(_ffi.c_u64p),
                score.ctypes.data_as(_ffi.c_f32p),
                qid.ctypes.data_as(_ffi.c_u32p),
                cap,
                C.byref(out_len),
                C.byref(st),
            )
        )
        return self._results(idx, score, qid, min(out_len.value, cap), st)

    def _results(self, idx, score, qid, m, st) -> MetaQueryResults:
        store = self._store
        store._last_stats = MetaQueryStats(
            st.total_chunks, st.pruned_chunks, st.evaluated_chunks, st.vectors_compared, st.prune_s, st.score_s, st.merge_s, st.total_s
        )
        indices = [int(i) for i in idx[:m]]
        names = sorted(store.schema().keys())  # src/meta.rs:723-724
        data = {n: store.gather(n, indices) for n in names}  # gathered on the device (store positions)
        if store._perm is not None:  # with_row_order: report the caller's row ids
            indices = [int(store._perm[i]) for i in indices]
        return MetaQueryResults(names, data, indices, [float(s) for s in score[:m]], [int(x) for x in qid[:m]])

    def collect_per_query(self) -> List[MetaQueryResults]:
        """Extension (``otters_metastore_query_batch``): one result per query of the batch instead of the reference's single
        merged list; entry i is exactly what ``store.query(queries[i], metric)...collect()`` returns."""
        if self._meta_error is not None:
            raise OttersError(self._meta_error)
        store = self._store
        vq, fp, q, k, nq = self.build_query()
        k_cap = max(min(k, store.len()), 1)
        vq.k = k_cap if k > 0 else 0
        idx = np.zeros((max(nq, 1), k_cap), np.uint64)
        score = np.zeros((max(nq, 1), k_cap), np.float32)
        lens = np.zeros(max(nq, 1), np.uint64)
        st = _ffi.QueryStats()
        check(_ffi.otters_metastore_query_batch(store.handle, C.byref(vq), fp.byref() if fp else None, idx.ctypes.data_as(_ffi.c_u64p),
                                                score.ctypes.data_as(_ffi.c_f32p), lens.ctypes.data_as(_ffi.c_u64p), C.byref(st)))
        return [self._results(idx[i], score[i], np.full(k_cap, i, np.uint32), int(lens[i]), st) for i in range(nq)]

    def submit(self) -> "PendingMetaQuery":
        """Non-blocking form of collect() (``otters_query_submit``): the query is enqueued on one of the context's two lanes and
        runs while the caller prepares the next one; ``wait()`` on the returned object gives what collect() would have returned.
        Keep at most two queries outstanding per context."""
        if self._meta_error is not None:
            raise OttersError(self._meta_error)
        vq, fp, q, k, nq = self.build_query()
        ticket = C.c_uint64(0)
        check(_ffi.otters_query_submit(None, self._store.handle, C.byref(vq), fp.byref() if fp else None, None, None, 0, C.byref(ticket)))
        return PendingMetaQuery(self, ticket.value, max(min(k, self._store.len() * max(nq, 1)), 1))


class PendingMetaQuery:
    def __init__(self, plan: MetaQueryPlan, ticket: int, cap: int):
        self._plan, self.ticket, self._cap = plan, ticket, cap

    def wait(self) -> MetaQueryResults:
        cap = self._cap
        idx, score, qid = np.zeros(cap, np.uint64), np.zeros(cap, np.float32), np.zeros(cap, np.uint32)
        out_len = C.c_uint64(0)
        st = _ffi.QueryStats()
        check(_ffi.otters_query_wait(self._plan._store.ctx.handle, self.ticket, idx.ctypes.data_as(_ffi.c_u64p),
                                     score.ctypes.data_as(_ffi.c_f32p), qid.ctypes.data_as(_ffi.c_u32p), cap, C.byref(out_len), C.byref(st)))
        return self._plan._results(idx, score, qid, min(out_len.value, cap), st)
