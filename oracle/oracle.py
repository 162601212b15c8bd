"""ctypes binding of the CPU oracle (oracle/libotters_oracle.so).

TEST INFRASTRUCTURE ONLY — see oracle/otters_oracle.h.  Imported by tests/, __graft_entry__.smoke()
and bench.py's cpu_baseline / --impl reference legs; never by the otters_b200 package.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from typing import List, Optional, Sequence, Tuple

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libotters_oracle.so")


def build(force: bool = False) -> str:
    """Compiles the C restatement with gcc (oracle/Makefile)."""
    src_m = max(os.path.getmtime(os.path.join(_HERE, f)) for f in ("otters_oracle.c", "otters_oracle.h", "Makefile"))
    if force or not os.path.exists(LIB_PATH) or os.path.getmtime(LIB_PATH) < src_m:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return LIB_PATH


_lib = None

u64p = C.POINTER(C.c_uint64)
u32p = C.POINTER(C.c_uint32)
u8p = C.POINTER(C.c_uint8)
f32p = C.POINTER(C.c_float)


class VecQuery(C.Structure):
    _fields_ = [
        ("queries", f32p),
        ("nq", C.c_uint32),
        ("dim", C.c_uint32),
        ("metric", C.c_int32),
        ("take_type", C.c_int32),
        ("k", C.c_uint64),
        ("has_filter", C.c_int32),
        ("thr", C.c_float),
        ("cmp", C.c_int32),
        ("row_mask_words", u64p),
        ("row_mask_bits", C.c_uint64),
    ]


class OraColumn(C.Structure):
    _fields_ = [
        ("dtype", C.c_int32),
        ("values", C.c_void_p),
        ("null_words", u64p),
        ("str_offsets", u64p),
        ("str_bytes", u8p),
    ]


class OraLeaf(C.Structure):
    _fields_ = [
        ("col", C.c_uint32),
        ("op", C.c_int32),
        ("kind", C.c_int32),
        ("i", C.c_int64),
        ("f", C.c_double),
        ("s", C.c_char_p),
        ("slen", C.c_uint64),
    ]


class OraFilter(C.Structure):
    _fields_ = [("n_clauses", C.c_uint32), ("clause_offsets", u32p), ("leaves", C.POINTER(OraLeaf))]


class OraStats(C.Structure):
    _fields_ = [
        ("total_chunks", C.c_uint64),
        ("pruned_chunks", C.c_uint64),
        ("evaluated_chunks", C.c_uint64),
        ("vectors_compared", C.c_uint64),
        ("prune_s", C.c_double),
        ("score_s", C.c_double),
        ("merge_s", C.c_double),
        ("total_s", C.c_double),
    ]


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(LIB_PATH)
        L.oracle_dot.restype = C.c_float
        L.oracle_dot.argtypes = [f32p, f32p, C.c_size_t]
        L.oracle_cosine.restype = C.c_float
        L.oracle_cosine.argtypes = [f32p, f32p, C.c_float, C.c_float, C.c_size_t]
        L.oracle_l2.restype = C.c_float
        L.oracle_l2.argtypes = [f32p, f32p, C.c_size_t]
        L.oracle_inv_norm.restype = C.c_float
        L.oracle_inv_norm.argtypes = [f32p, C.c_size_t]
        L.oracle_inv_norms.restype = None
        L.oracle_inv_norms.argtypes = [f32p, C.c_size_t, C.c_size_t, f32p]
        L.oracle_filter_mask_bits.restype = C.c_uint8
        L.oracle_filter_mask_bits.argtypes = [f32p, C.c_float, C.c_int]
        L.oracle_set_reduce_order.restype = None
        L.oracle_set_reduce_order.argtypes = [C.c_int]
        L.oracle_vecstore_query.restype = C.c_uint64
        L.oracle_vecstore_query.argtypes = [f32p, f32p, C.c_uint64, C.POINTER(VecQuery), C.c_int, u64p, f32p, u32p, C.c_uint64]
        L.oracle_meta_build.restype = C.c_void_p
        L.oracle_meta_build.argtypes = [
            f32p, C.c_uint64, C.c_uint32, C.c_uint64, C.POINTER(OraColumn), C.c_uint32, C.c_int, C.c_double, C.c_uint64,
        ]
        L.oracle_meta_free.restype = None
        L.oracle_meta_free.argtypes = [C.c_void_p]
        L.oracle_meta_n_chunks.restype = C.c_uint64
        L.oracle_meta_n_chunks.argtypes = [C.c_void_p]
        L.oracle_meta_zonemap_i64.restype = C.c_int
        L.oracle_meta_zonemap_i64.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_int64), C.POINTER(C.c_int64), u64p]
        L.oracle_meta_zonemap_f64.restype = C.c_int
        L.oracle_meta_zonemap_f64.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_double), C.POINTER(C.c_double), u64p]
        L.oracle_meta_chunk_mask.restype = None
        L.oracle_meta_chunk_mask.argtypes = [C.c_void_p, C.POINTER(OraFilter), u8p]
        L.oracle_meta_row_mask.restype = None
        L.oracle_meta_row_mask.argtypes = [C.c_void_p, C.POINTER(OraFilter), u8p]
        L.oracle_meta_query.restype = C.c_uint64
        L.oracle_meta_query.argtypes = [
            C.c_void_p, C.POINTER(VecQuery), C.POINTER(OraFilter), C.c_int, C.c_int, u64p, f32p, u32p, C.c_uint64, C.POINTER(OraStats),
        ]
        L.oracle_bloom_params.restype = None
        L.oracle_bloom_params.argtypes = [C.c_uint64, C.c_int, C.c_double, C.c_uint64, u64p, u32p]
        L.oracle_bloom_hash.restype = None
        L.oracle_bloom_hash.argtypes = [C.c_char_p, C.c_uint64, u64p, u64p]
        L.oracle_synth_fill.restype = None
        L.oracle_synth_fill.argtypes = [f32p, C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint64]
        L.oracle_num_threads.restype = C.c_int
        L.oracle_num_threads.argtypes = []
        _lib = L
    return _lib


FAITHFUL, CANONICAL = 0, 1


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _fp(a):
    return a.ctypes.data_as(f32p)


def dot(a, b) -> float:
    a, b = _f32(a), _f32(b)
    return float(lib().oracle_dot(_fp(a), _fp(b), len(a)))


def cosine(a, b, a_inv, b_inv) -> float:
    a, b = _f32(a), _f32(b)
    return float(lib().oracle_cosine(_fp(a), _fp(b), np.float32(a_inv), np.float32(b_inv), len(a)))


def l2(a, b) -> float:
    a, b = _f32(a), _f32(b)
    return float(lib().oracle_l2(_fp(a), _fp(b), len(a)))


def inv_norm(v) -> float:
    v = _f32(v)
    return float(lib().oracle_inv_norm(_fp(v), len(v)))


def inv_norms(rows) -> np.ndarray:
    rows = _f32(rows)
    out = np.zeros(rows.shape[0], np.float32)
    if rows.shape[0]:
        lib().oracle_inv_norms(_fp(rows), rows.shape[0], rows.shape[1], _fp(out))
    return out


def filter_mask_bits(scores8, thr, cmp) -> int:
    s = _f32(scores8)
    return int(lib().oracle_filter_mask_bits(_fp(s), np.float32(thr), int(cmp)))


def set_reduce_order(order: int) -> None:
    lib().oracle_set_reduce_order(int(order))


def synth_fill(row0: int, n_rows: int, dim: int, seed: int) -> np.ndarray:
    out = np.zeros((n_rows, dim), np.float32)
    if n_rows:
        lib().oracle_synth_fill(_fp(out), row0, n_rows, dim, seed)
    return out


def round_bf16(x) -> np.ndarray:
    """f32(bf16(x)) with round-to-nearest-even — the values a store created with OTTERS_VECTORS_FMT_BF16 holds.  The reference
    has no reduced-precision rows (they are a roadmap item, README.md:208); the parity contract of such a store is the
    reference's arithmetic on these rounded rows, so the oracle rounds its INPUT here and everything after is unchanged.
    IEEE rule restated directly: keep the upper 16 bits, round on the lower 16 (ties to the even upper half); NaN stays NaN."""
    a = np.ascontiguousarray(x, dtype=np.float32)
    bits = a.view(np.uint32).astype(np.uint64)
    upper, lower = bits >> 16, bits & 0xFFFF
    up = (lower > 0x8000) | ((lower == 0x8000) & ((upper & 1) == 1))
    out = ((upper + up.astype(np.uint64)) << 16).astype(np.uint32)
    nan = np.isnan(a)
    out = np.where(nan, (a.view(np.uint32) & np.uint32(0xFFFF0000)) | np.uint32(0x00400000), out).astype(np.uint32)
    return out.view(np.float32).reshape(a.shape)


def num_threads() -> int:
    return int(lib().oracle_num_threads())


def bloom_params(n_items, mode, fpr, bits) -> Tuple[int, int]:
    m, k = C.c_uint64(), C.c_uint32()
    lib().oracle_bloom_params(n_items, mode, fpr, bits, C.byref(m), C.byref(k))
    return m.value, k.value


def bloom_hash(s: bytes) -> Tuple[int, int]:
    h1, h2 = C.c_uint64(), C.c_uint64()
    lib().oracle_bloom_hash(s, len(s), C.byref(h1), C.byref(h2))
    return h1.value, h2.value


def _pack_mask(mask):
    m = np.asarray(mask, dtype=bool)
    n = len(m)
    padded = np.zeros((n + 63) // 64 * 64, dtype=np.uint8)
    padded[:n] = m
    w = np.packbits(padded, bitorder="little").view(np.uint64).copy()
    return w if len(w) else np.zeros(1, np.uint64)


def _make_query(queries, dim, metric, take_type, k, vec_filter, row_mask, keep):
    q = _f32(queries).reshape(-1, dim) if np.size(queries) else np.zeros((0, dim), np.float32)
    keep.append(q)
    vq = VecQuery()
    vq.queries = _fp(q)
    vq.nq, vq.dim = q.shape[0], dim
    vq.metric, vq.take_type, vq.k = int(metric), int(take_type), int(k)
    if vec_filter is not None:
        vq.has_filter, vq.thr, vq.cmp = 1, float(vec_filter[0]), int(vec_filter[1])
    if row_mask is not None:
        w = _pack_mask(row_mask)
        keep.append(w)
        vq.row_mask_words = w.ctypes.data_as(u64p)
        vq.row_mask_bits = len(row_mask)
    return vq


def vecstore_query(vectors, queries, metric, take_type, k, vec_filter=None, row_mask=None, mode=CANONICAL, inv=None):
    """VecQueryPlan::collect on the CPU.  Returns (idx u64, score f32, qid u32)."""
    vectors = _f32(vectors)
    n, dim = vectors.shape if vectors.ndim == 2 else (0, np.shape(queries)[-1])
    if inv is None:
        inv = inv_norms(vectors) if n else np.zeros(0, np.float32)
    keep = []
    vq = _make_query(queries, dim, metric, take_type, k, vec_filter, row_mask, keep)
    cap = int(min(k, max(n * vq.nq, 0)))
    idx, score, qid = np.zeros(max(cap, 1), np.uint64), np.zeros(max(cap, 1), np.float32), np.zeros(max(cap, 1), np.uint32)
    m = lib().oracle_vecstore_query(
        _fp(vectors) if n else None, _fp(inv) if n else None, n, C.byref(vq), mode,
        idx.ctypes.data_as(u64p), _fp(score), qid.ctypes.data_as(u32p), cap,
    )
    return idx[:m], score[:m], qid[:m]


class FilterPack:
    """clauses: list of clauses; leaf = (col_index, op, kind, value) with kind in {"i64","f64","str"}."""

    def __init__(self, clauses):
        leaves = [lf for cl in clauses for lf in cl]
        offs = [0]
        for cl in clauses:
            offs.append(offs[-1] + len(cl))
        self._offs = (C.c_uint32 * len(offs))(*offs)
        self._leaves = (OraLeaf * max(len(leaves), 1))()
        self._strs = []
        for i, (ci, op, kind, val) in enumerate(leaves):
            L = self._leaves[i]
            L.col, L.op = int(ci), int(op)
            if kind == "i64":
                L.kind, L.i = 0, int(val)
            elif kind == "f64":
                L.kind, L.f = 1, float(val)
            else:
                b = val.encode("utf-8")
                self._strs.append(b)
                L.kind, L.s, L.slen = 2, b, len(b)
        self.c = OraFilter(len(clauses), self._offs, self._leaves)

    @staticmethod
    def from_compiled(compiled, col_index):
        """compiled: an object with .clauses of leaves having .column/.cmp/.kind/.rhs (duck-typed)."""
        return FilterPack([[(col_index[lf.column], int(lf.cmp), lf.kind, lf.rhs) for lf in cl] for cl in compiled.clauses])


class MetaStore:
    """MetaStore on the CPU.  `columns` are duck-typed column objects exposing dtype(), numpy(),
    null_words(), string_buffers() (otters_b200.Column satisfies this), in ABI order."""

    def __init__(self, vectors, columns: Sequence, chunk_size=1024, bloom=("fpr", 0.01)):
        self.vectors = _f32(vectors)
        self.n, self.dim = self.vectors.shape
        self._keep = []
        self.cols = (OraColumn * max(len(columns), 1))()
        for i, c in enumerate(columns):
            oc = self.cols[i]
            oc.dtype = int(c.dtype())
            nw = c.null_words()
            if nw is not None:
                self._keep.append(nw)
                oc.null_words = nw.ctypes.data_as(u64p)
            if int(c.dtype()) == 4:
                offs, data = c.string_buffers()
                self._keep += [offs, data]
                oc.str_offsets = offs.ctypes.data_as(u64p)
                oc.str_bytes = data.ctypes.data_as(u8p)
            else:
                arr = c.numpy()
                self._keep.append(arr)
                oc.values = arr.ctypes.data if arr.size else None
        self.n_cols = len(columns)
        mode = 0 if bloom[0] == "fpr" else 1
        self.h = lib().oracle_meta_build(
            _fp(self.vectors) if self.n else None, self.n, self.dim, int(chunk_size), self.cols, self.n_cols, mode,
            float(bloom[1]) if mode == 0 else 0.01, int(bloom[1]) if mode == 1 else 0,
        )
        self.chunk_size = max(int(chunk_size), 1)

    def n_chunks(self) -> int:
        return int(lib().oracle_meta_n_chunks(self.h))

    def chunk_mask(self, flt: Optional[FilterPack]) -> np.ndarray:
        out = np.zeros(max(self.n_chunks(), 1), np.uint8)
        lib().oracle_meta_chunk_mask(self.h, C.byref(flt.c) if flt else None, out.ctypes.data_as(u8p))
        return out[: self.n_chunks()]

    def row_mask(self, flt: Optional[FilterPack]) -> np.ndarray:
        out = np.zeros(max(self.n, 1), np.uint8)
        lib().oracle_meta_row_mask(self.h, C.byref(flt.c) if flt else None, out.ctypes.data_as(u8p))
        return out[: self.n]

    def zonemap(self, col: int, is_float: bool):
        nc = self.n_chunks()
        nn = np.zeros(max(nc, 1), np.uint64)
        if is_float:
            mn, mx = np.zeros(max(nc, 1), np.float64), np.zeros(max(nc, 1), np.float64)
            rc = lib().oracle_meta_zonemap_f64(self.h, col, mn.ctypes.data_as(C.POINTER(C.c_double)), mx.ctypes.data_as(C.POINTER(C.c_double)), nn.ctypes.data_as(u64p))
        else:
            mn, mx = np.zeros(max(nc, 1), np.int64), np.zeros(max(nc, 1), np.int64)
            rc = lib().oracle_meta_zonemap_i64(self.h, col, mn.ctypes.data_as(C.POINTER(C.c_int64)), mx.ctypes.data_as(C.POINTER(C.c_int64)), nn.ctypes.data_as(u64p))
        assert rc == 0
        return mn[:nc], mx[:nc], nn[:nc]

    def query(self, queries, metric, take_type, k, vec_filter=None, flt: Optional[FilterPack] = None, mode=CANONICAL,
              n_threads=0, dim=None):
        keep = []
        vq = _make_query(queries, self.dim if dim is None else dim, metric, take_type, k, vec_filter, None, keep)
        cap = int(min(k, max(self.n * max(vq.nq, 1), 0)))
        idx, score, qid = np.zeros(max(cap, 1), np.uint64), np.zeros(max(cap, 1), np.float32), np.zeros(max(cap, 1), np.uint32)
        st = OraStats()
        m = lib().oracle_meta_query(
            self.h, C.byref(vq), C.byref(flt.c) if flt else None, mode, n_threads,
            idx.ctypes.data_as(u64p), _fp(score), qid.ctypes.data_as(u32p), cap, C.byref(st),
        )
        stats = {f: getattr(st, f) for f, _ in OraStats._fields_}
        return idx[:m], score[:m], qid[:m], stats

    def close(self):
        if self.h:
            lib().oracle_meta_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
