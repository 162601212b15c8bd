"""Typed metadata columns — host-side mirror of the reference's ``Column`` (src/col.rs).

Layout contract (what crosses the C ABI, include/otters_b200.h ``otters_column``): a typed value
array plus a null bitmap (bit = 1 means NULL, src/col.rs:21-28) with a sentinel in the value slot of
NULL rows (``i32::MIN`` / ``i64::MIN`` / ``NaN`` / ``""``, src/col.rs:238-326).
"""
from __future__ import annotations

import datetime as _dt
import re
from typing import Iterable, List, Optional

import numpy as np

from .types import DataType, OttersError

I32_MIN = -(2**31)
I64_MIN = -(2**63)


class ColumnError(OttersError):
    pass


_RFC3339 = re.compile(
    r"^(\d{4})-(\d{2})-(\d{2})[Tt ](\d{2}):(\d{2}):(\d{2})(\.\d+)?([Zz]|[+-]\d{2}:\d{2})$"
)
_DATE = re.compile(r"^(\d{4})-(\d{2})-(\d{2})$")
_DATETIME = re.compile(r"^(\d{4})-(\d{2})-(\d{2}) (\d{2}):(\d{2}):(\d{2})$")
_EPOCH = _dt.datetime(1970, 1, 1, tzinfo=_dt.timezone.utc)


def _millis(d: _dt.datetime) -> int:
    delta = d - _EPOCH
    return (delta.days * 86400 + delta.seconds) * 1000 + delta.microseconds // 1000


def parse_datetime_millis(s: str) -> Optional[int]:
    """RFC3339 | ``YYYY-MM-DD`` | ``YYYY-MM-DD HH:MM:SS`` -> epoch milliseconds UTC, else None.

    Mirrors ``parse_datetime`` (src/col.rs:506-529) and ``parse_datetime_literal_millis``
    (src/expr.rs:267-283), which use chrono in exactly this order.
    """
    try:
        m = _RFC3339.match(s)
        if m:
            y, mo, d, h, mi, sec = (int(m.group(i)) for i in range(1, 7))
            frac = m.group(7) or ""
            micros = int((frac[1:] + "000000")[:6]) if frac else 0
            tz = m.group(8)
            if tz in ("Z", "z"):
                off = _dt.timedelta(0)
            else:
                sign = 1 if tz[0] == "+" else -1
                off = sign * _dt.timedelta(hours=int(tz[1:3]), minutes=int(tz[4:6]))
            leap = sec == 60
            base = _dt.datetime(y, mo, d, h, mi, 59 if leap else sec, micros, tzinfo=_dt.timezone(off))
            return _millis(base) + (1000 if leap else 0)
        m = _DATE.match(s)
        if m:
            y, mo, d = (int(g) for g in m.groups())
            return _millis(_dt.datetime(y, mo, d, tzinfo=_dt.timezone.utc))
        m = _DATETIME.match(s)
        if m:
            y, mo, d, h, mi, sec = (int(g) for g in m.groups())
            return _millis(_dt.datetime(y, mo, d, h, mi, sec, tzinfo=_dt.timezone.utc))
    except ValueError:
        return None
    return None


def parse_datetime_fmt_millis(s: str, fmt: str) -> Optional[int]:
    """src/col.rs:531-545 — strptime with a caller format (datetime first, then date-only)."""
    try:
        d = _dt.datetime.strptime(s, fmt).replace(tzinfo=_dt.timezone.utc)
        return _millis(d)
    except ValueError:
        return None


_NP = {
    DataType.Int32: np.int32,
    DataType.Int64: np.int64,
    DataType.Float32: np.float32,
    DataType.Float64: np.float64,
    DataType.DateTime: np.int64,
}


class _CatSeq:
    """Lazy string sequence backed by a vocabulary and per-row codes (bulk synthetic columns)."""

    def __init__(self, vocab, codes):
        self.vocab = list(vocab)
        self.codes = np.ascontiguousarray(codes, dtype=np.int64)

    def __len__(self):
        return len(self.codes)

    def __getitem__(self, i):
        return self.vocab[int(self.codes[i])]

    def __iter__(self):
        v = self.vocab
        return (v[int(c)] for c in self.codes)


class Column:
    """A named, typed column with a null mask (src/col.rs:21-28)."""

    def __init__(self, name: str, dtype: DataType):
        self._name = name
        self._dtype = DataType(dtype)
        self._vals: list = []
        self._nulls: List[bool] = []
        self._fmt: Optional[str] = None
        self._np = None  # cached numpy view

    # ---- construction ---------------------------------------------------------------------
    def with_datetime_fmt(self, fmt: str) -> "Column":
        self._fmt = fmt
        return self

    def push(self, value) -> None:
        """Unified push (src/col.rs:358-389): ``None`` is NULL; DateTime accepts strings or millis."""
        self._np = None
        dt = self._dtype
        if value is None:
            self._nulls.append(True)
            self._vals.append(self._sentinel())
            return
        if dt in (DataType.Int32, DataType.Int64):
            if isinstance(value, (bool, float, str)) or not isinstance(value, (int, np.integer)):
                raise ColumnError(f"Type mismatch: expected {dt.name}, got incompatible type")
            self._vals.append(int(value))
        elif dt in (DataType.Float32, DataType.Float64):
            if isinstance(value, (str, bool)) or not isinstance(value, (int, float, np.integer, np.floating)):
                raise ColumnError(f"Type mismatch: expected {dt.name}, got incompatible type")
            self._vals.append(float(value))
        elif dt == DataType.String:
            if not isinstance(value, str):
                raise ColumnError(f"Type mismatch: expected {dt.name}, got incompatible type")
            self._vals.append(value)
        else:  # DateTime
            if isinstance(value, str):
                ms = parse_datetime_fmt_millis(value, self._fmt) if self._fmt else parse_datetime_millis(value)
                if ms is None:
                    if self._fmt:
                        raise ColumnError(f"Parse error: Cannot parse '{value}' with format '{self._fmt}'")
                    raise ColumnError(
                        f"Parse error: Cannot parse '{value}' as datetime. Supported formats: ISO 8601, "
                        "YYYY-MM-DD, YYYY-MM-DD HH:MM:SS"
                    )
                self._vals.append(ms)
            elif isinstance(value, (int, np.integer)) and not isinstance(value, bool):
                self._vals.append(int(value))
            else:
                raise ColumnError(f"Type mismatch: expected {dt.name}, got incompatible type")
        self._nulls.append(False)

    def from_values(self, values: Iterable) -> "Column":
        """The reference's ``Column::from(vec)`` (src/col.rs:392-401)."""
        for v in values:
            self.push(v)
        return self

    from_ = from_values

    @classmethod
    def from_numpy(cls, name: str, dtype: DataType, values, nulls=None) -> "Column":
        """Bulk constructor for large synthetic columns (no per-row Python work)."""
        c = cls(name, dtype)
        dtype = DataType(dtype)
        if dtype == DataType.String:
            c._vals = list(values)
        else:
            c._vals = np.ascontiguousarray(values, dtype=_NP[dtype])
        n = len(c._vals)
        c._nulls = np.zeros(n, dtype=bool) if nulls is None else np.ascontiguousarray(nulls, dtype=bool)
        if nulls is not None and dtype != DataType.String:
            c._vals = c._vals.copy()
            c._vals[c._nulls] = c._sentinel()
        return c

    @classmethod
    def from_categories(cls, name: str, vocab, codes, nulls=None) -> "Column":
        """String column given as vocabulary + per-row codes; NULL rows hold "" like the reference."""
        c = cls(name, DataType.String)
        vocab = list(vocab)
        codes = np.ascontiguousarray(codes, dtype=np.int64)
        n = len(codes)
        c._nulls = np.zeros(n, dtype=bool) if nulls is None else np.ascontiguousarray(nulls, dtype=bool)
        if c._nulls.any():
            if "" not in vocab:
                vocab = vocab + [""]
            codes = codes.copy()
            codes[c._nulls] = vocab.index("")
        c._vals = _CatSeq(vocab, codes)
        return c

    def _sentinel(self):
        dt = self._dtype
        if dt == DataType.Int32:
            return I32_MIN
        if dt in (DataType.Int64, DataType.DateTime):
            return I64_MIN
        if dt in (DataType.Float32, DataType.Float64):
            return float("nan")
        return ""

    # ---- accessors ------------------------------------------------------------------------
    def name(self) -> str:
        return self._name

    def dtype(self) -> DataType:
        return self._dtype

    def len(self) -> int:
        return len(self._vals)

    __len__ = len

    def is_empty(self) -> bool:
        return self.len() == 0

    def null_mask(self) -> np.ndarray:
        """Boolean array, True = NULL (src/col.rs:26; tests/column_tests.rs:33-37)."""
        return np.asarray(self._nulls, dtype=bool)

    def values(self):
        if self._dtype == DataType.String:
            return list(self._vals)
        return self.numpy()

    def numpy(self) -> np.ndarray:
        if self._dtype == DataType.String:
            raise ColumnError("String columns have no numeric view")
        if self._np is None or len(self._np) != len(self._vals):
            self._np = np.ascontiguousarray(self._vals, dtype=_NP[self._dtype])
        return self._np

    def i32_values(self):
        return self.numpy() if self._dtype == DataType.Int32 else None

    def i64_values(self):
        return self.numpy() if self._dtype == DataType.Int64 else None

    def f32_values(self):
        return self.numpy() if self._dtype == DataType.Float32 else None

    def f64_values(self):
        return self.numpy() if self._dtype == DataType.Float64 else None

    def datetime_values(self):
        return self.numpy() if self._dtype == DataType.DateTime else None

    def string_values(self):
        return list(self._vals) if self._dtype == DataType.String else None

    def get(self, i: int):
        """Value at row ``i`` or ``None`` when NULL."""
        if self._nulls[i]:
            return None
        v = self._vals[i]
        return v.item() if isinstance(v, np.generic) else v

    def gather(self, indices) -> "Column":
        """Rows at ``indices`` with NULLs preserved (MetaQueryPlan::collect, src/meta.rs:723-821)."""
        out = Column(self._name, self._dtype)
        nulls = self.null_mask()
        idx = np.asarray(indices, dtype=np.int64)
        if isinstance(self._vals, _CatSeq):
            out._vals = _CatSeq(self._vals.vocab, self._vals.codes[idx] if len(idx) else np.zeros(0, np.int64))
        elif self._dtype == DataType.String:
            out._vals = [self._vals[i] for i in idx]
        else:
            out._vals = self.numpy()[idx] if len(idx) else np.zeros(0, dtype=_NP[self._dtype])
        out._nulls = nulls[idx] if len(idx) else np.zeros(0, dtype=bool)
        return out

    # ---- C ABI layout ---------------------------------------------------------------------
    def null_words(self) -> Optional[np.ndarray]:
        """Lsb0 u64 words, bit = 1 NULL; None when the column has no NULLs."""
        nulls = self.null_mask()
        if not nulls.any():
            return None
        n = len(nulls)
        padded = np.zeros((n + 63) // 64 * 64, dtype=np.uint8)
        padded[:n] = nulls
        return np.packbits(padded, bitorder="little").view(np.uint64).copy()

    def string_buffers(self):
        """(offsets u64[n+1], bytes u8[]) for String columns."""
        if isinstance(self._vals, _CatSeq):
            enc_v = [s.encode("utf-8") for s in self._vals.vocab]
            vlen = np.array([len(b) for b in enc_v], dtype=np.uint64)
            width = max(int(vlen.max()) if len(vlen) else 0, 1)
            mat = np.zeros((len(enc_v), width), dtype=np.uint8)
            for i, b in enumerate(enc_v):
                mat[i, : len(b)] = np.frombuffer(b, dtype=np.uint8)
            codes = self._vals.codes
            lens = vlen[codes]
            offsets = np.zeros(len(codes) + 1, dtype=np.uint64)
            np.cumsum(lens, out=offsets[1:])
            if len(vlen) and int(vlen.min()) == width:
                data = mat[codes].reshape(-1)
            else:
                data = mat[codes][np.arange(width)[None, :] < lens[:, None].astype(np.int64)]
            data = np.ascontiguousarray(data, dtype=np.uint8)
            if data.size == 0:
                data = np.zeros(1, dtype=np.uint8)
            return offsets, data
        enc = [s.encode("utf-8") for s in self._vals]
        lens = np.fromiter((len(b) for b in enc), dtype=np.uint64, count=len(enc))
        offsets = np.zeros(len(enc) + 1, dtype=np.uint64)
        np.cumsum(lens, out=offsets[1:])
        data = np.frombuffer(b"".join(enc), dtype=np.uint8).copy() if enc else np.zeros(0, dtype=np.uint8)
        if data.size == 0:
            data = np.zeros(1, dtype=np.uint8)
        return offsets, data

    def __repr__(self):
        return f"Column({self._name!r}, {self._dtype.name}, len={self.len()})"
