"""Row ordering for better chunk pruning (the reference's roadmap: "Ability to reorder metadata for better pruning
(Something like Z-ordering)", README.md:212; the note under its demo, README.md:154: "sorting on common filter columns can
significantly improve pruning effectiveness").

Zonemaps (per-chunk min / max, src/meta_compute.rs:32-132) and per-chunk Bloom filters prune a chunk only when the values
inside it are clustered.  ``compute_row_order`` returns the permutation ``perm`` (store position -> input row) that
clusters the rows on the given filter columns; ``MetaStoreBuilder.with_row_order`` applies it to the vectors and to every
column before the build and the query plans map result rows back through it, so callers keep seeing their own row ids.

Two methods:
  * ``"sort"``   — stable lexicographic sort on the columns in the order given (NULLs last): the first column clusters
                   perfectly, the others only inside runs of equal leading values.
  * ``"zorder"`` — every column is turned into a 16-bit rank bucket (equal values share a bucket, NULLs take the last one)
                   and the buckets are bit-interleaved into a Morton code, most significant bits first, first column
                   first: every column gets narrow per-chunk ranges, none perfectly (up to 4 columns).
Both are stable: rows with equal keys keep their input order.  Host-side numpy — this is build-time bookkeeping like
``Expr::compile``, not part of the per-query path."""
from __future__ import annotations

from typing import Dict, List, Sequence

import numpy as np

from .column import Column, _CatSeq
from .types import DataType, OttersError

METHODS = ("sort", "zorder")


def _sort_key(col: Column) -> np.ndarray:
    """An int64 / float64 / uint32 array whose order is the column's value order (strings: lexicographic by code point)."""
    dt = col.dtype()
    if dt == DataType.String and isinstance(col._vals, _CatSeq):  # vocabulary + codes: rank the vocabulary, not the rows
        vocab = np.asarray(col._vals.vocab, dtype=object).astype(str)
        _, inv = np.unique(vocab, return_inverse=True)
        return inv.astype(np.int64)[col._vals.codes]
    if dt == DataType.String:
        vals = np.asarray(col.string_values(), dtype=object)
        if len(vals) == 0:
            return np.zeros(0, np.int64)
        _, inv = np.unique(vals.astype(str), return_inverse=True)
        return inv.astype(np.int64)
    return col.numpy()


def rank_buckets(col: Column, bits: int = 16) -> np.ndarray:
    """Rank bucket of every row in [0, 2^bits): rows with equal values share a bucket, NULLs (and NaN) take the last one."""
    n = col.len()
    top = (1 << bits) - 1
    out = np.full(n, top, np.uint64)
    key = _sort_key(col)
    live = ~col.null_mask()
    if key.dtype.kind == "f":
        live &= ~np.isnan(key)
    m = int(live.sum())
    if m == 0:
        return out
    k = key[live]
    srt = np.sort(k, kind="stable")
    first = np.searchsorted(srt, k, side="left").astype(np.uint64)  # rank of the first row holding this value
    out[live] = np.minimum(first * np.uint64(1 << bits) // np.uint64(m), np.uint64(top - 1 if m < n else top))
    return out


def morton_codes(buckets: Sequence[np.ndarray], bits: int = 16) -> np.ndarray:
    """Bit-interleaves up to four bucket arrays: bit b of column c lands at position b * ncols + (ncols - 1 - c)."""
    nc = len(buckets)
    if not 1 <= nc <= 4:
        raise OttersError("z-ordering takes one to four columns")
    code = np.zeros(len(buckets[0]), np.uint64)
    for c, bk in enumerate(buckets):
        bk = bk.astype(np.uint64)
        for b in range(bits):
            code |= ((bk >> np.uint64(b)) & np.uint64(1)) << np.uint64(b * nc + (nc - 1 - c))
    return code


def compute_row_order(columns: Dict[str, Column], by: Sequence[str], method: str = "sort") -> np.ndarray:
    """perm[i] = input row stored at position i."""
    by = list(by)
    if not by:
        raise OttersError("with_row_order needs at least one column")
    if method not in METHODS:
        raise OttersError(f"unknown row order method '{method}' (expected one of {', '.join(METHODS)})")
    for name in by:
        if name not in columns:
            raise OttersError(f"unknown column '{name}' not present in schema")
    n = columns[by[0]].len()
    if method == "zorder":
        code = morton_codes([rank_buckets(columns[name]) for name in by])
        return np.argsort(code, kind="stable").astype(np.uint64)
    keys: List[np.ndarray] = []
    for name in by:  # np.lexsort sorts by the LAST key first
        col = columns[name]
        key = _sort_key(col)
        null = col.null_mask().copy()
        if key.dtype.kind == "f":
            null |= np.isnan(key)
            key = np.where(null, 0.0, key)
        keys.append((null, key))
    flat = []
    for null, key in reversed(keys):
        flat += [key, null]  # within a column: NULL flag is the more significant key
    return np.lexsort(tuple(flat)).astype(np.uint64) if n else np.zeros(0, np.uint64)


def chunk_ranges_overlapping(values: np.ndarray, chunk_size: int, lo, hi) -> int:
    """How many chunks a range predicate lo <= v <= hi cannot prune from min / max alone (a planning aid and what the tests
    use to show the effect of an order)."""
    n = len(values)
    cnt = 0
    for s in range(0, n, chunk_size):
        c = values[s : s + chunk_size]
        c = c[~np.isnan(c)] if c.dtype.kind == "f" else c
        if len(c) and c.min() <= hi and c.max() >= lo:
            cnt += 1
    return cnt
