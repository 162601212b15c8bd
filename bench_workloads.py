"""Synthetic workloads of bench.py and of the full-size tests: NumPy only.

Nothing here imports otters_b200 or the oracle: both arms of bench.py (the CUDA library and the CPU reference arm) and
tests/test_gpu_fullsize.py build their stores from the SAME specs, each with its own column / filter types.

  * vectors:  counter-based generator x = (splitmix64(seed ^ (row*dim+col)) >> 40) * 2^-23 - 1  (U(-1,1), like the
              reference's examples/demo.rs:4-7), identical on host and device; config C5 additionally PLANTS ~5,000
              near-duplicates normalize(q + sigma*noise) of the query with cosine in (0.8, 0.99), since random 1536-d
              vectors never reach 0.8.
  * metadata: pure functions of the absolute row id, clustered by chunk like examples/demo.rs:29-77, ~1 % NULLs per column.
"""
from __future__ import annotations

import time
from typing import Dict, List, Optional, Tuple

import numpy as np

T0_MS = 1_700_000_000_000  # 2023-11-14T22:13:20Z
DATA_SEED = 0x07735
QUERY_SEED = 0xBEEF

# codes of include/otters_b200.h (reference declaration order)
DT_INT32, DT_INT64, DT_FLOAT32, DT_FLOAT64, DT_STRING, DT_DATETIME = range(6)
OP_EQ, OP_NEQ, OP_LT, OP_LTE, OP_GT, OP_GTE = range(6)
CMP_LT, CMP_GT, CMP_LTE, CMP_GTE, CMP_EQ = range(5)
METRIC_CODE = {"Cosine": 0, "Euclidean": 1, "DotProduct": 2}

WORKLOADS: Dict[str, dict] = {
    "target": dict(rows=10_000_000, dim=768, chunk=1024, metric="Cosine", k=100, meta="pit",
                   desc="MetaStore 10Mx768 fp32 Cosine top-100, chunk 1024, meta_filter price.gt & item.eq & ts.gte"),
    "c1": dict(rows=100_000, dim=128, chunk=0, metric="Cosine", k=10, meta=None, desc="VecStore 100kx128 fp32 Cosine top-10"),
    "c2": dict(rows=1_000_000, dim=768, chunk=0, metric="DotProduct", k=100, meta=None, nq=1024,
               desc="VecStore 1Mx768 fp32 Dot, batch of 1024 queries, top-100 (one merged list; tcgen05 tf32 selection + exact re-scoring)"),
    "c3": dict(rows=10_000_000, dim=128, chunk=1024, metric="Cosine", k=100, meta="pit",
               desc="MetaStore 10Mx128 Cosine top-100, chunk 1024, meta_filter price.gt & item.eq & ts.gte"),
    "c3u": dict(rows=10_000_000, dim=128, chunk=0, metric="Cosine", k=100, meta=None,
                desc="DIAGNOSTIC (not a BASELINE config): VecStore 10Mx128 Cosine top-100, C3's rows without its filter"),
    "c4": dict(rows=10_000_000, dim=768, chunk=0, metric="Euclidean", k=100, meta=None, desc="VecStore 10Mx768 fp32 L2 top-100"),
    "c5": dict(rows=5_000_000, dim=1536, chunk=1024, metric="Cosine", k=1000, meta="mixed", vec_filter=(0.8, CMP_GT), planted=5000,
               desc="MetaStore 5Mx1536 Cosine vec_filter(0.8,Gt) take(1000), chunk 1024, meta_filter qty.gte & (price.lt | item.eq) "
                    "& brand.neq over Int32/Float64/String columns, ~5000 planted near-duplicates of the query"),
}


# ---- vectors ---------------------------------------------------------------------------------------------
def synth_fill_np(row0: int, n_rows: int, dim: int, seed: int) -> np.ndarray:
    """NumPy form of the counter-based generator (bit-identical to the device fill and to oracle_synth_fill)."""
    idx = (np.arange(row0, row0 + n_rows, dtype=np.uint64)[:, None] * np.uint64(dim) + np.arange(dim, dtype=np.uint64)[None, :])
    x = (np.uint64(seed) ^ idx) + np.uint64(0x9E3779B97F4A7C15)
    x = (x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
    x = (x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
    x = x ^ (x >> np.uint64(31))
    return np.ascontiguousarray(((x >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / 8388608.0) - np.float32(1.0)))


def u01(row: np.ndarray, salt: int) -> np.ndarray:
    """Counter-based uniform in [0,1) per absolute row id."""
    x = (row.astype(np.uint64) + np.uint64(salt)) * np.uint64(0x9E3779B97F4A7C15)
    x ^= x >> np.uint64(30)
    x *= np.uint64(0xBF58476D1CE4E5B9)
    x ^= x >> np.uint64(27)
    x *= np.uint64(0x94D049BB133111EB)
    x ^= x >> np.uint64(31)
    return (x >> np.uint64(11)).astype(np.float64) * (1.0 / (1 << 53))


# ---- columns ---------------------------------------------------------------------------------------------
class SpecColumn:
    """A metadata column as plain NumPy: typed values (or vocabulary + codes for strings) + a NULL mask.  Satisfies the
    duck type oracle.MetaStore reads (dtype / numpy / null_words / string_buffers); `to_ob` makes the library's Column."""

    def __init__(self, name: str, dtype: int, values=None, nulls=None, vocab: Optional[List[str]] = None, codes=None):
        self._name, self._dtype = name, dtype
        self._nulls = None if nulls is None else np.ascontiguousarray(nulls, dtype=bool)
        self.vocab, self.codes = vocab, None
        if dtype == DT_STRING:
            vocab = list(vocab)
            codes = np.ascontiguousarray(codes, dtype=np.int64)
            if self._nulls is not None and self._nulls.any():  # NULL rows hold "" like the reference (src/col.rs:238-326)
                if "" not in vocab:
                    vocab = vocab + [""]
                codes = codes.copy()
                codes[self._nulls] = vocab.index("")
            self.vocab, self.codes = vocab, codes
            self._vals = None
        else:
            np_t = {DT_INT32: np.int32, DT_INT64: np.int64, DT_FLOAT32: np.float32, DT_FLOAT64: np.float64, DT_DATETIME: np.int64}[dtype]
            v = np.ascontiguousarray(values, dtype=np_t).copy()
            if self._nulls is not None and self._nulls.any():  # sentinel in the value slot of NULL rows
                v[self._nulls] = {DT_INT32: np.iinfo(np.int32).min, DT_INT64: np.iinfo(np.int64).min, DT_DATETIME: np.iinfo(np.int64).min,
                                  DT_FLOAT32: np.nan, DT_FLOAT64: np.nan}[dtype]
            self._vals = v

    def name(self) -> str:
        return self._name

    def dtype(self) -> int:
        return self._dtype

    def __len__(self) -> int:
        return len(self.codes) if self._dtype == DT_STRING else len(self._vals)

    def numpy(self) -> np.ndarray:
        return self._vals

    def null_mask(self) -> np.ndarray:
        return self._nulls if self._nulls is not None else np.zeros(len(self), bool)

    def null_words(self):
        nulls = self.null_mask()
        if not nulls.any():
            return None
        n = len(nulls)
        padded = np.zeros((n + 63) // 64 * 64, dtype=np.uint8)
        padded[:n] = nulls
        return np.packbits(padded, bitorder="little").view(np.uint64).copy()

    def string_buffers(self):
        enc = [s.encode("utf-8") for s in self.vocab]
        vlen = np.array([len(b) for b in enc], dtype=np.uint64)
        width = max(int(vlen.max()) if len(vlen) else 0, 1)
        mat = np.zeros((len(enc), width), dtype=np.uint8)
        for i, b in enumerate(enc):
            mat[i, : len(b)] = np.frombuffer(b, dtype=np.uint8)
        lens = vlen[self.codes]
        offsets = np.zeros(len(self.codes) + 1, dtype=np.uint64)
        np.cumsum(lens, out=offsets[1:])
        data = mat[self.codes][np.arange(width)[None, :] < lens[:, None].astype(np.int64)]
        data = np.ascontiguousarray(data, dtype=np.uint8)
        return offsets, (data if data.size else np.zeros(1, np.uint8))

    def to_ob(self, ob):
        if self._dtype == DT_STRING:
            return ob.Column.from_categories(self._name, self.vocab, self.codes, self._nulls)
        return ob.Column.from_numpy(self._name, ob.DataType(self._dtype), self._vals, self._nulls)

    def passes(self, op: int, value) -> np.ndarray:
        """Row mask of one leaf under the reference's semantics: NULL fails every leaf, NaN satisfies only Neq."""
        if self._dtype == DT_STRING:
            code = self.vocab.index(value) if value in self.vocab else -1
            eq = self.codes == code
            m = eq if op == OP_EQ else ~eq
        else:
            v = self._vals
            t = np.asarray(value).astype(v.dtype)
            with np.errstate(invalid="ignore"):
                m = {OP_EQ: v == t, OP_NEQ: v != t, OP_LT: v < t, OP_LTE: v <= t, OP_GT: v > t, OP_GTE: v >= t}[op]
        return m & ~self.null_mask()


def _columns_pit(row: np.ndarray, chunk: int) -> List[SpecColumn]:
    """price: Float64, item: String, ts: DateTime (target / C3)."""
    c = row // max(chunk, 1)
    # price: chunk groups of 4; 4 of every 5 groups are "expensive" (90..115), the fifth cheap (10..35)
    expensive = ((c // 4) % 5) != 0
    price = np.where(expensive, 90.0, 10.0) + 25.0 * u01(row, 1)
    # ts: monotone in the row id (1 s per row) with +-30 s jitter
    ts = T0_MS + row * 1000 + ((u01(row, 2) - 0.5) * 60_000).astype(np.int64)
    # item: 1000 categories; each chunk has a dominant one (85 % of its rows): item_0000 in 3 of every 4 groups of 8
    dominant = np.where(((c // 8) % 4) != 3, 0, 1 + (c // 8) % 7)
    other = (u01(row, 3) * 1000).astype(np.int64)
    code = np.where(u01(row, 4) < 0.85, dominant, other)
    vocab = [f"item_{i:04d}" for i in range(1000)]
    return [SpecColumn("price", DT_FLOAT64, price, u01(row, 5) < 0.01),
            SpecColumn("item", DT_STRING, None, u01(row, 6) < 0.01, vocab, code),
            SpecColumn("ts", DT_DATETIME, ts, u01(row, 7) < 0.01)]


def _columns_mixed(row: np.ndarray, chunk: int) -> List[SpecColumn]:
    """qty: Int32, price: Float64, item: String, brand: String (C5: mixed Int32/Float64/String predicates)."""
    c = row // max(chunk, 1)
    # qty: chunk groups of 4; one group in four holds small quantities (0..9), the others 10..59: qty >= 10 prunes by zonemap
    small = ((c // 4) % 4) == 0
    qty = np.where(small, 0, 10) + (u01(row, 11) * np.where(small, 10, 50)).astype(np.int64)
    # price: chunk groups of 2 alternate between cheap (5..45) and expensive (60..140) bands
    cheap = ((c // 2) % 2) == 0
    price = np.where(cheap, 5.0 + 40.0 * u01(row, 12), 60.0 + 80.0 * u01(row, 12))
    # item: 500 categories; groups of 16 chunks have a dominant one (80 % of their rows); widget_000 dominates every third group
    g = c // 16
    dominant = np.where(g % 3 == 0, 0, 1 + g % 11)
    code = np.where(u01(row, 13) < 0.8, dominant, (u01(row, 14) * 500).astype(np.int64))
    # brand: 20 brands, uniform per row (no pruning power; != drops ~5 % of the rows)
    brand = (u01(row, 15) * 20).astype(np.int64)
    return [SpecColumn("qty", DT_INT32, qty.astype(np.int32), u01(row, 16) < 0.01),
            SpecColumn("price", DT_FLOAT64, price, u01(row, 17) < 0.01),
            SpecColumn("item", DT_STRING, None, u01(row, 18) < 0.01, [f"widget_{i:03d}" for i in range(500)], code),
            SpecColumn("brand", DT_STRING, None, u01(row, 19) < 0.01, [f"brand_{i:02d}" for i in range(20)], brand)]


def round_bf16_np(x: np.ndarray) -> np.ndarray:
    """f32(bf16_rn(x)): the values a store built with --vector-format bf16 holds (round to nearest even)."""
    a = np.ascontiguousarray(x, dtype=np.float32)
    b = a.view(np.uint32).astype(np.uint64)
    return ((b + 0x7FFF + ((b >> 16) & 1)) & 0xFFFF0000).astype(np.uint32).view(np.float32).reshape(a.shape)


class Workload:
    def __init__(self, name: str, rows_override: int = 0, vector_format: str = "f32"):
        w = dict(WORKLOADS[name])
        self.name = name
        self.vector_format = vector_format  # "bf16": the store keeps bf16 rows; both arms score the ROUNDED rows
        self.rows = int(rows_override or w["rows"])
        self.dim, self.chunk, self.k = w["dim"], w["chunk"], w["k"]
        self.metric = w["metric"]
        self.metric_code = METRIC_CODE[self.metric]
        self.take_max = self.metric != "Euclidean"  # take() infers Min for Euclidean (src/vec.rs:92-101)
        self.meta = w["meta"]
        self.nq = w.get("nq", 1)
        self.desc = w["desc"]
        self.vec_filter = w.get("vec_filter")
        self.n_planted = min(int(w.get("planted", 0)), max(self.rows // 100, 1)) if w.get("planted") else 0
        self.block = self.chunk if self.chunk else 1024  # row block of the block-cyclic sharding

    # -- metadata ------------------------------------------------------------------------------------------
    def columns(self, rows) -> List[SpecColumn]:
        row = np.asarray(rows, dtype=np.int64)
        return _columns_pit(row, self.chunk) if self.meta == "pit" else _columns_mixed(row, self.chunk)

    def cut_ms(self) -> int:
        return T0_MS + int(0.10 * self.rows) * 1000  # ts.gte keeps the last ~90 % of the rows

    def clauses(self) -> List[List[Tuple[str, int, str, object]]]:
        """The compiled filter as CNF clauses of (column, op, literal kind, literal) — what Expr::compile produces."""
        if self.meta == "pit":
            return [[("price", OP_GT, "f64", 50.0)], [("item", OP_EQ, "str", "item_0000")], [("ts", OP_GTE, "i64", self.cut_ms())]]
        return [[("qty", OP_GTE, "i64", 10)], [("price", OP_LT, "f64", 50.0), ("item", OP_EQ, "str", "widget_000")],
                [("brand", OP_NEQ, "str", "brand_07")]]

    def expr(self, ob):
        """The same filter through the public expression API of the library's host mirror."""
        if self.meta == "pit":
            cut = time.strftime("%Y-%m-%d %H:%M:%S", time.gmtime(self.cut_ms() / 1000))
            return ob.col("price").gt(50.0) & ob.col("item").eq("item_0000") & ob.col("ts").gte(cut)
        return ob.col("qty").gte(10) & (ob.col("price").lt(50.0) | ob.col("item").eq("widget_000")) & ob.col("brand").neq("brand_07")

    def filter_desc(self) -> Optional[str]:
        if not self.meta:
            return None
        if self.meta == "pit":
            cut = time.strftime("%Y-%m-%d %H:%M:%S", time.gmtime(self.cut_ms() / 1000))
            return f"price.gt(50.0) & item.eq('item_0000') & ts.gte('{cut}')"
        return "qty.gte(10) & (price.lt(50.0) | item.eq('widget_000')) & brand.neq('brand_07')"

    def row_mask(self, cols: List[SpecColumn]) -> np.ndarray:
        """Rows of `cols` passing the filter (NumPy restatement of the CNF semantics; used for sample statistics only)."""
        by_name = {c.name(): c for c in cols}
        keep = np.ones(len(cols[0]), bool)
        for clause in self.clauses():
            any_ = np.zeros(len(cols[0]), bool)
            for name, op, _, val in clause:
                any_ |= by_name[name].passes(op, val)
            keep &= any_
        return keep

    # -- queries and planted rows ----------------------------------------------------------------------------
    def n_query_variants(self) -> int:
        if self.n_planted:
            return 1  # the planted rows are near-duplicates of THE query
        return 16 if self.nq == 1 else 2

    def queries(self) -> np.ndarray:
        """[variants][nq][dim]"""
        v = self.n_query_variants()
        return synth_fill_np(0, v * self.nq, self.dim, QUERY_SEED).reshape(v, self.nq, self.dim)

    def planted(self) -> Tuple[np.ndarray, np.ndarray]:
        """(global row ids ascending, vectors): normalize(q^ + tan(theta) * n^) with cos(theta) uniform in (0.8, 0.99) and
        n^ unit noise orthogonal to the query, so the cosine of every planted row is known by construction."""
        if not self.n_planted:
            return np.zeros(0, np.int64), np.zeros((0, self.dim), np.float32)
        rng = np.random.default_rng(0xC5)
        rows = np.sort(rng.choice(self.rows, self.n_planted, replace=False)).astype(np.int64)
        q = self.queries()[0, 0].astype(np.float64)
        qh = q / np.linalg.norm(q)
        noise = rng.standard_normal((self.n_planted, self.dim))
        noise -= (noise @ qh)[:, None] * qh[None, :]
        noise /= np.linalg.norm(noise, axis=1)[:, None]
        cos = rng.uniform(0.8, 0.99, self.n_planted)
        v = qh[None, :] + np.tan(np.arccos(cos))[:, None] * noise
        v *= (0.5 + rng.random(self.n_planted))[:, None] / np.linalg.norm(v, axis=1)[:, None]  # norms in [0.5, 1.5): cosine must not care
        return rows, np.ascontiguousarray(v, dtype=np.float32)

    # -- the `config` object of the JSON line: identical in both arms -------------------------------------------
    def config(self) -> dict:
        return {"workload": f"{self.name}: {self.desc}", "rows": self.rows, "dim": self.dim, "k": self.k, "nq": self.nq,
                "chunk_size": self.chunk, "metric": self.metric, "filter": self.filter_desc(),
                "vec_filter": ([self.vec_filter[0], ["Lt", "Gt", "Lte", "Gte", "Eq"][self.vec_filter[1]]] if self.vec_filter else None),
                "planted_rows": self.n_planted, "vector_format": self.vector_format}

    def stored(self, vectors: np.ndarray) -> np.ndarray:
        """The rows as the store holds them (rounded to bf16 for a bf16 store)."""
        return round_bf16_np(vectors) if self.vector_format == "bf16" else vectors


# ---- block-cyclic sharding (same formulas as otters_shard_map) ---------------------------------------------
def cyclic_local_rows(n_rows: int, block_rows: int, world: int, rank: int) -> int:
    block_rows = max(int(block_rows), 1)
    n_blocks = (n_rows + block_rows - 1) // block_rows
    mine = (n_blocks - rank + world - 1) // world if n_blocks > rank else 0
    if mine == 0:
        return 0
    last_block = (mine - 1) * world + rank
    tail = n_rows - last_block * block_rows
    return (mine - 1) * block_rows + min(block_rows, tail)


def cyclic_global_rows(n_rows: int, block_rows: int, world: int, rank: int) -> np.ndarray:
    n_local = cyclic_local_rows(n_rows, block_rows, world, rank)
    local = np.arange(n_local, dtype=np.int64)
    return (local // block_rows * world + rank) * block_rows + local % block_rows


def global_to_local(global_rows: np.ndarray, block_rows: int, world: int, rank: int):
    """(mask of the rows this rank holds, their local row ids) under the block-cyclic deal."""
    g = np.asarray(global_rows, dtype=np.int64)
    b = g // block_rows
    mine = (b % world) == rank
    return mine, (b[mine] // world) * block_rows + g[mine] % block_rows


def sample_blocks(n_rows: int, chunk: int, budget_rows: int, n_blocks: int) -> List[Tuple[int, int]]:
    """n_blocks chunk-aligned row ranges [r0, r1) of about budget_rows rows in total, centred at (i + 0.5) / n_blocks of the store."""
    chunk = max(chunk, 1)
    if n_rows <= budget_rows:
        return [(0, n_rows)]
    block_rows = max(budget_rows // n_blocks // chunk, 1) * chunk
    out = set()
    for i in range(n_blocks):
        centre = (i + 0.5) / n_blocks * n_rows
        r0 = int(max(centre - block_rows / 2, 0)) // chunk * chunk
        r0 = min(r0, max((n_rows - block_rows) // chunk * chunk, 0))
        out.add((r0, min(r0 + block_rows, n_rows)))
    return sorted(out)


def stratified_sample_blocks(n_rows: int, chunk: int, dim: int, nq: int) -> List[Tuple[int, int]]:
    """Row ranges of the bounded CPU sample: ten or more chunk-aligned blocks spread evenly over the store (so range
    predicates such as ts >= cut and the chunk-group clustering of the metadata are represented in proportion), sized for
    a few GB of host memory and tens of seconds of CPU work."""
    if nq > 1:
        return sample_blocks(n_rows, chunk, max(4_194_304 // nq, 1), 1)  # a batch re-scores every sampled row nq times
    budget_rows = 2_129_920 if dim <= 256 else (983_040 if dim <= 768 else 491_520)
    return sample_blocks(n_rows, chunk, budget_rows, 10 if dim > 256 else 13)
