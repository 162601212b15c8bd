//! otters-gpu — the device side of otters' exact-search hot path, behind the reference's own types.
//!
//! The reference (AtharvBhat/otters) has no FFI seam: the drop-in boundary is its public Rust API.  This crate is what the
//! bodies of `VecQueryPlan::collect` (src/vec.rs:206-311), `MetaStoreBuilder::build` (src/meta.rs:151-305) and
//! `MetaQueryPlan::collect` (src/meta.rs:632-829) call once the plan has been validated and its defaults resolved; every
//! builder method, every error string, `Expr::compile`, `Column`, `Metric` / `Cmp` / `TakeType` stay untouched Rust in the
//! otters crate.  INTEGRATION.md shows the three call sites.
//!
//! SOURCE ONLY: the authoring image has no Rust toolchain, so this crate has never been compiled.  What can be checked
//! without one is checked: `otters-sys` is generated from include/otters_b200.h and tests/test_rust_sys.py verifies every
//! repr(C) layout against gcc and that this file only uses declared items.
use otters::col::Column;
use otters::expr::{CmpOp, ColumnFilter, CompiledFilter, NumericLiteral};
use otters::type_utils::DataType;
use otters_sys as sys;
use std::collections::HashMap;
use std::ffi::{CStr, CString};
use std::os::raw::c_void;
use std::ptr;
use std::time::Duration;

fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::otters_last_error()).to_string_lossy().into_owned() }
}

fn check(rc: i32) -> Result<(), String> {
    if rc == sys::OTTERS_OK { Ok(()) } else { Err(last_error()) }
}

/// One CUDA device + stream + scratch (`otters_ctx`).  Stores borrow it; queries on one context are serialised, or
/// pipelined two deep through `submit` / `wait`.
pub struct DeviceContext {
    raw: *mut sys::otters_ctx,
}

impl DeviceContext {
    pub fn new(device: i32) -> Result<Self, String> {
        let mut raw = ptr::null_mut();
        check(unsafe { sys::otters_ctx_create(device, ptr::null_mut(), &mut raw) })?;
        Ok(Self { raw })
    }
    pub fn synchronize(&self) -> Result<(), String> {
        check(unsafe { sys::otters_ctx_synchronize(self.raw) })
    }
}

impl Drop for DeviceContext {
    fn drop(&mut self) {
        unsafe { sys::otters_ctx_destroy(self.raw) };
    }
}

/// The resolved fields of a `VecQueryPlan` / `MetaQueryPlan` (after `validate()` and the defaults of src/vec.rs:213-214,
/// src/meta.rs:638-644): exactly what crosses the boundary.
pub struct ResolvedQuery<'a> {
    pub queries: &'a [Vec<f32>],
    pub metric: i32,    // Metric as declared: Cosine 0, Euclidean 1, DotProduct 2
    pub take_type: i32, // TakeType: Min 0, Max 1
    pub k: usize,
    pub filter: Option<(f32, i32)>, // (.filter / .vec_filter threshold, Cmp: Lt 0, Gt 1, Lte 2, Gte 3, Eq 4)
    pub row_mask: Option<(&'a [usize], usize)>, // BitVec<usize, Lsb0> raw words + length in bits (VecStore only)
}

struct FlatQuery {
    flat: Vec<f32>,
    q: sys::otters_vec_query,
}

fn flatten(r: &ResolvedQuery<'_>) -> FlatQuery {
    let dim = r.queries.first().map_or(0, |q| q.len());
    // mixed dimensions are reported by the plan's validate() (VecStore) or swallowed per chunk (MetaStore): a ragged
    // batch travels as dim 0, which the library treats exactly like the reference's per-chunk error
    let ragged = r.queries.iter().any(|q| q.len() != dim);
    let flat: Vec<f32> = if ragged { Vec::new() } else { r.queries.iter().flatten().copied().collect() };
    let q = sys::otters_vec_query {
        queries: if flat.is_empty() { ptr::null() } else { flat.as_ptr() },
        nq: r.queries.len() as u32,
        dim: if ragged { 0 } else { dim as u32 },
        metric: r.metric,
        take_type: r.take_type,
        k: r.k as u64,
        has_filter: r.filter.is_some() as i32,
        thr: r.filter.map_or(0.0, |f| f.0),
        cmp: r.filter.map_or(0, |f| f.1),
        row_mask_words: r.row_mask.map_or(ptr::null(), |m| m.0.as_ptr() as *const u64),
        row_mask_bits: r.row_mask.map_or(0, |m| m.1 as u64),
    };
    FlatQuery { flat, q }
}

// =====================================================================================================================
// VecStore
// =====================================================================================================================
/// How a store keeps its rows in HBM (OTTERS_VECTORS_FMT_* of include/otters_b200.h).
#[derive(Clone, Copy, Debug, PartialEq, Eq)]
#[repr(i32)]
pub enum VectorFormat {
    F32 = 0,
    Bf16 = 1,
}

pub struct DeviceVecStore<'c> {
    ctx: &'c DeviceContext,
    vs: *mut sys::otters_vecstore,
    dim: usize,
}

impl<'c> DeviceVecStore<'c> {
    /// VecStore::new (src/vec.rs:346-355)
    pub fn new(ctx: &'c DeviceContext, dim: usize) -> Result<Self, String> {
        Self::with_format(ctx, dim, VectorFormat::F32)
    }

    /// A store whose rows are kept as `format` (see VectorFormat).
    pub fn with_format(ctx: &'c DeviceContext, dim: usize, format: VectorFormat) -> Result<Self, String> {
        let mut vs = ptr::null_mut();
        check(unsafe { sys::otters_vecstore_create_fmt(ctx.raw, dim as u32, format as i32, &mut vs) })?;
        Ok(Self { ctx, vs, dim })
    }

    /// VecStore::add_vectors (src/vec.rs:374-376): the rows are flattened once and copied to HBM, where the inverse norms
    /// are computed exactly as add_vector does (src/vec.rs:357-371).
    pub fn add_vectors(&mut self, rows: &[Vec<f32>]) -> Result<(), String> {
        let mut flat = Vec::with_capacity(rows.len() * self.dim);
        for r in rows {
            if r.len() != self.dim {
                return Err(format!("Input vector length {} does not match expected dimension {}", r.len(), self.dim));
            }
            flat.extend_from_slice(r);
        }
        check(unsafe { sys::otters_vecstore_add(self.vs, flat.as_ptr(), rows.len() as u64) })
    }

    pub fn len(&self) -> usize {
        unsafe { sys::otters_vecstore_len(self.vs) as usize }
    }

    /// The body of VecQueryPlan::collect after validate(): one merged list for a batch (src/vec.rs:217-219).
    pub fn collect(&self, r: &ResolvedQuery<'_>) -> Result<Vec<(usize, f32)>, String> {
        let fq = flatten(r);
        let cap = r.k.min(self.len() * r.queries.len());
        let (mut idx, mut score) = (vec![0u64; cap.max(1)], vec![0f32; cap.max(1)]);
        let mut len = 0u64;
        check(unsafe {
            sys::otters_vecstore_query(self.vs, &fq.q, idx.as_mut_ptr(), score.as_mut_ptr(), ptr::null_mut(), cap as u64, &mut len)
        })?;
        drop(fq.flat);
        Ok(idx.into_iter().zip(score).take((len as usize).min(cap)).map(|(i, s)| (i as usize, s)).collect())
    }

    /// Extension: one list per query of the batch (`otters_vecstore_query_batch`).
    pub fn collect_per_query(&self, r: &ResolvedQuery<'_>) -> Result<Vec<Vec<(usize, f32)>>, String> {
        let mut fq = flatten(r);
        let (nq, k) = (r.queries.len(), r.k.min(self.len()).max(1));
        fq.q.k = if r.k == 0 { 0 } else { k as u64 };
        let (mut idx, mut score, mut lens) = (vec![0u64; nq * k], vec![0f32; nq * k], vec![0u64; nq]);
        check(unsafe { sys::otters_vecstore_query_batch(self.vs, &fq.q, idx.as_mut_ptr(), score.as_mut_ptr(), lens.as_mut_ptr()) })?;
        Ok((0..nq)
            .map(|i| (0..lens[i] as usize).map(|j| (idx[i * k + j] as usize, score[i * k + j])).collect())
            .collect())
    }

    /// Non-blocking form: at most two tickets outstanding per context (`otters_query_submit` / `otters_query_wait`).
    pub fn submit(&self, r: &ResolvedQuery<'_>) -> Result<u64, String> {
        let fq = flatten(r);
        let mut ticket = 0u64;
        check(unsafe { sys::otters_query_submit(self.vs, ptr::null_mut(), &fq.q, ptr::null(), ptr::null(), ptr::null(), 0, &mut ticket) })?;
        Ok(ticket) // the library has copied the queries: `fq` may go
    }

    pub fn wait(&self, ticket: u64, k: usize) -> Result<Vec<(usize, f32)>, String> {
        let (mut idx, mut score) = (vec![0u64; k.max(1)], vec![0f32; k.max(1)]);
        let mut len = 0u64;
        check(unsafe {
            sys::otters_query_wait(self.ctx.raw, ticket, idx.as_mut_ptr(), score.as_mut_ptr(), ptr::null_mut(), k as u64, &mut len, ptr::null_mut())
        })?;
        Ok(idx.into_iter().zip(score).take((len as usize).min(k)).map(|(i, s)| (i as usize, s)).collect())
    }
}

impl Drop for DeviceVecStore<'_> {
    fn drop(&mut self) {
        unsafe { sys::otters_vecstore_destroy(self.vs) };
    }
}

// =====================================================================================================================
// CompiledFilter -> otters_filter (src/expr.rs:192-226)
// =====================================================================================================================
fn cmp_code(c: CmpOp) -> i32 {
    match c {
        CmpOp::Eq => sys::OTTERS_OP_EQ,
        CmpOp::Neq => sys::OTTERS_OP_NEQ,
        CmpOp::Lt => sys::OTTERS_OP_LT,
        CmpOp::Lte => sys::OTTERS_OP_LTE,
        CmpOp::Gt => sys::OTTERS_OP_GT,
        CmpOp::Gte => sys::OTTERS_OP_GTE,
    }
}

/// Owns everything the `otters_filter` points to.
pub struct LoweredFilter {
    offsets: Vec<u32>,
    leaves: Vec<sys::otters_leaf>,
    _strings: Vec<Vec<u8>>,
}

impl LoweredFilter {
    /// Flattens the CNF (outer AND of inner ORs) into clause offsets + leaves; columns are resolved through the store's
    /// schema order.  An unknown column cannot occur here: `Expr::compile` has already checked the schema (src/expr.rs:385-466).
    pub fn new(f: &CompiledFilter, col_index: &HashMap<String, usize>) -> Result<Self, String> {
        let mut offsets = vec![0u32];
        let mut leaves = Vec::new();
        let mut strings: Vec<Vec<u8>> = Vec::new();
        for clause in &f.clauses {
            for leaf in clause {
                let (column, cmp) = match leaf {
                    ColumnFilter::Numeric { column, cmp, .. } | ColumnFilter::String { column, cmp, .. } => (column, *cmp),
                };
                let col = *col_index.get(column).ok_or_else(|| format!("Unknown column '{}'", column))? as u32;
                let mut l = sys::otters_leaf { col, op: cmp_code(cmp), kind: 0, i: 0, f: 0.0, s: ptr::null(), slen: 0 };
                match leaf {
                    ColumnFilter::Numeric { rhs: NumericLiteral::I64(v), .. } => {
                        l.kind = sys::OTTERS_LIT_I64;
                        l.i = *v;
                    }
                    ColumnFilter::Numeric { rhs: NumericLiteral::F64(v), .. } => {
                        l.kind = sys::OTTERS_LIT_F64;
                        l.f = *v;
                    }
                    ColumnFilter::String { rhs, .. } => {
                        strings.push(rhs.as_bytes().to_vec());
                        let b = strings.last().unwrap();
                        l.kind = sys::OTTERS_LIT_STR;
                        l.s = b.as_ptr(); // a Vec's heap buffer does not move when the outer Vec grows
                        l.slen = b.len() as u64;
                    }
                }
                leaves.push(l);
            }
            offsets.push(leaves.len() as u32);
        }
        Ok(Self { offsets, leaves, _strings: strings })
    }

    fn as_ffi(&self) -> sys::otters_filter {
        sys::otters_filter {
            n_clauses: (self.offsets.len() - 1) as u32,
            clause_offsets: self.offsets.as_ptr(),
            leaves: if self.leaves.is_empty() { ptr::null() } else { self.leaves.as_ptr() },
        }
    }
}

// =====================================================================================================================
// MetaStore
// =====================================================================================================================
/// MetaQueryStats (src/meta.rs:832-842) as the library reports it.
#[derive(Debug, Clone, Default)]
pub struct DeviceQueryStats {
    pub total_chunks: usize,
    pub pruned_chunks: usize,
    pub evaluated_chunks: usize,
    pub vectors_compared: usize,
    pub prune_duration: Duration,
    pub score_duration: Duration,
    pub merge_duration: Duration,
    pub total_duration: Duration,
}

impl From<sys::otters_query_stats> for DeviceQueryStats {
    fn from(s: sys::otters_query_stats) -> Self {
        Self {
            total_chunks: s.total_chunks as usize,
            pruned_chunks: s.pruned_chunks as usize,
            evaluated_chunks: s.evaluated_chunks as usize,
            vectors_compared: s.vectors_compared as usize,
            prune_duration: Duration::from_secs_f64(s.prune_s.max(0.0)),
            score_duration: Duration::from_secs_f64(s.score_s.max(0.0)),
            merge_duration: Duration::from_secs_f64(s.merge_s.max(0.0)),
            total_duration: Duration::from_secs_f64(s.total_s.max(0.0)),
        }
    }
}

/// MetaBuildStats (src/meta.rs:844-852)
#[derive(Debug, Clone, Default)]
pub struct DeviceBuildStats {
    pub n_rows: usize,
    pub dim: usize,
    pub n_chunks: usize,
    pub vectors_ingest_duration: Duration,
    pub zonemap_build_duration: Duration,
    pub build_total_duration: Duration,
}

/// Bloom sizing knob of MetaStoreBuilder (src/meta.rs:92-110)
pub enum BloomSpec {
    Fpr(f64),
    Bits(usize),
}

fn dtype_code(d: DataType) -> i32 {
    match d {
        DataType::Int32 => sys::OTTERS_DTYPE_INT32,
        DataType::Int64 => sys::OTTERS_DTYPE_INT64,
        DataType::Float32 => sys::OTTERS_DTYPE_FLOAT32,
        DataType::Float64 => sys::OTTERS_DTYPE_FLOAT64,
        DataType::String => sys::OTTERS_DTYPE_STRING,
        DataType::DateTime => sys::OTTERS_DTYPE_DATETIME,
    }
}

pub struct DeviceMetaStore<'c> {
    ctx: &'c DeviceContext,
    ms: *mut sys::otters_metastore,
    col_index: HashMap<String, usize>,
    dtypes: Vec<DataType>,
    n_rows: usize,
    pub build_stats: DeviceBuildStats,
}

impl<'c> DeviceMetaStore<'c> {
    /// The body of MetaStoreBuilder::build (src/meta.rs:151-305) after its own validation: vectors and columns go to HBM,
    /// zonemaps / Bloom filters / dictionary codes are built there (otters_b200/csrc/build.cu).  `columns` in schema order.
    pub fn build(ctx: &'c DeviceContext, vectors: &[Vec<f32>], columns: &[&Column], chunk_size: usize, bloom: BloomSpec) -> Result<Self, String> {
        Self::build_with_format(ctx, vectors, columns, chunk_size, bloom, VectorFormat::F32)
    }

    /// Same, with the rows kept in HBM as `format` (VectorFormat::Bf16: half the bytes per scan; scores are the
    /// reference's arithmetic on the rounded rows — the roadmap's "Quantization for vectors", README.md:208).
    pub fn build_with_format(ctx: &'c DeviceContext, vectors: &[Vec<f32>], columns: &[&Column], chunk_size: usize, bloom: BloomSpec,
                             format: VectorFormat) -> Result<Self, String> {
        let dim = vectors.first().map_or(0, |v| v.len());
        let mut flat = Vec::with_capacity(vectors.len() * dim);
        for v in vectors {
            if v.len() != dim {
                return Err(format!("Input vector length {} does not match expected dimension {}", v.len(), dim));
            }
            flat.extend_from_slice(v);
        }
        // per-column views in the ABI's layout; string columns are flattened to offsets + bytes
        let mut names: Vec<CString> = Vec::new();
        let mut str_offsets: Vec<Vec<u64>> = Vec::new();
        let mut str_bytes: Vec<Vec<u8>> = Vec::new();
        let mut ffi_cols: Vec<sys::otters_column> = Vec::new();
        let mut col_index = HashMap::new();
        let mut dtypes = Vec::new();
        for (i, c) in columns.iter().enumerate() {
            col_index.insert(c.name().to_string(), i);
            dtypes.push(c.dtype());
            names.push(CString::new(c.name()).map_err(|e| e.to_string())?);
            // BitVec<usize, Lsb0>: bit = 1 NULL (src/col.rs:26) — the raw words are what the ABI takes
            let nulls = c.null_mask();
            let null_words = if nulls.any() { nulls.as_raw_slice().as_ptr() as *const u64 } else { ptr::null() };
            let (mut values, mut offs, mut bytes): (*const c_void, *const u64, *const u8) = (ptr::null(), ptr::null(), ptr::null());
            match c.dtype() {
                DataType::Int32 => values = c.i32_values().unwrap().as_ptr() as *const c_void,
                DataType::Int64 => values = c.i64_values().unwrap().as_ptr() as *const c_void,
                DataType::Float32 => values = c.f32_values().unwrap().as_ptr() as *const c_void,
                DataType::Float64 => values = c.f64_values().unwrap().as_ptr() as *const c_void,
                DataType::DateTime => values = c.datetime_values().unwrap().as_ptr() as *const c_void,
                DataType::String => {
                    let strs = c.string_values().unwrap();
                    let mut o = Vec::with_capacity(strs.len() + 1);
                    let mut b = Vec::new();
                    o.push(0u64);
                    for s in strs {
                        b.extend_from_slice(s.as_bytes());
                        o.push(b.len() as u64);
                    }
                    if b.is_empty() {
                        b.push(0);
                    }
                    str_offsets.push(o);
                    str_bytes.push(b);
                    offs = str_offsets.last().unwrap().as_ptr();
                    bytes = str_bytes.last().unwrap().as_ptr();
                }
            }
            ffi_cols.push(sys::otters_column {
                name: names.last().unwrap().as_ptr(),
                dtype: dtype_code(c.dtype()),
                values,
                null_words,
                str_offsets: offs,
                str_bytes: bytes,
            });
        }
        let (bloom_mode, bloom_fpr, bloom_bits) = match bloom {
            BloomSpec::Fpr(p) => (0, p, 0u64),
            BloomSpec::Bits(b) => (1, 0.01, b as u64),
        };
        let params = sys::otters_build_params {
            n_rows: vectors.len() as u64,
            dim: dim as u32,
            chunk_size: chunk_size as u64,
            bloom_mode,
            bloom_fpr,
            bloom_bits,
            vectors_kind: sys::OTTERS_VECTORS_HOST,
            vectors: flat.as_ptr(),
            synthetic_seed: 0,
            synthetic_first_row: 0,
            synthetic_map: ptr::null(),
            columns: ffi_cols.as_ptr(),
            n_columns: ffi_cols.len() as u32,
            vector_format: format as i32,
        };
        let mut ms = ptr::null_mut();
        let mut st = sys::otters_build_stats::default();
        check(unsafe { sys::otters_metastore_build(ctx.raw, &params, &mut ms, &mut st) })?;
        Ok(Self {
            ctx,
            ms,
            col_index,
            dtypes,
            n_rows: vectors.len(),
            build_stats: DeviceBuildStats {
                n_rows: st.n_rows as usize,
                dim: st.dim as usize,
                n_chunks: st.n_chunks as usize,
                vectors_ingest_duration: Duration::from_secs_f64(st.vectors_ingest_s),
                zonemap_build_duration: Duration::from_secs_f64(st.zonemap_build_s),
                build_total_duration: Duration::from_secs_f64(st.build_total_s),
            },
        })
    }

    pub fn n_chunks(&self) -> usize {
        unsafe { sys::otters_metastore_n_chunks(self.ms) as usize }
    }

    /// The body of MetaQueryPlan::collect (src/meta.rs:632-721) after the `meta_filter compile error` check: chunk pruning,
    /// row predicate, scoring, top-k and stats in one call.  Returns (indices, scores, stats); the caller stores the stats
    /// in its `RefCell` and builds `MetaQueryResults.data` from `gather`.
    pub fn collect(&self, r: &ResolvedQuery<'_>, filter: Option<&CompiledFilter>) -> Result<(Vec<usize>, Vec<f32>, DeviceQueryStats), String> {
        let fq = flatten(r);
        let lowered = filter.map(|f| LoweredFilter::new(f, &self.col_index)).transpose()?;
        let ffi_filter = lowered.as_ref().map(|l| l.as_ffi());
        let cap = r.k.min(self.n_rows * r.queries.len().max(1)).max(1);
        let (mut idx, mut score) = (vec![0u64; cap], vec![0f32; cap]);
        let mut len = 0u64;
        let mut st = sys::otters_query_stats::default();
        check(unsafe {
            sys::otters_metastore_query(self.ms, &fq.q, ffi_filter.as_ref().map_or(ptr::null(), |f| f as *const _), idx.as_mut_ptr(),
                                        score.as_mut_ptr(), ptr::null_mut(), cap as u64, &mut len, &mut st)
        })?;
        let n = (len as usize).min(cap);
        Ok((idx[..n].iter().map(|&i| i as usize).collect(), score[..n].to_vec(), st.into()))
    }

    /// Extension: one result per query of the batch (`otters_metastore_query_batch`).
    pub fn collect_per_query(&self, r: &ResolvedQuery<'_>, filter: Option<&CompiledFilter>) -> Result<(Vec<Vec<(usize, f32)>>, DeviceQueryStats), String> {
        let mut fq = flatten(r);
        let lowered = filter.map(|f| LoweredFilter::new(f, &self.col_index)).transpose()?;
        let ffi_filter = lowered.as_ref().map(|l| l.as_ffi());
        let (nq, k) = (r.queries.len(), r.k.min(self.n_rows).max(1));
        fq.q.k = if r.k == 0 { 0 } else { k as u64 };
        let (mut idx, mut score, mut lens) = (vec![0u64; nq.max(1) * k], vec![0f32; nq.max(1) * k], vec![0u64; nq.max(1)]);
        let mut st = sys::otters_query_stats::default();
        check(unsafe {
            sys::otters_metastore_query_batch(self.ms, &fq.q, ffi_filter.as_ref().map_or(ptr::null(), |f| f as *const _), idx.as_mut_ptr(),
                                              score.as_mut_ptr(), lens.as_mut_ptr(), &mut st)
        })?;
        let lists = (0..nq).map(|i| (0..lens[i] as usize).map(|j| (idx[i * k + j] as usize, score[i * k + j])).collect()).collect();
        Ok((lists, st.into()))
    }

    /// MetaStore::last_query_stats (src/meta.rs:395-397)
    /// Roadmap "Persistence (save/load MetaStore to/from disk)" (README.md:206): writes the store's HBM image.  `user` is an
    /// opaque blob that comes back from `load` (the crate would keep its Columns or a row-order permutation there).
    pub fn save(&self, path: &str, user: &[u8]) -> Result<(), String> {
        let p = CString::new(path).map_err(|e| e.to_string())?;
        let up = if user.is_empty() { ptr::null() } else { user.as_ptr() as *const c_void };
        check(unsafe { sys::otters_metastore_save(self.ms, p.as_ptr(), up, user.len() as u64) })
    }

    /// Loads a store written by `save`: device arrays are copied back as they were, nothing is rebuilt.  Returns the store
    /// and the caller blob.
    pub fn load(ctx: &'c DeviceContext, path: &str) -> Result<(Self, Vec<u8>), String> {
        let p = CString::new(path).map_err(|e| e.to_string())?;
        let mut ms = ptr::null_mut();
        check(unsafe { sys::otters_metastore_load(ctx.raw, p.as_ptr(), &mut ms) })?;
        let mut col_index = HashMap::new();
        let mut dtypes = Vec::new();
        for i in 0..unsafe { sys::otters_metastore_n_columns(ms) } {
            let (mut name, mut dt) = (ptr::null(), 0i32);
            check(unsafe { sys::otters_metastore_column_info(ms, i, &mut name, &mut dt) })?;
            let name = unsafe { std::ffi::CStr::from_ptr(name) }.to_string_lossy().into_owned();
            col_index.insert(name, i as usize);
            dtypes.push(match dt {
                sys::OTTERS_DTYPE_INT32 => DataType::Int32,
                sys::OTTERS_DTYPE_INT64 => DataType::Int64,
                sys::OTTERS_DTYPE_FLOAT32 => DataType::Float32,
                sys::OTTERS_DTYPE_FLOAT64 => DataType::Float64,
                sys::OTTERS_DTYPE_STRING => DataType::String,
                _ => DataType::DateTime,
            });
        }
        let (mut bp, mut bl) = (ptr::null(), 0u64);
        check(unsafe { sys::otters_metastore_user_blob(ms, &mut bp, &mut bl) })?;
        let blob = if bl == 0 { Vec::new() } else { unsafe { std::slice::from_raw_parts(bp as *const u8, bl as usize) }.to_vec() };
        let n_rows = unsafe { sys::otters_metastore_len(ms) } as usize;
        let n_chunks = unsafe { sys::otters_metastore_n_chunks(ms) } as usize;
        let dim = unsafe { sys::otters_metastore_dim(ms) } as usize;
        Ok((Self { ctx, ms, col_index, dtypes, n_rows,
                   build_stats: DeviceBuildStats { n_rows, dim, n_chunks, ..Default::default() } }, blob))
    }

    pub fn last_query_stats(&self) -> Option<DeviceQueryStats> {
        let mut st = sys::otters_query_stats::default();
        if unsafe { sys::otters_metastore_last_stats(self.ms, &mut st) } == sys::OTTERS_OK { Some(st.into()) } else { None }
    }

    /// MetaQueryResults.data for one column (src/meta.rs:723-821): gathered on the device, NULLs preserved.
    pub fn gather(&self, column: &str, rows: &[usize]) -> Result<Column, String> {
        let ci = *self.col_index.get(column).ok_or_else(|| format!("Unknown column '{}'", column))?;
        let dt = self.dtypes[ci];
        let rows64: Vec<u64> = rows.iter().map(|&r| r as u64).collect();
        let n = rows.len();
        let mut raw = vec![0u64; n.max(1)]; // wide enough for every value type
        let mut nulls = vec![0u8; n.max(1)];
        check(unsafe { sys::otters_metastore_gather(self.ms, ci as u32, rows64.as_ptr(), n as u64, raw.as_mut_ptr() as *mut c_void, nulls.as_mut_ptr()) })?;
        let mut out = Column::new(column, dt);
        let push_err = |e: otters::col::ColumnError| format!("{:?}", e);
        for i in 0..n {
            let null = nulls[i] != 0;
            match dt {
                DataType::Int32 => {
                    let v = unsafe { *(raw.as_ptr() as *const i32).add(i) };
                    out.push(if null { None } else { Some(v) }).map_err(push_err)?;
                }
                DataType::Int64 => out.push(if null { None } else { Some(raw[i] as i64) }).map_err(push_err)?,
                DataType::DateTime => out.push(if null { None } else { Some(raw[i] as i64) }).map_err(push_err)?,
                DataType::Float32 => {
                    let v = unsafe { *(raw.as_ptr() as *const f32).add(i) };
                    out.push(if null { None } else { Some(v) }).map_err(push_err)?;
                }
                DataType::Float64 => out.push(if null { None } else { Some(f64::from_bits(raw[i])) }).map_err(push_err)?,
                DataType::String => {
                    let code = unsafe { *(raw.as_ptr() as *const u32).add(i) };
                    if null {
                        out.push(None::<String>).map_err(push_err)?;
                    } else {
                        let (mut p, mut len) = (ptr::null(), 0u64);
                        check(unsafe { sys::otters_metastore_dict_entry(self.ms, ci as u32, code, &mut p, &mut len) })?;
                        let bytes = unsafe { std::slice::from_raw_parts(p, len as usize) };
                        out.push(Some(String::from_utf8_lossy(bytes).into_owned())).map_err(push_err)?;
                    }
                }
            }
        }
        Ok(out)
    }

    pub fn context(&self) -> &DeviceContext {
        self.ctx
    }
}

impl Drop for DeviceMetaStore<'_> {
    fn drop(&mut self) {
        unsafe { sys::otters_metastore_destroy(self.ms) };
    }
}
