//! How the reference's `collect()` bodies are replaced (sketch kept in sync with include/otters_b200.h).
//!
//! In the otters crate, `VecStore` gains a lazily created device handle and `VecQueryPlan::collect`
//! (src/vec.rs:206-311) keeps `validate()` and the default resolution, then makes ONE FFI call.  All builder
//! methods, error strings, `Metric`/`Cmp`/`TakeType`, `Expr::compile`, `Column` and the result-column gather
//! (src/meta.rs:723-828) stay untouched Rust.
use otters_sys as sys;
use std::ffi::CStr;

fn last_error() -> String {
    unsafe { CStr::from_ptr(sys::otters_last_error()).to_string_lossy().into_owned() }
}

pub struct DeviceVecStore {
    ctx: *mut sys::otters_ctx,
    vs: *mut sys::otters_vecstore,
    dim: usize,
}

impl DeviceVecStore {
    pub fn new(dim: usize) -> Result<Self, String> {
        let mut ctx = std::ptr::null_mut();
        let mut vs = std::ptr::null_mut();
        unsafe {
            if sys::otters_ctx_create(0, std::ptr::null_mut(), &mut ctx) != 0 { return Err(last_error()); }
            if sys::otters_vecstore_create(ctx, dim as u32, &mut vs) != 0 { return Err(last_error()); }
        }
        Ok(Self { ctx, vs, dim })
    }

    /// VecStore::add_vectors (src/vec.rs:374-376) with the rows flattened once.
    pub fn add_vectors(&mut self, rows: &[Vec<f32>]) -> Result<(), String> {
        let mut flat = Vec::with_capacity(rows.len() * self.dim);
        for r in rows {
            if r.len() != self.dim {
                return Err(format!("Input vector length {} does not match expected dimension {}", r.len(), self.dim));
            }
            flat.extend_from_slice(r);
        }
        if unsafe { sys::otters_vecstore_add(self.vs, flat.as_ptr(), rows.len() as u64) } != 0 { return Err(last_error()); }
        Ok(())
    }

    /// The body of VecQueryPlan::collect after validate(): metric/take_type/k/filter are the plan's resolved fields.
    #[allow(clippy::too_many_arguments)]
    pub fn collect(&self, queries: &[Vec<f32>], metric: i32, take_type: i32, k: usize, filter: Option<(f32, i32)>,
                   row_mask_words: Option<(&[u64], usize)>) -> Result<Vec<(usize, f32)>, String> {
        let flat: Vec<f32> = queries.iter().flatten().copied().collect();
        let n = unsafe { sys::otters_vecstore_len(self.vs) } as usize;
        let cap = k.min(n * queries.len());
        let (mut idx, mut score) = (vec![0u64; cap], vec![0f32; cap]);
        let q = sys::otters_vec_query {
            queries: flat.as_ptr(), nq: queries.len() as u32, dim: queries.first().map_or(0, |q| q.len()) as u32,
            metric, take_type, k: k as u64,
            has_filter: filter.is_some() as i32, thr: filter.map_or(0.0, |f| f.0), cmp: filter.map_or(0, |f| f.1),
            row_mask_words: row_mask_words.map_or(std::ptr::null(), |m| m.0.as_ptr()),
            row_mask_bits: row_mask_words.map_or(0, |m| m.1 as u64),
        };
        let mut len = 0u64;
        let rc = unsafe { sys::otters_vecstore_query(self.vs, &q, idx.as_mut_ptr(), score.as_mut_ptr(), std::ptr::null_mut(), cap as u64, &mut len) };
        if rc != 0 { return Err(last_error()); }
        Ok(idx.into_iter().zip(score).take(len as usize).map(|(i, s)| (i as usize, s)).collect())
    }
}

impl Drop for DeviceVecStore {
    fn drop(&mut self) {
        unsafe {
            sys::otters_vecstore_destroy(self.vs);
            sys::otters_ctx_destroy(self.ctx);
        }
    }
}
