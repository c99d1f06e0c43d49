//! Rust shim over the C ABI of include/probly_b200.h — NOT BUILT OR TESTED HERE (the image has no
//! Rust toolchain); it documents exactly what the reference-side binding looks like.
//!
//! Surface kept from probly-search 2.0.1:
//!   `Index::<T>::new(fields_num)`                      src/index.rs:37
//!   `add_document(&[FieldAccessor<D>], Tokenizer, key, &doc)`   src/index.rs:77
//!   `remove_document(key)` / `vacuum()`                src/index.rs:161 / :194
//!   `query(&self, &str, &mut S, Tokenizer, &[f64]) -> Vec<QueryResult<T>>`   src/query.rs:21
//!   `score::bm25::new()`, `score::zero_to_one::new()`  bm25.rs:21, zero_to_one.rs:35
//! plus `query_batch` (many queries, top-k each).
//!
//! `ScoreCalculator` is arbitrary user code in the reference; only the two built-in calculators
//! exist as device code.  The trait bound on the GPU entry points is therefore the SEALED marker
//! `DeviceScorer`: any other implementor is a compile-time error (there is no CPU fallback).
use std::borrow::Cow;
use std::collections::HashMap;
use std::ffi::CStr;
use std::hash::Hash;
use std::os::raw::{c_char, c_int};

pub type FieldAccessor<D> = fn(&D) -> Vec<&str>; // src/lib.rs:11
pub type Tokenizer = fn(&str) -> Vec<Cow<'_, str>>; // src/lib.rs:14

#[derive(Debug, PartialEq)]
pub struct QueryResult<T> {
    pub key: T,
    pub score: f64,
} // src/query.rs:9-15

// ------------------------------------------------------------------------------------------------
// FFI (mirrors include/probly_b200.h field by field)
// ------------------------------------------------------------------------------------------------
#[repr(C)]
struct PbDocTokens {
    tok_bytes: *const u8,
    tok_off: *const u64,
    value_tok_count: *const u32,
    field_value_count: *const u32,
}
#[repr(C)]
struct PbIndexImage {
    // filled by pb_builder_flatten, passed straight to pb_index_create.  248 bytes, 8-aligned: pinned by
    // PB_STATIC_ASSERT(sizeof(pb_index_image) == 248) in include/probly_b200.h (use bindgen in a real build).
    _opaque: [u64; 31],
}
#[repr(C)]
struct PbQueryBatchDesc {
    n_queries: u64,
    query_term_off: *const u64,
    term_byte_off: *const u64,
    term_bytes: *const u8,
    scorer: u32,
    bm25_k1: f64,
    bm25_b: f64,
    fields_boost: *const f64,
    n_fields_boost: u32,
    top_k: u32,
}
#[repr(C)]
struct PbQueryResults {
    n_results: *mut u64,
    doc_digest: *mut u64,
    score_digest: *mut u64,
    topk_n: *mut u32,
    topk_doc: *mut u32,
    topk_score: *mut f64,
}
#[repr(C)] struct PbBuilder { _p: [u8; 0] }
#[repr(C)] struct PbIndex { _p: [u8; 0] }
#[repr(C)] struct PbImageFile { _p: [u8; 0] }
#[repr(C)] struct PbComm { _p: [u8; 0] }
#[repr(C)] struct PbGroup { _p: [u8; 0] }

extern "C" {
    fn pb_builder_create(num_fields: u32, out: *mut *mut PbBuilder) -> c_int;
    fn pb_builder_destroy(b: *mut PbBuilder);
    fn pb_builder_add_document(b: *mut PbBuilder, key: u64, doc: *const PbDocTokens) -> c_int;
    fn pb_builder_remove_document(b: *mut PbBuilder, key: u64) -> c_int;
    fn pb_builder_vacuum(b: *mut PbBuilder) -> c_int;
    fn pb_builder_flatten(b: *mut PbBuilder, out: *mut PbIndexImage) -> c_int;
    fn pb_index_create(image: *const PbIndexImage, device: c_int, out: *mut *mut PbIndex) -> c_int;
    fn pb_index_destroy(ix: *mut PbIndex);
    // incremental maintenance: a delta segment = the rows of the docs added since `from_doc_ordinal` under the
    // current trie; each segment is told the other's per-term live counts (BM25 df is over the whole index)
    fn pb_builder_flatten_from(b: *mut PbBuilder, from_doc_ordinal: u64, out: *mut PbIndexImage) -> c_int;
    fn pb_builder_flatten_term_ids(b: *const PbBuilder, out: *mut u32, cap: u64, n_terms: *mut u64) -> c_int;
    fn pb_index_set_df_extra(ix: *mut PbIndex, df_extra: *const u64, n: u64) -> c_int;
    // multi-GPU: one process per GPU (pb_comm + pb_batch_set_gather) or one process, several devices (pb_group)
    fn pb_comm_unique_id(id: *mut u8) -> c_int; // [128]
    fn pb_comm_create(id: *const u8, rank: c_int, world: c_int, device: c_int, out: *mut *mut PbComm) -> c_int;
    fn pb_comm_destroy(c: *mut PbComm);
    fn pb_group_create(image: *const PbIndexImage, devices: *const c_int, n: c_int, out: *mut *mut PbGroup) -> c_int;
    fn pb_group_query_batch(g: *mut PbGroup, q: *const PbQueryBatchDesc, out: *mut PbQueryResults) -> c_int;
    fn pb_group_destroy(g: *mut PbGroup);
    // on-disk image (the reference has no serialisation): save what `flatten` produced, serve from a file
    fn pb_image_save(image: *const PbIndexImage, path: *const c_char) -> c_int;
    fn pb_image_load(path: *const c_char, out: *mut *mut PbImageFile) -> c_int;
    fn pb_image_file_image(f: *const PbImageFile) -> *const PbIndexImage;
    fn pb_image_file_free(f: *mut PbImageFile);
    fn pb_query_batch(ix: *mut PbIndex, q: *const PbQueryBatchDesc, out: *mut PbQueryResults) -> c_int;
    fn pb_query_full(ix: *mut PbIndex, q: *const PbQueryBatchDesc, cap: u64, out_query: *mut u32,
                     out_doc: *mut u32, out_score: *mut f64, n_total: *mut u64) -> c_int;
    fn pb_last_error() -> *const c_char;
}

fn check(rc: c_int) {
    if rc != 0 {
        // the reference panics on its error paths too (unwrap() at src/query.rs:46,63,70)
        let msg = unsafe { CStr::from_ptr(pb_last_error()) }.to_string_lossy().into_owned();
        panic!("probly_b200 error {rc}: {msg}");
    }
}

// ------------------------------------------------------------------------------------------------
// score module
// ------------------------------------------------------------------------------------------------
pub mod score {
    mod sealed { pub trait Sealed {} }
    /// Calculators that exist as device code.  Sealed: user types cannot implement it.
    pub trait DeviceScorer: sealed::Sealed {
        #[doc(hidden)] fn pb_params(&self) -> (u32, f64, f64);
    }
    pub mod bm25 {
        pub struct BM25 { pub bm25k1: f64, pub bm25b: f64 } // bm25.rs:14-20
        pub fn new() -> BM25 { BM25 { bm25k1: 1.2, bm25b: 0.75 } } // bm25.rs:21-26
        impl super::sealed::Sealed for BM25 {}
        impl super::DeviceScorer for BM25 { fn pb_params(&self) -> (u32, f64, f64) { (0, self.bm25k1, self.bm25b) } }
    }
    pub mod zero_to_one {
        pub struct ZeroToOne; // zero_to_one.rs:24-26 — the per-query state lives in the device workspace
        pub fn new() -> ZeroToOne { ZeroToOne } // zero_to_one.rs:35-39
        impl super::sealed::Sealed for ZeroToOne {}
        impl super::DeviceScorer for ZeroToOne { fn pb_params(&self) -> (u32, f64, f64) { (1, 1.2, 0.75) } }
    }
}

// ------------------------------------------------------------------------------------------------
// Index
// ------------------------------------------------------------------------------------------------
pub struct Index<T> {
    builder: *mut PbBuilder,
    device_index: std::cell::Cell<*mut PbIndex>,
    dirty: std::cell::Cell<bool>,
    fields_num: usize,
    key_to_id: HashMap<T, u64>,
    id_to_key: Vec<T>,
    ord_to_id: std::cell::RefCell<Vec<u64>>,
}
unsafe impl<T: Send> Send for Index<T> {} // the reference Index is used inside a Mutex (tests/integrations_tests.rs:151-168)

impl<T: Eq + Hash + Copy + std::fmt::Debug> Index<T> {
    pub fn new(fields_num: usize) -> Self {
        let mut b = std::ptr::null_mut();
        check(unsafe { pb_builder_create(fields_num as u32, &mut b) });
        Index { builder: b, device_index: std::cell::Cell::new(std::ptr::null_mut()), dirty: std::cell::Cell::new(true),
                fields_num, key_to_id: HashMap::new(), id_to_key: Vec::new(), ord_to_id: Default::default() }
    }

    pub fn add_document<D>(&mut self, field_accessors: &[FieldAccessor<D>], tokenizer: Tokenizer, key: T, doc: &D) {
        let (mut bytes, mut off, mut vcount, mut fcount) = (Vec::<u8>::new(), vec![0u64], Vec::<u32>::new(), Vec::<u32>::new());
        for i in 0..self.fields_num {
            let values = field_accessors[i](doc);
            fcount.push(values.len() as u32);
            for v in values {
                let terms = tokenizer(v);
                vcount.push(terms.len() as u32);
                for t in terms { bytes.extend_from_slice(t.as_bytes()); off.push(bytes.len() as u64); }
            }
        }
        vcount.push(0);
        let id = *self.key_to_id.entry(key).or_insert_with(|| { self.id_to_key.push(key); (self.id_to_key.len() - 1) as u64 });
        let d = PbDocTokens { tok_bytes: bytes.as_ptr(), tok_off: off.as_ptr(), value_tok_count: vcount.as_ptr(), field_value_count: fcount.as_ptr() };
        check(unsafe { pb_builder_add_document(self.builder, id, &d) });
        self.dirty.set(true);
    }

    pub fn remove_document(&mut self, key: T) {
        if let Some(&id) = self.key_to_id.get(&key) { check(unsafe { pb_builder_remove_document(self.builder, id) }); self.dirty.set(true); }
    }

    pub fn vacuum(&mut self) { check(unsafe { pb_builder_vacuum(self.builder) }); self.dirty.set(true); }

    fn sync_device(&self) {
        if !self.dirty.get() { return; }
        let mut image = PbIndexImage { _opaque: [0; 256] };
        check(unsafe { pb_builder_flatten(self.builder, &mut image) });
        let old = self.device_index.replace(std::ptr::null_mut());
        if !old.is_null() { unsafe { pb_index_destroy(old) }; }
        let mut ix = std::ptr::null_mut();
        check(unsafe { pb_index_create(&image, 0, &mut ix) });
        self.device_index.set(ix);
        // doc_key (ordinal -> id) sits behind a pointer inside the image; a real build would use
        // bindgen for pb_index_image instead of the opaque blob and copy it here.
        self.dirty.set(false);
    }

    /// src/query.rs:21-106 — same signature except the calculator bound (`DeviceScorer`).
    pub fn query<S: score::DeviceScorer>(&self, query: &str, score_calculator: &mut S, tokenizer: Tokenizer,
                                         fields_boost: &[f64]) -> Vec<QueryResult<T>> {
        self.sync_device();
        let terms = tokenizer(query);
        let (mut bytes, mut toff) = (Vec::<u8>::new(), vec![0u64]);
        for t in &terms { bytes.extend_from_slice(t.as_bytes()); toff.push(bytes.len() as u64); }
        let qoff = [0u64, terms.len() as u64];
        let (scorer, k1, b) = score_calculator.pb_params();
        let d = PbQueryBatchDesc { n_queries: 1, query_term_off: qoff.as_ptr(), term_byte_off: toff.as_ptr(), term_bytes: bytes.as_ptr(),
                                   scorer, bm25_k1: k1, bm25_b: b, fields_boost: fields_boost.as_ptr(),
                                   n_fields_boost: fields_boost.len() as u32, top_k: 0 };
        let mut cap = 1024u64;
        loop {
            let (mut oq, mut od, mut os) = (vec![0u32; cap as usize], vec![0u32; cap as usize], vec![0f64; cap as usize]);
            let mut n = 0u64;
            let rc = unsafe { pb_query_full(self.device_index.get(), &d, cap, oq.as_mut_ptr(), od.as_mut_ptr(), os.as_mut_ptr(), &mut n) };
            if rc == -4 { cap = n + 16; continue; } // PB_ERR_CAPACITY
            check(rc);
            let ords = self.ord_to_id.borrow();
            let mut res: Vec<QueryResult<T>> = (0..n as usize)
                .map(|i| QueryResult { key: self.id_to_key[ords[od[i] as usize] as usize], score: os[i] }).collect();
            res.sort_by(|a, b| b.score.partial_cmp(&a.score).unwrap()); // src/query.rs:103
            return res;
        }
    }
}

impl<T> Drop for Index<T> {
    fn drop(&mut self) {
        unsafe {
            if !self.device_index.get().is_null() { pb_index_destroy(self.device_index.get()); }
            pb_builder_destroy(self.builder);
        }
    }
}

#[allow(dead_code)]
fn _uses(_: PbQueryResults) { let _ = pb_query_batch; }
