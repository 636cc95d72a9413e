//! Port + adapters for the reference crate, written against its conventions:
//!   * ports are `#[async_trait] pub trait X: Send + Sync` returning `Result<_, DomainError>`
//!     (src/domain/ports/post_analyzer.rs:7-11), borrowed slices in, owned Vec out;
//!   * adapter faults map to `DomainError::SourceFailure { name, message }` (src/domain/error.rs:17-18);
//!   * adapters are built at the composition roots (src/main.rs:16-47, src/mcp/server.rs:237-259).
//! UNCOMPILED in the source repository (no Rust toolchain there); the tested equivalents are
//! openintel_b200/host/openintel_host.hpp and openintel_b200/capi.py.
use async_trait::async_trait;
use openintel_gpu_sys as sys;
use std::collections::HashMap;
use std::ffi::CStr;

// ---- stand-ins for crate::domain types of the reference (delete when vendored into the crate) ----
#[derive(Debug)]
pub enum DomainError {
    SourceFailure { name: String, message: String },
    AnalyzerMismatch { expected: usize, got: usize },
}
pub struct SocialPost { pub id: String, pub text: String }
pub struct PostSignal { pub polarity: f64, pub speculative: bool }
#[async_trait]
pub trait PostAnalyzer: Send + Sync {
    async fn analyze(&self, posts: &[SocialPost]) -> Result<Vec<PostSignal>, DomainError>;
}

// ---- the new port --------------------------------------------------------------------------------
pub struct SearchQuery { pub embedding: Vec<f32>, pub terms: Vec<u32> }
#[derive(Debug, Clone, PartialEq)]
pub struct Hit { pub doc_id: u32, pub rrf: f32, pub rank_cosine: u32, pub rank_bm25: u32 }

#[async_trait]
pub trait HybridSearch: Send + Sync {
    /// One ranked list (RRF desc, doc id asc, <= k hits) per query, aligned to input order.
    async fn search(&self, queries: &[SearchQuery], k: usize) -> Result<Vec<Vec<Hit>>, DomainError>;
}

fn fail(name: &str, h: *const sys::oi_index) -> DomainError {
    let message = unsafe { CStr::from_ptr(sys::oi_last_error(h)) }.to_string_lossy().into_owned();
    DomainError::SourceFailure { name: name.into(), message }
}

// ---- tokenizer (src/adapters/analyzer/lexicon.rs:54-58) -------------------------------------------
pub fn tokenize(text: &str) -> Vec<String> {
    text.to_lowercase().split(|c: char| !c.is_ascii_alphanumeric()).filter(|w| !w.is_empty()).map(str::to_owned).collect()
}

// ---- index builder: posts -> vocabulary (lexicographic term ids) + CSR ----------------------------
#[derive(Default)]
pub struct IndexBuilder {
    vocab: Vec<String>,
    term_offsets: Vec<u64>,
    doc_ids: Vec<u32>,
    tfs: Vec<u32>,
    doc_len: Vec<u32>,
    raw: Vec<(String, u32, u32)>, // (token, doc, tf), docs ascending
}
impl IndexBuilder {
    pub fn add(&mut self, text: &str) {
        let doc = self.doc_len.len() as u32;
        let toks = tokenize(text);
        self.doc_len.push(toks.len() as u32);
        let mut tf: HashMap<String, u32> = HashMap::new();
        for t in toks { *tf.entry(t).or_insert(0) += 1; }
        for (t, f) in tf { self.raw.push((t, doc, f)); }
    }
    pub fn finish(&mut self) {
        self.raw.sort_by(|a, b| a.0.cmp(&b.0).then(a.1.cmp(&b.1)));
        self.vocab.clear(); self.term_offsets.clear(); self.doc_ids.clear(); self.tfs.clear();
        for (t, d, f) in &self.raw {
            if self.vocab.last() != Some(t) { self.vocab.push(t.clone()); self.term_offsets.push(self.doc_ids.len() as u64); }
            self.doc_ids.push(*d); self.tfs.push(*f);
        }
        self.term_offsets.push(self.doc_ids.len() as u64);
    }
    pub fn term_id(&self, token: &str) -> Option<u32> { self.vocab.binary_search_by(|v| v.as_str().cmp(token)).ok().map(|i| i as u32) }
    pub fn query_terms(&self, text: &str) -> Vec<u32> { tokenize(text).iter().filter_map(|t| self.term_id(t)).collect() }
}

// ---- GPU adapter ----------------------------------------------------------------------------------
pub struct GpuHybridSearch { h: *mut sys::oi_index, dim: usize, rrf_k: u32 }
unsafe impl Send for GpuHybridSearch {} // the library serialises calls on one handle (openintel_gpu.h, "Conventions")
unsafe impl Sync for GpuHybridSearch {}
impl Drop for GpuHybridSearch { fn drop(&mut self) { unsafe { sys::oi_index_destroy(self.h) } } }

impl GpuHybridSearch {
    /// `embeddings`: n_docs x dim f32, L2-normalised, doc order = the builder's.
    pub fn new(device: i32, dim: usize, embeddings: &[f32], ix: &mut IndexBuilder, max_k: u32, max_batch: u32) -> Result<Self, DomainError> {
        ix.finish();
        let n_docs = ix.doc_len.len() as u64;
        let desc = sys::oi_index_desc { struct_size: std::mem::size_of::<sys::oi_index_desc>() as u32, device, n_docs, doc_base: 0,
                                        dim: dim as u32, dtype: sys::OI_DTYPE_F32, max_k, max_batch };
        let mut h = std::ptr::null_mut();
        if unsafe { sys::oi_index_create(&desc, &mut h) } != sys::OI_OK { return Err(fail("gpu-search", std::ptr::null())); }
        let me = GpuHybridSearch { h, dim, rrf_k: 60 };
        let p = sys::oi_bm25_params { struct_size: std::mem::size_of::<sys::oi_bm25_params>() as u32, k1: 1.2, b: 0.75, avgdl: 0.0,
                                      n_docs_global: 0, global_df: std::ptr::null() };
        unsafe {
            if sys::oi_index_load_embeddings(h, embeddings.as_ptr().cast(), 0, n_docs) != sys::OI_OK
                || sys::oi_index_load_bm25(h, ix.term_offsets.as_ptr(), ix.doc_ids.as_ptr(), ix.tfs.as_ptr(), ix.doc_len.as_ptr(), ix.vocab.len() as u32) != sys::OI_OK
                || sys::oi_index_bm25_finalize(h, &p) != sys::OI_OK { return Err(fail("gpu-search", h)); }
        }
        Ok(me)
    }
}

#[async_trait]
impl HybridSearch for GpuHybridSearch {
    async fn search(&self, queries: &[SearchQuery], k: usize) -> Result<Vec<Vec<Hit>>, DomainError> {
        let nq = queries.len();
        let mut emb = Vec::with_capacity(nq * self.dim);
        let (mut terms, mut offs) = (Vec::<u32>::new(), vec![0u32]);
        for q in queries { emb.extend_from_slice(&q.embedding); terms.extend_from_slice(&q.terms); offs.push(terms.len() as u32); }
        if terms.is_empty() { terms.push(0); }
        let (h, rrf_k) = (self.h as usize, self.rrf_k);
        let out = tokio::task::spawn_blocking(move || { // blocking FFI call off the async executor
            let n = nq * k;
            let (mut ids, mut rrf, mut rc, mut rb) = (vec![0u32; n], vec![0f32; n], vec![0u32; n], vec![0u32; n]);
            let st = unsafe { sys::oi_search_hybrid(h as *mut _, emb.as_ptr(), terms.as_ptr(), offs.as_ptr(), nq as u32, k as u32, rrf_k,
                                                    ids.as_mut_ptr(), rrf.as_mut_ptr(), rc.as_mut_ptr(), rb.as_mut_ptr()) };
            (st, ids, rrf, rc, rb)
        }).await.map_err(|e| DomainError::SourceFailure { name: "gpu-search".into(), message: e.to_string() })?;
        let (st, ids, rrf, rc, rb) = out;
        if st != sys::OI_OK { return Err(fail("gpu-search", self.h)); }
        Ok((0..nq).map(|j| (0..k).map(|i| j * k + i).take_while(|&a| ids[a] != sys::OI_NO_DOC)
            .map(|a| Hit { doc_id: ids[a], rrf: rrf[a], rank_cosine: rc[a], rank_bm25: rb[a] }).collect()).collect())
    }
}

/// Replaces `LexiconAnalyzer::analyze` (src/adapters/analyzer/lexicon.rs:82-87) with the batched GPU scorer.
pub struct GpuLexiconAnalyzer { pub device: i32 }
#[async_trait]
impl PostAnalyzer for GpuLexiconAnalyzer {
    async fn analyze(&self, posts: &[SocialPost]) -> Result<Vec<PostSignal>, DomainError> {
        let mut blob = Vec::new();
        let mut offs = vec![0u64];
        for p in posts { blob.extend_from_slice(p.text.as_bytes()); offs.push(blob.len() as u64); }
        if blob.is_empty() { blob.push(0); }
        let n = posts.len();
        let (mut pol, mut spec) = (vec![0f64; n], vec![0u8; n]);
        let st = unsafe { sys::oi_lexicon_analyze(self.device, blob.as_ptr(), offs.as_ptr(), n as u64, pol.as_mut_ptr(), spec.as_mut_ptr(),
                                                  std::ptr::null_mut(), std::ptr::null_mut()) };
        if st != sys::OI_OK { return Err(fail("gpu-lexicon", std::ptr::null())); }
        if pol.len() != n { return Err(DomainError::AnalyzerMismatch { expected: n, got: pol.len() }); }
        Ok(pol.into_iter().zip(spec).map(|(p, s)| PostSignal { polarity: p, speculative: s != 0 }).collect())
    }
}
