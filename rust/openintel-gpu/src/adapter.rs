//! src/adapters/search/gpu.rs — `GpuHybridSearch: HybridSearch` over libopenintel_gpu.so, and
//! src/adapters/analyzer/gpu.rs — `GpuLexiconAnalyzer: PostAnalyzer`
//! (replaces `LexiconAnalyzer::analyze`, src/adapters/analyzer/lexicon.rs:82-87).
use crate::domain_stubs::{DomainError, PostAnalyzer, PostSignal, SocialPost};
use crate::index_builder::IndexBuilder;
use crate::ports::{Hit, HybridSearch, SearchQuery};
use async_trait::async_trait;
use openintel_gpu_sys as sys;
use std::ffi::CStr;
use std::sync::Arc;

const NAME: &str = "gpu-search";

fn source_failure(message: impl Into<String>) -> DomainError {
    // same helper idiom as src/adapters/market/yahoo/response.rs:74
    DomainError::SourceFailure { name: NAME.into(), message: message.into() }
}

/// Owns the `oi_index*`.  Held in an `Arc` that every in-flight blocking call clones, so `oi_index_destroy`
/// cannot run while `oi_search_hybrid` is still inside the library — not even when the awaiting future is
/// cancelled and the adapter dropped.
struct Handle(*mut sys::oi_index);
unsafe impl Send for Handle {} // the library serialises calls on one handle (include/openintel_gpu.h, "Conventions")
unsafe impl Sync for Handle {}
impl Drop for Handle {
    fn drop(&mut self) {
        unsafe { sys::oi_index_destroy(self.0) }
    }
}
impl Handle {
    fn last_error(&self) -> DomainError {
        let msg = unsafe { CStr::from_ptr(sys::oi_last_error(self.0)) }.to_string_lossy().into_owned();
        source_failure(msg)
    }
}

pub struct GpuHybridSearch {
    h: Arc<Handle>,
    dim: usize,
    max_k: usize,
    max_batch: usize,
    rrf_k: u32,
}

impl GpuHybridSearch {
    /// `embeddings`: n_docs x dim f32, L2-normalised, doc order = the builder's.
    pub fn new(device: i32, dim: usize, embeddings: &[f32], ix: &mut IndexBuilder, max_k: u32, max_batch: u32) -> Result<Self, DomainError> {
        ix.finish();
        let n_docs = ix.n_docs();
        // the library reads n_docs * dim floats from `embeddings`: a short slice must never reach it
        if dim == 0 || embeddings.len() != n_docs * dim {
            return Err(source_failure(format!("embeddings hold {} floats, expected {} docs x {} dims", embeddings.len(), n_docs, dim)));
        }
        let desc = sys::oi_index_desc {
            struct_size: std::mem::size_of::<sys::oi_index_desc>() as u32,
            device,
            n_docs: n_docs as u64,
            doc_base: 0,
            dim: dim as u32,
            dtype: sys::OI_DTYPE_F32,
            max_k,
            max_batch,
        };
        let mut raw = std::ptr::null_mut();
        if unsafe { sys::oi_index_create(&desc, &mut raw) } != sys::OI_OK {
            let msg = unsafe { CStr::from_ptr(sys::oi_last_error(std::ptr::null())) }.to_string_lossy().into_owned();
            return Err(source_failure(msg));
        }
        let h = Arc::new(Handle(raw));
        let p = sys::oi_bm25_params {
            struct_size: std::mem::size_of::<sys::oi_bm25_params>() as u32,
            k1: 1.2,
            b: 0.75,
            avgdl: 0.0,
            n_docs_global: 0,
            global_df: std::ptr::null(),
        };
        let ok = unsafe {
            sys::oi_index_load_embeddings(h.0, embeddings.as_ptr().cast(), 0, n_docs as u64) == sys::OI_OK
                && sys::oi_index_load_bm25(h.0, ix.term_offsets.as_ptr(), ix.doc_ids.as_ptr(), ix.tfs.as_ptr(), ix.doc_len.as_ptr(),
                                           ix.vocab.len() as u32) == sys::OI_OK
                && sys::oi_index_bm25_finalize(h.0, &p) == sys::OI_OK
        };
        if !ok {
            return Err(h.last_error());
        }
        Ok(GpuHybridSearch { h, dim, max_k: max_k as usize, max_batch: max_batch as usize, rrf_k: 60 })
    }
}

#[async_trait]
impl HybridSearch for GpuHybridSearch {
    async fn search(&self, queries: &[SearchQuery], k: usize) -> Result<Vec<Vec<Hit>>, DomainError> {
        let nq = queries.len();
        if nq == 0 {
            return Ok(Vec::new());
        }
        if k == 0 || k > self.max_k || nq > self.max_batch {
            return Err(source_failure(format!("k = {k} (max {}) / {nq} queries (max {})", self.max_k, self.max_batch)));
        }
        let mut emb = Vec::with_capacity(nq * self.dim);
        let (mut terms, mut offs) = (Vec::<u32>::new(), vec![0u32]);
        for (j, q) in queries.iter().enumerate() {
            // oi_search_hybrid reads nq * dim floats: every embedding must have exactly `dim` of them
            if q.embedding.len() != self.dim {
                return Err(source_failure(format!("query {j}: embedding has {} dims, the index has {}", q.embedding.len(), self.dim)));
            }
            emb.extend_from_slice(&q.embedding);
            terms.extend_from_slice(&q.terms);
            offs.push(terms.len() as u32);
        }
        if terms.is_empty() {
            terms.push(0); // a valid pointer for an empty array
        }
        let (h, rrf_k) = (Arc::clone(&self.h), self.rrf_k);
        let out = tokio::task::spawn_blocking(move || {
            // blocking FFI call off the async executor; `h` keeps the index alive until the call returns
            let n = nq * k;
            let (mut ids, mut rrf, mut rc, mut rb) = (vec![0u32; n], vec![0f32; n], vec![0u32; n], vec![0u32; n]);
            let st = unsafe {
                sys::oi_search_hybrid(h.0, emb.as_ptr(), terms.as_ptr(), offs.as_ptr(), nq as u32, k as u32, rrf_k, ids.as_mut_ptr(),
                                      rrf.as_mut_ptr(), rc.as_mut_ptr(), rb.as_mut_ptr())
            };
            if st != sys::OI_OK {
                return Err(h.last_error());
            }
            Ok((ids, rrf, rc, rb))
        })
        .await
        .map_err(|e| source_failure(e.to_string()))??;
        let (ids, rrf, rc, rb) = out;
        Ok((0..nq)
            .map(|j| {
                (0..k)
                    .map(|i| j * k + i)
                    .take_while(|&a| ids[a] != sys::OI_NO_DOC) // padding of short lists
                    .map(|a| Hit { doc_id: ids[a], rrf: rrf[a], rank_cosine: rc[a], rank_bm25: rb[a] })
                    .collect()
            })
            .collect())
    }
    fn dim(&self) -> usize {
        self.dim
    }
}

/// The batched GPU lexicon scorer behind the reference's existing `PostAnalyzer` port.
pub struct GpuLexiconAnalyzer {
    pub device: i32,
}
#[async_trait]
impl PostAnalyzer for GpuLexiconAnalyzer {
    async fn analyze(&self, posts: &[SocialPost]) -> Result<Vec<PostSignal>, DomainError> {
        let mut blob = Vec::new();
        let mut offs = vec![0u64];
        for p in posts {
            blob.extend_from_slice(p.text.as_bytes());
            offs.push(blob.len() as u64);
        }
        if blob.is_empty() {
            blob.push(0);
        }
        let n = posts.len();
        let device = self.device;
        let (pol, spec) = tokio::task::spawn_blocking(move || {
            let (mut pol, mut spec) = (vec![0f64; n], vec![0u8; n]);
            let st = unsafe {
                sys::oi_lexicon_analyze(device, blob.as_ptr(), offs.as_ptr(), n as u64, pol.as_mut_ptr(), spec.as_mut_ptr(), std::ptr::null_mut(),
                                        std::ptr::null_mut())
            };
            if st != sys::OI_OK {
                let msg = unsafe { CStr::from_ptr(sys::oi_last_error(std::ptr::null())) }.to_string_lossy().into_owned();
                return Err(DomainError::SourceFailure { name: "gpu-lexicon".into(), message: msg });
            }
            Ok((pol, spec))
        })
        .await
        .map_err(|e| DomainError::SourceFailure { name: "gpu-lexicon".into(), message: e.to_string() })??;
        if pol.len() != n {
            return Err(DomainError::AnalyzerMismatch { expected: n, got: pol.len() });
        }
        Ok(pol.into_iter().zip(spec).map(|(p, s)| PostSignal { polarity: p, speculative: s != 0 }).collect())
    }
}
