//! src/domain/ports/hybrid_search.rs — the new port, shaped like `PostAnalyzer`
//! (src/domain/ports/post_analyzer.rs:7-11).
use crate::domain_stubs::DomainError;
use async_trait::async_trait;

/// One query: an L2-normalised embedding of the index dimension and the query text's term ids
/// (index vocabulary; duplicates are harmless, the library de-duplicates: docs/SPEC.md §3).
pub struct SearchQuery {
    pub embedding: Vec<f32>,
    pub terms: Vec<u32>,
}

/// Ranks are 1-based; 0 = the document is absent from that modality's top-k.
#[derive(Debug, Clone, PartialEq)]
pub struct Hit {
    pub doc_id: u32,
    pub rrf: f32,
    pub rank_cosine: u32,
    pub rank_bm25: u32,
}

#[async_trait]
pub trait HybridSearch: Send + Sync {
    /// One ranked list (RRF desc, doc id asc, <= k hits) per query, aligned to input order
    /// (`len == queries.len()`, the contract the engine enforces for analyzers:
    /// src/domain/engine/speculation_engine.rs:29-34).
    async fn search(&self, queries: &[SearchQuery], k: usize) -> Result<Vec<Vec<Hit>>, DomainError>;
    /// Embedding dimension the index was built with.
    fn dim(&self) -> usize;
}
