//! src/application/search.rs — the use case, in the shape of `application::analyze`
//! (src/application/analyze.rs:16-73): async orchestration at the IO edge, ports injected as `&dyn`,
//! `DomainError` out.  The reference has no embedding model, so the caller supplies the query vector
//! (an embedder port would slot in exactly like `PostAnalyzer` does).
use crate::domain_stubs::DomainError;
use crate::index_builder::IndexBuilder;
use crate::ports::{Hit, HybridSearch, SearchQuery};

pub struct SearchRequest {
    pub text: String,
    /// same dimension as the index; normalised here if it is not yet
    pub embedding: Vec<f32>,
    pub k: usize,
}

#[derive(Debug, Clone, PartialEq)]
pub struct PostHit {
    pub post_id: String,
    pub doc_id: u32,
    pub rrf: f32,
    pub rank_cosine: u32,
    pub rank_bm25: u32,
}

pub struct SearchReport {
    pub hits: Vec<PostHit>,
    /// query tokens that are not in the index vocabulary (reported as a note, not an error: the reference's
    /// graceful-degradation convention, src/application/analyze.rs:40-45)
    pub notes: Vec<String>,
}

pub async fn search(req: &SearchRequest, index: &IndexBuilder, searcher: &dyn HybridSearch) -> Result<SearchReport, DomainError> {
    if req.k == 0 || index.n_docs() == 0 {
        return Err(DomainError::NoData);
    }
    if req.embedding.len() != searcher.dim() {
        return Err(DomainError::SourceFailure {
            name: "search".into(),
            message: format!("query embedding has {} dims, the index has {}", req.embedding.len(), searcher.dim()),
        });
    }
    let mut notes = Vec::new();
    let tokens = crate::index_builder::tokenize(&req.text);
    let unknown: Vec<&String> = tokens.iter().filter(|t| index.term_id(t).is_none()).collect();
    if !unknown.is_empty() {
        notes.push(format!("{} query token(s) not in the index vocabulary", unknown.len()));
    }
    let norm = req.embedding.iter().map(|x| (*x as f64) * (*x as f64)).sum::<f64>().sqrt();
    let embedding: Vec<f32> = if norm > 0.0 { req.embedding.iter().map(|x| (*x as f64 / norm) as f32).collect() } else { req.embedding.clone() };
    let query = SearchQuery { embedding, terms: index.query_terms(&req.text) };
    let mut lists = searcher.search(std::slice::from_ref(&query), req.k).await?;
    // the port's contract: one list per query (the check the engine applies to analyzers,
    // src/domain/engine/speculation_engine.rs:29-34)
    if lists.len() != 1 {
        return Err(DomainError::AnalyzerMismatch { expected: 1, got: lists.len() });
    }
    let hits: Vec<Hit> = lists.pop().unwrap();
    Ok(SearchReport {
        hits: hits
            .into_iter()
            .map(|h| PostHit {
                post_id: index.post_id(h.doc_id).unwrap_or("").to_owned(),
                doc_id: h.doc_id,
                rrf: h.rrf,
                rank_cosine: h.rank_cosine,
                rank_bm25: h.rank_bm25,
            })
            .collect(),
        notes,
    })
}
