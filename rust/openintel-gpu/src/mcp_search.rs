//! The `search_posts` MCP tool: args + `run_search_posts` for src/mcp/tools.rs (pattern: `ScanArgs` / `run_scan`,
//! src/mcp/tools.rs:164-200) and the `#[tool]` method for src/mcp/server.rs (pattern: src/mcp/server.rs:82-94).
use crate::application_search::{search, PostHit, SearchRequest};
use crate::index_builder::IndexBuilder;
use crate::ports::HybridSearch;
use schemars::JsonSchema;
use serde::{Deserialize, Serialize};

#[derive(Debug, Deserialize, JsonSchema)]
pub struct SearchPostsArgs {
    /// Free-text query; tokenized like post text (lower-cased, split on non-alphanumerics).
    pub query: String,
    /// Query embedding, same dimension as the index (any scale: it is L2-normalised).
    pub embedding: Vec<f32>,
    /// Hits to return (default 10, at most the index's max_k).
    pub k: Option<usize>,
}

#[derive(Debug, Serialize)]
pub struct SearchHitOut {
    pub post_id: String,
    pub rrf: f32,
    /// 1-based rank in the cosine list, absent when the post is not in its top-k
    #[serde(skip_serializing_if = "Option::is_none")]
    pub rank_cosine: Option<u32>,
    #[serde(skip_serializing_if = "Option::is_none")]
    pub rank_bm25: Option<u32>,
}

#[derive(Debug, Serialize)]
pub struct SearchPostsOutput {
    pub hits: Vec<SearchHitOut>,
    pub notes: Vec<String>,
    pub disclaimer: &'static str,
}

const DISCLAIMER: &str = "Read-only retrieval over stored posts. Not investment advice.";

fn rank(r: u32) -> Option<u32> {
    if r == 0 { None } else { Some(r) }
}

pub async fn run_search_posts(args: SearchPostsArgs, index: &IndexBuilder, searcher: &dyn HybridSearch) -> Result<SearchPostsOutput, crate::domain_stubs::DomainError> {
    let req = SearchRequest { text: args.query, embedding: args.embedding, k: args.k.unwrap_or(10) };
    let report = search(&req, index, searcher).await?;
    Ok(SearchPostsOutput {
        hits: report
            .hits
            .into_iter()
            .map(|PostHit { post_id, rrf, rank_cosine, rank_bm25, .. }| SearchHitOut { post_id, rrf, rank_cosine: rank(rank_cosine), rank_bm25: rank(rank_bm25) })
            .collect(),
        notes: report.notes,
        disclaimer: DISCLAIMER,
    })
}

// The method to add inside `#[tool_router] impl OpenIntelServer` (src/mcp/server.rs:40-213); the server struct gains
// `search: Option<Arc<SearchDeps>>` built in `serve()` (see wiring.rs).  Kept as text because the macro needs the
// reference's server type:
//
//     #[tool(
//         description = "Hybrid search over stored posts: BM25 + cosine similarity fused with reciprocal \
//                        rank fusion (GPU). Returns post ids with their fused score and per-modality ranks. \
//                        Read-only."
//     )]
//     async fn search_posts(
//         &self,
//         Parameters(args): Parameters<tools::SearchPostsArgs>,
//     ) -> Result<CallToolResult, ErrorData> {
//         let deps = self.search.as_ref().ok_or_else(|| ErrorData::internal_error("search index not configured", None))?;
//         let out = tools::run_search_posts(args, &deps.index, deps.searcher.as_ref())
//             .await
//             .map_err(|e| ErrorData::internal_error(e.to_string(), None))?;
//         let json = serde_json::to_string_pretty(&out).map_err(|e| ErrorData::internal_error(e.to_string(), None))?;
//         Ok(CallToolResult::success(vec![ContentBlock::text(json)]))
//     }
