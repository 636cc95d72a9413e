//! The edits at the reference's two composition roots, written out as functions so they read as code rather than prose.
//!
//! 1. src/application/analyze.rs:16-20 and :62-63 — the analyzer stops being hard-wired.  Today:
//!        let analyzer = LexiconAnalyzer::new();
//!        let signals = analyzer.analyze(&posts).await?;
//!    becomes a parameter, so that `GpuLexiconAnalyzer` (adapter.rs) can be injected:
//!        pub async fn analyze(req: &AnalysisRequest, social_sources: &[Box<dyn SocialDataSource>],
//!                             market_source: Option<&dyn MarketDataSource>, analyzer: &dyn PostAnalyzer) -> ...
//!        let signals = analyzer.analyze(&posts).await?;
//!    and the callers (src/cli/run.rs:8-23, src/mcp/tools.rs `run_analyze`, src/application/dip.rs `sentiment_for`)
//!    pass `&LexiconAnalyzer::new()` or the GPU adapter built at the root.
//!
//! 2. src/main.rs:16-47 — the CLI root builds the adapters and matches the new subcommand (`build_search_deps` +
//!    `main_search_arm` below).
//!
//! 3. src/mcp/server.rs:14-38 and :237-259 — `OpenIntelServer` gains `search: Option<Arc<SearchDeps>>`, `serve()` builds
//!    it once next to the other adapters (`serve_search_deps` below) and the `search_posts` tool (mcp_search.rs) uses it.
use crate::adapter::{GpuHybridSearch, GpuLexiconAnalyzer};
use crate::cli_search::SearchArgs;
use crate::domain_stubs::{DomainError, PostAnalyzer, SocialPost};
use crate::index_builder::IndexBuilder;
use crate::ports::HybridSearch;
use std::sync::Arc;

/// What both roots hold: the vocabulary / post-id side of the index (host) and the GPU adapter behind the port.
pub struct SearchDeps {
    pub index: IndexBuilder,
    pub searcher: Box<dyn HybridSearch>,
}

/// Lift stored posts + their embeddings into the GPU index (the store itself is outside the reference today:
/// SURVEY.md §0; openintel_b200/store.py is the tested SQLite stand-in with the `SocialPost` schema).
pub fn build_search_deps(device: i32, dim: usize, posts: &[SocialPost], embeddings: &[f32], max_k: u32, max_batch: u32) -> Result<SearchDeps, DomainError> {
    let mut index = IndexBuilder::default();
    for p in posts {
        index.add(&p.id, &p.text);
    }
    let searcher = GpuHybridSearch::new(device, dim, embeddings, &mut index, max_k, max_batch)?;
    Ok(SearchDeps { index, searcher: Box::new(searcher) })
}

/// The analyzer both roots inject into `application::analyze` once edit 1 is made.
pub fn build_post_analyzer(device: Option<i32>) -> Box<dyn PostAnalyzer> {
    match device {
        Some(device) => Box::new(GpuLexiconAnalyzer { device }),
        None => unimplemented!("Box::new(LexiconAnalyzer::new()) inside the reference crate"),
    }
}

/// The arm `Command::Search(args) => ...` of src/main.rs (same outcome handling as the Analyze arm, :31-42).
pub async fn main_search_arm(args: SearchArgs, deps: &SearchDeps) -> std::process::ExitCode {
    match crate::cli_search::run(&args, &deps.index, deps.searcher.as_ref()).await {
        Ok((_report, rendered)) => {
            println!("{rendered}");
            std::process::ExitCode::SUCCESS
        }
        Err(e) => {
            eprintln!("error: {e}");
            std::process::ExitCode::FAILURE
        }
    }
}

/// What `serve()` adds before `OpenIntelServer::new(...)` (src/mcp/server.rs:237-259): a failed GPU index disables the
/// tool with a warning instead of failing the server — the convention used for the optional X pulse feed (:245-254).
pub fn serve_search_deps(device: i32, dim: usize, posts: &[SocialPost], embeddings: &[f32]) -> Option<Arc<SearchDeps>> {
    match build_search_deps(device, dim, posts, embeddings, 100, 64) {
        Ok(d) => Some(Arc::new(d)),
        Err(e) => {
            eprintln!("warning: search_posts disabled: {e}");
            None
        }
    }
}
