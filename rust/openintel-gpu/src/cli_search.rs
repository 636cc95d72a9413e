//! The `openintel search` subcommand: clap args for src/cli/args.rs (pattern: `AnalyzeArgs`, src/cli/args.rs:37-57) and
//! the leaf for src/cli/search.rs (pattern: src/cli/run.rs:8-23 — run the use case, render, return the string).
use crate::application_search::{search, SearchReport, SearchRequest};
use crate::domain_stubs::DomainError;
use crate::index_builder::IndexBuilder;
use crate::ports::HybridSearch;

// In src/cli/args.rs, `enum Command` gains:
//     /// Hybrid BM25 + cosine search over stored posts (GPU)
//     Search(SearchArgs),
#[derive(clap::Args, Debug)]
pub struct SearchArgs {
    /// Query text
    pub query: String,

    /// File with the query embedding: little-endian f32 values, index dimension
    #[arg(long)]
    pub embedding: std::path::PathBuf,

    /// Hits to print
    #[arg(long, default_value_t = 10)]
    pub k: usize,

    /// CUDA device ordinal
    #[arg(long, default_value_t = 0)]
    pub device: i32,

    #[arg(long, value_enum, default_value_t = SearchFormat::Table)]
    pub format: SearchFormat,
}

#[derive(clap::ValueEnum, Clone, Copy, Debug, PartialEq, Eq)]
pub enum SearchFormat {
    Table,
    Json,
}

fn read_f32s(path: &std::path::Path) -> Result<Vec<f32>, DomainError> {
    let bytes = std::fs::read(path).map_err(|e| DomainError::SourceFailure { name: "search".into(), message: format!("{}: {e}", path.display()) })?;
    if bytes.len() % 4 != 0 {
        return Err(DomainError::SourceFailure { name: "search".into(), message: "embedding file is not a whole number of f32 values".into() });
    }
    Ok(bytes.chunks_exact(4).map(|c| f32::from_le_bytes([c[0], c[1], c[2], c[3]])).collect())
}

pub fn render(report: &SearchReport, format: SearchFormat) -> String {
    match format {
        SearchFormat::Json => {
            let hits: Vec<String> = report
                .hits
                .iter()
                .map(|h| format!("{{\"post_id\":{:?},\"rrf\":{},\"rank_cosine\":{},\"rank_bm25\":{}}}", h.post_id, h.rrf, h.rank_cosine, h.rank_bm25))
                .collect();
            format!("{{\"hits\":[{}],\"notes\":{:?}}}", hits.join(","), report.notes)
        }
        SearchFormat::Table => {
            let mut out = String::from("rank  rrf        cos  bm25  post\n");
            for (i, h) in report.hits.iter().enumerate() {
                out.push_str(&format!("{:<5} {:<10.6} {:<4} {:<5} {}\n", i + 1, h.rrf, h.rank_cosine, h.rank_bm25, h.post_id));
            }
            for n in &report.notes {
                out.push_str(&format!("note: {n}\n"));
            }
            out
        }
    }
}

/// `cli::search::run` — returns (report, rendered) like `cli::run::analyze` does.
pub async fn run(args: &SearchArgs, index: &IndexBuilder, searcher: &dyn HybridSearch) -> Result<(SearchReport, String), DomainError> {
    let req = SearchRequest { text: args.query.clone(), embedding: read_f32s(&args.embedding)?, k: args.k };
    let report = search(&req, index, searcher).await?;
    let rendered = render(&report, args.format);
    Ok((report, rendered))
}
