//! Stand-ins for `crate::domain::*` items of the reference so that this crate is self-contained.
//! Delete this module when the files are vendored into the reference crate and import the real ones:
//!   DomainError   src/domain/error.rs:3-22
//!   SocialPost    src/domain/entities/social_post.rs:30-38
//!   PostSignal    src/domain/values/post_signal.rs:3-7
//!   PostAnalyzer  src/domain/ports/post_analyzer.rs:7-11
use async_trait::async_trait;

#[derive(Debug)]
pub enum DomainError {
    SourceFailure { name: String, message: String },
    AnalyzerMismatch { expected: usize, got: usize },
    NoData,
}
impl std::fmt::Display for DomainError {
    fn fmt(&self, f: &mut std::fmt::Formatter<'_>) -> std::fmt::Result {
        match self {
            DomainError::SourceFailure { name, message } => write!(f, "data source '{name}' failed: {message}"),
            DomainError::AnalyzerMismatch { expected, got } => write!(f, "analyzer returned {got} signals for {expected} posts"),
            DomainError::NoData => write!(f, "no data"),
        }
    }
}
impl std::error::Error for DomainError {}

#[derive(Clone, Debug)]
pub struct SocialPost {
    pub id: String,
    pub text: String,
}
pub struct PostSignal {
    pub polarity: f64,
    pub speculative: bool,
}
#[async_trait]
pub trait PostAnalyzer: Send + Sync {
    async fn analyze(&self, posts: &[SocialPost]) -> Result<Vec<PostSignal>, DomainError>;
}
