//! Port, adapters, use case, MCP tool and CLI subcommand for the hybrid search path, written against the
//! conventions of Kloudy-Sky/openintel:
//!   * ports are `#[async_trait] pub trait X: Send + Sync` returning `Result<_, DomainError>`
//!     (src/domain/ports/post_analyzer.rs:7-11), borrowed slices in, owned Vec out;
//!   * adapter faults map to `DomainError::SourceFailure { name, message }` (src/domain/error.rs:17-18);
//!   * use cases live in src/application/ and take `&dyn Port` (src/application/analyze.rs:16-20);
//!   * adapters are built at the two composition roots (src/main.rs:16-47, src/mcp/server.rs:237-259).
//!
//! UNCOMPILED in the source repository (no Rust toolchain there).  The tested equivalents are
//! openintel_b200/host/openintel_host.hpp (C++) and openintel_b200/capi.py (Python); tests/test_capi_symbols.py
//! checks that every `sys::oi_*` item these files name is declared by include/openintel_gpu.h.
//!
//! Module -> file it becomes inside the reference crate:
//!   ports              src/domain/ports/hybrid_search.rs
//!   adapter            src/adapters/search/gpu.rs            (+ GpuLexiconAnalyzer: src/adapters/analyzer/gpu.rs)
//!   index_builder      src/adapters/search/index_builder.rs
//!   application_search src/application/search.rs
//!   mcp_search         src/mcp/tools.rs (args + run_*) and the #[tool] method of src/mcp/server.rs
//!   cli_search         src/cli/args.rs (subcommand) and src/cli/search.rs (leaf)
//!   wiring             the edits at both composition roots, as compilable functions
pub mod adapter;
pub mod application_search;
pub mod cli_search;
pub mod domain_stubs;
pub mod index_builder;
pub mod mcp_search;
pub mod ports;
pub mod wiring;
