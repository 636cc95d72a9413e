"""`python -m openintel_b200.search` — the running twin of the `openintel search` subcommand / `search_posts` MCP tool
whose Rust source is in rust/openintel-gpu/src/{cli_search,mcp_search,application_search}.rs (patterns:
src/cli/args.rs:37-57, src/cli/run.rs:8-23, src/mcp/tools.rs:164-176 of the reference).

Same request (free-text query tokenized like post text + a query embedding of the index dimension + k), same report
(hits = post id, fused score, 1-based rank in each modality or none; notes), same two renderings (table / json), over a
SQLite post store lifted onto the GPU by openintel_b200.store.  The search itself is oi_search_hybrid; there is no CPU
path -- without the CUDA library or a GPU the command fails with the library's error.
"""
import argparse
import json
import sys

import numpy as np

DISCLAIMER = "Informational only; not investment advice."  # the reference appends its disclaimer to every tool output


def search_posts(index, query, embedding, k=10):
    """application::search: validate -> tokenize -> hybrid search -> report (dict with hits / notes / disclaimer).

    index: an open store.StoreIndex; embedding: any scale, L2-normalised by the index."""
    notes = []
    emb = np.asarray(embedding, dtype=np.float32).reshape(-1)
    if emb.shape[0] != index.dim:
        raise ValueError("query embedding has %d components, the index has %d" % (emb.shape[0], index.dim))
    if not np.all(np.isfinite(emb)) or not float(np.dot(emb, emb)) > 0.0:
        raise ValueError("query embedding must be finite and non-zero")
    if k < 1:
        raise ValueError("k must be at least 1")
    if len(index.query_terms([query])[0]) == 0:
        notes.append("no query token is in the index vocabulary: ranking is by embedding similarity only")
    hits = index.search([query], emb[None, :], int(k))[0]
    out = []
    for h in hits:
        out.append({"post_id": h["id"], "rrf": h["rrf"],
                    "rank_cosine": h["rank_cosine"] if h["rank_cosine"] > 0 else None,
                    "rank_bm25": h["rank_bm25"] if h["rank_bm25"] > 0 else None})
    return {"hits": out, "notes": notes, "disclaimer": DISCLAIMER}


def render(report, fmt="table"):
    if fmt == "json":
        hits = [{k: v for k, v in h.items() if v is not None} for h in report["hits"]]
        return json.dumps({"hits": hits, "notes": report["notes"], "disclaimer": report["disclaimer"]})
    lines = ["rank  rrf        cos  bm25  post"]
    for i, h in enumerate(report["hits"]):
        lines.append("%-5d %-10.6f %-4s %-5s %s" % (i + 1, h["rrf"], h["rank_cosine"] or "-", h["rank_bm25"] or "-", h["post_id"]))
    lines += ["note: " + n for n in report["notes"]]
    return "\n".join(lines) + "\n"


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m openintel_b200.search", description="Hybrid BM25 + cosine search over stored posts (GPU)")
    ap.add_argument("query", help="query text")
    ap.add_argument("--store", required=True, help="SQLite post store (openintel_b200.store schema)")
    ap.add_argument("--embedding", required=True, help="file with the query embedding: little-endian f32 values, index dimension")
    ap.add_argument("--k", type=int, default=10, help="hits to print")
    ap.add_argument("--device", type=int, default=0, help="CUDA device ordinal")
    ap.add_argument("--format", choices=["table", "json"], default="table")
    args = ap.parse_args(argv)
    from . import store
    raw = open(args.embedding, "rb").read()
    if len(raw) % 4:
        ap.error("embedding file is not a whole number of f32 values")
    emb = np.frombuffer(raw, dtype="<f4")
    conn = store.open_store(args.store)
    with store.StoreIndex(conn, device=args.device, max_k=max(args.k, 1), max_batch=1) as ix:
        report = search_posts(ix, args.query, emb, args.k)
    sys.stdout.write(render(report, args.format) if args.format == "table" else render(report, "json") + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
