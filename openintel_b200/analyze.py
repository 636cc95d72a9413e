"""`python -m openintel_b200.analyze` — the reference's `analyze` use case (src/application/analyze.rs:16-73) over the posts
of a SQLite post store, with the GPU PostAnalyzer in the seat of `LexiconAnalyzer`:

    posts (store) -> GpuLexicon: per-post (polarity, speculative) + the social summary, one device call
                  -> fusion signals on the host (fusion.py: crowding, alignment, confidence)
                  -> a report shaped like the reference's SpeculationReport (src/domain/entities/speculation_report.rs)

The reference fetches a ticker's posts from its HTTP sources; here they are already stored, so the request selects them
(all, one source, or those that mention a cashtag / word).  An optional market snapshot (last price, previous close,
volume, average volume, IV rank) feeds the fusion exactly as `SpeculationEngine::aggregate` does.  No CPU analyzer path.
"""
import argparse
import json
import sys
from collections import OrderedDict

from . import capi, fusion

DISCLAIMER = "Informational only; not investment advice."


class NoData(ValueError):
    """DomainError::NoData (src/application/analyze.rs:58-60): neither posts nor a market snapshot"""


def select_posts(conn, source=None, mention=None, limit=None):
    """(ids, sources, texts) of the stored posts in doc order; `mention`: keep posts one of whose tokens equals it
    (case-insensitive, a leading '$' ignored: the reference's tokenizer drops it too)"""
    sql, args = "SELECT id, source, text FROM posts", []
    if source:
        sql += " WHERE source = ?"
        args.append(source)
    sql += " ORDER BY doc_id"
    rows = conn.execute(sql, args).fetchall()
    if mention:
        from . import hostlib
        want = mention.lstrip("$").lower()
        rows = [r for r in rows if want in hostlib.tokenize(r[2])]
    if limit is not None:
        rows = rows[:limit]
    return [r[0] for r in rows], [r[1] for r in rows], [r[2] for r in rows]


def analyze_posts(lexicon, sources, texts, market=None, cfg=fusion.EngineConfig()):
    """SpeculationEngine::aggregate (src/domain/engine/speculation_engine.rs:21-66) with the signals and the social
    summary coming from ONE GPU call.  market: fusion.MarketSummary or None."""
    if len(sources) != len(texts):
        raise ValueError("analyzer mismatch: %d sources for %d posts" % (len(sources), len(texts)))
    if not texts and market is None:
        raise NoData("no posts and no market snapshot")
    notes = []
    if texts:
        pol, spec, _, _, sm = lexicon.analyze(texts, summary=True, threshold=cfg.bull_bear_threshold)
    else:
        pol, spec = [], []
        sm = dict(total=0, bullish=0, bearish=0, neutral=0, net_sentiment=0.0, speculation_index=0.0, bull_bear_ratio=-1.0)
    by_source = OrderedDict()
    for s in sorted(set(sources)):  # the reference keeps a BTreeMap: sorted keys
        by_source[s] = sources.count(s)
    if market is None:
        notes.append("social-only, no price reference")
    social = {"total_mentions": int(sm["total"]), "mentions_by_source": by_source, "net_sentiment": sm["net_sentiment"],
              "bullish": int(sm["bullish"]), "bearish": int(sm["bearish"]), "neutral": int(sm["neutral"]),
              "bull_bear_ratio": None if sm["bull_bear_ratio"] < 0 else sm["bull_bear_ratio"],
              "speculation_index": sm["speculation_index"]}

    class _S:
        total, net_sentiment, speculation_index = social["total_mentions"], social["net_sentiment"], social["speculation_index"]
    fz = fusion.fuse(_S, market, cfg)
    report = {"social": social,
              "market": None if market is None else {"pct_change": market.pct_change, "rvol": market.rvol, "iv_rank": market.iv_rank},
              "fusion": {"alignment": fz["alignment"], "crowding": fz["crowding"], "notes": notes},
              "social_confidence": fz["social_confidence"], "disclaimer": DISCLAIMER}
    return report, pol, spec


def main(argv=None):
    ap = argparse.ArgumentParser(prog="python -m openintel_b200.analyze", description="speculation report over stored posts (GPU PostAnalyzer)")
    ap.add_argument("--store", required=True, help="SQLite post store (openintel_b200.store schema)")
    ap.add_argument("--source", choices=["reddit", "bluesky"], help="only this source's posts")
    ap.add_argument("--mention", help="only posts that mention this cashtag / word (e.g. GME)")
    ap.add_argument("--limit", type=int, help="at most this many posts")
    ap.add_argument("--market", help="last_price,previous_close,volume,avg_volume[,iv_rank]")
    ap.add_argument("--device", type=int, default=0)
    args = ap.parse_args(argv)
    from . import store
    market = None
    if args.market:
        f = args.market.split(",")
        if len(f) not in (4, 5):
            ap.error("--market takes last_price,previous_close,volume,avg_volume[,iv_rank]")
        market = fusion.MarketSummary.from_snapshot(float(f[0]), float(f[1]), int(f[2]), int(f[3]), float(f[4]) if len(f) == 5 else None)
    conn = store.open_store(args.store)
    _, sources, texts = select_posts(conn, args.source, args.mention, args.limit)
    with capi.GpuLexicon(args.device) as lx:
        try:
            report, _, _ = analyze_posts(lx, sources, texts, market)
        except NoData as e:
            sys.stderr.write("error: %s\n" % e)
            return 1
    sys.stdout.write(json.dumps(report) + "\n")
    return 0


if __name__ == "__main__":
    sys.exit(main())
