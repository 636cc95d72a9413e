"""GPU parity: the batched lexicon scorer (GPU PostAnalyzer) vs the reference's goldens and the
oracle's restatement of src/adapters/analyzer/lexicon.rs:53-73.  Integer hit counts and the f64
polarity are compared exactly."""
import json
import os
import random

import numpy as np
import pytest

import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oi():
    import __graft_entry__ as ge  # noqa: F401
    import openintel_b200
    openintel_b200.load_library()
    return openintel_b200


@pytest.fixture(scope="module")
def G(golden_dir):
    with open(os.path.join(golden_dir, "reference_lexicon_goldens.json"), encoding="utf-8") as f:
        return json.load(f)


def _check(oi, texts):
    pol, spec, bull, bear = oi.lexicon_analyze(texts)
    for i, t in enumerate(texts):
        wp, ws, wb, we = O.lexicon_score(t)
        assert (pol[i], bool(spec[i]), int(bull[i]), int(bear[i])) == (wp, ws, wb, we), repr(t)
    return pol, spec


def test_reference_goldens(oi, G):
    cases = G["lexicon_unit"]["cases"]
    pol, spec = _check(oi, [c["text"] for c in cases])
    for i, c in enumerate(cases):
        assert int(pol[i] > 0) - int(pol[i] < 0) == c["polarity_sign"] and bool(spec[i]) == c["speculative"]
    posts = G["fixture_posts"]["posts"]
    pol, spec = _check(oi, [p["text"] for p in posts])
    assert [float(x) for x in pol] == [p["polarity"] for p in posts]
    assert [bool(x) for x in spec] == [p["speculative"] for p in posts]
    s = O.social_summary(pol, spec.astype(np.int32))
    d = G["fixture_posts"]["summary"]["derived"]
    assert (s["net_sentiment"], s["speculation_index"], s["bullish"], s["bearish"]) == \
        (d["net_sentiment"], d["speculation_index"], d["bullish"], d["bearish"])


def test_tokenizer_edge_cases_and_random_text(oi, G):
    texts = [c["text"] for c in G["tokenizer_rules"]["cases"]]
    texts += ["", " ", "moon", "MOON", "moon moon dump", "$CALLS!", "0dte", "0DTE.", "xmoon", "moonx", "mo on",
              "Kalls calls", "pumpİt", "rİp", "İv iv", "buK", "long" * 3000, "a" * 10000,
              "up " * 3333, "é" * 500 + "bear", "🚀rocket🚀", "iv̇", "bagholder contracts", "contractsx"]
    rnd = random.Random(11)
    alphabet = list("abcXYZ019 $_-.,!\n\t") + ["é", "İ", "K", "ß", "Σ", "日", "🚀", "̇"]
    words = ["moon", "calls", "puts", "yolo", "Dump", "RIP", "iv", "red", "GREEN", "strike", "bear", "up"]
    for _ in range(400):
        parts = [rnd.choice(words) if rnd.random() < 0.4 else "".join(rnd.choice(alphabet) for _ in range(rnd.randint(0, 12)))
                 for _ in range(rnd.randint(0, 30))]
        texts.append(rnd.choice([" ", "", ",", "K", "İ"]).join(parts))
    _check(oi, texts)


def test_large_batch_is_aligned_to_input_order(oi):
    rnd = random.Random(5)
    words = ["moon", "calls", "puts", "sell", "the", "stock", "0dte", "report", "up", "down", "x"]
    texts = [" ".join(rnd.choice(words) for _ in range(rnd.randint(1, 60))) for _ in range(20000)]
    pol, spec, bull, bear = oi.lexicon_analyze(texts)
    assert len(pol) == len(texts)
    for i in range(0, len(texts), 37):
        wp, ws, wb, we = O.lexicon_score(texts[i])
        assert (pol[i], bool(spec[i]), int(bull[i]), int(bear[i])) == (wp, ws, wb, we)


def test_handle_api_and_fused_social_summary(oi, G):
    """round 2: the handle form (persistent stream / buffers) returns what the one-shot call returns, and the fused
    social summary equals the oracle's restatement of SpeculationEngine::social_summary (pinned to the reference's
    fixture values: net 0.5, 7 / 2 / 1, speculation index 0.3, bull/bear 3.5) bit for bit."""
    posts = [p["text"] for p in G["fixture_posts"]["posts"]] if "fixture_posts" in G else None
    rnd = random.Random(11)
    words = ["moon", "calls", "puts", "yolo", "dump", "rally", "the", "AAPL", "0dte", "x", "bagholder", "Ⅻ", "K", "İ"]
    many = [" ".join(rnd.choice(words) for _ in range(rnd.randint(0, 40))) for _ in range(3000)]
    with oi.GpuLexicon(reserve_bytes=64, reserve_posts=2) as lx:       # tiny reservation: the buffers must grow
        for texts in ([t for t in (posts or [])], many, ["", "moon"], many[:7]):
            if not texts:
                continue
            pol, spec, bull, bear, summ = lx.analyze(texts, summary=True)
            p2, s2, b2, e2 = oi.lexicon_analyze(texts)
            assert np.array_equal(pol, p2) and np.array_equal(spec, s2) and np.array_equal(bull, b2) and np.array_equal(bear, e2)
            for i, t in enumerate(texts[:200]):
                assert (pol[i], bool(spec[i]), int(bull[i]), int(bear[i])) == O.lexicon_score(t), repr(t)
            want = O.social_summary(pol, spec.astype(np.int32))
            for f in ("total", "bullish", "bearish", "neutral"):
                assert summ[f] == want[f], f
            for f in ("net_sentiment", "speculation_index", "bull_bear_ratio"):
                assert np.float64(summ[f]).view(np.uint64) == np.float64(want[f]).view(np.uint64), f
        assert lx.launch_count() >= 6
        # a post longer than the shared-memory staging area takes the in-place path
        long_post = " ".join(rnd.choice(words) for _ in range(3000))
        assert len(long_post.encode()) > 4096
        pol, spec, bull, bear = lx.analyze([long_post, "calls"])
        assert (pol[0], bool(spec[0]), int(bull[0]), int(bear[0])) == O.lexicon_score(long_post)
    if posts:
        with oi.GpuLexicon() as lx:
            _, _, _, _, summ = lx.analyze(posts, summary=True)
        assert summ["total"] == 10 and (summ["bullish"], summ["bearish"], summ["neutral"]) == (7, 2, 1)
        assert summ["net_sentiment"] == 0.5 and summ["speculation_index"] == 0.3 and summ["bull_bear_ratio"] == 3.5


def test_analyze_flow_over_a_post_store_matches_the_reference_fixture(tmp_path):
    """The reference's `analyze` use case with the GPU PostAnalyzer injected (openintel_b200/analyze.py): the ten
    fixture posts of the reference's mock sources + its mock market snapshot must give the report the reference's own
    flow test asserts (tests/analyze_flow.rs:128-131: 10 mentions, ConfirmingBullish) and the values derived from its
    fixtures (tests/golden/reference_lexicon_goldens.json)."""
    import json
    import subprocess
    import sys
    import oracle as O
    from openintel_b200 import analyze, capi, fusion, store
    G = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "reference_lexicon_goldens.json")))
    fx = G["fixture_posts"]
    posts = [dict(id=p["id"], source=p["source"], author=p["author"], text=p["text"], created_at=1700000000 + i, engagement=p["engagement"])
             for i, p in enumerate(fx["posts"])]
    db = str(tmp_path / "fixture.db")
    conn = store.open_store(db, dim=4)
    store.insert_posts(conn, posts, np.ones((len(posts), 4), np.float32))
    m = fx["mock_market"]
    mk = fusion.MarketSummary.from_snapshot(m["last_price"], m["previous_close"], m["volume"], m["avg_volume"], m["iv_rank"])
    ids, sources, texts = analyze.select_posts(conn)
    assert ids == [p["id"] for p in fx["posts"]]
    with capi.GpuLexicon() as lx:
        rep, pol, spec = analyze.analyze_posts(lx, sources, texts, mk)
        d = fx["summary"]["derived"]
        assert list(pol) == [p["polarity"] for p in fx["posts"]] and list(spec) == [p["speculative"] for p in fx["posts"]]
        s = rep["social"]
        assert s["total_mentions"] == fx["summary"]["asserted"]["total_mentions"] == 10
        assert (s["bullish"], s["bearish"], s["neutral"]) == (d["bullish"], d["bearish"], d["neutral"])
        assert s["net_sentiment"] == d["net_sentiment"] and s["speculation_index"] == d["speculation_index"]
        assert s["bull_bear_ratio"] == d["bull_bear_ratio"] and dict(s["mentions_by_source"]) == {"bluesky": 6, "reddit": 4}
        assert rep["fusion"]["alignment"] == fx["summary"]["asserted"]["alignment"] == "ConfirmingBullish"
        assert abs(rep["fusion"]["crowding"] - m["derived"]["crowding"]) < 1e-12 and rep["social_confidence"] == "Medium"
        assert rep["fusion"]["crowding"] == O.crowding(10, d["speculation_index"], rvol=mk.rvol, iv=mk.iv_rank)
        # social-only run: Quiet + the reference's note; one source only; a mention filter; nothing at all
        solo, _, _ = analyze.analyze_posts(lx, sources, texts, None)
        assert solo["fusion"]["alignment"] == "Quiet" and solo["fusion"]["notes"] == ["social-only, no price reference"] and solo["market"] is None
        _, src_r, txt_r = analyze.select_posts(conn, source="reddit")
        assert len(txt_r) == 4 and set(src_r) == {"reddit"}
        _, _, txt_m = analyze.select_posts(conn, mention="$aapl")
        assert 0 < len(txt_m) <= 10 and all("aapl" in t.lower() for t in txt_m)
        with pytest.raises(analyze.NoData):
            analyze.analyze_posts(lx, [], [], None)
        empty, _, _ = analyze.analyze_posts(lx, [], [], mk)
        assert empty["social"]["total_mentions"] == 0 and empty["fusion"]["alignment"] == "Quiet" and empty["social_confidence"] == "Low"
    conn.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, "-m", "openintel_b200.analyze", "--store", db, "--market",
                          "%r,%r,%d,%d,%r" % (m["last_price"], m["previous_close"], m["volume"], m["avg_volume"], m["iv_rank"])],
                         cwd=root, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    cli = json.loads(out.stdout)
    assert cli["fusion"]["alignment"] == "ConfirmingBullish" and cli["social"]["total_mentions"] == 10
    assert abs(cli["fusion"]["crowding"] - m["derived"]["crowding"]) < 1e-12 and abs(cli["market"]["pct_change"] - m["derived"]["pct_change"]) < 1e-12
    # the compiled-host twin: host_demo --analyze (C++ PostAnalyzer port + fusion functions of openintel_host.hpp)
    from openintel_b200.host import build as host_build
    demo = host_build.build()[1]
    cpp = subprocess.run([demo, "--analyze", db, repr(m["last_price"]), repr(m["previous_close"]), str(m["volume"]), str(m["avg_volume"]),
                          repr(m["iv_rank"])], capture_output=True, text=True, timeout=120)
    assert cpp.returncode == 0, cpp.stderr[-2000:]
    f = cpp.stdout.split()
    assert f[0] == "report" and [int(x) for x in f[1:5]] == [10, d["bullish"], d["bearish"], d["neutral"]]
    assert float(f[5]) == d["net_sentiment"] and float(f[6]) == d["speculation_index"]
    assert abs(float(f[7]) - m["derived"]["crowding"]) < 1e-12 and f[8] == "ConfirmingBullish" and f[9] == "Medium"
    solo = subprocess.run([demo, "--analyze", db], capture_output=True, text=True, timeout=120)
    assert solo.returncode == 0 and solo.stdout.split()[8] == "Quiet"
