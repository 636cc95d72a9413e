"""Pins the oracle's tokenizer / lexicon / engine-summary restatements to the reference's own
goldens.  This is the only reference-pinned parity in the repo."""
import json
import os

import pytest

import oracle as O


@pytest.fixture(scope="module")
def G(golden_dir):
    with open(os.path.join(golden_dir, "reference_lexicon_goldens.json"), encoding="utf-8") as f:
        return json.load(f)


def test_lexicon_unit_cases(G):
    for c in G["lexicon_unit"]["cases"]:
        pol, spec, _, _ = O.lexicon_score(c["text"])
        assert (pol > 0) - (pol < 0) == c["polarity_sign"], c
        assert spec == c["speculative"], c


def test_fixture_posts_per_post_and_summary(G):
    fx = G["fixture_posts"]
    pols, specs = [], []
    for p in fx["posts"]:
        pol, spec, _, _ = O.lexicon_score(p["text"])
        assert pol == p["polarity"], p["id"]
        assert spec == p["speculative"], p["id"]
        pols.append(pol)
        specs.append(spec)
    s = O.social_summary(pols, specs)
    d = fx["summary"]["derived"]
    assert s["total"] == fx["summary"]["asserted"]["total_mentions"] == 10
    assert (s["bullish"], s["bearish"], s["neutral"]) == (d["bullish"], d["bearish"], d["neutral"])
    assert s["net_sentiment"] == d["net_sentiment"]
    assert s["speculation_index"] == d["speculation_index"]
    assert s["bull_bear_ratio"] == d["bull_bear_ratio"]
    m = fx["mock_market"]
    pct = (m["last_price"] - m["previous_close"]) / m["previous_close"] * 100.0
    rvol = m["volume"] / m["avg_volume"]
    assert abs(pct - m["derived"]["pct_change"]) < 1e-12
    assert abs(rvol - m["derived"]["rvol"]) < 1e-12
    cr = O.crowding(s["total"], s["speculation_index"], rvol=rvol, iv=m["iv_rank"])
    assert abs(cr - m["derived"]["crowding"]) < 1e-12
    assert O.alignment(True, s["total"], s["net_sentiment"], pct) == fx["summary"]["asserted"]["alignment"]
    # single-source runs (src/application/analyze.rs:135, tests/analyze_flow.rs:144): 4 reddit / 6 bluesky
    assert sum(p["source"] == "reddit" for p in fx["posts"]) == 4
    assert sum(p["source"] == "bluesky" for p in fx["posts"]) == 6


def test_engine_known_answers(G):
    E = G["engine_known_answers"]
    for c in E["crowding"]:
        got = O.crowding(c["total"], c["spec_index"], rvol=c["rvol"], iv=c["iv"])
        assert abs(got - c["expect"]) < 1e-9, c
    b = E["bullish_batch"]
    s = O.social_summary(b["polarity"], [1] * 9 + [0] * 3)
    assert s["bullish"] == b["bullish"]
    assert O.alignment(True, s["total"], s["net_sentiment"], b["pct_up"]) == b["alignment_up"]
    assert O.alignment(True, s["total"], s["net_sentiment"], b["pct_down"]) == b["alignment_down"]
    assert O.alignment(False, s["total"], s["net_sentiment"], b["pct_up"]) == "Quiet"
    assert O.alignment(True, 9, 0.6, 10.0) == "Quiet"  # below min_sample


def test_whole_word_matching(G):
    W = G["whole_word"]
    hits = []
    for h in W["headlines"]:
        for t in O.tokenize(h):
            if t in W["keywords_probe"] and t not in hits:
                hits.append(t)
    assert hits == W["expect_hits"]
    assert "miss" not in O.tokenize("dismissal of claims")


def test_tokenizer_rules(G):
    for c in G["tokenizer_rules"]["cases"]:
        assert O.tokenize(c["text"]) == c["tokens"], c


def test_tokenizer_matches_python_model_on_random_text():
    """Independent model of the same rule: str.lower() agrees with Rust's to_lowercase on every
    char whose lowercase touches ASCII (K -> k, U+0130 -> i + U+0307)."""
    import random
    import re
    rnd = random.Random(7)
    alphabet = list("abcXYZ019 $_-.,!\n\t") + ["é", "İ", "K", "ß", "Σ", "日", "🚀", "̇"]
    for _ in range(300):
        s = "".join(rnd.choice(alphabet) for _ in range(rnd.randint(0, 40)))
        want = [t for t in re.split(r"[^a-z0-9]", s.lower()) if t]
        assert O.tokenize(s) == want, repr(s)
