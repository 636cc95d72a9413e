"""Host-side fusion signals (openintel_b200/fusion.py and the C++ twin in host/openintel_host.hpp) against the values
the reference's own tests hold (tests/golden/reference_lexicon_goldens.json: src/domain/engine/speculation_engine.rs
test module, tests/test_fixtures.rs) and against the CPU oracle's restatement of the same functions."""
import ctypes as C
import json
import os

import pytest

import oracle as O
from openintel_b200 import fusion as F

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def G():
    return json.load(open(os.path.join(HERE, "golden", "reference_lexicon_goldens.json")))


@pytest.fixture(scope="module")
def host():
    from openintel_b200.host import build
    so = build.build()[0]
    L = C.CDLL(so)
    L.oih_crowding.restype = C.c_double
    L.oih_crowding.argtypes = [C.c_uint64, C.c_double, C.c_int, C.c_int, C.c_double, C.c_int, C.c_double]
    L.oih_alignment.restype = C.c_int
    L.oih_alignment.argtypes = [C.c_uint64, C.c_double, C.c_int, C.c_double]
    L.oih_confidence.restype = C.c_int
    L.oih_confidence.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64]
    return L


def _cpp_crowding(L, total, si, m):
    has_m = m is not None
    rv = m.rvol if has_m and m.rvol is not None else 0.0
    iv = m.iv_rank if has_m and m.iv_rank is not None else 0.0
    return L.oih_crowding(total, si, int(has_m), int(has_m and m.rvol is not None), rv, int(has_m and m.iv_rank is not None), iv)


def _cpp_alignment(L, total, net, m):
    return F.ALIGNMENTS[L.oih_alignment(total, net, int(m is not None), m.pct_change if m is not None else 0.0)]


def test_fixture_report_matches_the_reference(G, host):
    fx = G["fixture_posts"]
    d, m = fx["summary"]["derived"], fx["mock_market"]
    mk = F.MarketSummary.from_snapshot(m["last_price"], m["previous_close"], m["volume"], m["avg_volume"], m["iv_rank"])
    assert abs(mk.pct_change - m["derived"]["pct_change"]) < 1e-12 and abs(mk.rvol - m["derived"]["rvol"]) < 1e-12
    total = fx["summary"]["asserted"]["total_mentions"]
    for cr in (F.crowding(total, d["speculation_index"], mk), _cpp_crowding(host, total, d["speculation_index"], mk)):
        assert abs(cr - m["derived"]["crowding"]) < 1e-12
    want = fx["summary"]["asserted"]["alignment"]
    assert F.alignment(total, d["net_sentiment"], mk) == want == _cpp_alignment(host, total, d["net_sentiment"], mk)

    class S:  # the shape of GpuLexicon's SocialSummary
        pass
    s = S()
    s.total, s.net_sentiment, s.speculation_index = total, d["net_sentiment"], d["speculation_index"]
    out = F.fuse(s, mk)
    assert out["alignment"] == want and abs(out["crowding"] - m["derived"]["crowding"]) < 1e-12
    assert out["social_confidence"] == "Medium"  # 10 mentions: low <= n < high


def test_engine_known_answers(G, host):
    E = G["engine_known_answers"]
    for c in E["crowding"]:
        mk = F.MarketSummary(0.0, c["rvol"], c["iv"]) if (c["rvol"] is not None or c["iv"] is not None) else None
        for got in (F.crowding(c["total"], c["spec_index"], mk), _cpp_crowding(host, c["total"], c["spec_index"], mk)):
            assert abs(got - c["expect"]) < 1e-9, c
        assert F.crowding(c["total"], c["spec_index"], mk) == O.crowding(c["total"], c["spec_index"], rvol=c["rvol"], iv=c["iv"])
    b = E["bullish_batch"]
    s = O.social_summary(b["polarity"], [1] * 9 + [0] * 3)
    for pct, want in ((b["pct_up"], b["alignment_up"]), (b["pct_down"], b["alignment_down"])):
        mk = F.MarketSummary(pct)
        assert F.alignment(s["total"], s["net_sentiment"], mk) == want == _cpp_alignment(host, s["total"], s["net_sentiment"], mk)
    assert F.alignment(s["total"], s["net_sentiment"], None) == "Quiet" == _cpp_alignment(host, s["total"], s["net_sentiment"], None)
    assert F.alignment(9, 0.6, F.MarketSummary(10.0)) == "Quiet" == _cpp_alignment(host, 9, 0.6, F.MarketSummary(10.0))


def test_alignment_and_confidence_edges(host):
    cfg = F.EngineConfig()
    # thresholds are inclusive (>=), zero sentiment / zero move are "not positive"
    assert F.alignment(10, 0.05, F.MarketSummary(1.0)) == "ConfirmingBullish"
    assert F.alignment(10, -0.05, F.MarketSummary(-1.0)) == "ConfirmingBearish"
    assert F.alignment(10, 0.05, F.MarketSummary(-1.0)) == "Diverging"
    assert F.alignment(10, 0.049, F.MarketSummary(5.0)) == "Quiet" and F.alignment(10, 0.5, F.MarketSummary(0.99)) == "Quiet"
    assert F.alignment(10, float("nan"), F.MarketSummary(5.0)) == "Quiet"
    for n, net, pct in ((10, 0.05, 1.0), (10, -0.05, -1.0), (10, 0.05, -1.0), (10, 0.049, 5.0), (9, 0.9, 9.0), (50, -0.3, 2.5)):
        assert F.alignment(n, net, F.MarketSummary(pct)) == _cpp_alignment(host, n, net, F.MarketSummary(pct))
        assert F.alignment(n, net, F.MarketSummary(pct)) == O.alignment(True, n, net, pct)
    # values/speculation.rs tests: confidence_buckets, reversed_thresholds_match_ordered
    assert [F.confidence(n) for n in (5, 10, 49, 50)] == ["Low", "Medium", "Medium", "High"]
    assert F.confidence(30, F.EngineConfig(confidence_low=50, confidence_high=10)) == F.confidence(30)
    for n in (0, 5, 10, 49, 50, 1000):
        assert F.CONFIDENCES[host.oih_confidence(n, cfg.confidence_low, cfg.confidence_high)] == F.confidence(n)
    assert F.CONFIDENCES[host.oih_confidence(30, 50, 10)] == "Medium"
    # crowding: nothing present -> 0; components clamp
    assert F.crowding(0, 0.9, None) == 0.0 == _cpp_crowding(host, 0, 0.9, None)
    assert F.crowding(0, 0.0, F.MarketSummary(0.0, rvol=30.0)) == 1.0
    assert F.MarketSummary.from_snapshot(10.0, 0.0, 5, 0).pct_change == 0.0 and F.MarketSummary.from_snapshot(10.0, 0.0, 5, 0).rvol is None
