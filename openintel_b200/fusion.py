"""Host-side fusion signals on top of the GPU PostAnalyzer's social summary.

`GpuLexicon.run(..., want_summary=True)` returns the reference's `SocialSummary` (computed on the device, see
csrc/lexicon.cu); what `SpeculationEngine::aggregate` derives from it next is O(1) scalar arithmetic on that summary and a
market snapshot -- crowding, alignment, confidence (src/domain/engine/speculation_engine.rs:43-66, 151-208;
src/domain/values/speculation.rs:29-42; defaults: src/domain/engine/config.rs:18-33).  It lives here, on the host, written
against the reference's definitions; nothing in this module touches the GPU or the CPU oracle.
"""
from dataclasses import dataclass
from typing import Optional

ALIGNMENTS = ("ConfirmingBullish", "ConfirmingBearish", "Diverging", "Quiet")
CONFIDENCES = ("Low", "Medium", "High")


@dataclass(frozen=True)
class EngineConfig:
    """src/domain/engine/config.rs:2-33 (same names, same defaults)"""
    bull_bear_threshold: float = 0.2
    net_sentiment_threshold: float = 0.05
    price_move_threshold: float = 1.0
    crowding_weight_spec: float = 0.5
    crowding_weight_rvol: float = 0.3
    crowding_weight_iv: float = 0.2
    rvol_cap: float = 3.0
    min_sample: int = 10
    confidence_low: int = 10
    confidence_high: int = 50


@dataclass(frozen=True)
class MarketSummary:
    """the fields of the reference's MarketSummary the fusion reads (speculation_engine.rs:127-148): percent change of
    the last price against the previous close, relative volume (None without an average volume), IV rank in [0, 1]"""
    pct_change: float
    rvol: Optional[float] = None
    iv_rank: Optional[float] = None

    @staticmethod
    def from_snapshot(last_price, previous_close, volume, avg_volume, iv_rank=None):
        """speculation_engine.rs:127-148: a zero previous close gives pct_change 0, a zero average volume no rvol"""
        pct = 0.0 if previous_close == 0.0 else (last_price - previous_close) / previous_close * 100.0
        rvol = None if avg_volume == 0 else float(volume) / float(avg_volume)
        return MarketSummary(pct, rvol, iv_rank)


def _clamp01(x):
    return 0.0 if x < 0.0 else (1.0 if x > 1.0 else x)


def crowding(total_mentions, speculation_index, market: Optional[MarketSummary] = None, cfg: EngineConfig = EngineConfig()):
    """weighted blend of the components that are present, renormalised over their weights (speculation_engine.rs:151-176)"""
    weighted = weight_sum = 0.0
    if total_mentions > 0:
        weighted += cfg.crowding_weight_spec * speculation_index
        weight_sum += cfg.crowding_weight_spec
    if market is not None:
        if market.rvol is not None:
            weighted += cfg.crowding_weight_rvol * _clamp01(market.rvol / cfg.rvol_cap)
            weight_sum += cfg.crowding_weight_rvol
        if market.iv_rank is not None:
            weighted += cfg.crowding_weight_iv * _clamp01(market.iv_rank)
            weight_sum += cfg.crowding_weight_iv
    return 0.0 if weight_sum == 0.0 else _clamp01(weighted / weight_sum)


def alignment(total_mentions, net_sentiment, market: Optional[MarketSummary] = None, cfg: EngineConfig = EngineConfig()):
    """does the crowd agree with the tape (speculation_engine.rs:178-208)"""
    if market is None or total_mentions < cfg.min_sample:
        return "Quiet"
    s, p = net_sentiment, market.pct_change
    if not (abs(s) >= cfg.net_sentiment_threshold) or not (abs(p) >= cfg.price_move_threshold):
        return "Quiet"
    if s > 0.0 and p > 0.0:
        return "ConfirmingBullish"
    if not (s > 0.0) and not (p > 0.0):
        return "ConfirmingBearish"
    return "Diverging"


def confidence(total_mentions, cfg: EngineConfig = EngineConfig()):
    """sample-size bucket; reversed thresholds are normalised first (values/speculation.rs:29-42)"""
    low, high = min(cfg.confidence_low, cfg.confidence_high), max(cfg.confidence_low, cfg.confidence_high)
    return "Low" if total_mentions < low else ("Medium" if total_mentions < high else "High")


def fuse(summary, market: Optional[MarketSummary] = None, cfg: EngineConfig = EngineConfig()):
    """summary: the SocialSummary of GpuLexicon (any object with total / net_sentiment / speculation_index) ->
    the FusionSignals + confidence of the reference's report (speculation_engine.rs:43-66)"""
    return {"crowding": crowding(summary.total, summary.speculation_index, market, cfg),
            "alignment": alignment(summary.total, summary.net_sentiment, market, cfg),
            "social_confidence": confidence(summary.total, cfg)}
