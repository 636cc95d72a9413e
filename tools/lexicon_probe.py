#!/usr/bin/env python3
"""Throughput of the GPU PostAnalyzer (oi_lexicon_analyze, the batched replacement of the reference's
LexiconAnalyzer::analyze, src/adapters/analyzer/lexicon.rs:53-87) on synthetic posts, end to end through the
host-buffer C ABI, next to the CPU oracle's restatement of the reference scorer on a sample of the same posts."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import openintel_b200 as oi
    import oracle as O
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    rng = np.random.RandomState(5)
    words = ["moon", "calls", "puts", "dump", "$AAPL", "tsla", "YOLO", "0dte", "earnings", "beat", "miss", "rocket", "bag", "short",
             "squeeze", "long", "the", "a", "of", "to", "bullish", "bearish", "buy", "sell", "crash", "rally", "gamma", "tendies", "hodl"]
    base = [" ".join(rng.choice(words, size=rng.randint(5, 40))) for _ in range(5000)]
    texts = [base[i % 5000] for i in range(n)]
    nbytes = sum(len(t) for t in texts)
    oi.lexicon_analyze(texts[:1000])
    t0 = time.perf_counter()
    pol, spec, bull, bear = oi.lexicon_analyze(texts)
    gpu_s = time.perf_counter() - t0
    m = 20000
    t0 = time.perf_counter()
    ref = [O.lexicon_score(t) for t in texts[:m]]
    cpu_s = (time.perf_counter() - t0) / m
    same = all(abs(pol[i] - ref[i][0]) == 0 and bool(spec[i]) == bool(ref[i][1]) for i in range(m))
    print(json.dumps({"workload": "GPU PostAnalyzer over %d posts (%.0f MB of text), host buffers in and out, incl. the Python packing" % (n, nbytes / 1e6),
                      "gpu_posts_per_s": n / gpu_s, "gpu_text_MBps": nbytes / gpu_s / 1e6, "gpu_s": gpu_s,
                      "cpu_oracle_posts_per_s": 1.0 / cpu_s, "identical_to_oracle_on_sample": same}), flush=True)


if __name__ == "__main__":
    main()
