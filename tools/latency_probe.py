#!/usr/bin/env python3
"""Single-query latencies through the host-buffer C ABI (one query per call): cosine, BM25, hybrid."""
import json, os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def main():
    import openintel_b200 as oi
    import oracle as O
    n, dim, vocab, k = 1_000_000, 384, 1_000_000, 100
    cdf = O.zipf_cdf(vocab)
    ix = oi.GpuIndex(n_docs=n, dim=dim, max_k=k, max_batch=4)
    ix.synth_embeddings(20261018)
    ix.synth_bm25(20261018, vocab, cdf)
    ix.bm25_finalize()
    q = O.synth_rows_f32(64, dim, stream=1)
    qt = O.synth_query_terms(64, 8, cdf)
    def t(fn, n=200):
        for i in range(10): fn(i)
        t0 = time.perf_counter()
        for i in range(n): fn(i)
        return (time.perf_counter() - t0) / n * 1e6
    res = {"cosine_us": t(lambda i: ix.search_cosine(q[i % 64:i % 64 + 1], k)),
           "bm25_us": t(lambda i: ix.search_bm25(qt[i % 64:i % 64 + 1], k)),
           "hybrid_us": t(lambda i: ix.search_hybrid(q[i % 64:i % 64 + 1], qt[i % 64:i % 64 + 1], k))}
    print(json.dumps({"workload": "single query per call, host buffers, 1M docs x 384 f32 + 1M-term Zipf BM25, top-100", **res}))
    ix.close()

if __name__ == "__main__":
    main()
