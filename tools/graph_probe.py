#!/usr/bin/env python3
"""tools/graph_probe.py — the hybrid `_dev` call captured in a CUDA graph against plain stream launches, one GPU
(SURVEY §5 names graph capture as the alternative to measure for launch-bound small steps): ms per call for a few shard
sizes / batch sizes, and a check that the replayed graph returns the same lists."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
import openintel_b200 as oi

dev = torch.device("cuda", 0)
K = bench.TOPK
cdf = bench._zipf_cdf(bench.VOCAB)
out = []
for docs, dim, dtype, B in ((6_250_000, 768, oi.DTYPE_BF16, 256), (1_000_000, 768, oi.DTYPE_BF16, 16), (1_000_000, 384, oi.DTYPE_F32, 1),
                            (200_000, 384, oi.DTYPE_F32, 1)):
    ix = oi.GpuIndex(n_docs=docs, dim=dim, dtype=dtype, max_k=K, max_batch=B)
    ix.synth_embeddings(bench.SEED)
    ix.synth_bm25(bench.SEED, bench.VOCAB, cdf)
    ix.bm25_finalize()
    q = bench._unit_queries(1, B, dim, 1234).to(dev)[0].contiguous()
    terms = torch.from_numpy(bench._zipf_queries(B, bench.QTERMS, cdf, 100).astype(np.int32).reshape(-1)).to(dev)
    offs = torch.arange(0, B * bench.QTERMS + 1, bench.QTERMS, dtype=torch.int32, device=dev)
    o = [torch.empty(B, K, dtype=torch.int32, device=dev) for _ in range(3)]
    rrf = torch.empty(B, K, dtype=torch.float32, device=dev)
    s = torch.cuda.Stream()

    def call():
        ix.search_hybrid_dev(q, terms, offs, B, K, bench.RRF_K, o[0], rrf, o[1], o[2], s.cuda_stream)

    with torch.cuda.stream(s):
        for _ in range(3):
            call()
        s.synchronize()
        want = (o[0].clone(), rrf.clone())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            call()
        o[0].zero_()
        g.replay()
        s.synchronize()
        same = bool(torch.equal(o[0], want[0]) and torch.equal(rrf, want[1]))

        def timed(fn, n):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            for _ in range(5):
                fn()
            e0.record(s)
            for _ in range(n):
                fn()
            e1.record(s)
            e1.synchronize()
            return e0.elapsed_time(e1) / n

        n = 200 if docs <= 1_000_000 else 40
        r = {"docs": docs, "dim": dim, "batch": B, "graph_equals_stream": same,
             "stream_ms": [round(timed(call, n), 4) for _ in range(3)], "graph_ms": [round(timed(g.replay, n), 4) for _ in range(3)]}
    print(json.dumps(r), flush=True)
    out.append(r)
    del g
    ix.close()
