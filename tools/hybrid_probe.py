#!/usr/bin/env python3
"""tools/hybrid_probe.py — the hybrid call alone on one GPU at a given shard size (default: the 6.25M-document shard the
50M-document headline corpus leaves on each of 8 GPUs): ms per batch for the step and for each leg, BM25 first (an
issue-bound kernel timed right after the power-capped tensor-core leg looks 30 % slower than it is).  `--once` = a few
calls only (the command ncu wraps for the per-kernel launch list)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
import openintel_b200 as oi

ap = argparse.ArgumentParser()
ap.add_argument("--docs", type=int, default=6_250_000)
ap.add_argument("--once", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda", 0)
B, K = bench.BATCH, bench.TOPK
ix = oi.GpuIndex(n_docs=args.docs, dim=bench.DIM, dtype=oi.DTYPE_BF16, max_k=K, max_batch=B)
ix.synth_embeddings(bench.SEED)
cdf = bench._zipf_cdf(bench.VOCAB)
ix.synth_bm25(bench.SEED, bench.VOCAB, cdf)
ix.bm25_finalize()
pool = bench._unit_queries(4, B, bench.DIM, 1234).to(dev)
terms = [torch.from_numpy(bench._zipf_queries(B, bench.QTERMS, cdf, 100 + p).astype(np.int32).reshape(-1)).to(dev) for p in range(4)]
offs = torch.arange(0, B * bench.QTERMS + 1, bench.QTERMS, dtype=torch.int32, device=dev)
o = [torch.empty(B, K, dtype=torch.int32, device=dev) for _ in range(3)]
rrf = torch.empty(B, K, dtype=torch.float32, device=dev)
stream = torch.cuda.current_stream().cuda_stream


def hyb(i):
    ix.search_hybrid_dev(pool[i % 4], terms[i % 4], offs, B, K, bench.RRF_K, o[0], rrf, o[1], o[2], stream)


if args.once:
    for i in range(3):
        hyb(i)
    torch.cuda.synchronize()
    print("ok")
else:
    out = {"n_docs": args.docs}
    out["bm25_ms"] = [bench._dev_time(lambda i: ix.search_bm25_dev(terms[i % 4], offs, B, K, o[0], rrf, stream), 20, 3) for _ in range(2)]
    out["cosine_ms"] = [bench._dev_time(lambda i: ix.search_cosine_dev(pool[i % 4], B, K, o[0], rrf, stream), 20, 3) for _ in range(2)]
    out["hybrid_ms"] = [bench._dev_time(hyb, 40, 5) for _ in range(3)]
    print(json.dumps(out))
ix.close()
