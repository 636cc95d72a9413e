#!/usr/bin/env python3
"""tools/bm25_probe.py — the BM25 call alone on one GPU (10M docs, 1M-term Zipf vocabulary, 8-term queries): ms per
batch, batch 1024 and 256, Zipf and uniform queries (OI_GPU_LIB selects another build of the library for A/B runs).
`--once` = a few calls only (the command ncu wraps)."""
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
ap.add_argument("--docs", type=int, default=10_000_000)
ap.add_argument("--once", action="store_true")
ap.add_argument("--batch", type=int, default=1024)
args = ap.parse_args()
dev = torch.device("cuda", 0)
K = bench.TOPK
ix = oi.GpuIndex(n_docs=args.docs, dim=8, max_k=K, max_batch=1024)
cdf = bench._zipf_cdf(bench.VOCAB)
ix.synth_bm25(bench.SEED, bench.VOCAB, cdf)
ix.bm25_finalize()
stream = torch.cuda.current_stream().cuda_stream
rng = np.random.default_rng(5)


def opt(name, v):
    try:
        ix.set_option(name, v)
        return True
    except Exception:
        return False


def pools(nq, uniform):
    if uniform:
        return [np.stack([rng.choice(bench.VOCAB, 8, replace=False) for _ in range(nq)]).astype(np.uint32) for _ in range(4)]
    return [bench._zipf_queries(nq, 8, cdf, 100 + p) for p in range(4)]


if "OI_IPW" in os.environ:
    opt("bm25_items_per_warp", int(os.environ["OI_IPW"]))
out = {"lib": os.environ.get("OI_GPU_LIB", "in-tree"), "n_docs": args.docs, "ipw": os.environ.get("OI_IPW", "default")}
for nq in ((args.batch,) if (args.once or "OI_IPW" in os.environ) else (1024, 256)):
    for uniform in ((False,) if (args.once or "OI_IPW" in os.environ) else (False, True)):
        ps = pools(nq, uniform)
        d_t = [torch.from_numpy(p.astype(np.int32).reshape(-1)).to(dev) for p in ps]
        d_o = torch.arange(0, nq * 8 + 1, 8, dtype=torch.int32, device=dev)
        ids = torch.empty(nq, K, dtype=torch.int32, device=dev)
        sc = torch.empty(nq, K, dtype=torch.float32, device=dev)

        def bm(i):
            ix.search_bm25_dev(d_t[i % 4], d_o, nq, K, ids, sc, stream)

        if args.once:
            for i in range(3):
                bm(i)
            torch.cuda.synchronize()
            continue
        key = "batch%d_%s" % (nq, "uniform" if uniform else "zipf")
        ms = [bench._dev_time(bm, 10, 3) for _ in range(3)]
        out[key] = {"ms": ms, "queries_per_s_best": nq / (min(ms) * 1e-3)}
print(json.dumps(out))
ix.close()
