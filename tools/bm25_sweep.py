#!/usr/bin/env python3
"""tools/bm25_sweep.py — tuning sweep of the BM25 call on one GPU: one index (--docs), batch 256 Zipf queries, a list of
option sets (`name=val,name=val;...`), ms per batch for each.  Options are the oi_index_set_option tuning knobs
(bm25_items_per_warp, bm25_warps, bm25_block_docs, bm25_stage_slots)."""
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
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--sets", type=str, default="")
args = ap.parse_args()
dev = torch.device("cuda", 0)
K = bench.TOPK
nq = args.batch
ix = oi.GpuIndex(n_docs=args.docs, dim=8, max_k=K, max_batch=nq)
cdf = bench._zipf_cdf(bench.VOCAB)
ix.synth_bm25(bench.SEED, bench.VOCAB, cdf)
ix.bm25_finalize()
stream = torch.cuda.current_stream().cuda_stream
ps = [bench._zipf_queries(nq, 8, cdf, 100 + p) for p in range(4)]
d_t = [torch.from_numpy(p.astype(np.int32).reshape(-1)).to(dev) for p in ps]
d_o = torch.arange(0, nq * 8 + 1, 8, dtype=torch.int32, device=dev)
ids = torch.empty(nq, K, dtype=torch.int32, device=dev)
sc = torch.empty(nq, K, dtype=torch.float32, device=dev)
DEFAULTS = {"bm25_items_per_warp": 0, "bm25_warps": 0, "bm25_block_docs": 0, "bm25_stage_slots": -1}


def bm(i):
    ix.search_bm25_dev(d_t[i % 4], d_o, nq, K, ids, sc, stream)


ref = None
for spec in [""] + [s for s in args.sets.split(";") if s]:
    opts = dict(DEFAULTS)
    for kv in spec.split(","):
        if kv:
            k_, v_ = kv.split("=")
            opts[k_] = int(v_)
    try:
        for k_, v_ in opts.items():
            ix.set_option(k_, v_)
        bm(0)
        torch.cuda.synchronize()
        got = ids.cpu().numpy().copy()
        if ref is None:
            ref = got
        same = bool(np.array_equal(ref, got))
        ms = [bench._dev_time(bm, 10, 3) for _ in range(3)]
        print(json.dumps({"docs": args.docs, "batch": nq, "opts": spec or "default", "ms": [round(m, 4) for m in ms], "same_ids": same}), flush=True)
    except Exception as e:  # an option set that does not fit an SM is reported, not fatal
        print(json.dumps({"opts": spec, "error": str(e)[:200]}), flush=True)
ix.close()
