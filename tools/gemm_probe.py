#!/usr/bin/env python3
"""tools/gemm_probe.py — the batched cosine call (tcgen05 path) alone on one GPU: ms per batch for a few probe sizes.
`--once` runs a handful of calls only (the command ncu wraps for the launch list)."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

import bench
import openintel_b200 as oi

ap = argparse.ArgumentParser()
ap.add_argument("--docs", type=int, default=10_000_000)
ap.add_argument("--batch", type=int, default=256)
ap.add_argument("--once", action="store_true")
ap.add_argument("--lite", type=int, default=0)
args = ap.parse_args()
dev = torch.device("cuda", 0)
B, K = args.batch, bench.TOPK
ix = oi.GpuIndex(n_docs=args.docs, dim=bench.DIM, dtype=oi.DTYPE_BF16, max_k=K, max_batch=B)
ix.synth_embeddings(bench.SEED)
ix.set_option("cosine_gemm_lite", args.lite)
pool = bench._unit_queries(4, B, bench.DIM, 1234).to(dev)
ids = torch.empty(B, K, dtype=torch.int32, device=dev)
sc = torch.empty(B, K, dtype=torch.float32, device=dev)
stream = torch.cuda.current_stream().cuda_stream


def cos(i):
    ix.search_cosine_dev(pool[i % 4], B, K, ids, sc, stream)


if args.once:
    for i in range(3):
        cos(i)
    torch.cuda.synchronize()
    print("ok")
else:
    out = {}
    for pt in (0, 2, 4, 9, 17, 32):
        ix.set_option("cosine_gemm_sample_tiles", pt)
        out["probe_tiles_%d" % pt] = [bench._dev_time(cos, 40, 5) for _ in range(3)]
    print(json.dumps(out))
ix.close()
