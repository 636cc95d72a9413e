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
ap.add_argument("--debug", type=int, default=0)
args = ap.parse_args()
dev = torch.device("cuda", 0)
B, K = args.batch, bench.TOPK
ix = oi.GpuIndex(n_docs=args.docs, dim=bench.DIM, dtype=oi.DTYPE_BF16, max_k=K, max_batch=B)
ix.synth_embeddings(bench.SEED)


def opt(name, v):
    try:
        ix.set_option(name, v)
        return True
    except Exception:
        return False   # an older build of the library (A/B runs with OI_GPU_LIB)


opt("cosine_gemm_lite", args.lite)
if "OI_PAIR" in os.environ:
    opt("cosine_gemm_pair", int(os.environ["OI_PAIR"]))
if "OI_PAIR_RING" in os.environ:
    opt("cosine_gemm_pair_ring", int(os.environ["OI_PAIR_RING"]))
pool = bench._unit_queries(4, B, bench.DIM, 1234).to(dev)
ids = torch.empty(B, K, dtype=torch.int32, device=dev)
sc = torch.empty(B, K, dtype=torch.float32, device=dev)
stream = torch.cuda.current_stream().cuda_stream


def cos(i):
    ix.search_cosine_dev(pool[i % 4], B, K, ids, sc, stream)


if args.once:
    opt("cosine_gemm_debug", args.debug)
    for i in range(3):
        cos(i)
    torch.cuda.synchronize()
    print("ok")
else:
    out = {"lib": os.environ.get("OI_GPU_LIB", "in-tree")}
    out["default"] = [bench._dev_time(cos, 40, 5) for _ in range(3)]
    for dbg in (4, 8, 1, 2):
        if opt("cosine_gemm_debug", dbg):
            out["debug_%d" % dbg] = [bench._dev_time(cos, 40, 5) for _ in range(2)]
    opt("cosine_gemm_debug", 0)
    for pt in (4, 32):
        if opt("cosine_gemm_sample_tiles", pt):
            out["probe_tiles_%d" % pt] = [bench._dev_time(cos, 40, 5) for _ in range(2)]
    print(json.dumps(out))
ix.close()
