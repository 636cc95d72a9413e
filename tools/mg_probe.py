#!/usr/bin/env python3
"""Per-rank timing probe for the sharded hybrid path (run under torchrun): times the cosine and BM25 legs on every
rank with and without the all-gather, and samples each GPU's SM clock while the BM25 leg runs.  Rank 0 prints the
per-rank table; used to separate kernel time, exchange time and rank skew at 8 GPUs."""
import argparse
import json
import os
import subprocess
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SEED = 20261018


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=100_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--vocab", type=int, default=1_000_000)
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--steps", type=int, default=10)
    a = ap.parse_args()
    import torch
    import torch.distributed as dist
    import openintel_b200 as oi
    from openintel_b200 import sharding
    import oracle as O
    rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(lr)
    dev = torch.device("cuda", lr)
    dist.init_process_group("nccl", device_id=dev)
    cdf = O.zipf_cdf(a.vocab)
    sh = sharding.ShardedIndex(a.docs, a.dim, dtype=oi.DTYPE_BF16, dist=dist, device_index=lr, max_k=a.k, max_batch=a.batch)
    sh.ix.synth_embeddings(SEED)
    sh.ix.synth_bm25(SEED, a.vocab, cdf)
    sh.finalize_bm25()
    ix = sh.ix
    g = torch.Generator().manual_seed(7)
    qv = torch.randn(4, a.batch, a.dim, generator=g)
    qv = (qv / qv.norm(dim=2, keepdim=True)).to(dev)
    qt = [torch.from_numpy(O.synth_query_terms(a.batch, 8, cdf, first=p * a.batch).astype(np.int32).reshape(-1)).to(dev) for p in range(4)]
    offs = torch.arange(0, a.batch * 8 + 1, 8, dtype=torch.int32, device=dev)
    o = torch.empty(a.batch, a.k, dtype=torch.int32, device=dev)
    sc = torch.empty(a.batch, a.k, dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def timed(fn):
        for i in range(3):
            fn(i)
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(a.steps):
            fn(3 + i)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / a.steps

    def clock():
        try:
            out = subprocess.run(["nvidia-smi", "-i", str(lr), "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits"],
                                 capture_output=True, text=True, timeout=10).stdout.strip()
            return out
        except Exception as e:
            return str(e)

    res = {}
    for skip in (0, 1):
        ix.set_option("comm_debug_skip_gather", skip)
        res["bm25_skip%d" % skip] = timed(lambda i: ix.search_bm25_dev(qt[i % 4], offs, a.batch, a.k, o, sc, stream))
        res["cos_skip%d" % skip] = timed(lambda i: ix.search_cosine_dev(qv[i % 4], a.batch, a.k, o, sc, stream))
    ix.set_option("comm_debug_skip_gather", 1)
    # clocks while the BM25 leg runs back to back
    for i in range(30):
        ix.search_bm25_dev(qt[i % 4], offs, a.batch, a.k, o, sc, stream)
    res["clock_power_bm25"] = clock()
    torch.cuda.synchronize()
    for i in range(30):
        ix.search_cosine_dev(qv[i % 4], a.batch, a.k, o, sc, stream)
    res["clock_power_cos"] = clock()
    torch.cuda.synchronize()
    ix.set_option("comm_debug_skip_gather", 0)
    allres = [None] * world
    dist.all_gather_object(allres, res)
    if rank == 0:
        for r, x in enumerate(allres):
            print(json.dumps({"rank": r, **x}), flush=True)
    sh.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
