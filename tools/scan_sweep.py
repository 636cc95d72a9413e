#!/usr/bin/env python3
"""Tuning experiment (not part of the product): per-query time of the single-query cosine scan for
several shard sizes and ring shapes, device-resident queries, one GPU."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import openintel_b200 as oi
    dev = torch.device("cuda", 0)
    dim, k, nq = 384, 100, 64
    g = torch.Generator().manual_seed(1)
    q = torch.randn(4, nq, dim, generator=g)
    q = (q / q.norm(dim=2, keepdim=True)).to(dev)
    ids = torch.empty(nq, k, dtype=torch.int32, device=dev)
    sc = torch.empty(nq, k, dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    for n in (1_000_000, 500_000, 250_000, 125_000):
        ix = oi.GpuIndex(n_docs=n, dim=dim, max_k=k, max_batch=nq)
        ix.synth_embeddings(20261018)
        for rows, stages in ((0, 0), (32, 3), (32, 2), (64, 2), (32, 4) if False else (32, 3)):
            ix.set_option("cosine_scan_shape", rows * 256 + stages)
            for i in range(3):
                ix.search_cosine_dev(q[i % 4], nq, k, ids, sc, st)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(20):
                ix.search_cosine_dev(q[i % 4], nq, k, ids, sc, st)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 / nq * 1e3
            print(json.dumps({"n_docs": n, "tile_rows": rows, "stages": stages, "us_per_query": round(us, 2),
                              "GBps": round(n * dim * 4 / us / 1e3, 1)}), flush=True)
        ix.set_option("cosine_scan_shape", 0)
        ix.close()


def bf16_single():
    """single-query scan over a bf16 matrix (the path a bf16 index takes below the tensor-core batch threshold)"""
    import torch
    import openintel_b200 as oi
    dev = torch.device("cuda", 0)
    n, dim, k, nq = 5_000_000, 768, 100, 3
    g = torch.Generator().manual_seed(1)
    q = torch.randn(4, nq, dim, generator=g)
    q = (q / q.norm(dim=2, keepdim=True)).to(dev)
    ids = torch.empty(nq, k, dtype=torch.int32, device=dev)
    sc = torch.empty(nq, k, dtype=torch.float32, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    ix = oi.GpuIndex(n_docs=n, dim=dim, dtype=oi.DTYPE_BF16, max_k=k, max_batch=nq)
    ix.synth_embeddings(20261018)
    for i in range(3):
        ix.search_cosine_dev(q[i % 4], nq, k, ids, sc, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(20):
        ix.search_cosine_dev(q[i % 4], nq, k, ids, sc, st)
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 20 / nq * 1e3
    print(json.dumps({"workload": "single-query scan, 5M x 768 bf16", "us_per_query": round(us, 1), "GBps": round(n * dim * 2 / us / 1e3, 1)}), flush=True)
    ix.close()


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "bf16":
        bf16_single()
    else:
        main()
