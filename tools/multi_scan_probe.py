#!/usr/bin/env python3
"""Device-timed multi-query cosine scan (one matrix pass per group of 4 queries) against the single-query path:
1M x 384 f32, batches of 4 / 16 / 64 queries in one call."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import openintel_b200 as oi
    n, dim, k = 1_000_000, 384, 100
    dev = torch.device("cuda", 0)
    ix = oi.GpuIndex(n_docs=n, dim=dim, max_k=k, max_batch=64)
    ix.synth_embeddings(20261018)
    g = torch.Generator().manual_seed(7)
    qv = torch.randn(4, 64, dim, generator=g)
    qv = (qv / qv.norm(dim=2, keepdim=True)).to(dev)
    ids = torch.empty(64, k, dtype=torch.int32, device=dev)
    sc = torch.empty(64, k, dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream

    def t(nb, steps=20, warm=3):
        for i in range(warm):
            ix.search_cosine_dev(qv[i % 4], nb, k, ids, sc, stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(steps):
            ix.search_cosine_dev(qv[i % 4], nb, k, ids, sc, stream)
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / steps

    for multi in (2, 1, 0):
        ix.set_option("cosine_multi_query", multi)
        for nb in (4, 16, 64):
            ms = t(nb)
            passes = (nb + 3) // 4 if multi else nb
            print(json.dumps({"multi_query": multi, "batch": nb, "ms_per_batch": ms, "queries_per_s": nb / ms * 1e3,
                              "us_per_matrix_pass": ms * 1e3 / passes, "hbm_gbs": passes * n * dim * 4 / (ms * 1e-3) / 1e9}), flush=True)
    ix.close()


if __name__ == "__main__":
    main()
