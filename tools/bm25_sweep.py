"""Tuning sweep of the BM25 scoring kernel (BASELINE configs[2] shape): builds the synthetic index once and times
oi_search_bm25_dev for a list of (warps per CTA, documents per block, staged chunks per warp) settings.
Usage: python tools/bm25_sweep.py [--docs N] [--batch B] [--configs "16:1024:8,24:1024:0,..."]"""
import argparse
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
SEED = 20261018


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=10000000)
    ap.add_argument("--vocab", type=int, default=1000000)
    ap.add_argument("--batch", type=int, default=1024)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--steps", type=int, default=4)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--dense-div", type=int, default=0, help="bm25_dense_div (0 = library default)")
    ap.add_argument("--configs", default="16:1024:8,16:1024:0,24:1024:0,20:1024:4,18:1024:6,16:1024:4,16:2048:2,12:1024:8")
    a = ap.parse_args()
    import torch
    import openintel_b200 as oi
    import oracle as O
    dev = torch.device("cuda", 0)
    cdf = O.zipf_cdf(a.vocab)
    ix = oi.GpuIndex(n_docs=a.docs, dim=8, max_k=a.k, max_batch=a.batch)
    ix.synth_bm25(SEED, a.vocab, cdf)
    if a.dense_div:
        ix.set_option("bm25_dense_div", a.dense_div)
    ix.bm25_finalize()
    stream = torch.cuda.current_stream().cuda_stream
    pools = [O.synth_query_terms(a.batch, 8, cdf, first=p * a.batch) for p in range(4)]
    d_terms = [torch.from_numpy(p.astype(np.int32).reshape(-1)).to(dev) for p in pools]
    d_offs = torch.arange(0, a.batch * 8 + 1, 8, dtype=torch.int32, device=dev)
    d_ids = torch.empty(a.batch, a.k, dtype=torch.int32, device=dev)
    d_sc = torch.empty(a.batch, a.k, dtype=torch.float32, device=dev)
    ref = None
    for cfg in a.configs.split(","):
        f = [int(x) for x in cfg.split(":")]
        w, r, s = f[:3]
        ipw = f[3] if len(f) > 3 else 0
        ix.set_option("bm25_items_per_warp", ipw)
        ix.set_option("bm25_warps", w)
        ix.set_option("bm25_block_docs", r)
        ix.set_option("bm25_stage_slots", s)
        try:
            for i in range(a.warmup):
                ix.search_bm25_dev(d_terms[i % 4], d_offs, a.batch, a.k, d_ids, d_sc, stream)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(a.steps):
                ix.search_bm25_dev(d_terms[i % 4], d_offs, a.batch, a.k, d_ids, d_sc, stream)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / a.steps
            # same result whatever the tuning (last step used pool (steps - 1) % 4)
            ix.search_bm25_dev(d_terms[0], d_offs, a.batch, a.k, d_ids, d_sc, stream)
            torch.cuda.synchronize()
            got = (d_ids.cpu().numpy().copy(), d_sc.cpu().numpy().copy())
            if ref is None:
                ref = got
            same = bool(np.array_equal(ref[0], got[0]) and np.array_equal(ref[1].view(np.uint32), got[1].view(np.uint32)))
            print(json.dumps({"dense_div": a.dense_div, "warps": w, "block_docs": r, "slots": s, "items_per_warp": ipw, "ms_per_batch": ms, "queries_per_s": a.batch / ms * 1e3, "same_as_first": same}), flush=True)
        except Exception as e:  # a setting that does not fit one SM
            print(json.dumps({"warps": w, "block_docs": r, "slots": s, "error": str(e)}), flush=True)
    ix.close()


if __name__ == "__main__":
    main()
