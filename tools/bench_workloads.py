#!/usr/bin/env python3
"""Secondary workloads of BASELINE.json (the headline bench is bench.py): device-timed throughput
of the BM25 kernel at config 3 scale, the hybrid path and the batched cosine path, one JSON line
per workload.  Single GPU.  Results are copied into profiles/ by hand.

    python tools/bench_workloads.py bm25 [--docs 10000000 --vocab 1000000 --batch 1024]
    python tools/bench_workloads.py hybrid [--docs 1000000 --dim 384 --batch 16]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
SEED = 20261018


def _time(fn, steps, warmup):
    import torch
    for i in range(warmup):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(steps):
        fn(warmup + i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def bench_bm25(a):
    import torch
    import openintel_b200 as oi
    import oracle as O
    dev = torch.device("cuda", 0)
    cdf = O.zipf_cdf(a.vocab)
    t0 = time.perf_counter()
    ix = oi.GpuIndex(n_docs=a.docs, dim=8, max_k=a.k, max_batch=a.batch)
    ix.synth_bm25(SEED, a.vocab, cdf)
    ix.bm25_finalize()
    build_s = time.perf_counter() - t0
    df, sdl, npost = ix.bm25_local_stats()
    if a.groups:
        ix.set_option("bm25_variant", a.groups)
    if a.warps:
        ix.set_option("bm25_warps", a.warps)
    if a.block_docs:
        ix.set_option("bm25_block_docs", a.block_docs)
    if a.slots >= 0:
        ix.set_option("bm25_stage_slots", a.slots)
    stream = torch.cuda.current_stream().cuda_stream
    for name, uniform in (("zipf", False), ("uniform", True))[:1 if a.zipf_only else 2]:
        n_pool = 4
        pools = [O.synth_query_terms(a.batch, 8, cdf, uniform=uniform, first=p * a.batch) for p in range(n_pool)]
        touched = float(np.mean([df[p].astype(np.int64).sum(axis=1).mean() for p in pools]))
        d_terms = [torch.from_numpy(p.astype(np.int32).reshape(-1)).to(dev) for p in pools]
        d_offs = torch.arange(0, a.batch * 8 + 1, 8, dtype=torch.int32, device=dev)
        d_ids = torch.empty(a.batch, a.k, dtype=torch.int32, device=dev)
        d_sc = torch.empty(a.batch, a.k, dtype=torch.float32, device=dev)
        l0 = ix.launch_count()
        ms = _time(lambda i: ix.search_bm25_dev(d_terms[i % n_pool], d_offs, a.batch, a.k, d_ids, d_sc, stream), a.steps, a.warmup)
        launches = ix.launch_count() - l0
        # e2e through the host-buffer call
        t0 = time.perf_counter()
        for i in range(a.steps):
            ix.search_bm25(pools[i % n_pool], a.k)
        e2e_ms = (time.perf_counter() - t0) / a.steps * 1e3
        qps = a.batch / (ms * 1e-3)
        print(json.dumps({
            "workload": "configs[2]: BM25, %d docs, %d-term Zipf vocab, 8-term %s queries, batch %d, top-%d" % (a.docs, a.vocab, name, a.batch, a.k),
            "value": qps, "unit": "queries/s", "ms_per_batch": ms, "e2e_queries_per_s": a.batch / (e2e_ms * 1e-3),
            "postings": int(npost), "postings_touched_per_query": touched,
            "algorithmic_GBps": touched * 8 * qps / 1e9, "index_build_s": build_s, "launches_per_batch": launches / (a.steps + a.warmup),
            "groups_override": a.groups}), flush=True)
    if a.cpu_queries:
        # the self-written CPU oracle on the same index (read back from the GPU), one thread, a handful of queries:
        # a baseline next to the GPU number, not a reference implementation (the reference has none)
        t0 = time.perf_counter()
        csr = ix.read_bm25(int(npost))
        w = O.bm25_weights(csr["term_offsets"], csr["doc_ids"], csr["tfs"], csr["doc_len"], O.bm25_idf(a.docs, df))
        prep_s = time.perf_counter() - t0
        for name, uniform in (("zipf", False), ("uniform", True)):
            qs = O.synth_query_terms(a.cpu_queries, 8, cdf, uniform=uniform)
            got_ids, got_sc = ix.search_bm25(qs, a.k)
            t0 = time.perf_counter()
            same = 0
            for j in range(a.cpu_queries):
                sc = O.bm25_score_dense(csr["term_offsets"], csr["doc_ids"], w, qs[j], a.docs)
                wi, ws, _ = O.topk_f32(sc, a.k, only_positive=True)
                same += int(np.array_equal(wi, got_ids[j]) and np.array_equal(ws.view(np.uint32), got_sc[j].view(np.uint32)))
            cpu_s = (time.perf_counter() - t0) / a.cpu_queries
            print(json.dumps({"cpu_oracle": "BM25 %s queries, %d docs, 1 thread (self-written oracle; no reference implementation exists)" % (name, a.docs),
                              "cpu_queries_per_s": 1.0 / cpu_s, "cpu_ms_per_query": cpu_s * 1e3, "queries": a.cpu_queries,
                              "bit_identical_to_gpu_lists": "%d / %d" % (same, a.cpu_queries), "oracle_weight_prep_s": prep_s}), flush=True)
    ix.close()


def bench_hybrid(a):
    import torch
    import openintel_b200 as oi
    import oracle as O
    dev = torch.device("cuda", 0)
    cdf = O.zipf_cdf(a.vocab)
    ix = oi.GpuIndex(n_docs=a.docs, dim=a.dim, max_k=a.k, max_batch=a.batch)
    ix.synth_embeddings(SEED)
    ix.synth_bm25(SEED, a.vocab, cdf)
    ix.bm25_finalize()
    g = torch.Generator().manual_seed(7)
    qv = torch.randn(4, a.batch, a.dim, generator=g)
    qv = (qv / qv.norm(dim=2, keepdim=True)).to(dev)
    qt = [torch.from_numpy(O.synth_query_terms(a.batch, 8, cdf, first=p * a.batch).astype(np.int32).reshape(-1)).to(dev) for p in range(4)]
    d_offs = torch.arange(0, a.batch * 8 + 1, 8, dtype=torch.int32, device=dev)
    outs = [torch.empty(a.batch, a.k, dtype=torch.int32, device=dev) for _ in range(3)]
    d_rrf = torch.empty(a.batch, a.k, dtype=torch.float32, device=dev)
    stream = torch.cuda.current_stream().cuda_stream
    ms = _time(lambda i: ix.search_hybrid_dev(qv[i % 4], qt[i % 4], d_offs, a.batch, a.k, 60, outs[0], d_rrf, outs[1], outs[2], stream), a.steps, a.warmup)
    ms_cos = _time(lambda i: ix.search_cosine_dev(qv[i % 4], a.batch, a.k, outs[0], d_rrf, stream), a.steps, a.warmup)
    ms_bm = _time(lambda i: ix.search_bm25_dev(qt[i % 4], d_offs, a.batch, a.k, outs[0], d_rrf, stream), a.steps, a.warmup)
    print(json.dumps({"workload": "hybrid BM25+cosine+RRF top-%d, %d docs x %d f32, vocab %d, batch %d" % (a.k, a.docs, a.dim, a.vocab, a.batch),
                      "value": a.batch / (ms * 1e-3), "unit": "queries/s", "ms_per_batch": ms, "ms_cosine_only": ms_cos, "ms_bm25_only": ms_bm}), flush=True)
    ix.close()


def bench_gemm(a):
    """configs[3]: 10M docs x 768 bf16, batch 256, cosine top-100 on the tcgen05 path"""
    import torch
    import openintel_b200 as oi
    dev = torch.device("cuda", 0)
    ix = oi.GpuIndex(n_docs=a.docs, dim=a.dim, dtype=oi.DTYPE_BF16, max_k=a.k, max_batch=a.batch)
    ix.synth_embeddings(SEED)
    if a.debug:
        ix.set_option("cosine_gemm_debug", a.debug)
    g = torch.Generator().manual_seed(7)
    qv = torch.randn(4, a.batch, a.dim, generator=g)
    qv = (qv / qv.norm(dim=2, keepdim=True))
    d_qv = qv.to(dev)
    h_qv = qv.pin_memory()
    d_ids = torch.empty(a.batch, a.k, dtype=torch.int32, device=dev)
    d_sc = torch.empty(a.batch, a.k, dtype=torch.float32, device=dev)
    h_ids = torch.empty(a.batch, a.k, dtype=torch.int32).pin_memory()
    h_sc = torch.empty(a.batch, a.k, dtype=torch.float32).pin_memory()
    stream = torch.cuda.current_stream().cuda_stream
    l0 = ix.launch_count()
    from bench import ClockSampler
    cs = ClockSampler(0)
    cs.start()
    ms = _time(lambda i: ix.search_cosine_dev(d_qv[i % 4], a.batch, a.k, d_ids, d_sc, stream), a.steps, a.warmup)
    clocks = cs.stop()
    launches = (ix.launch_count() - l0) / (a.steps + a.warmup)
    t0 = time.perf_counter()
    for i in range(a.steps):
        ix.search_cosine(h_qv[i % 4], a.k, h_ids, h_sc)
    e2e_ms = (time.perf_counter() - t0) / a.steps * 1e3
    bytes_ = a.docs * a.dim * 2
    flops = 2.0 * a.docs * a.dim * a.batch
    sc = h_sc.numpy()
    assert a.debug or np.all(np.diff(sc, axis=1) <= 0)
    print(json.dumps({"debug": a.debug, "workload": "configs[3]: %d docs x %d bf16, batch %d, cosine top-%d (tcgen05 GEMM path)" % (a.docs, a.dim, a.batch, a.k),
                      "value": a.batch / (ms * 1e-3), "unit": "queries/s", "ms_per_batch": ms, "e2e_queries_per_s": a.batch / (e2e_ms * 1e-3),
                      "hbm_GBps": bytes_ / (ms * 1e-3) / 1e9, "tensor_TFLOPs": flops / (ms * 1e-3) / 1e12,
                      "launches_per_batch": launches, "clocks": clocks}), flush=True)
    ix.close()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("workload", choices=["bm25", "hybrid", "gemm"])
    ap.add_argument("--docs", type=int, default=None)
    ap.add_argument("--vocab", type=int, default=1000000)
    ap.add_argument("--dim", type=int, default=384)
    ap.add_argument("--batch", type=int, default=None)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--groups", type=int, default=0)
    ap.add_argument("--warps", type=int, default=0, help="bm25_warps tuning option")
    ap.add_argument("--block-docs", type=int, default=0, help="bm25_block_docs tuning option")
    ap.add_argument("--slots", type=int, default=-1, help="bm25_stage_slots tuning option")
    ap.add_argument("--zipf-only", action="store_true")
    ap.add_argument("--cpu-queries", type=int, default=0, help="bm25: also time this many queries on the CPU oracle (1 thread)")
    ap.add_argument("--debug", type=int, default=0)
    a = ap.parse_args()
    if a.workload == "gemm":
        a.docs = a.docs or 10_000_000
        a.batch = a.batch or 256
        if a.dim == 384:
            a.dim = 768
        bench_gemm(a)
    elif a.workload == "bm25":
        a.docs = a.docs or 10_000_000
        a.batch = a.batch or 1024
        bench_bm25(a)
    else:
        a.docs = a.docs or 1_000_000
        a.batch = a.batch or 16
        bench_hybrid(a)


if __name__ == "__main__":
    main()
