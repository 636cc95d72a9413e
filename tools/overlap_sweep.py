#!/usr/bin/env python3
"""tools/overlap_sweep.py — one GPU: the hybrid call's two legs alone, back to back, and overlapped (co-resident lite
kernels / SM partition) on a 10M x 768 bf16 corpus with its BM25 index, batch 256; plus GEMM variants (probe size,
lite ring).  Every overlapped result is compared bit for bit with the serial one.  Prints one JSON object."""
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
import openintel_b200 as oi


class Power:
    def __init__(self):
        self.rows = []
        self.p = subprocess.Popen(["nvidia-smi", "-i", "0", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits", "-lms", "50"],
                                  stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        self.t = threading.Thread(target=self._read, daemon=True)
        self.t.start()

    def _read(self):
        for line in self.p.stdout:
            try:
                a, b = line.split(",")
                self.rows.append((time.perf_counter(), float(a), float(b)))
            except Exception:
                pass

    def window(self, t0, t1):
        r = [x for x in self.rows if t0 <= x[0] <= t1]
        if not r:
            return None
        return {"sm_mhz": statistics.median(x[1] for x in r), "power_w": statistics.median(x[2] for x in r), "samples": len(r)}

    def stop(self):
        self.p.terminate()


def timed(fn, steps, warm, power=None, min_s=0.0):
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    reps = steps
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    n = 0
    while True:
        for i in range(reps):
            fn(i)
        n += reps
        torch.cuda.synchronize()
        if time.perf_counter() - t0 >= min_s:
            break
    e1.record()
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    out = {"ms": e0.elapsed_time(e1) / n}
    if power is not None:
        out.update(power.window(t0 + 0.1, t1) or {})
    return out


def main():
    n = int(os.environ.get("SWEEP_DOCS", "10000000"))
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    B, K = bench.BATCH, bench.TOPK
    ix = oi.GpuIndex(n_docs=n, dim=bench.DIM, dtype=oi.DTYPE_BF16, max_k=K, max_batch=B)
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
    pw = Power()
    res = {"n_docs": n, "batch": B}

    def cos(i):
        ix.search_cosine_dev(pool[i % 4], B, K, o[0], rrf, stream)

    def bm(i):
        ix.search_bm25_dev(terms[i % 4], offs, B, K, o[0], rrf, stream)

    def hyb(i):
        ix.search_hybrid_dev(pool[i % 4], terms[i % 4], offs, B, K, bench.RRF_K, o[0], rrf, o[1], o[2], stream)

    def snapshot():
        hyb(0)
        torch.cuda.synchronize()
        return [t.clone() for t in (o[0], rrf, o[1], o[2])]

    # --- legs alone (>= 1.5 s each so that nvidia-smi sees the steady state)
    res["cosine_full_ring"] = timed(cos, 10, 3, pw, 1.5)
    res["bm25_20warps"] = timed(bm, 10, 3, pw, 1.5)
    ix.set_option("cosine_gemm_lite", 1)
    res["cosine_lite_ring"] = timed(cos, 10, 3, pw, 1.0)
    ix.set_option("cosine_gemm_lite", 0)
    for pt in (4, 8, 32, 56):
        ix.set_option("cosine_gemm_sample_tiles", pt)
        res["cosine_probe_tiles_%d" % pt] = timed(cos, 10, 3)
    ix.set_option("cosine_gemm_sample_tiles", 0)
    for w in (8, 10, 12):
        ix.set_option("bm25_warps", w)
        ix.set_option("bm25_stage_slots", 1)
        res["bm25_%dwarps_alone" % w] = timed(bm, 5, 2)
    ix.set_option("bm25_warps", 0)
    ix.set_option("bm25_stage_slots", -1)
    # --- hybrid
    ix.set_option("hybrid_overlap", 0)
    want = snapshot()
    res["hybrid_serial"] = timed(hyb, 10, 3, pw, 1.5)
    for w in (8, 10, 12):
        for slots in (1, 2):
            ix.set_option("hybrid_overlap", 1)
            ix.set_option("overlap_bm25_warps", w)
            ix.set_option("overlap_bm25_slots", slots)
            try:
                got = snapshot()
                same = all(torch.equal(a, b) for a, b in zip(want, got))
                r = timed(hyb, 10, 3, pw, 1.0)
                r["bit_identical_to_serial"] = same
                res["hybrid_coresident_w%d_s%d" % (w, slots)] = r
            except Exception as e:
                res["hybrid_coresident_w%d_s%d" % (w, slots)] = {"error": str(e)[:200]}
    for sms in (60, 74, 90, 104, 118):
        ix.set_option("hybrid_overlap", 2)
        ix.set_option("overlap_gemm_sms", sms)
        try:
            got = snapshot()
            same = all(torch.equal(a, b) for a, b in zip(want, got))
            r = timed(hyb, 10, 3, pw, 1.0)
            r["bit_identical_to_serial"] = same
            res["hybrid_partition_gemm%d" % sms] = r
        except Exception as e:
            res["hybrid_partition_gemm%d" % sms] = {"error": str(e)[:200]}
    pw.stop()
    ix.close()
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
