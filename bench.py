#!/usr/bin/env python3
"""bench.py — the driver's benchmark contract for the hybrid-retrieval hot path.

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...   # CPU baseline arm (rank 0 only)

Workload (BASELINE.json configs[1]): 1,000,000 docs x 384-dim f32 unit rows, cosine top-100,
single-query GEMV path.  A *step* submits QUERIES_PER_STEP independent single-query scans in one
C-ABI call (each query is its own full pass over the matrix; one persistent launch walks them back
to back).  `latency_us_nq1` is the same path called with ONE query per call.
N > 1: the same fixed corpus is sharded by document over the N GPUs (strong scaling); every rank
scans its shard, local top-k lists are all-gathered with NCCL and merged on device.

Prints ONE JSON line on rank 0.  `value` = device-timed queries/s with queries resident in HBM;
`e2e` = the same through the host-buffer C-ABI call (pinned host queries in, host results out);
`roofline` = algorithmic bytes of one scan launch / its average duration vs the measured HBM
copy bandwidth; `cpu_baseline` = the self-written CPU oracle (no reference implementation of
this path exists: SURVEY.md §0) timed on this box's host cores.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_DOCS, DIM, TOPK = 1_000_000, 384, 100
QUERIES_PER_STEP = 64
SEED = 20261018
METRIC = "cosine top-100 queries/sec (single-query GEMV path, 1M x 384 f32)"
UNIT = "queries/s"
WORKLOAD = "configs[1]: 1M docs x 384-dim f32, single query, cosine top-100 (GEMV bandwidth path)"


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, copy)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.rows, self.proc, self.gpu = [], None, gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for i, nme in enumerate(names):
                    if r[5 + i].lower().startswith("active"):
                        reasons.add(nme)
            except Exception:
                continue
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def _cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def _synth_rows_parallel(O, n, dim, threads):
    """oracle rows, generated in parallel chunks (ctypes releases the GIL)."""
    import numpy as np
    out = np.empty((n, dim), dtype=np.float32)
    chunk = (n + threads - 1) // threads

    def work(t):
        lo, hi = t * chunk, min(n, (t + 1) * chunk)
        if lo < hi:
            out[lo:hi] = O.synth_rows_f32(hi - lo, dim, seed=SEED, first=lo)
    ts = [threading.Thread(target=work, args=(t,)) for t in range(threads)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    return out


def _traffic(key):
    """dram__bytes_read+write per launch of the dominant kernel, from the committed ncu --set full capture"""
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))[key]
        # the capture was one launch of `queries_per_launch` queries; a bench launch walks QUERIES_PER_STEP
        return t["traffic_bytes_per_launch"] / t["queries_per_launch"] * QUERIES_PER_STEP
    except Exception:
        return None


def cpu_baseline_leg(rows, budget_s=12.0):
    """Times the oracle's multi-threaded f32 scan + top-k on this box's host cores, on a bounded
    sample: as many full single-query passes over the same 1M x 384 matrix as fit the budget."""
    import numpy as np
    import oracle as O
    threads = O.max_threads()
    rng = np.random.RandomState(1)
    q = rng.standard_normal((64, DIM)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    O.cosine_topk_f32_fast(rows, q[0], TOPK, threads)  # warm-up
    n, t0 = 0, time.perf_counter()
    while True:
        O.cosine_topk_f32_fast(rows, q[n % 64], TOPK, threads)
        n += 1
        dt = time.perf_counter() - t0
        if dt > budget_s or n >= 2000:
            break
    return {"value": n / dt, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "%d full single-query passes over the 1M x 384 f32 matrix in %.1f s, %d OpenMP threads on '%s' "
                      "(self-written CPU oracle; no reference implementation of this path exists)" % (n, dt, threads, _cpu_model()),
            "gbs": n * N_DOCS * DIM * 4 / dt / 1e9}


def _dev_time(fn, steps, warmup):
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


def _zipf_cdf(vocab):
    """cumulative Zipf(s = 1) distribution over term ranks 1..vocab (numpy; inputs of the secondary legs only)"""
    import numpy as np
    w = 1.0 / np.arange(1, vocab + 1, dtype=np.float64)
    c = np.cumsum(w)
    c /= c[-1]
    c[-1] = 1.0
    return c


def _zipf_queries(nq, terms, cdf, seed):
    """nq queries of `terms` DISTINCT term ids drawn from the Zipf distribution (head-heavy, the realistic case)"""
    import numpy as np
    rng = np.random.default_rng(seed)
    out = np.empty((nq, terms), dtype=np.uint32)
    for j in range(nq):
        seen = []
        while len(seen) < terms:
            t = int(np.searchsorted(cdf, rng.random(), side="right"))
            t = min(t, len(cdf) - 1)
            if t not in seen:
                seen.append(t)
        out[j] = seen
    return out


def secondary_workloads(dev):
    """The other single-GPU configurations of BASELINE.json, measured in the same run (device-timed,
    inputs resident in HBM).  Informational: the headline line above stays configs[1]."""
    import numpy as np
    import torch
    import openintel_b200 as oi
    out = {}
    stream = torch.cuda.current_stream().cuda_stream
    # one 10M-document index serves configs[3] (768-dim bf16 rows, batch 256, tcgen05 path), configs[2] (BM25 over a
    # 1M-term Zipf vocabulary, 8-term queries, batch 1024) and the hybrid call on both (batch 256)
    try:
        n, dim, nb, nbm, vocab = 10_000_000, 768, 256, 1024, 1_000_000
        ix = oi.GpuIndex(n_docs=n, dim=dim, dtype=oi.DTYPE_BF16, max_k=TOPK, max_batch=nbm)
        ix.synth_embeddings(SEED)
        g = torch.Generator().manual_seed(7)
        qv = torch.randn(4, nb, dim, generator=g)
        qv = (qv / qv.norm(dim=2, keepdim=True)).to(dev)
        ids = torch.empty(nbm, TOPK, dtype=torch.int32, device=dev)
        sc = torch.empty(nbm, TOPK, dtype=torch.float32, device=dev)
        ms = _dev_time(lambda i: ix.search_cosine_dev(qv[i % 4], nb, TOPK, ids, sc, stream), 20, 3)
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
        tf = 2.0 * n * dim * nb / (ms * 1e-3) / 1e12
        out["configs[3] cosine 10M x 768 bf16 batch 256 (tcgen05)"] = {
            "queries_per_s": nb / (ms * 1e-3), "ms_per_batch": ms, "tensor_tflops": tf, "hbm_gbs": n * dim * 2 / (ms * 1e-3) / 1e9,
            "frac_of_measured_bf16_sustained": tf / peaks["bf16_tflops_sustained"] if peaks else None,
            "frac_of_measured_bf16_burst": tf / peaks["bf16_tflops"] if peaks else None}
        cdf = _zipf_cdf(vocab)
        ix.synth_bm25(SEED, vocab, cdf)
        ix.bm25_finalize()
        df, _, npost = ix.bm25_local_stats()
        pools = [_zipf_queries(nbm, 8, cdf, 100 + p) for p in range(4)]
        touched = float(np.mean([df[p].astype(np.int64).sum(axis=1).mean() for p in pools]))
        qt = [torch.from_numpy(p.astype(np.int32).reshape(-1)).to(dev) for p in pools]
        offs = torch.arange(0, nbm * 8 + 1, 8, dtype=torch.int32, device=dev)
        ms_b = _dev_time(lambda i: ix.search_bm25_dev(qt[i % 4], offs, nbm, TOPK, ids, sc, stream), 10, 3)
        out["configs[2] BM25 10M docs, 1M-term Zipf vocab, 8-term queries, batch 1024"] = {
            "queries_per_s": nbm / (ms_b * 1e-3), "ms_per_batch": ms_b, "postings": int(npost), "postings_touched_per_query": touched,
            "algorithmic_posting_gbs": touched * 8 * nbm / (ms_b * 1e-3) / 1e9}
        o = [torch.empty(nb, TOPK, dtype=torch.int32, device=dev) for _ in range(3)]
        rrf = torch.empty(nb, TOPK, dtype=torch.float32, device=dev)
        qt256 = [t[: nb * 8].contiguous() for t in qt]
        offs256 = offs[: nb + 1].contiguous()
        ms_h = _dev_time(lambda i: ix.search_hybrid_dev(qv[i % 4], qt256[i % 4], offs256, nb, TOPK, 60, o[0], rrf, o[1], o[2], stream), 10, 3)
        # the same call end to end through the host-buffer C ABI (numpy inputs and outputs, blocking)
        h_qv = qv.cpu().numpy()
        h_qt = [p[:nb] for p in pools]
        for i in range(2):
            ix.search_hybrid(h_qv[i % 4], h_qt[i % 4], TOPK)
        t0 = time.perf_counter()
        for i in range(10):
            ix.search_hybrid(h_qv[i % 4], h_qt[i % 4], TOPK)
        e2e_ms = (time.perf_counter() - t0) / 10 * 1e3
        out["hybrid BM25+cosine+RRF top-100, 10M x 768 bf16, 1M-term Zipf vocab, batch 256"] = {
            "queries_per_s": nb / (ms_h * 1e-3), "ms_per_batch": ms_h, "e2e_queries_per_s": nb / (e2e_ms * 1e-3), "e2e_ms_per_batch": e2e_ms}
        ix.close()
    except Exception as e:  # informational leg: never take the headline down
        out["configs[3]/[2]"] = {"error": str(e)[:200]}
    # hybrid BM25 + cosine + RRF on the configs[1] corpus (1M x 384 f32, 1M-term Zipf vocabulary), batch 16
    try:
        n, vocab, nb = N_DOCS, 1_000_000, 16
        cdf = _zipf_cdf(vocab)
        ix = oi.GpuIndex(n_docs=n, dim=DIM, max_k=TOPK, max_batch=nb)
        ix.synth_embeddings(SEED)
        ix.synth_bm25(SEED, vocab, cdf)
        ix.bm25_finalize()
        g = torch.Generator().manual_seed(7)
        qv = torch.randn(4, nb, DIM, generator=g)
        qv = (qv / qv.norm(dim=2, keepdim=True)).to(dev)
        qt = [torch.from_numpy(_zipf_queries(nb, 8, cdf, 200 + p).astype(np.int32).reshape(-1)).to(dev) for p in range(4)]
        offs = torch.arange(0, nb * 8 + 1, 8, dtype=torch.int32, device=dev)
        o = [torch.empty(nb, TOPK, dtype=torch.int32, device=dev) for _ in range(3)]
        rrf = torch.empty(nb, TOPK, dtype=torch.float32, device=dev)
        ms = _dev_time(lambda i: ix.search_hybrid_dev(qv[i % 4], qt[i % 4], offs, nb, TOPK, 60, o[0], rrf, o[1], o[2], stream), 20, 3)
        ms_b = _dev_time(lambda i: ix.search_bm25_dev(qt[i % 4], offs, nb, TOPK, o[0], rrf, stream), 20, 3)
        ms_c = _dev_time(lambda i: ix.search_cosine_dev(qv[i % 4], nb, TOPK, o[0], rrf, stream), 20, 3)
        out["hybrid BM25+cosine+RRF top-100, 1M x 384 f32, 1M-term Zipf vocab, batch 16"] = {
            "queries_per_s": nb / (ms * 1e-3), "ms_per_batch": ms, "ms_bm25_only": ms_b}
        out["cosine top-100, 1M x 384 f32, batch 16 in one call (multi-query scan: one matrix pass per 4 queries)"] = {
            "queries_per_s": nb / (ms_c * 1e-3), "ms_per_batch": ms_c, "matrix_passes": (nb + 3) // 4,
            "hbm_gbs": ((nb + 3) // 4) * n * DIM * 4 / (ms_c * 1e-3) / 1e9}
        ix.close()
    except Exception as e:
        out["hybrid"] = {"error": str(e)[:200]}
    return out


def run_reference(args, rank):
    """--impl reference: the CPU arm.  The reference has no implementation of this path (SURVEY.md
    §0) and is Rust (no toolchain here), so the arm times the oracle port on the host cores."""
    if rank != 0:
        return
    import numpy as np
    import oracle as O
    threads = O.max_threads()
    rows = _synth_rows_parallel(O, N_DOCS, DIM, threads)
    rng = np.random.RandomState(1)
    qper = 2  # bounded sample: 2 of the QUERIES_PER_STEP queries per step
    q = rng.standard_normal((64, DIM)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    for w in range(max(args.warmup, 1)):
        O.cosine_topk_f32_fast(rows, q[w % 64], TOPK, threads)
    t0 = time.perf_counter()
    for s in range(args.steps):
        for j in range(qper):
            O.cosine_topk_f32_fast(rows, q[(s * qper + j) % 64], TOPK, threads)
    dt = time.perf_counter() - t0
    val = args.steps * qper / dt
    sample = ("%d steps x %d full single-query passes over the 1M x 384 f32 matrix, %d OpenMP threads on '%s'"
              % (args.steps, qper, threads, _cpu_model()))
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic (counter-hash unit rows, seed 20261018)",
        "config": {"workload": WORKLOAD, "queries_per_step": QUERIES_PER_STEP, "n_docs": N_DOCS, "dim": DIM, "k": TOPK,
                   "parallelism": "host cores (OpenMP)", "kernel_variant": "CPU oracle port",
                   "note": "the reference has no implementation of this path (SURVEY.md §0); the CPU oracle port is timed; "
                           "each step is a bounded sample: %d of the step's %d queries" % (qper, QUERIES_PER_STEP)},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--variant", type=int, default=None, help="cosine kernel variant override (0 ldg, 1 bulk)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true", help="skip the informational configs[3] / hybrid legs")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)
    if args.warmup < 3:
        args.warmup = 3

    import numpy as np
    import torch
    import openintel_b200 as oi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path")
    torch.cuda.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    dev = torch.device("cuda", local_rank)

    # ---- shard of the fixed corpus held by this rank (strong scaling) --------------------------
    per = (N_DOCS + world - 1) // world
    base = rank * per
    n_local = max(0, min(N_DOCS, base + per) - base)
    ix = oi.GpuIndex(n_docs=n_local, dim=DIM, dtype=oi.DTYPE_F32, device=local_rank, doc_base=base,
                     max_k=TOPK, max_batch=QUERIES_PER_STEP)
    ix.synth_embeddings(SEED)
    if args.variant is not None:
        ix.set_option("cosine_variant", args.variant)
    # configs[1] is the SINGLE-query GEMV path: every query of a step scans the matrix on its own.  (Left on, the
    # library serves a multi-query call with one pass per group of 4 queries; that is measured as a secondary leg.)
    ix.set_option("cosine_multi_query", 0)
    if world > 1:
        uid = torch.zeros(128, dtype=torch.uint8)
        if rank == 0:
            uid = torch.from_numpy(oi.GpuIndex.comm_unique_id().copy())
        uid = uid.to(dev)
        dist.broadcast(uid, 0)
        ix.comm_init(rank, world, uid.cpu().numpy())

    # ---- queries: a pool of distinct unit vectors, identical on every rank ----------------------
    n_pool = 64
    g = torch.Generator().manual_seed(1234)
    pool = torch.randn(n_pool, QUERIES_PER_STEP, DIM, generator=g, dtype=torch.float32)
    pool = pool / pool.norm(dim=2, keepdim=True)
    d_pool = pool.to(dev)
    h_pool = pool.pin_memory()
    d_ids = torch.empty(QUERIES_PER_STEP, TOPK, dtype=torch.int32, device=dev)
    d_sc = torch.empty(QUERIES_PER_STEP, TOPK, dtype=torch.float32, device=dev)
    h_ids = torch.empty(QUERIES_PER_STEP, TOPK, dtype=torch.int32).pin_memory()
    h_sc = torch.empty(QUERIES_PER_STEP, TOPK, dtype=torch.float32).pin_memory()
    stream = torch.cuda.current_stream().cuda_stream

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def step_dev(i):
        ix.search_cosine_dev(d_pool[i % n_pool], QUERIES_PER_STEP, TOPK, d_ids, d_sc, stream)

    def step_host(i):
        ix.search_cosine(h_pool[i % n_pool], TOPK, h_ids, h_sc)

    # ---- device-resident leg ("value") ------------------------------------------------------------
    for i in range(args.warmup):
        step_dev(i)
    barrier()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    l0 = ix.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        step_dev(args.warmup + i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    launches = ix.launch_count() - l0
    clocks = sampler.stop() if rank == 0 else None
    if dist is not None:
        t = torch.tensor([ms], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    n_queries = args.steps * QUERIES_PER_STEP
    value = n_queries / (ms * 1e-3)

    # ---- end-to-end leg: host buffers through the C ABI -------------------------------------------
    for i in range(args.warmup):
        step_host(i)
    barrier()
    t0 = time.perf_counter()
    for i in range(args.steps):
        step_host(args.warmup + i)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if dist is not None:
        t = torch.tensor([e2e_s], device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t.item())
    barrier()
    e2e_value = n_queries / e2e_s

    # ---- single-query latency: one query per call, device-resident --------------------------------
    lat_ms = _dev_time(lambda i: ix.search_cosine_dev(d_pool[i % n_pool][:1], 1, TOPK, d_ids[:1], d_sc[:1], stream), 50, 5)

    # ---- sanity outside the timed region: results are ranked lists of real docs -------------------
    ids = h_ids.numpy().view(np.uint32)
    sc = h_sc.numpy()
    assert np.all(np.diff(sc, axis=1) <= 0) and ids.max() < N_DOCS, "bench result is not a ranked list"

    if rank == 0:
        peak, peak_src = _peaks()
        # variant 1 (default): ONE persistent scan launch per step walks the step's queries back to back
        per_query_launch = args.variant == 0
        scan_launches = args.steps * (QUERIES_PER_STEP if per_query_launch else 1)
        bytes_per_launch = n_local * DIM * 4 * (1 if per_query_launch else QUERIES_PER_STEP)
        avg_launch_s = ms * 1e-3 / scan_launches  # includes the per-step unpack/merge share: conservative
        achieved = bytes_per_launch / avg_launch_s / 1e9
        cpu = None
        if world == 1 and not args.no_cpu_baseline:
            rows = ix.read_embeddings(0, n_local)  # bit-identical to the oracle's rows (tests/test_gpu_cosine.py)
            cpu = cpu_baseline_leg(rows)
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic (counter-hash unit rows generated on device, seed 20261018; random unit queries)",
            "config": {"workload": WORKLOAD, "queries_per_step": QUERIES_PER_STEP, "n_docs": N_DOCS, "dim": DIM, "k": TOPK,
                       "parallelism": "doc-sharded x%d, NCCL all-gather of local top-k + device merge" % world if world > 1 else "1 GPU",
                       "l2": "each query streams %.2f GB per GPU, larger than the 126 MB L2; no flush needed" % (n_local * DIM * 4 / 1e9),
                       "kernel_variant": ix_variant_name(args.variant),
                       "multi_query_scan": "off: the step's queries are independent single-query scans (one matrix pass each)"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": QUERIES_PER_STEP * DIM * 4,
                    "d2h_bytes_per_step": QUERIES_PER_STEP * TOPK * 8, "timing": "wall clock around blocking C-ABI calls, max over ranks"},
            "gpu_launches": launches,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": _traffic("cosine_scan_bulk_f32_1Mx384_q16") if (world == 1 and not per_query_launch) else None,
                         "traffic_source": "profiles/r01_ncu_cosine_scan_bulk.md (ncu --set full, same kernel and shape; per-query DRAM bytes x queries per launch)",
                         "peak_source": peak_src, "frac_of_8TBs_nominal": achieved / 8000.0,
                         "kernel": "cosine_scan_*_kernel", "bytes_per_launch": bytes_per_launch,
                         "avg_launch_us": avg_launch_s * 1e6,
                         "note": "avg launch = timed region / scan launches (includes unpack + launch gaps)"},
            "cpu_baseline": cpu,
            "clocks": clocks,
            "latency_us_nq1": lat_ms * 1e3,
        }
        if world == 1 and not args.no_secondary:
            ix.close()
            out["secondary"] = secondary_workloads(dev)
        print(json.dumps(out))
    ix.close()
    if dist is not None:
        dist.destroy_process_group()


def ix_variant_name(v):
    return {None: "default (bulk, dynamic tiles, persistent over the step's queries)", 0: "ldg (128-bit direct loads)", 1: "bulk (cp.async.bulk + mbarrier ring)"}[v]


if __name__ == "__main__":
    main()
