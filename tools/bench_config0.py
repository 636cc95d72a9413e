#!/usr/bin/env python3
"""BASELINE.json configs[0]: hybrid search over 10k synthetic posts held in a SQLite store, 384-dim f32 embeddings,
BM25 + cosine + RRF top-10, one query per call.  Times (a) the store -> GPU index lift, (b) the GPU path through
StoreIndex.search (query tokenisation, host buffers, blocking C-ABI call), (c) the same query on the CPU oracle
(self-written: the reference has no retrieval code), and checks that both return the same posts."""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--docs", type=int, default=10000)
    ap.add_argument("--vocab", type=int, default=50000)
    ap.add_argument("--dim", type=int, default=384)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--queries", type=int, default=200)
    a = ap.parse_args()
    import oracle as O
    from openintel_b200 import store
    posts, _ = O.synth_posts(a.docs, a.vocab, O.SEED)
    emb = O.synth_rows_f32(a.docs, a.dim)
    conn = store.open_store(":memory:", dim=a.dim)
    t0 = time.perf_counter()
    store.insert_posts(conn, posts, emb)
    t_ins = time.perf_counter() - t0
    t0 = time.perf_counter()
    sx = store.StoreIndex(conn, max_k=a.k, max_batch=1)
    t_lift = time.perf_counter() - t0
    texts = [" ".join(posts[(i * 7919) % a.docs]["text"].split()[:8]) for i in range(a.queries)]
    qv = O.synth_rows_f32(a.queries, a.dim, stream=1)
    for i in range(5):
        sx.search(texts[i:i + 1], qv[i:i + 1], a.k)
    t0 = time.perf_counter()
    got = [sx.search(texts[i:i + 1], qv[i:i + 1], a.k)[0] for i in range(a.queries)]
    gpu_us = (time.perf_counter() - t0) / a.queries * 1e6
    # CPU oracle on the CSR the store produced (lift not timed), same queries
    b, csr, _ = store.lift_csr(conn)
    w = O.bm25_weights(csr["term_offsets"], csr["doc_ids"], csr["tfs"], csr["doc_len"], O.bm25_idf(a.docs, np.diff(csr["term_offsets"])))
    rows = store.normalise_rows_f32(emb)
    qn = store.normalise_rows_f32(qv)
    t0 = time.perf_counter()
    same = 0
    for i in range(a.queries):
        qt = b.query_terms(texts[i])
        c_ids, _, _ = O.topk_f64(O.cosine_scores_f32(rows, qn[i]), a.k)
        s = O.bm25_score_dense(csr["term_offsets"], csr["doc_ids"], w, qt, a.docs)
        b_ids, _, _ = O.topk_f32(s, a.k, only_positive=True)
        e_ids, _, _, _, m = O.rrf(c_ids, b_ids, a.k)
        same += int([h["doc_id"] for h in got[i]] == e_ids[:m].tolist())
    cpu_us = (time.perf_counter() - t0) / a.queries * 1e6
    print(json.dumps({"workload": "configs[0]: %d posts in SQLite, %d-dim f32, BM25+cosine+RRF top-%d, single query per call" % (a.docs, a.dim, a.k),
                      "gpu_us_per_query_e2e": gpu_us, "gpu_queries_per_s": 1e6 / gpu_us,
                      "cpu_oracle_us_per_query": cpu_us, "cpu_note": "self-written CPU oracle (numpy + C, 1 thread); no reference implementation exists",
                      "identical_result_lists": "%d / %d" % (same, a.queries), "store_insert_s": t_ins, "store_to_gpu_index_s": t_lift,
                      "n_terms": sx.n_terms}), flush=True)
    sx.close()


if __name__ == "__main__":
    main()
