#!/usr/bin/env python3
"""Multi-GPU parity check (run under torchrun, one rank per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 \
        --master-port 29511 tools/multigpu_check.py

Builds a doc-sharded hybrid index (SPEC §5) through openintel_b200.sharding, runs cosine / BM25 /
hybrid searches whose local top-k lists are all-gathered with NCCL and merged on device, and
compares every rank's result with the UNSHARDED CPU oracle.  Exit code 0 = parity."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import openintel_b200 as oi
    from openintel_b200 import sharding
    import oracle as O
    from gpu_util import assert_ranked_close

    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    n, dim, vocab, k, nq = 120001, 128, 8000, 100, 12
    cdf = O.zipf_cdf(vocab)
    for dtype, tol, exchange in ((oi.DTYPE_F32, 1e-5, "nccl"), (oi.DTYPE_BF16, 2e-3, "p2p"), (oi.DTYPE_F32, 1e-5, "p2p")):
        with sharding.ShardedIndex(n, dim, dtype=dtype, dist=dist, device_index=lr, max_k=k, max_batch=nq, exchange=exchange) as sh:
            sh.synth(O.SEED, vocab, cdf)
            qv = np.concatenate([O.synth_rows_f32(nq // 2, dim, stream=1), O.synth_planted_queries(nq - nq // 2, dim, n)[0]])
            qt = O.synth_query_terms(nq, 8, cdf)
            cos_ids, cos_sc = sh.ix.search_cosine(qv, k)
            bm_ids, bm_sc = sh.ix.search_bm25(qt, k)
            ids, rrf, rc, rb = sh.ix.search_hybrid(qv, qt, k)
            if exchange == "p2p":
                for _ in range(5):  # the two-slot protocol over several consecutive batches
                    ids2, rrf2, _, _ = sh.ix.search_hybrid(qv, qt, k)
                    assert np.array_equal(ids2, ids) and np.array_equal(rrf2, rrf)
                batches, timed_out = sh.ix.p2p_status()
                assert batches >= 8 and not timed_out, (batches, timed_out)
        # unsharded oracle
        corp = O.synth_bm25_corpus(n, vocab)
        idf = O.bm25_idf(n, np.diff(corp["term_offsets"]))
        w = O.bm25_weights(corp["term_offsets"], corp["doc_ids"], corp["tfs"], corp["doc_len"], idf)
        rows = O.synth_rows_f32(n, dim) if dtype == oi.DTYPE_F32 else O.synth_rows_bf16(n, dim)
        for j in range(nq):
            s = O.bm25_score_dense(corp["term_offsets"], corp["doc_ids"], w, qt[j], n)
            wi, ws, _ = O.topk_f32(s, k, only_positive=True)
            assert np.array_equal(bm_ids[j], wi), "rank %d: BM25 ids differ for query %d" % (rank, j)
            assert np.array_equal(bm_sc[j].view(np.uint32), ws.view(np.uint32))
            allsc = (O.cosine_scores_f32 if dtype == oi.DTYPE_F32 else O.cosine_scores_bf16)(rows, qv[j])
            ci, cs, _ = O.topk_f64(allsc, k)
            assert_ranked_close(cos_ids[j], cos_sc[j], ci, cs, allsc, tol)
            e_ids, e_val, e_rc, e_rb, _ = O.rrf(cos_ids[j], bm_ids[j], k)
            assert np.array_equal(ids[j], e_ids) and np.array_equal(rrf[j].view(np.uint32), e_val.view(np.uint32))
            assert np.array_equal(rc[j], e_rc) and np.array_equal(rb[j], e_rb)
        # every rank holds the same global lists
        t = torch.from_numpy(ids.astype(np.int64)).cuda()
        ref = t.clone()
        dist.broadcast(ref, 0)
        assert torch.equal(t, ref), "rank %d: hybrid lists differ from rank 0's" % rank
    dist.barrier()
    if rank == 0:
        print("multigpu_check ok: world=%d, %d docs sharded, cosine(f32+bf16)/BM25/hybrid equal the unsharded oracle (NCCL all-gather and peer-to-peer push)" % (world, n), flush=True)
    full_size_oracle(rank, world, lr, dist, oi, sharding, O)
    dist.destroy_process_group()


def full_size_oracle(rank, world, lr, dist, oi, sharding, O):
    """BASELINE configs[4] at its 8-GPU shard size (12.5M documents x 768 bf16 per GPU, i.e. 12.5M x world documents in
    all): the hybrid call's global lists of two queries against the oracle run on the WHOLE corpus -- cosine by chunked
    brute force over regenerated rows, BM25 from the oracle's own CSR of the touched terms -- computed on rank 0's host."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from test_gpu_full_size_oracle import assert_list_matches_oracle
    per, dim, vocab, k, nq = 12_500_000, 768, 1_000_000, 100, 256
    n = per * world
    cdf = O.zipf_cdf(vocab)
    pq, tgt = O.synth_planted_queries(1, dim, n)
    q = np.concatenate([pq, O.synth_rows_f32(nq - 1, dim, stream=1)])
    qt = O.synth_query_terms(nq, 8, cdf, first=777)
    with sharding.ShardedIndex(n, dim, dtype=oi.DTYPE_BF16, dist=dist, device_index=lr, max_k=k, max_batch=nq, exchange="p2p") as sh:
        sh.synth(O.SEED, vocab, cdf)
        ids, rrf, rc, rb = sh.ix.search_hybrid(q, qt, k)
        c_ids, c_sc = sh.ix.search_cosine(q, k)
        b_ids, b_sc = sh.ix.search_bm25(qt, k)
        gdf, avgdl = sh.global_df, sh.avgdl
    if rank == 0:
        check = [0, 1]
        o_ids, o_sc = O.scale_cosine_topk(n, dim, q[check], k, bf16=True)
        assert o_ids[0][0] == tgt[0]
        mini = O.scale_bm25_mini_index(n, vocab, qt[check].reshape(-1), cdf=cdf)
        for c, j in enumerate(check):
            swaps = assert_list_matches_oracle(c_ids[j], c_sc[j], o_ids[c], o_sc[c], q[j], dim, True, 2e-3)
            w_ids, w_sc, _ = O.scale_bm25_topk(mini, qt[j], k, n_docs_global=n, avgdl=float(avgdl), df_global=gdf[mini["terms"]])
            assert np.array_equal(w_ids, b_ids[j]) and np.array_equal(w_sc.view(np.uint32), b_sc[j].view(np.uint32)), "BM25 list differs from the oracle"
            e_ids, e_val, e_rc, e_rb, _ = O.rrf(o_ids[c] if swaps == 0 else c_ids[j], w_ids, k)
            assert np.array_equal(ids[j], e_ids) and np.array_equal(rrf[j].view(np.uint32), e_val.view(np.uint32))
            assert np.array_equal(rc[j], e_rc) and np.array_equal(rb[j], e_rb)
        print("multigpu full_size_oracle ok: %d x %d documents x 768 bf16 (configs[4] shard size), hybrid top-%d of 2 queries equals the "
              "oracle on the whole %dM-document corpus" % (world, per, k, n // 1_000_000), flush=True)
    dist.barrier()


if __name__ == "__main__":
    main()
