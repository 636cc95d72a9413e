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
    for dtype, tol in ((oi.DTYPE_F32, 1e-5), (oi.DTYPE_BF16, 2e-3)):
        with sharding.ShardedIndex(n, dim, dtype=dtype, dist=dist, device_index=lr, max_k=k, max_batch=nq) as sh:
            sh.synth(O.SEED, vocab, cdf)
            qv = np.concatenate([O.synth_rows_f32(nq // 2, dim, stream=1), O.synth_planted_queries(nq - nq // 2, dim, n)[0]])
            qt = O.synth_query_terms(nq, 8, cdf)
            cos_ids, cos_sc = sh.ix.search_cosine(qv, k)
            bm_ids, bm_sc = sh.ix.search_bm25(qt, k)
            ids, rrf, rc, rb = sh.ix.search_hybrid(qv, qt, k)
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
        print("multigpu_check ok: world=%d, %d docs sharded, cosine(f32+bf16)/BM25/hybrid equal the unsharded oracle" % (world, n))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
