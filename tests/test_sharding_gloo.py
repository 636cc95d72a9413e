"""N > 1 host logic on CPU: two gloo ranks shard a synthetic corpus by document (SPEC §5), agree
on GLOBAL BM25 statistics through openintel_b200.sharding, score their shards with the oracle,
exchange local top-k lists and merge — the result must equal the unsharded oracle lists bit for
bit.  (The per-query exchange of the product is NCCL + a device merge: tools/multigpu_check.py.)"""
import os
import socket

import numpy as np
import pytest

import oracle as O
from openintel_b200 import sharding

N_DOCS, VOCAB, K, NQ, DIM = 9001, 1500, 25, 5, 64


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import torch
        base, n_local = sharding.shard_range(N_DOCS, world, rank)
        corp = O.synth_bm25_corpus(n_local, VOCAB, first=base)
        df = np.diff(corp["term_offsets"]).astype(np.uint32)
        gdf, avgdl, n = sharding.global_bm25_stats(dist, df, int(corp["doc_len"].sum()), n_local)
        idf = O.bm25_idf(n, gdf)
        w = O.bm25_weights(corp["term_offsets"], corp["doc_ids"], corp["tfs"], corp["doc_len"], idf, avgdl=float(avgdl))
        qt = O.synth_query_terms(NQ, 8, O.zipf_cdf(VOCAB))
        rows = O.synth_rows_f32(n_local, DIM, first=base)
        qv = O.synth_rows_f32(NQ, DIM, stream=1)
        ids = np.empty((2, NQ, K), dtype=np.uint32)
        sc = np.empty((2, NQ, K), dtype=np.float32)
        for j in range(NQ):
            s = O.bm25_score_dense(corp["term_offsets"], corp["doc_ids"], w, qt[j], n_local)
            ids[0, j], sc[0, j], _ = O.topk_f32(s, K, only_positive=True, doc_base=base)
            c = O.cosine_scores_f32(rows, qv[j]).astype(np.float32)
            ids[1, j], sc[1, j], _ = O.topk_f32(c, K, doc_base=base)
        t_ids = torch.from_numpy(ids.astype(np.int64))
        t_sc = torch.from_numpy(sc)
        g_ids = [torch.empty_like(t_ids) for _ in range(world)]
        g_sc = [torch.empty_like(t_sc) for _ in range(world)]
        dist.all_gather(g_ids, t_ids)
        dist.all_gather(g_sc, t_sc)
        out = {}
        for m, name in enumerate(("bm25", "cos")):
            out[name] = sharding.merge_lists_host([g[m].numpy().astype(np.uint32) for g in g_ids],
                                                  [g[m].numpy() for g in g_sc], K)
        q.put((rank, base, n_local, gdf, float(avgdl), n, out))
    finally:
        dist.destroy_process_group()


def test_shard_range_covers_the_corpus():
    for n, g in [(10, 3), (9001, 2), (100, 8), (7, 8), (0, 2), (1000000, 8)]:
        spans = [sharding.shard_range(n, g, r) for r in range(g)]
        assert sum(s[1] for s in spans) == n
        pos = 0
        for base, cnt in spans:
            if cnt:
                assert base == pos
            pos += cnt
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def test_two_gloo_ranks_reproduce_the_unsharded_lists():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)

    corp = O.synth_bm25_corpus(N_DOCS, VOCAB)
    gdf = np.diff(corp["term_offsets"]).astype(np.uint32)
    avgdl = O.bm25_avgdl(corp["doc_len"])
    idf = O.bm25_idf(N_DOCS, gdf)
    w = O.bm25_weights(corp["term_offsets"], corp["doc_ids"], corp["tfs"], corp["doc_len"], idf)
    qt = O.synth_query_terms(NQ, 8, corp["cdf"])
    rows = O.synth_rows_f32(N_DOCS, DIM)
    qv = O.synth_rows_f32(NQ, DIM, stream=1)
    for rank, base, n_local, r_gdf, r_avgdl, r_n, out in res:
        assert (base, n_local) == sharding.shard_range(N_DOCS, 2, rank)
        assert np.array_equal(r_gdf, gdf) and r_avgdl == avgdl and r_n == N_DOCS
        for j in range(NQ):
            s = O.bm25_score_dense(corp["term_offsets"], corp["doc_ids"], w, qt[j], N_DOCS)
            wi, ws, _ = O.topk_f32(s, K, only_positive=True)
            assert np.array_equal(out["bm25"][0][j], wi)
            assert np.array_equal(out["bm25"][1][j].view(np.uint32), ws.view(np.uint32))
            c = O.cosine_scores_f32(rows, qv[j]).astype(np.float32)
            wi, ws, _ = O.topk_f32(c, K)
            assert np.array_equal(out["cos"][0][j], wi)
            assert np.array_equal(out["cos"][1][j].view(np.uint32), ws.view(np.uint32))


# ---- document-sharded lift out of the SQLite post store (openintel_b200.store.lift_shard) ----------------------
N_POSTS = 1501


def _store_worker(rank, world, port, path, q):
    import torch.distributed as dist
    from openintel_b200 import store
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        conn = store.open_store(path)
        sh = store.lift_shard(conn, dist)
        conn.close()
        q.put((rank, sh))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_store_lift_equals_the_unsharded_lift(tmp_path, world):
    """every rank tokenises only its doc range; the union vocabulary, the re-indexed shard CSRs and the all-reduced
    statistics must add up to exactly the unsharded lift (global term ids, global df / avgdl / N)"""
    import runpy
    import torch.multiprocessing as mp
    from openintel_b200 import hostlib, store
    runpy.run_path(os.path.join(os.path.dirname(hostlib.__file__), "host", "build.py"), run_name="__build__")
    path = str(tmp_path / "posts.db")
    posts, _ = O.synth_posts(N_POSTS, 700, O.SEED)
    conn = store.open_store(path, dim=8)
    store.insert_posts(conn, posts, np.ones((N_POSTS, 8), np.float32))
    b, whole, ids = store.lift_csr(conn)
    vocab = [b.term(i) for i in range(whole["n_terms"])]
    conn.close()

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_store_worker, args=(r, world, port, path, q)) for r in range(world)]
    [p.start() for p in procs]
    res = sorted([q.get(timeout=180) for _ in procs], key=lambda t: t[0])
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)

    df = np.diff(whole["term_offsets"]).astype(np.uint32)
    avgdl = float(np.float32(np.float64(whole["doc_len"].sum()) / N_POSTS))
    lists = [[] for _ in vocab]
    tfs = [[] for _ in vocab]
    all_ids, all_dl = [], []
    for rank, sh in res:
        assert (sh["doc_base"], sh["n_local"]) == sharding.shard_range(N_POSTS, world, rank)
        assert sh["vocab"] == vocab and sh["n_docs_global"] == N_POSTS
        assert np.array_equal(sh["global_df"], df) and sh["avgdl"] == avgdl
        c = sh["csr"]
        assert c["n_terms"] == len(vocab) and c["n_docs"] == sh["n_local"] and len(c["doc_len"]) == sh["n_local"]
        off = c["term_offsets"].astype(np.int64)
        for t in range(len(vocab)):
            lists[t].extend((c["doc_ids"][off[t]:off[t + 1]].astype(np.int64) + sh["doc_base"]).tolist())
            tfs[t].extend(c["tfs"][off[t]:off[t + 1]].tolist())
        all_ids.extend(sh["post_ids"])
        all_dl.extend(c["doc_len"].tolist())
    w_off = whole["term_offsets"].astype(np.int64)
    for t in range(len(vocab)):
        assert lists[t] == whole["doc_ids"][w_off[t]:w_off[t + 1]].tolist(), vocab[t]
        assert tfs[t] == whole["tfs"][w_off[t]:w_off[t + 1]].tolist()
    assert all_ids == ids and all_dl == whole["doc_len"].tolist()
