"""Host-side plumbing of the document-sharded index (docs/SPEC.md §5): one process per GPU,
contiguous doc ranges, GLOBAL BM25 statistics, and the NCCL communicator the C library uses for
the all-gather of local top-k lists.  torch.distributed is used only as the control plane
(all-reduce of df / doc-length sums at build time, broadcast of the NCCL unique id); the
per-query exchange itself is libopenintel_gpu.so's ncclAllGather + device merge.

The same functions run over gloo on CPU tensors (tests/test_sharding_gloo.py) and over NCCL.
"""
import numpy as np

from . import capi


def shard_range(n_docs, world, rank):
    """SPEC §5: shard r of G holds [r*ceil(N/G), min(N, (r+1)*ceil(N/G))).  -> (doc_base, n_local)"""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("rank %d / world %d" % (rank, world))
    per = (n_docs + world - 1) // world
    base = min(n_docs, rank * per)
    return base, max(0, min(n_docs, base + per) - base)


def global_bm25_stats(dist, df_local, sum_doc_len_local, n_docs_local, device="cpu"):
    """All-reduce the shard statistics into the GLOBAL (df[n_terms] u32, avgdl f32, N) every shard
    must score with, so that a sharded index ranks exactly like an unsharded one (SPEC §3/§5).
    `dist` is torch.distributed (initialised) or None for a single shard."""
    df = np.ascontiguousarray(df_local, dtype=np.uint32)
    if dist is None or dist.get_world_size() == 1:
        n = int(n_docs_local)
        return df, np.float32(np.float64(sum_doc_len_local) / n) if n else np.float32(1.0), n
    import torch
    t_df = torch.from_numpy(df.astype(np.int64)).to(device)
    t_sc = torch.tensor([int(sum_doc_len_local), int(n_docs_local)], dtype=torch.int64, device=device)
    dist.all_reduce(t_df, op=dist.ReduceOp.SUM)
    dist.all_reduce(t_sc, op=dist.ReduceOp.SUM)
    gdf = t_df.cpu().numpy()
    if gdf.max(initial=0) > 0xFFFFFFFF:
        raise OverflowError("global df does not fit u32")
    s, n = (int(x) for x in t_sc.cpu().tolist())
    avgdl = np.float32(np.float64(s) / np.float64(n)) if n else np.float32(1.0)
    return gdf.astype(np.uint32), avgdl, n


def broadcast_unique_id(dist, device="cpu"):
    """rank 0 creates the NCCL unique id (C ABI), everybody receives it."""
    import torch
    uid = torch.zeros(capi.UNIQUE_ID_BYTES, dtype=torch.uint8)
    if dist.get_rank() == 0:
        uid = torch.from_numpy(capi.GpuIndex.comm_unique_id().copy())
    uid = uid.to(device)
    dist.broadcast(uid, 0)
    return uid.cpu().numpy()


def attach_p2p(dist, ix, device="cpu"):
    """Switch a shard's exchange from ncclAllGather to the peer-to-peer push (all ranks on one NVLink box): every rank
    exports the IPC handle of its exchange buffer, the handles are all-gathered through torch.distributed (control
    plane), every rank maps its peers' buffers."""
    import torch
    mine = torch.from_numpy(ix.p2p_export().copy()).to(device)
    allh = [torch.zeros_like(mine) for _ in range(dist.get_world_size())]
    dist.all_gather(allh, mine)
    ix.p2p_attach(np.stack([t.cpu().numpy() for t in allh]))


def merge_lists_host(ids_per_rank, scores_per_rank, k):
    """Reference model of the device merge for tests: G lists [nq][k] -> best k by SPEC §1 order
    (score desc, doc id asc; NO_DOC padding dropped).  Not used by the product path."""
    ids = np.concatenate(ids_per_rank, axis=1).astype(np.uint32)
    sc = np.concatenate(scores_per_rank, axis=1).astype(np.float32)
    nq = ids.shape[0]
    out_ids = np.full((nq, k), capi.NO_DOC, dtype=np.uint32)
    out_sc = np.zeros((nq, k), dtype=np.float32)
    for j in range(nq):
        keep = ids[j] != capi.NO_DOC
        i, s = ids[j][keep], sc[j][keep] + np.float32(0.0)
        order = np.lexsort((i, -s.astype(np.float64)))[:k]
        out_ids[j, :len(order)] = i[order]
        out_sc[j, :len(order)] = s[order]
    return out_ids, out_sc


class ShardedIndex:
    """One rank's shard of a doc-sharded hybrid index.  Builds the shard (synthetic, SPEC §9, or
    from caller-provided arrays), agrees on global BM25 statistics and wires the communicator."""

    def __init__(self, n_docs_global, dim, dtype=capi.DTYPE_F32, dist=None, device_index=0, max_k=100, max_batch=1, exchange="nccl"):
        self.dist = dist if (dist is not None and dist.get_world_size() > 1) else None
        self.rank = self.dist.get_rank() if self.dist else 0
        self.world = self.dist.get_world_size() if self.dist else 1
        self.n_docs_global = int(n_docs_global)
        self.doc_base, self.n_local = shard_range(self.n_docs_global, self.world, self.rank)
        self.device_index = device_index
        self.ix = capi.GpuIndex(n_docs=self.n_local, dim=dim, dtype=dtype, device=device_index, doc_base=self.doc_base,
                                max_k=max_k, max_batch=max_batch)
        if self.dist:
            uid = broadcast_unique_id(self.dist, device="cuda:%d" % device_index)
            self.ix.comm_init(self.rank, self.world, uid)
            if exchange == "p2p":
                attach_p2p(self.dist, self.ix, device="cuda:%d" % device_index)

    def synth(self, seed, vocab=None, zipf_cdf=None, k1=1.2, b=0.75):
        self.ix.synth_embeddings(seed)
        if vocab:
            self.ix.synth_bm25(seed, vocab, zipf_cdf)
            self.finalize_bm25(k1, b)
        return self

    def finalize_bm25(self, k1=1.2, b=0.75):
        df, sdl, _ = self.ix.bm25_local_stats()
        gdf, avgdl, n = global_bm25_stats(self.dist, df, sdl, self.n_local, device="cuda:%d" % self.device_index)
        assert n == self.n_docs_global
        self.global_df, self.avgdl = gdf, avgdl
        self.ix.bm25_finalize(k1=k1, b=b, avgdl=float(avgdl), n_docs_global=n, global_df=gdf)

    def close(self):
        self.ix.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
