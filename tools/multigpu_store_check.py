#!/usr/bin/env python3
"""Multi-GPU check of the store -> sharded index path (run under torchrun, one rank per GPU): rank 0 writes a SQLite
post store, every rank lifts its doc range out of it (openintel_b200.store.ShardedStoreIndex: local tokenisation,
union vocabulary, global BM25 statistics, NCCL communicator) and answers hybrid text + vector queries; the lists must
equal, bit for bit, those of the unsharded StoreIndex built from the same file on one GPU.  Exit code 0 = parity."""
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist
    import oracle as O
    from openintel_b200 import store
    rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(lr)
    dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    n, vocab, dim, k = 20001, 5000, 128, 20
    path = os.path.join(tempfile.gettempdir(), "oi_store_check_%s.db" % os.environ.get("MASTER_PORT", "0"))
    posts, _ = O.synth_posts(n, vocab, O.SEED)
    if rank == 0:
        if os.path.exists(path):
            os.remove(path)
        conn = store.open_store(path, dim=dim)
        store.insert_posts(conn, posts, O.synth_rows_f32(n, dim) * np.float32(1.7))
        conn.close()
    dist.barrier()
    texts = [" ".join(posts[i]["text"].split()[:7]) for i in (5, n // 3, n - 2)] + ["zzz nothing known"]
    qv = O.synth_rows_f32(len(texts), dim, stream=1)
    conn = store.open_store(path)
    with store.ShardedStoreIndex(conn, dist, device_index=lr, max_k=k, max_batch=8) as sx:
        got = sx.search(texts, qv, k)
        one = sx.search(texts[:1], qv[:1], k)
    ok = all(np.array_equal(a[:1], b) for a, b in zip(got, one))
    if rank == 0:
        with store.StoreIndex(conn, device=lr, max_k=k, max_batch=8) as whole:
            want = whole.ix.search_hybrid(store.normalise_rows_f32(qv), whole.query_terms(texts), k)
        for a, b in zip(got, want):
            ok = ok and np.array_equal(a.view(np.uint32), b.view(np.uint32))
    conn.close()
    flag = torch.tensor([1 if ok else 0], device="cuda:%d" % lr)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    dist.barrier()
    if rank == 0:
        os.remove(path)
        print("multigpu_store_check %s: world=%d, %d posts lifted from SQLite in %d shards, hybrid lists equal the unsharded index"
              % ("ok" if int(flag.item()) else "FAILED", world, n, world), flush=True)
    dist.destroy_process_group()
    sys.exit(0 if int(flag.item()) else 1)


if __name__ == "__main__":
    main()
