"""SQLite post store and the index builder that lifts a GPU hybrid index out of it
(BASELINE.json configs[0] "SQLite-backed").

The reference keeps no store of any kind, so the schema is ours; it mirrors the
reference's document type field for field -- `SocialPost {id, source, author, text, created_at,
engagement}` (src/domain/entities/social_post.rs:30-38), `PostText::parse` (trimmed, non-empty,
at most 10 000 chars: social_post.rs:7-23), `SourceKind {Reddit, Bluesky}` rendered as
"reddit" / "bluesky" (src/domain/values/source_kind.rs:5-20) -- plus one embedding per post.

    store  --posts, in doc_id order-->  host tokenizer + IndexBuilder (C++, the reference tokenizer
                                        src/adapters/analyzer/lexicon.rs:54-58)  -->  CSR  -->  oi_index_load_bm25
           --embeddings (f32 BLOBs)-->  L2-normalise in f32 (SPEC §2)            -->  oi_index_load_embeddings

Scoring happens in libopenintel_gpu.so only; nothing here computes a score.
"""
import sqlite3

import numpy as np

from . import capi, hostlib

MAX_POST_LEN = 10_000  # chars, not bytes (social_post.rs:7, :19)
SOURCES = ("reddit", "bluesky")

SCHEMA = """
CREATE TABLE IF NOT EXISTS posts (
  doc_id     INTEGER PRIMARY KEY,                              -- dense row id 0..N-1 = the index's u32 doc id
  id         TEXT    NOT NULL UNIQUE,                          -- SocialPost.id
  source     TEXT    NOT NULL CHECK (source IN ('reddit', 'bluesky')),
  author     TEXT    NOT NULL,
  text       TEXT    NOT NULL,                                 -- PostText: trimmed, non-empty, <= 10 000 chars
  created_at TEXT    NOT NULL,                                 -- RFC 3339, UTC
  engagement INTEGER NOT NULL CHECK (engagement BETWEEN 0 AND 4294967295)
);
CREATE TABLE IF NOT EXISTS embeddings (
  doc_id INTEGER PRIMARY KEY REFERENCES posts(doc_id),
  vec    BLOB NOT NULL                                         -- dim x f32, little endian
);
CREATE TABLE IF NOT EXISTS meta (key TEXT PRIMARY KEY, value TEXT NOT NULL);
"""


class InvalidPostText(ValueError):
    """DomainError::InvalidPostText (src/domain/error.rs) for rows that PostText::parse would reject"""


# Rust's str::trim() strips the Unicode White_Space set; Python's bare str.strip() also strips U+001C..U+001F
_WHITE_SPACE = "\t\n\x0b\x0c\r \x85\xa0\u1680\u2000\u2001\u2002\u2003\u2004\u2005\u2006\u2007\u2008\u2009\u200a\u2028\u2029\u202f\u205f\u3000"


def _dedupe(ids):
    """first-seen-order de-duplication of a query's term ids (the 64-term limit counts distinct terms)"""
    return np.array(list(dict.fromkeys(int(t) for t in ids)), dtype=np.uint32)


def parse_post_text(raw):
    """PostText::parse (social_post.rs:13-23): Unicode-whitespace trim, reject empty and > 10 000 chars."""
    t = raw.strip(_WHITE_SPACE)
    if not t:
        raise InvalidPostText("empty")
    if len(t) > MAX_POST_LEN:
        raise InvalidPostText("exceeds max length")
    return t


def open_store(path=":memory:", dim=None):
    """Opens (creating if needed) a post store.  `dim` fixes the embedding dimension of a new store."""
    conn = sqlite3.connect(path)
    conn.executescript(SCHEMA)
    if dim is not None:
        cur = conn.execute("SELECT value FROM meta WHERE key = 'dim'").fetchone()
        if cur is None:
            conn.execute("INSERT INTO meta (key, value) VALUES ('dim', ?)", (str(int(dim)),))
        elif int(cur[0]) != int(dim):
            raise ValueError("store holds %s-dim embeddings, not %d" % (cur[0], dim))
    conn.commit()
    return conn


def store_dim(conn):
    cur = conn.execute("SELECT value FROM meta WHERE key = 'dim'").fetchone()
    return int(cur[0]) if cur else None


def insert_posts(conn, posts, embeddings=None):
    """Appends posts (dicts with the SocialPost fields) and, optionally, their embeddings [n][dim] f32.
    doc ids continue from the current row count so they stay dense.  Text is validated like PostText::parse."""
    n0 = conn.execute("SELECT COUNT(*) FROM posts").fetchone()[0]
    rows = []
    for i, p in enumerate(posts):
        if p["source"] not in SOURCES:
            raise ValueError("unknown source %r" % (p["source"],))
        rows.append((n0 + i, p["id"], p["source"], p["author"], parse_post_text(p["text"]), p["created_at"], int(p["engagement"])))
    emb, dim = None, store_dim(conn)
    if embeddings is not None:
        emb = np.ascontiguousarray(embeddings, dtype="<f4")
        if emb.ndim != 2 or emb.shape[0] != len(rows) or (dim is not None and emb.shape[1] != dim):
            raise ValueError("embeddings must be [%d][%s] f32" % (len(rows), dim))
    with conn:  # one transaction: a rejected batch leaves the store unchanged
        conn.executemany("INSERT INTO posts (doc_id, id, source, author, text, created_at, engagement) VALUES (?,?,?,?,?,?,?)", rows)
        if emb is not None:
            if dim is None:
                conn.execute("INSERT INTO meta (key, value) VALUES ('dim', ?)", (str(emb.shape[1]),))
            conn.executemany("INSERT INTO embeddings (doc_id, vec) VALUES (?, ?)", ((n0 + i, emb[i].tobytes()) for i in range(len(rows))))
    return n0, len(rows)


def iter_posts(conn, chunk=4096):
    """(doc_id, id, text) in doc_id order, `chunk` rows at a time"""
    cur = conn.execute("SELECT doc_id, id, text FROM posts ORDER BY doc_id")
    while True:
        rows = cur.fetchmany(chunk)
        if not rows:
            return
        yield rows


def normalise_rows_f32(rows):
    """SPEC §2: stored rows are L2-normalised at build time -- f32 sum of squares in index order is not
    required (any association), so numpy's f32 pairwise sum is fine; a zero row stays zero."""
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    n = np.sqrt(np.einsum("ij,ij->i", rows, rows, dtype=np.float32)).astype(np.float32)
    n[n == 0] = 1.0
    return rows / n[:, None]


def to_bf16_rne(rows):
    """f32 -> bf16 bit patterns (uint16), round to nearest even (SPEC §2 bf16 path); finite inputs"""
    u = np.ascontiguousarray(rows, dtype=np.float32).view(np.uint32)
    return ((u + np.uint32(0x7FFF) + ((u >> np.uint32(16)) & np.uint32(1))) >> np.uint32(16)).astype(np.uint16)


def lift_csr(conn, chunk=4096):
    """Tokenises every post with the reference tokenizer and builds vocabulary + CSR on the host.
    -> (IndexBuilder, csr dict, post ids in doc order)"""
    b = hostlib.IndexBuilder()
    ids = []
    expect = 0
    for rows in iter_posts(conn, chunk):
        for doc_id, _, _ in rows:
            if doc_id != expect:
                raise ValueError("posts.doc_id must be dense 0..N-1 (gap at %d)" % expect)
            expect += 1
        ids.extend(r[1] for r in rows)
        b.add([r[2] for r in rows])
    csr = b.finish()
    return b, csr, ids


def lift_embeddings(conn, n_docs, dim, chunk=8192):
    """[n_docs][dim] f32, L2-normalised, in doc_id order; a post without an embedding is an error."""
    out = np.empty((n_docs, dim), dtype=np.float32)
    cur = conn.execute("SELECT doc_id, vec FROM embeddings ORDER BY doc_id")
    got = 0
    while True:
        rows = cur.fetchmany(chunk)
        if not rows:
            break
        for doc_id, blob in rows:
            if doc_id != got:
                raise ValueError("embedding missing for doc %d" % got)
            v = np.frombuffer(blob, dtype="<f4")
            if v.shape[0] != dim:
                raise ValueError("doc %d: embedding has %d components, store says %d" % (doc_id, v.shape[0], dim))
            out[got] = v
            got += 1
    if got != n_docs:
        raise ValueError("embedding missing for doc %d" % got)
    return normalise_rows_f32(out)


class StoreIndex:
    """A GPU hybrid index built from a post store: the drop-in search side of the path.

    search(texts, vectors, k) tokenises the query texts with the same tokenizer as the documents, maps tokens
    to term ids (unknown tokens are dropped), normalises the query vectors and calls oi_search_hybrid."""

    def __init__(self, conn, dtype=capi.DTYPE_F32, device=0, max_k=100, max_batch=16, k1=1.2, b=0.75):
        self.conn = conn
        self.builder, csr, self.post_ids = lift_csr(conn)
        self.n_docs, self.n_terms = csr["n_docs"], csr["n_terms"]
        self.dim = store_dim(conn)
        if self.dim is None:
            raise ValueError("store has no embeddings")
        rows = lift_embeddings(conn, self.n_docs, self.dim)
        self.ix = capi.GpuIndex(n_docs=self.n_docs, dim=self.dim, dtype=dtype, device=device, max_k=max_k, max_batch=max_batch)
        try:
            # a bf16 index stores the normalised rows rounded to nearest even (SPEC §2)
            self.ix.load_embeddings(to_bf16_rne(rows) if dtype == capi.DTYPE_BF16 else rows)
            self.ix.load_bm25(csr["term_offsets"], csr["doc_ids"], csr["tfs"], csr["doc_len"])
            self.ix.bm25_finalize(k1=k1, b=b)
        except Exception:
            self.ix.close()
            raise

    def query_terms(self, texts):
        """term ids of each query text: unknown tokens dropped, duplicates dropped in first-seen order (SPEC §3)"""
        return [_dedupe(self.builder.query_terms(t)) for t in texts]

    def search(self, texts, vectors, k, rrf_k=60):
        q = normalise_rows_f32(np.asarray(vectors, dtype=np.float32).reshape(len(texts), self.dim))
        ids, rrf, rc, rb = self.ix.search_hybrid(q, self.query_terms(texts), k, rrf_k)
        hits = []
        for j in range(len(texts)):
            hits.append([dict(doc_id=int(d), id=self.post_ids[int(d)], rrf=float(s), rank_cosine=int(c), rank_bm25=int(bm))
                         for d, s, c, bm in zip(ids[j], rrf[j], rc[j], rb[j]) if d != capi.NO_DOC])
        return hits

    def close(self):
        self.ix.close()
        self.builder.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()


# ---- document-sharded lift (one process per GPU, SPEC §5) ----------------------------------------
def lift_shard(conn, dist, rank=None, world=None, device="cpu"):
    """This rank's share of a document-sharded index, lifted out of the store every rank can read:

      * the rank tokenises ONLY its contiguous doc range (sharding.shard_range),
      * the vocabularies are unioned over the ranks (all_gather_object) so term ids are GLOBAL: the sorted union,
        i.e. exactly the ids the unsharded lift assigns,
      * the shard's CSR is re-indexed to the global vocabulary (absent terms = empty lists, doc ids shard-local),
      * df / doc-length sums are all-reduced (sharding.global_bm25_stats), so every shard scores with global
        N, avgdl and idf.

    `dist` = torch.distributed (initialised, gloo or nccl) or None.  -> dict(doc_base, n_local, n_docs_global, vocab,
    csr (global vocabulary), global_df, avgdl, post_ids (this shard's))"""
    from . import sharding
    n_total = conn.execute("SELECT COUNT(*) FROM posts").fetchone()[0]
    if dist is not None and dist.get_world_size() > 1:
        rank, world = dist.get_rank(), dist.get_world_size()
    else:
        dist, rank, world = None, 0, 1
    base, n_local = sharding.shard_range(n_total, world, rank)
    b = hostlib.IndexBuilder()
    ids = []
    cur = conn.execute("SELECT doc_id, id, text FROM posts WHERE doc_id >= ? AND doc_id < ? ORDER BY doc_id", (base, base + n_local))
    expect = base
    while True:
        rows = cur.fetchmany(4096)
        if not rows:
            break
        for doc_id, _, _ in rows:
            if doc_id != expect:
                raise ValueError("posts.doc_id must be dense 0..N-1 (gap at %d)" % expect)
            expect += 1
        ids.extend(r[1] for r in rows)
        b.add([r[2] for r in rows])
    if expect != base + n_local:
        raise ValueError("posts.doc_id must be dense 0..N-1 (gap at %d)" % expect)
    local = b.finish()
    local_vocab = [b.term(i) for i in range(local["n_terms"])]
    b.close()
    if dist is not None:
        gathered = [None] * world
        dist.all_gather_object(gathered, local_vocab)
        vocab = sorted(set().union(*gathered))
    else:
        vocab = local_vocab
    # local term id -> global term id (both lists are sorted, so the map is increasing)
    pos = {w: i for i, w in enumerate(vocab)}
    gid = np.array([pos[w] for w in local_vocab], dtype=np.int64)
    lens = np.zeros(len(vocab), dtype=np.uint64)
    if len(gid):
        lens[gid] = np.diff(local["term_offsets"])
    off = np.zeros(len(vocab) + 1, dtype=np.uint64)
    np.cumsum(lens, out=off[1:])
    csr = dict(term_offsets=off, doc_ids=local["doc_ids"], tfs=local["tfs"], doc_len=local["doc_len"],
               n_docs=n_local, n_terms=len(vocab))  # postings keep their order: lists are concatenated by term id
    gdf, avgdl, n = sharding.global_bm25_stats(dist, lens.astype(np.uint32), int(local["doc_len"].sum()), n_local, device=device)
    if n != n_total:
        raise ValueError("the ranks see %d posts in total, the store holds %d" % (n, n_total))
    return dict(doc_base=base, n_local=n_local, n_docs_global=n_total, vocab=vocab, csr=csr, global_df=gdf, avgdl=float(avgdl),
                post_ids=ids)


class ShardedStoreIndex:
    """StoreIndex over the GPUs of one box: every rank lifts its doc range (lift_shard), loads it into its GPU with
    GLOBAL BM25 statistics and joins the NCCL communicator; search() returns the same global lists on every rank."""

    def __init__(self, conn, dist, device_index=0, max_k=100, max_batch=16, k1=1.2, b=0.75):
        from . import sharding
        self.dist = dist if (dist is not None and dist.get_world_size() > 1) else None
        dev = "cuda:%d" % device_index
        sh = lift_shard(conn, self.dist, device=dev if self.dist else "cpu")
        self.vocab, self.doc_base, self.n_local, self.n_docs = sh["vocab"], sh["doc_base"], sh["n_local"], sh["n_docs_global"]
        self.term_id = {w: i for i, w in enumerate(self.vocab)}
        self.dim = store_dim(conn)
        if self.dim is None:
            raise ValueError("store has no embeddings")
        rows = np.empty((self.n_local, self.dim), dtype=np.float32)
        cur = conn.execute("SELECT doc_id, vec FROM embeddings WHERE doc_id >= ? AND doc_id < ? ORDER BY doc_id",
                           (self.doc_base, self.doc_base + self.n_local))
        got = 0
        for doc_id, blob in cur:
            if doc_id != self.doc_base + got:
                raise ValueError("embedding missing for doc %d" % (self.doc_base + got))
            rows[got] = np.frombuffer(blob, dtype="<f4")
            got += 1
        if got != self.n_local:
            raise ValueError("embedding missing for doc %d" % (self.doc_base + got))
        rows = normalise_rows_f32(rows)
        self.ix = capi.GpuIndex(n_docs=self.n_local, dim=self.dim, device=device_index, doc_base=self.doc_base, max_k=max_k, max_batch=max_batch)
        try:
            self.ix.load_embeddings(rows)
            c = sh["csr"]
            self.ix.load_bm25(c["term_offsets"], c["doc_ids"], c["tfs"], c["doc_len"])
            self.ix.bm25_finalize(k1=k1, b=b, avgdl=sh["avgdl"], n_docs_global=self.n_docs, global_df=sh["global_df"])
            if self.dist:
                self.ix.comm_init(self.dist.get_rank(), self.dist.get_world_size(), sharding.broadcast_unique_id(self.dist, device=dev))
        except Exception:
            self.ix.close()
            raise

    def query_terms(self, texts):
        return [_dedupe([self.term_id[t] for t in hostlib.tokenize(x) if t in self.term_id]) for x in texts]

    def search(self, texts, vectors, k, rrf_k=60):
        q = normalise_rows_f32(np.asarray(vectors, dtype=np.float32).reshape(len(texts), self.dim))
        return self.ix.search_hybrid(q, self.query_terms(texts), k, rrf_k)

    def close(self):
        self.ix.close()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()
