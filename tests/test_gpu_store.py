"""BASELINE.json configs[0] end to end on the GPU: 10k synthetic posts in a SQLite store, 384-dim f32 embeddings,
BM25 + cosine + RRF top-10 for a single query.  The index is lifted out of the store by openintel_b200.store
(reference tokenizer -> CSR -> C ABI) and the answer is compared with the CPU oracle run on an independent
restatement of the same lift (oracle tokenizer + plain Python CSR): BM25 and RRF bit-exact, cosine within the
f32 tolerance of SPEC §2.  Self-written oracle: the reference has no retrieval code."""
import os
from collections import Counter

import numpy as np
import pytest

import oracle as O
from gpu_util import assert_ranked_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def oi():
    import __graft_entry__ as ge  # noqa: F401
    import openintel_b200
    openintel_b200.load_library()
    import runpy
    from openintel_b200 import hostlib
    runpy.run_path(os.path.join(os.path.dirname(hostlib.__file__), "host", "build.py"), run_name="__build__")
    return openintel_b200


def _python_csr(texts):
    docs = [O.tokenize(t) for t in texts]
    vocab = sorted({w for d in docs for w in d})
    tid = {w: i for i, w in enumerate(vocab)}
    lists = [[] for _ in vocab]
    for d, toks in enumerate(docs):
        for w, tf in sorted(Counter(toks).items()):
            lists[tid[w]].append((d, tf))
    off = np.zeros(len(vocab) + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(l) for l in lists])
    di = np.array([d for l in lists for d, _ in l], dtype=np.uint32)
    tf = np.array([f for l in lists for _, f in l], dtype=np.uint32)
    dl = np.array([len(d) for d in docs], dtype=np.uint32)
    return tid, off, di, tf, dl


@pytest.mark.parametrize("n,vocab,dim,k", [(10000, 50000, 384, 10), (3000, 500, 64, 100)])
def test_store_hybrid_matches_oracle(oi, tmp_path, n, vocab, dim, k):
    from openintel_b200 import store
    posts, _ = O.synth_posts(n, vocab, O.SEED)
    emb = O.synth_rows_f32(n, dim) * np.float32(3.5)  # un-normalised in the store: the lift normalises
    conn = store.open_store(str(tmp_path / "posts.db"), dim=dim)
    store.insert_posts(conn, posts, emb)
    # queries: words taken from documents (mixed case / punctuation), one unknown word, and a duplicate
    texts = ["%s %s, $%s zzzunknown %s" % tuple(posts[37]["text"].split()[:3] + [posts[37]["text"].split()[0]]),
             " ".join(posts[n // 2]["text"].split()[:8]),
             "nothingmatches here"]
    qv = O.synth_rows_f32(len(texts), dim, stream=1) * np.float32(0.25)
    with store.StoreIndex(conn, max_k=k, max_batch=4) as sx:
        assert sx.n_docs == n
        got = sx.search(texts, qv, k)
        one = sx.search(texts[:1], qv[:1], k)  # the single-query call of configs[0]
        q_terms = sx.query_terms(texts)
        terms_words = [[sx.builder.term(int(t)) for t in q] for q in q_terms]
    assert one[0] == got[0]
    # ---- oracle on an independent restatement of the lift
    tid, off, di, tf, dl = _python_csr([store.parse_post_text(p["text"]) for p in posts])
    w = O.bm25_weights(off, di, tf, dl, O.bm25_idf(n, np.diff(off)))
    rows = store.normalise_rows_f32(emb)
    qn = store.normalise_rows_f32(qv)
    for j, text in enumerate(texts):
        want_terms = list(dict.fromkeys(tid[t] for t in O.tokenize(text) if t in tid))  # de-duplicated, first-seen order
        assert [tid[x] for x in terms_words[j]] == want_terms
        cs = O.cosine_scores_f32(rows, qn[j])
        c_ids, c_sc, _ = O.topk_f64(cs, k)
        s = O.bm25_score_dense(off, di, w, np.asarray(want_terms, dtype=np.uint32), n)
        b_ids, _, _ = O.topk_f32(s, k, only_positive=True)
        e_ids, e_val, e_rc, e_rb, m = O.rrf(c_ids, b_ids, k)
        hits = got[j]
        assert len(hits) == m
        # cosine ranks may swap inside the f32 tie band (SPEC §2); when they do not, everything is identical
        g_ids = np.array([h["doc_id"] for h in hits], dtype=np.uint32)
        if np.array_equal(g_ids, e_ids[:m]):
            assert np.array_equal(np.array([h["rrf"] for h in hits], dtype=np.float32).view(np.uint32), e_val[:m].view(np.uint32))
            assert [h["rank_cosine"] for h in hits] == e_rc[:m].tolist() and [h["rank_bm25"] for h in hits] == e_rb[:m].tolist()
        else:  # a tie-band swap in the cosine list: same set of fused documents
            assert sorted(g_ids.tolist()) == sorted(e_ids[:m].tolist())
        assert [h["id"] for h in hits] == ["post-%d" % d for d in g_ids]
        if j == 2:
            assert all(h["rank_bm25"] == 0 for h in hits)  # no known term: the fused list is the cosine list
    conn.close()


def test_cpp_host_layer_lifts_the_store_and_searches(oi, tmp_path):
    """host_demo --store: SqlitePostStore -> IndexBuilder -> GpuHybridSearch in C++ returns the lists the Python
    lift returns on the same store file (BM25 / RRF bit-exact; the cosine list may differ inside the f32 tie band
    because the two lifts sum the row norms in a different order)"""
    import subprocess
    from openintel_b200 import hostlib, store
    n, vocab, dim, k = 4000, 3000, 96, 10
    posts, _ = O.synth_posts(n, vocab, O.SEED)
    emb = O.synth_rows_f32(n, dim) * np.float32(2.0)
    path = str(tmp_path / "posts.db")
    conn = store.open_store(path, dim=dim)
    store.insert_posts(conn, posts, emb)
    texts = [" ".join(posts[11]["text"].split()[:6]), " ".join(posts[n - 5]["text"].split()[:4]) + " zzzunknown"]
    qv = O.synth_rows_f32(len(texts), dim, stream=1)
    with store.StoreIndex(conn, max_k=k, max_batch=4) as sx:
        want = sx.search(texts, qv, k)
    conn.close()
    (tmp_path / "queries.txt").write_text("\n".join(texts) + "\n")
    qv.tofile(tmp_path / "qemb.f32")
    demo = os.path.join(os.path.dirname(hostlib.__file__), "host", "host_demo")
    r = subprocess.run([demo, "--store", path, str(tmp_path / "queries.txt"), str(tmp_path / "qemb.f32"), str(k)],
                       capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    got = {}
    for line in r.stdout.splitlines():
        f = line.split()
        if f[0] == "hit":
            got.setdefault(int(f[1]), []).append(dict(doc_id=int(f[3]), rrf=float(np.float32(f[4])), rank_cosine=int(f[5]), rank_bm25=int(f[6]), id=f[7]))
    assert "store %d posts dim %d" % (n, dim) in r.stdout
    for j in range(len(texts)):
        if [h["doc_id"] for h in got[j]] == [h["doc_id"] for h in want[j]]:
            assert got[j] == want[j]
        else:
            assert sorted(h["doc_id"] for h in got[j]) == sorted(h["doc_id"] for h in want[j])
        assert all(h["id"] == "post-%d" % h["doc_id"] for h in got[j])
    # a missing store is a DomainError::SourceFailure, not a crash
    r = subprocess.run([demo, "--store", str(tmp_path / "nope.db"), str(tmp_path / "queries.txt"), str(tmp_path / "qemb.f32"), str(k)],
                       capture_output=True, text=True, timeout=60)
    assert r.returncode == 1 and "post-store" in r.stderr


def test_store_bf16_index(oi, tmp_path):
    """a bf16 index lifted from the store: rows normalised in f32, rounded to nearest even, scanned on the bf16 path"""
    from openintel_b200 import store
    n, vocab, dim, k = 6000, 800, 128, 10
    posts, _ = O.synth_posts(n, vocab, O.SEED)
    emb = O.synth_rows_f32(n, dim) * np.float32(0.5)
    conn = store.open_store(str(tmp_path / "p.db"), dim=dim)
    store.insert_posts(conn, posts, emb)
    texts = [" ".join(posts[3]["text"].split()[:5])]
    qv = O.synth_planted_queries(1, dim, n)[0]
    with store.StoreIndex(conn, dtype=oi.DTYPE_BF16, max_k=k, max_batch=2) as sx:
        got = sx.search(texts, qv, k)[0]
        cos_ids, cos_sc = sx.ix.search_cosine(store.normalise_rows_f32(qv), k)
    conn.close()
    rows16 = O.f32_to_bf16(store.normalise_rows_f32(emb))
    allsc = O.cosine_scores_bf16(rows16, store.normalise_rows_f32(qv)[0])
    wi, ws, _ = O.topk_f64(allsc, k)
    assert_ranked_close(cos_ids[0], cos_sc[0], wi, ws, allsc, 2e-3)
    assert len(got) == k and len({h["doc_id"] for h in got}) == k
    assert {h["doc_id"] for h in got if h["rank_cosine"]} <= set(int(i) for i in cos_ids[0])


def test_search_posts_cli(oi, tmp_path):
    """`python -m openintel_b200.search`: the running twin of the Rust `openintel search` subcommand / `search_posts`
    MCP tool (rust/openintel-gpu/src/{cli_search,mcp_search}.rs): same request, same report, both renderings."""
    import json
    import subprocess
    import sys
    from openintel_b200 import search, store
    n, vocab, dim, k = 2000, 800, 64, 7
    posts, _ = O.synth_posts(n, vocab, O.SEED)
    emb = O.synth_rows_f32(n, dim)
    db = str(tmp_path / "posts.db")
    conn = store.open_store(db, dim=dim)
    store.insert_posts(conn, posts, emb)
    query = " ".join(posts[11]["text"].split()[:5])
    qv = (O.synth_rows_f32(1, dim, stream=1)[0] * np.float32(2.0)).astype("<f4")
    qfile = tmp_path / "q.f32"
    qfile.write_bytes(qv.tobytes())
    with store.StoreIndex(conn, max_k=k, max_batch=1) as sx:
        want = sx.search([query], qv[None, :], k)[0]
        rep = search.search_posts(sx, query, qv, k)
        assert [h["post_id"] for h in rep["hits"]] == [h["id"] for h in want] and rep["notes"] == []
        only_vec = search.search_posts(sx, "zzzunknownzzz", qv, k)
        assert only_vec["notes"] and all(h["rank_bm25"] is None for h in only_vec["hits"])
        with pytest.raises(ValueError):
            search.search_posts(sx, query, qv[:-1], k)
        with pytest.raises(ValueError):
            search.search_posts(sx, query, np.zeros(dim, np.float32), k)
    conn.close()
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "openintel_b200.search", query, "--store", db, "--embedding", str(qfile), "--k", str(k)]
    out = subprocess.run(cmd + ["--format", "json"], cwd=root, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    got = json.loads(out.stdout)
    assert [h["post_id"] for h in got["hits"]] == [h["id"] for h in want]
    assert [h.get("rank_cosine") for h in got["hits"]] == [h["rank_cosine"] or None for h in want]
    assert abs(got["hits"][0]["rrf"] - want[0]["rrf"]) < 1e-7 and got["disclaimer"]
    table = subprocess.run(cmd, cwd=root, capture_output=True, text=True, timeout=300)
    assert table.returncode == 0 and table.stdout.splitlines()[0].startswith("rank") and len(table.stdout.splitlines()) == 1 + len(want)
