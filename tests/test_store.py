"""SQLite post store + index lift on CPU: the schema mirrors the reference's
SocialPost (src/domain/entities/social_post.rs:7-38), text validation replays the reference's own PostText tests
(social_post.rs:44-60), and the lifted CSR equals a plain Python restatement built with the oracle tokenizer."""
import os
from collections import Counter

import numpy as np
import pytest

import oracle as O
from openintel_b200 import hostlib, store


@pytest.fixture(scope="module", autouse=True)
def _built():
    import runpy
    runpy.run_path(os.path.join(os.path.dirname(hostlib.__file__), "host", "build.py"), run_name="__build__")


def test_post_text_rules_match_reference_tests():
    # social_post.rs:44-49
    assert store.parse_post_text("  hello  ") == "hello"
    with pytest.raises(store.InvalidPostText):
        store.parse_post_text("   ")
    with pytest.raises(store.InvalidPostText):
        store.parse_post_text("x" * 10_001)
    # social_post.rs:52-59: the limit counts chars, not bytes
    assert store.parse_post_text("é" * 10_000) == "é" * 10_000
    with pytest.raises(store.InvalidPostText):
        store.parse_post_text("é" * 10_001)
    # Rust's trim() strips White_Space only: U+001F is kept, U+3000 and U+0085 go
    assert store.parse_post_text("　\x85 a\x1f ") == "a\x1f"


def _posts(n, vocab=400):
    return O.synth_posts(n, vocab, O.SEED)[0]


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
    return vocab, off, di, tf, dl


def test_store_roundtrip_and_lift(tmp_path):
    path = str(tmp_path / "posts.db")
    posts = _posts(700)
    rng = np.random.RandomState(5)
    emb = rng.randn(700, 48).astype(np.float32)
    emb[13] = 0.0  # a zero embedding stays zero
    conn = store.open_store(path, dim=48)
    assert store.insert_posts(conn, posts[:300], emb[:300]) == (0, 300)
    assert store.insert_posts(conn, posts[300:], emb[300:]) == (300, 400)  # doc ids stay dense across inserts
    conn.close()
    conn = store.open_store(path, dim=48)  # reopen: everything is on disk
    assert store.store_dim(conn) == 48
    b, csr, ids = store.lift_csr(conn, chunk=128)
    assert ids == [p["id"] for p in posts]
    vocab, off, di, tf, dl = _python_csr([store.parse_post_text(p["text"]) for p in posts])
    assert csr["n_docs"] == 700 and csr["n_terms"] == len(vocab)
    assert [b.term(i) for i in range(len(vocab))] == vocab
    assert np.array_equal(csr["term_offsets"], off) and np.array_equal(csr["doc_ids"], di)
    assert np.array_equal(csr["tfs"], tf) and np.array_equal(csr["doc_len"], dl)
    rows = store.lift_embeddings(conn, 700, 48, chunk=100)
    norms = np.linalg.norm(rows.astype(np.float64), axis=1)
    assert np.allclose(np.delete(norms, 13), 1.0, atol=1e-6) and norms[13] == 0.0
    want = emb / np.maximum(np.linalg.norm(emb.astype(np.float64), axis=1), 1e-30)[:, None]
    assert np.allclose(np.delete(rows, 13, axis=0), np.delete(want, 13, axis=0), atol=1e-6)
    # the tokens of the generated text are the generator's term ids: "$W17," -> "w17"
    assert all(w.startswith("w") and w[1:].isdigit() for w in vocab)
    conn.close()


def test_store_rejects_bad_rows(tmp_path):
    conn = store.open_store(":memory:", dim=8)
    ok = dict(id="a", source="reddit", author="u", text="hello world", created_at="2026-10-18T00:00:00Z", engagement=1)
    with pytest.raises(ValueError):
        store.insert_posts(conn, [dict(ok, source="x")])
    with pytest.raises(store.InvalidPostText):
        store.insert_posts(conn, [dict(ok, text=" \n ")])
    with pytest.raises(ValueError):
        store.insert_posts(conn, [ok], np.zeros((1, 9), np.float32))
    store.insert_posts(conn, [ok], np.ones((1, 8), np.float32))
    with pytest.raises(Exception):  # SocialPost.id is unique; the rejected batch leaves nothing behind
        store.insert_posts(conn, [dict(ok, id="c"), ok], np.ones((2, 8), np.float32))
    assert conn.execute("SELECT COUNT(*) FROM posts").fetchone()[0] == 1
    store.insert_posts(conn, [dict(ok, id="b")])  # a post without an embedding ...
    with pytest.raises(ValueError):               # ... cannot be lifted into a hybrid index
        store.lift_embeddings(conn, 2, 8)
    conn.close()
    path = str(tmp_path / "d.db")
    store.open_store(path, dim=4).close()
    with pytest.raises(ValueError):  # the store's dimension is fixed at creation
        store.open_store(path, dim=5)


def test_cpp_store_reader_lifts_the_same_index(tmp_path):
    """host/openintel_store.hpp (C++, libsqlite3 bound with dlopen) vs openintel_b200.store (Python) on one store file"""
    path = str(tmp_path / "posts.db")
    posts = _posts(500)
    posts[7] = dict(posts[7], text="Ünïcode café — $AAPL 0dte \U0001F680 calls", id="post-unicode")
    emb = np.random.RandomState(9).randn(500, 40).astype(np.float32) * 7
    emb[3] = 0.0
    conn = store.open_store(path, dim=40)
    store.insert_posts(conn, posts, emb)
    b_py, csr_py, ids_py = store.lift_csr(conn)
    rows_py = store.lift_embeddings(conn, 500, 40)
    conn.close()
    cs = hostlib.CppPostStore(path)
    assert (cs.n_posts, cs.dim) == (500, 40)
    b = hostlib.IndexBuilder()
    assert cs.lift_posts(b) == 500
    csr = b.finish()
    for name in ("term_offsets", "doc_ids", "tfs", "doc_len"):
        assert np.array_equal(csr[name], csr_py[name]), name
    assert [b.term(i) for i in range(csr["n_terms"])] == [b_py.term(i) for i in range(csr_py["n_terms"])]
    assert [b.post_id(d) for d in (0, 7, 499)] == [ids_py[0], "post-unicode", ids_py[499]]
    rows = cs.embeddings()
    assert np.all(rows[3] == 0) and np.allclose(rows, rows_py, atol=2e-7)  # f32 norms summed in a different order
    cs.close()
    with pytest.raises(OSError):
        hostlib.CppPostStore(str(tmp_path / "missing.db"))


def test_bf16_rounding_is_nearest_even():
    x = np.random.RandomState(2).randn(500, 32).astype(np.float32)
    x[0, :4] = np.array([0x3F808000, 0x3F818000, 0x3F80FFFF, 0x3F807FFF], dtype=np.uint32).view(np.float32)  # ties and neighbours
    got = store.to_bf16_rne(x)
    assert got.dtype == np.uint16 and np.array_equal(got, O.f32_to_bf16(x))
    assert [hex(v) for v in got[0, :4]] == ["0x3f80", "0x3f82", "0x3f81", "0x3f80"]
