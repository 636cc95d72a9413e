"""C++ host layer on CPU: the tokenizer must equal the oracle's restatement of the reference
tokenizer (src/adapters/analyzer/lexicon.rs:54-58, pinned to the reference's goldens by
test_oracle_lexicon.py), and the index builder must produce the CSR a plain Python restatement
produces from the same tokens."""
import json
import os
from collections import Counter

import numpy as np
import pytest

import oracle as O
from openintel_b200 import hostlib


@pytest.fixture(scope="module", autouse=True)
def _built():
    import runpy
    runpy.run_path(os.path.join(os.path.dirname(hostlib.__file__), "host", "build.py"), run_name="__build__")


@pytest.fixture(scope="module")
def posts(golden_dir):
    with open(os.path.join(golden_dir, "reference_lexicon_goldens.json"), encoding="utf-8") as f:
        G = json.load(f)
    return [p["text"] for p in G["fixture_posts"]["posts"]]


EDGE = ["", "   ", "$AAPL 0dte calls!!", "Earnings miss shocks", "dismissal of claims", "MISS again",
        "café naïve Über", "Kelvin 300K", "İstanbul İ", "ȧb", "x" * 300,
        "emoji \U0001F680\U0001F680 moon", "tabs\tand\nnewlines\r\nhere", "MiXeD CaSe 123abc ABC123", "é", "42"]


def test_tokenizer_matches_oracle(posts):
    for t in posts + EDGE:
        assert hostlib.tokenize(t) == O.tokenize(t), repr(t)
    # the reference's whole-word idiom (src/domain/dip.rs:851-855)
    assert hostlib.tokenize("Earnings miss shocks") == ["earnings", "miss", "shocks"]
    assert "miss" not in hostlib.tokenize("dismissal of claims")
    assert hostlib.tokenize("$AAPL") == ["aapl"] and hostlib.tokenize("0dte") == ["0dte"]


def _python_index(texts):
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


def test_index_builder_matches_python_restatement(posts):
    rng = np.random.RandomState(3)
    words = ["moon", "calls", "puts", "dump", "AAPL", "tsla", "yolo", "0dte", "Earnings", "beat", "miss", "the", "a"]
    synth = [" ".join(rng.choice(words, size=rng.randint(0, 30))) for _ in range(500)]
    texts = posts + EDGE + synth
    b = hostlib.IndexBuilder()
    b.add(texts[:7])
    b.add(texts[7:])  # incremental adds keep doc ids in input order
    ix = b.finish()
    vocab, off, di, tf, dl = _python_index(texts)
    assert ix["n_docs"] == len(texts) and ix["n_terms"] == len(vocab)
    assert [b.term(i) for i in range(len(vocab))] == vocab
    assert np.array_equal(ix["term_offsets"], off)
    assert np.array_equal(ix["doc_ids"], di) and np.array_equal(ix["tfs"], tf) and np.array_equal(ix["doc_len"], dl)
    # CSR invariants the C ABI checks at load time
    for t in range(len(vocab)):
        l = ix["doc_ids"][int(off[t]):int(off[t + 1])]
        assert len(l) > 0 and np.all(np.diff(l.astype(np.int64)) > 0)
    assert int(tf.sum()) == int(dl.sum())
    # query side: tokens -> ids, unknown tokens dropped
    q = b.query_terms("AAPL to the MOON, zzzunknown calls")
    assert [vocab[i] for i in q] == ["aapl", "to", "the", "moon", "calls"]
    assert b.term_id("zzzunknown") == 0xFFFFFFFF
    b.close()


def test_empty_builder():
    b = hostlib.IndexBuilder()
    ix = b.finish()
    assert ix["n_docs"] == 0 and ix["n_terms"] == 0 and list(ix["term_offsets"]) == [0]
    b.add(["", "!!!"])
    ix = b.finish()
    assert ix["n_docs"] == 2 and ix["n_terms"] == 0 and list(ix["doc_len"]) == [0, 0]
    b.close()
