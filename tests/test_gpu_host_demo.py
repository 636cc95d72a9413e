"""GPU: the C++ host layer end to end (openintel_b200/host/host_demo.cpp): IndexBuilder ->
GpuHybridSearch (HybridSearch port) -> hits, and GpuLexiconAnalyzer (PostAnalyzer port), checked
against the CPU oracle on the same posts."""
import json
import os
import subprocess

import numpy as np
import pytest

import oracle as O
from openintel_b200 import hostlib

pytestmark = pytest.mark.gpu


def test_cpp_host_layer_end_to_end(tmp_path, golden_dir):
    import runpy
    host = os.path.join(os.path.dirname(hostlib.__file__), "host")
    runpy.run_path(os.path.join(host, "build.py"), run_name="__build__")
    demo = os.path.join(host, "host_demo")
    assert os.path.exists(demo), "host_demo was not built (libopenintel_gpu.so missing?)"
    with open(os.path.join(golden_dir, "reference_lexicon_goldens.json"), encoding="utf-8") as f:
        fixture = [p["text"] for p in json.load(f)["fixture_posts"]["posts"]]
    rng = np.random.RandomState(11)
    words = ["moon", "calls", "puts", "dump", "aapl", "tsla", "yolo", "0dte", "earnings", "beat", "miss", "rocket", "bag",
             "short", "squeeze", "long", "the", "a", "of", "to"]
    texts = fixture + [" ".join(rng.choice(words, size=rng.randint(3, 25))) for _ in range(3000)]
    n, dim, k = len(texts), 64, 10
    emb = O.synth_rows_f32(n, dim)
    qtexts = ["AAPL calls to the moon", "short squeeze yolo", "earnings miss dump puts", "zzz unknown words only"]
    qemb = O.synth_rows_f32(len(qtexts), dim, stream=1)
    (tmp_path / "posts.txt").write_text("\n".join(texts) + "\n")
    (tmp_path / "queries.txt").write_text("\n".join(qtexts) + "\n")
    emb.tofile(tmp_path / "emb.f32")
    qemb.tofile(tmp_path / "qemb.f32")
    r = subprocess.run([demo, str(tmp_path / "posts.txt"), str(tmp_path / "emb.f32"), str(dim), str(tmp_path / "queries.txt"),
                        str(tmp_path / "qemb.f32"), str(k)], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0, r.stderr
    hits = {}
    signals = {}
    for line in r.stdout.splitlines():
        f = line.split()
        if f[0] == "hit":
            hits.setdefault(int(f[1]), []).append((int(f[3]), np.float32(f[4]), int(f[5]), int(f[6])))
        elif f[0] == "signal":
            signals[int(f[1])] = (float(f[2]), bool(int(f[3])))

    # oracle: same tokenizer -> same CSR (test_host_index_builder.py) -> BM25, cosine, RRF
    b = hostlib.IndexBuilder()
    b.add(texts)
    ix = b.finish()
    idf = O.bm25_idf(n, np.diff(ix["term_offsets"]))
    w = O.bm25_weights(ix["term_offsets"], ix["doc_ids"], ix["tfs"], ix["doc_len"], idf)
    for j, qt in enumerate(qtexts):
        terms = b.query_terms(qt)
        s = O.bm25_score_dense(ix["term_offsets"], ix["doc_ids"], w, terms, n)
        bm_ids, _, _ = O.topk_f32(s, k, only_positive=True)
        cos = O.cosine_scores_f32(emb, qemb[j])
        cos_ids, _, _ = O.topk_f64(cos, k)
        e_ids, e_val, e_rc, e_rb, m = O.rrf(cos_ids, bm_ids, k)
        got = hits.get(j, [])
        assert [h[0] for h in got] == [int(x) for x in e_ids[:m]], (j, got, e_ids)
        assert np.array_equal(np.array([h[1] for h in got], dtype=np.float32).view(np.uint32), e_val[:m].view(np.uint32))
        assert [h[2] for h in got] == list(e_rc[:m]) and [h[3] for h in got] == list(e_rb[:m])
    assert len(hits[3]) == k and all(h[3] == 0 for h in hits[3])  # no known term: cosine only
    # PostAnalyzer port: one signal per post, input order, equal to the reference-pinned oracle
    assert len(signals) == n
    for i in list(range(len(fixture))) + [100, 2999]:
        pol, spec, _, _ = O.lexicon_score(texts[i])
        assert signals[i] == (pol, spec), (i, texts[i])
    b.close()
