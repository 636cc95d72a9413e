"""Self-checks of the hybrid-path oracle (parity unpinned: SPEC.md is the only authority).
Each oracle function is cross-checked against an independent numpy / pure-Python restatement
of the same SPEC paragraph on small seeded inputs."""
import math

import numpy as np
import pytest

import oracle as O

M64 = (1 << 64) - 1


def _mix(z):
    z ^= z >> 30
    z = (z * 0xBF58476D1CE4E5B9) & M64
    z ^= z >> 27
    z = (z * 0x94D049BB133111EB) & M64
    z ^= z >> 31
    return z


def _hash(seed, stream, row, col):
    g = 0x9E3779B97F4A7C15
    base = _mix((seed + g * (stream + 1)) & M64)
    rk = _mix(base ^ ((row * 0xD1B54A32D192ED03) & M64))
    return _mix((rk + g * (col + 1)) & M64)


def test_hash_matches_python_model():
    for s, st, r, c in [(O.SEED, 0, 0, 0), (O.SEED, 1, 12345, 383), (1, 3, 2**40 + 7, 255), (0, 0, 0, 0)]:
        assert O.hash64(s, st, r, c) == _hash(s, st, r, c)


def test_synth_rows_normalised_and_reproducible():
    a = O.synth_rows_f32(64, 384)
    b = O.synth_rows_f32(16, 384, first=40)
    assert np.array_equal(a[40:56], b)  # any row regenerates independently
    assert np.allclose(np.linalg.norm(a.astype(np.float64), axis=1), 1.0, atol=1e-6)
    # python model of one row
    v = []
    for c in range(384):
        h = _hash(O.SEED, 0, 5, c)
        v.append(((h & 0xFFFF) + ((h >> 16) & 0xFFFF) + ((h >> 32) & 0xFFFF) + (h >> 48)) - 131070)
    norm = math.sqrt(float(sum(x * x for x in v)))
    want = np.array([np.float32(x / norm) for x in v], dtype=np.float32)
    assert np.array_equal(a[5], want)


def test_bf16_rounding_is_rne():
    x = np.array([1.0, 1.00390625, 1.01171875, -0.3333333, 3.0e-39, 65504.0], dtype=np.float32)
    got = np.array([O.lib().oio_f32_to_bf16(float(v)) for v in x], dtype=np.uint16)
    assert np.array_equal(got, O.f32_to_bf16(x))
    # 1 + 2^-8 is a tie between 0x3F80 and 0x3F81 -> even (0x3F80); 1 + 3*2^-8 ties -> 0x3F82
    assert got[1] == 0x3F80 and got[2] == 0x3F82
    rows = O.synth_rows_bf16(4, 64)
    assert np.array_equal(rows, O.f32_to_bf16(O.synth_rows_f32(4, 64)))


def test_planted_queries_hit_target():
    n, dim = 2000, 128
    rows = O.synth_rows_f32(n, dim)
    q, tgt = O.synth_planted_queries(8, dim, n)
    for j in range(8):
        sc = O.cosine_scores_f32(rows, q[j])
        assert int(np.argmax(sc)) == int(tgt[j])
        assert sc[tgt[j]] > 0.8


def test_cosine_and_topk_against_numpy():
    n, dim, k = 5000, 96, 50
    rows = O.synth_rows_f32(n, dim)
    q = O.synth_rows_f32(1, dim, stream=1)[0]
    sc = O.cosine_scores_f32(rows, q)
    assert np.allclose(sc, rows.astype(np.float64) @ q.astype(np.float64), atol=1e-12)
    ids, top, m = O.topk_f64(sc, k, doc_base=100)
    order = sorted(range(n), key=lambda i: (-sc[i], i))[:k]
    assert m == k and list(ids) == [100 + i for i in order]
    assert np.array_equal(top, sc[order])
    # bf16 path: the query is rounded too
    rb = O.synth_rows_bf16(n, dim)
    scb = O.cosine_scores_bf16(rb, q)
    want = O.bf16_to_f32(rb).astype(np.float64) @ O.bf16_to_f32(O.f32_to_bf16(q)).astype(np.float64)
    assert np.allclose(scb, want, atol=1e-12)


def test_topk_f32_ties_and_padding():
    s = np.array([0.5, 0.0, 0.5, -0.0, 0.25, 0.5, -1.0], dtype=np.float32)
    ids, sc, m = O.topk_f32(s, 4)
    assert list(ids) == [0, 2, 5, 4] and m == 4  # ties by ascending doc id
    ids, sc, m = O.topk_f32(s, 7)
    assert list(ids) == [0, 2, 5, 4, 1, 3, 6]  # -0.0 == +0.0, broken by id
    ids, sc, m = O.topk_f32(s, 6, only_positive=True)
    assert m == 4 and list(ids[4:]) == [O.NO_DOC] * 2 and list(sc[4:]) == [0.0, 0.0]
    assert O.key(0.0, 3) == O.key(-0.0, 3)
    assert O.key(1.0, 3) > O.key(1.0, 4) > O.key(0.5, 0) > O.key(-0.5, 0) > O.key(-1.0, 0)


def test_fast_baseline_agrees_with_checker():
    n, dim, k = 20000, 384, 100
    rows = O.synth_rows_f32(n, dim)
    q = O.synth_rows_f32(1, dim, stream=1)[0]
    ids_f, sc_f = O.cosine_topk_f32_fast(rows, q, k, n_threads=3)
    ids, sc, _ = O.topk_f64(O.cosine_scores_f32(rows, q), k)
    assert np.allclose(sc_f, sc, atol=1e-5)
    assert len(set(ids_f) ^ set(ids)) <= 2  # tie band only


def _np_bm25(corp, n_docs, q_terms, k1=np.float32(1.2), b=np.float32(0.75)):
    """numpy float32 restatement of SPEC §3, operation by operation."""
    to, di, tf, dl = corp["term_offsets"], corp["doc_ids"], corp["tfs"], corp["doc_len"]
    df = np.diff(to).astype(np.float64)
    idf = np.where(df > 0, np.log(1.0 + (n_docs - df + 0.5) / (df + 0.5)), 0.0).astype(np.float32)
    avgdl = np.float32(np.float64(dl.sum()) / n_docs)
    scores = np.zeros(n_docs, dtype=np.float32)
    for t in sorted(set(int(x) for x in q_terms)):
        if t >= len(to) - 1:
            continue
        sl = slice(int(to[t]), int(to[t + 1]))
        ratio = dl[di[sl]].astype(np.float32) / avgdl
        norm = k1 * ((np.float32(1.0) - b) + b * ratio)
        tff = tf[sl].astype(np.float32)
        w = idf[t] * ((tff * (k1 + np.float32(1.0))) / (tff + norm))
        scores[di[sl]] = scores[di[sl]] + w
    return scores


def test_bm25_bit_exact_against_numpy_float32():
    n_docs, vocab = 3000, 2000
    corp = O.synth_bm25_corpus(n_docs, vocab)
    to, di, tf, dl = corp["term_offsets"], corp["doc_ids"], corp["tfs"], corp["doc_len"]
    # CSR sanity: tf sums to doc lengths, doc ids ascending inside each list
    per_doc = np.zeros(n_docs, dtype=np.int64)
    np.add.at(per_doc, di, tf)
    assert np.array_equal(per_doc, dl)
    for t in range(0, vocab, 97):
        seg = di[int(to[t]):int(to[t + 1])]
        assert np.all(np.diff(seg.astype(np.int64)) > 0)
    idf = O.bm25_idf(n_docs, np.diff(to))
    w = O.bm25_weights(to, di, tf, dl, idf)
    qs = O.synth_query_terms(6, 8, corp["cdf"])
    qu = O.synth_query_terms(6, 8, corp["cdf"], uniform=True)
    for q in list(qs) + list(qu) + [np.array([0, 0, 1, 5000], dtype=np.uint32)]:
        got = O.bm25_score_dense(to, di, w, q, n_docs)
        want = _np_bm25(corp, n_docs, q)
        assert np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert all(len(set(r)) == 8 for r in qs) and all(len(set(r)) == 8 for r in qu)


def test_zipf_cdf_and_lengths():
    cdf = O.zipf_cdf(1000)
    h = np.cumsum(1.0 / np.arange(1, 1001))
    assert cdf[-1] == 1.0 and np.allclose(cdf[:-1], (h / h[-1])[:-1], rtol=1e-15)
    dl = O.synth_doc_lens(20000)
    assert dl.min() >= 3 and dl.max() <= 256 and 22.0 < dl.mean() < 26.0


def _py_rrf(a, b, k, rrf_k=60):
    ra = {d: i + 1 for i, d in enumerate(a) if d != O.NO_DOC}
    rb = {d: i + 1 for i, d in enumerate(b) if d != O.NO_DOC}
    out = []
    for d in set(ra) | set(rb):
        s = np.float32(0.0)
        x = np.float32(1.0) / np.float32(rrf_k + ra[d]) if d in ra else np.float32(0.0)
        y = np.float32(1.0) / np.float32(rrf_k + rb[d]) if d in rb else np.float32(0.0)
        s = np.float32(x + y)
        out.append((-float(s), d, s, ra.get(d, 0), rb.get(d, 0)))
    out.sort()
    return out[:k]


def test_rrf_against_python_model():
    rnd = np.random.RandomState(3)
    for k in (1, 10, 100):
        a = rnd.permutation(3 * k)[:k].astype(np.uint32)
        b = rnd.permutation(3 * k)[:k].astype(np.uint32)
        if k == 10:
            b[7:] = O.NO_DOC  # short BM25 list
        ids, val, rc, rb, m = O.rrf(a, b, k)
        want = _py_rrf(list(a), list(b), k)
        assert m == len(want)
        for i, (_, d, s, x, y) in enumerate(want):
            assert (ids[i], rc[i], rb[i]) == (d, x, y)
            assert val[i].view(np.uint32) == s.view(np.uint32)
        assert all(ids[m:] == O.NO_DOC)


# ---- full-size drivers (oracle/oracle_scale.c) and the timed CPU arm (oracle/oracle_fast.c) ----------------------
def test_scale_rows_and_cosine_topk_equal_the_direct_oracle():
    """the chunked, multi-threaded drivers return exactly what oracle.c returns in one piece"""
    assert np.array_equal(O.scale_synth_rows(5000, 128, bf16=True, first=7), O.synth_rows_bf16(5000, 128, first=7))
    assert np.array_equal(O.scale_synth_rows(3000, 64, first=3, n_threads=3), O.synth_rows_f32(3000, 64, first=3))
    q = O.synth_rows_f32(3, 128, stream=1)
    for bf16 in (True, False):
        rows = O.synth_rows_bf16(30000, 128, first=11) if bf16 else O.synth_rows_f32(30000, 128, first=11)
        ids, sc = O.scale_cosine_topk(30000, 128, q, 10, bf16=bf16, first=11)
        for j in range(3):
            all_sc = O.cosine_scores_bf16(rows, q[j]) if bf16 else O.cosine_scores_f32(rows, q[j])
            wi, ws, _ = O.topk_f64(all_sc, 10, doc_base=11)
            assert np.array_equal(wi, ids[j]) and np.array_equal(ws, sc[j])


@pytest.mark.parametrize("n_threads", [1, 3, 8])
def test_mini_index_scores_like_the_whole_csr(n_threads):
    """the CSR restricted to the touched terms (what the full-size GPU tests check BM25 against) holds exactly the
    whole CSR's lists of those terms, and scores / ranks bit for bit like the whole index"""
    n, vocab = 30000, 5000
    corp = O.synth_bm25_corpus(n, vocab)
    qt = np.concatenate([O.synth_query_terms(3, 8, corp["cdf"]), O.synth_query_terms(1, 8, corp["cdf"], uniform=True)])
    mini = O.scale_bm25_mini_index(n, vocab, qt.reshape(-1), cdf=corp["cdf"], n_threads=n_threads)
    assert mini["sum_doc_len"] == int(corp["doc_len"].sum()) and np.array_equal(mini["doc_len"], corp["doc_len"])
    for i, t in enumerate(mini["terms"]):
        a, b = int(corp["term_offsets"][t]), int(corp["term_offsets"][t + 1])
        c, d = int(mini["term_offsets"][i]), int(mini["term_offsets"][i + 1])
        assert np.array_equal(corp["doc_ids"][a:b], mini["doc_ids"][c:d]) and np.array_equal(corp["tfs"][a:b], mini["tfs"][c:d])
    w = O.bm25_weights(corp["term_offsets"], corp["doc_ids"], corp["tfs"], corp["doc_len"], O.bm25_idf(n, np.diff(corp["term_offsets"])))
    for j in range(4):
        s = O.bm25_score_dense(corp["term_offsets"], corp["doc_ids"], w, qt[j], n)
        wi, ws, _ = O.topk_f32(s, 20, only_positive=True)
        gi, gs, _ = O.scale_bm25_topk(mini, qt[j], 20)
        assert np.array_equal(wi, gi) and np.array_equal(ws.view(np.uint32), gs.view(np.uint32))


def test_cpu_arm_hybrid_batch_ranks_like_the_oracle():
    """the timed CPU baseline (f32 accumulation, blocked, OpenMP) returns the oracle's fused lists on a corpus where
    no two cosine scores are closer than the f32 / f64 accumulation difference"""
    n, vocab, dim, k = 20000, 3000, 128, 10
    corp = O.synth_bm25_corpus(n, vocab)
    w = O.bm25_weights(corp["term_offsets"], corp["doc_ids"], corp["tfs"], corp["doc_len"], O.bm25_idf(n, np.diff(corp["term_offsets"])))
    rows = O.synth_rows_bf16(n, dim)
    q = O.synth_rows_f32(4, dim, stream=1)
    qt = O.synth_query_terms(4, 8, corp["cdf"])
    ids, rrf, rc, rb = O.hybrid_batch_fast(rows, q, corp["term_offsets"], corp["doc_ids"], w, qt, k, n_threads=3)
    for j in range(4):
        ci, _, _ = O.topk_f64(O.cosine_scores_bf16(rows, q[j]), k)
        bi, _, _ = O.topk_f32(O.bm25_score_dense(corp["term_offsets"], corp["doc_ids"], w, qt[j], n), k, only_positive=True)
        e = O.rrf(ci, bi, k)
        assert np.array_equal(e[0], ids[j]) and np.array_equal(e[2], rc[j]) and np.array_equal(e[3], rb[j])
