"""The CPU oracle against the PUBLISHED definitions, computed with independent arithmetic (decimal at 50 digits,
exact fractions) on hand-sized inputs.  The reference has no retrieval code, so these known-answer
tests are what anchors the oracle -- and through it the GPU -- to the algorithms docs/SPEC.md names:
Robertson / Sparck Jones BM25 with the Lucene idf, reciprocal rank fusion with k = 60 (Cormack, Clarke, Buettcher,
SIGIR 2009), cosine of L2-normalised vectors, and the (score desc, doc id asc) order north_star prescribes."""
from decimal import Decimal, getcontext
from fractions import Fraction

import numpy as np

import oracle as O

getcontext().prec = 50


def _csr(docs, vocab):
    """docs: list of token-id lists -> CSR like the index builder's"""
    lists = [[] for _ in range(vocab)]
    for d, toks in enumerate(docs):
        for t in sorted(set(toks)):
            lists[t].append((d, toks.count(t)))
    off = np.zeros(vocab + 1, dtype=np.uint64)
    off[1:] = np.cumsum([len(l) for l in lists])
    di = np.array([d for l in lists for d, _ in l], dtype=np.uint32)
    tf = np.array([f for l in lists for _, f in l], dtype=np.uint32)
    dl = np.array([len(x) for x in docs], dtype=np.uint32)
    return off, di, tf, dl


def _bm25_decimal(docs, vocab, query, k1="1.2", b="0.75"):
    k1, b = Decimal(k1), Decimal(b)
    n = len(docs)
    avgdl = Decimal(sum(len(d) for d in docs)) / Decimal(n)
    out = []
    for d in docs:
        s = Decimal(0)
        for t in sorted(set(query)):
            if t >= vocab:
                continue
            df = sum(1 for x in docs if t in x)
            tf = d.count(t)
            if df == 0 or tf == 0:
                continue
            idf = (Decimal(1) + (Decimal(n) - Decimal(df) + Decimal("0.5")) / (Decimal(df) + Decimal("0.5"))).ln()
            s += idf * (Decimal(tf) * (k1 + 1)) / (Decimal(tf) + k1 * (Decimal(1) - b + b * Decimal(len(d)) / avgdl))
        out.append(s)
    return out


def test_bm25_matches_the_published_formula():
    # 6 tiny "posts" over a 7-term vocabulary; term 0 is in every document, term 6 in none
    docs = [[0, 1, 1, 2], [0, 2, 2, 2, 3, 4, 4], [0, 1, 5], [0, 0, 0, 3], [0, 4, 5, 5, 5, 5, 1, 2, 3], [0, 3]]
    vocab = 7
    off, di, tf, dl = _csr(docs, vocab)
    idf = O.bm25_idf(len(docs), np.diff(off))
    w = O.bm25_weights(off, di, tf, dl, idf)
    for query in ([1, 2], [0], [5, 5, 3], [4, 6, 9], [0, 1, 2, 3, 4, 5], [6]):
        got = O.bm25_score_dense(off, di, w, np.array(query, dtype=np.uint32), len(docs))
        want = _bm25_decimal(docs, vocab, query)
        for g, x in zip(got, want):
            assert abs(Decimal(float(g)) - x) <= Decimal("2e-6") * max(x, Decimal(1)), (query, g, x)
        # and the ranking the oracle derives from its f32 scores is the exact one (no near-ties in this corpus)
        ids, sc, m = O.topk_f32(got, 3, only_positive=True)
        exact = sorted([(-x, d) for d, x in enumerate(want) if x > 0])[:3]
        assert [int(i) for i in ids[:m]] == [d for _, d in exact]
    # Lucene idf of a term in every document is ln(1 + 0.5 / (N + 0.5)) > 0: never negative, never zero
    assert idf[0] > 0 and abs(Decimal(float(idf[0])) - (Decimal(1) + Decimal("0.5") / Decimal("6.5")).ln()) < Decimal("1e-7")
    assert idf[6] == 0  # df = 0


def test_rrf_matches_cormack_et_al():
    # RRFscore(d) = sum over the rankings that contain d of 1 / (k + rank), k = 60, ranks from 1
    cos = [11, 22, 33, 44]
    bm = [33, 11, 55]
    ids, val, rc, rb, m = O.rrf(np.array(cos, np.uint32), np.array(bm, np.uint32), 5)
    exact = {}
    for lst in (cos, bm):
        for r, d in enumerate(lst, 1):
            exact[d] = exact.get(d, Fraction(0)) + Fraction(1, 60 + r)
    order = sorted(exact, key=lambda d: (-exact[d], d))
    assert m == 5 and [int(x) for x in ids[:m]] == order == [11, 33, 22, 55, 44]
    for d, v in zip(ids[:m], val[:m]):
        assert abs(Fraction(float(v)) - exact[int(d)]) < Fraction(1, 10 ** 8)
    assert list(rc[:m]) == [1, 3, 2, 0, 4] and list(rb[:m]) == [2, 1, 0, 3, 0]  # 0 = absent from that list
    # equal fused scores order by ascending doc id: d at cosine rank 2 only vs d' at BM25 rank 2 only
    ids, val, _, _, m = O.rrf(np.array([9, 7], np.uint32), np.array([8, 3], np.uint32), 4)
    assert [int(x) for x in ids[:m]] == [8, 9, 3, 7] and val[0] == val[1] and val[2] == val[3]


def test_cosine_and_order_known_answers():
    # rows of a 4-d orthonormal frame and two 3-4-5 combinations: every dot product is exact in f32
    rows = np.array([[1, 0, 0, 0], [0, 1, 0, 0], [0.6, 0.8, 0, 0], [0.8, 0.6, 0, 0], [0, 0, 1, 0], [0.6, 0.8, 0, 0]], dtype=np.float32)
    q = np.array([0, 1, 0, 0], dtype=np.float32)
    sc = O.cosine_scores_f32(rows, q)
    assert np.allclose(sc, [0, 1, 0.8, 0.6, 0, 0.8], atol=1e-7)
    ids, s, m = O.topk_f64(sc, 4)
    assert [int(x) for x in ids] == [1, 2, 5, 3]  # the tie between docs 2 and 5 goes to the smaller id
    ids, s, m = O.topk_f64(sc, 6, doc_base=100)
    assert [int(x) for x in ids] == [101, 102, 105, 103, 100, 104]  # zero scores are ranked too (cosine keeps them)
