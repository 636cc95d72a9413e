"""GPU parity: BM25 (index load / device synth / weights / blocked scoring + fused top-k), RRF and
the hybrid call through the C ABI vs the CPU oracle.  SPEC §3/§4: bit-exact, no tie band."""
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
    return openintel_b200


def _oracle_lists(corp, w, queries, n_docs, k, doc_base=0):
    ids, scs = [], []
    for q in queries:
        sc = O.bm25_score_dense(corp["term_offsets"], corp["doc_ids"], w, np.asarray(q, dtype=np.uint32), n_docs)
        i, s, _ = O.topk_f32(sc, k, only_positive=True, doc_base=doc_base)
        ids.append(i)
        scs.append(s)
    return np.array(ids), np.array(scs)


def _weights(corp, n_docs):
    idf = O.bm25_idf(n_docs, np.diff(corp["term_offsets"]))
    return O.bm25_weights(corp["term_offsets"], corp["doc_ids"], corp["tfs"], corp["doc_len"], idf)


def test_device_synth_csr_is_bit_identical_to_oracle(oi):
    n, vocab = 6000, 5000
    corp = O.synth_bm25_corpus(n, vocab, first=250)
    with oi.GpuIndex(n_docs=n, dim=64, doc_base=250) as ix:
        ix.synth_bm25(O.SEED, vocab, corp["cdf"])
        df, sdl, npost = ix.bm25_local_stats()
        assert npost == len(corp["doc_ids"]) and sdl == int(corp["doc_len"].sum())
        assert np.array_equal(df, np.diff(corp["term_offsets"]).astype(np.uint32))
        ix.bm25_finalize()
        got = ix.read_bm25(npost)
    for name in ("term_offsets", "doc_ids", "tfs", "doc_len"):
        assert np.array_equal(got[name], corp[name]), name
    w = _weights(corp, n)
    assert np.array_equal(got["weights"].view(np.uint32), w.view(np.uint32))


@pytest.mark.parametrize("n,vocab,k,nq,groups", [(20000, 3000, 100, 24, 0), (70000, 20000, 10, 5, 0), (5000, 300, 1, 3, 0),
                                                 (40000, 1000, 100, 1, 0), (9000, 2000, 1000, 2, 0), (50000, 5000, 50, 40, 4),
                                                 (50000, 5000, 50, 7, 1), (33000, 4000, 100, 16, 16)])
def test_bm25_parity_bit_exact(oi, n, vocab, k, nq, groups):
    corp = O.synth_bm25_corpus(n, vocab)
    w = _weights(corp, n)
    qz = O.synth_query_terms(nq, 8, corp["cdf"])
    qu = O.synth_query_terms(nq, 8, corp["cdf"], uniform=True)
    with oi.GpuIndex(n_docs=n, dim=64, max_k=k, max_batch=nq, doc_base=11) as ix:
        ix.load_bm25(corp["term_offsets"], corp["doc_ids"], corp["tfs"], corp["doc_len"])
        ix.bm25_finalize()
        if groups:
            ix.set_option("bm25_variant", groups)
        for qs in (qz, qu):
            ids, sc = ix.search_bm25(qs, k)
            wi, ws = _oracle_lists(corp, w, qs, n, k, doc_base=11)
            assert np.array_equal(ids, wi)
            assert np.array_equal(sc.view(np.uint32), ws.view(np.uint32))


@pytest.mark.parametrize("warps,block_docs,slots", [(16, 1024, 8), (24, 1024, 0), (8, 2048, 3), (16, 1024, 1), (20, 1024, 2), (3, 4096, 16)])
def test_bm25_tuning_options_bit_exact(oi, warps, block_docs, slots):
    """warps per CTA, documents per block and staged chunks per warp change the schedule, never the result"""
    n, vocab, k, nq = 60000, 2500, 100, 33
    corp = O.synth_bm25_corpus(n, vocab)
    w = _weights(corp, n)
    with oi.GpuIndex(n_docs=n, dim=64, max_k=k, max_batch=nq) as ix:
        ix.load_bm25(corp["term_offsets"], corp["doc_ids"], corp["tfs"], corp["doc_len"])
        ix.bm25_finalize()
        ix.set_option("bm25_warps", warps)
        ix.set_option("bm25_block_docs", block_docs)
        ix.set_option("bm25_stage_slots", slots)
        for uniform in (False, True):
            qs = O.synth_query_terms(nq, 12, corp["cdf"], uniform=uniform)
            ids, sc = ix.search_bm25(qs, k)
            wi, ws = _oracle_lists(corp, w, qs, n, k)
            assert np.array_equal(ids, wi)
            assert np.array_equal(sc.view(np.uint32), ws.view(np.uint32))


@pytest.mark.parametrize("k", [1, 7, 100, 128, 256, 300])
def test_bm25_cold_start_bound_bit_exact(oi, k):
    """single-query calls start every block without a threshold: the group-maxima bound (k <= 256) and the plain
    overflow path (bound disabled, or k > 256) must return the same exact lists; few-term vocabularies give many ties"""
    n, vocab = 50000, 60
    corp = O.synth_bm25_corpus(n, vocab)
    w = _weights(corp, n)
    qs = [[0], [1, 2], [59], [3, 30, 58, 59], [40, 41, 42, 43, 44, 45, 46, 47]]
    with oi.GpuIndex(n_docs=n, dim=64, max_k=k, max_batch=1) as ix:
        ix.load_bm25(corp["term_offsets"], corp["doc_ids"], corp["tfs"], corp["doc_len"])
        ix.bm25_finalize()
        for off in (0, 1):
            ix.set_option("bm25_no_cold_bound", off)
            for q in qs:
                ids, sc = ix.search_bm25([q], k)
                wi, ws = _oracle_lists(corp, w, [q], n, k)
                assert np.array_equal(ids, wi), (off, q)
                assert np.array_equal(sc.view(np.uint32), ws.view(np.uint32))


@pytest.mark.parametrize("n,vocab,k,nq", [(30000, 800, 100, 20), (70000, 20000, 10, 5), (9000, 2000, 1000, 2)])
def test_bm25_sparse_only_path_bit_exact(oi, n, vocab, k, nq):
    """bm25_variant = 100 builds no dense weight columns: every term, however common, walks its postings"""
    corp = O.synth_bm25_corpus(n, vocab)
    w = _weights(corp, n)
    qs = O.synth_query_terms(nq, 8, corp["cdf"])
    with oi.GpuIndex(n_docs=n, dim=64, max_k=k, max_batch=nq) as ix:
        ix.load_bm25(corp["term_offsets"], corp["doc_ids"], corp["tfs"], corp["doc_len"])
        ix.set_option("bm25_variant", 100)
        ix.bm25_finalize()
        ids, sc = ix.search_bm25(qs, k)
    wi, ws = _oracle_lists(corp, w, qs, n, k)
    assert np.array_equal(ids, wi)
    assert np.array_equal(sc.view(np.uint32), ws.view(np.uint32))


def test_bm25_edge_cases(oi):
    n, vocab, k = 12000, 1500, 20
    corp = O.synth_bm25_corpus(n, vocab)
    w = _weights(corp, n)
    queries = [
        [],                                  # empty query: all padding
        [vocab + 5, 4000000000],             # unknown terms only
        [3, 3, 3, 3],                        # repeated term counts once
        [0, 1, 2, vocab + 1, 2, 1],          # mixed
        list(range(64)),                     # the maximum number of terms
        [vocab - 1],                         # rarest term: list shorter than k?
    ]
    with oi.GpuIndex(n_docs=n, dim=64, max_k=k, max_batch=len(queries)) as ix:
        ix.load_bm25(corp["term_offsets"], corp["doc_ids"], corp["tfs"], corp["doc_len"])
        ix.bm25_finalize()
        ids, sc = ix.search_bm25(queries, k)
        wi, ws = _oracle_lists(corp, w, queries, n, k)
        assert np.array_equal(ids, wi)
        assert np.array_equal(sc.view(np.uint32), ws.view(np.uint32))
        assert np.all(ids[0] == oi.NO_DOC) and np.all(sc[0] == 0)
        # more than 64 distinct known terms: the first 64 (first-seen order) are scored (SPEC §3), on both paths
        ids65, sc65 = ix.search_bm25([list(range(65))], k)
        assert np.array_equal(ids65[0], ids[4]) and np.array_equal(sc65[0], sc[4])
        # a call may not carry more raw term ids than the staging array holds (64 x max_batch)
        with pytest.raises(oi.OiError) as e:
            ix.search_bm25([[1] * (64 * len(queries) + 1)], k)
        assert e.value.status == 1
        # nq == 0 is a no-op
        ids0, _ = ix.search_bm25([], k)
        assert ids0.shape == (0, k)


def test_load_bm25_rejects_a_malformed_csr(oi):
    """The postings are validated on the device after the copy: doc ids inside the shard, strictly ascending per list."""
    n, vocab = 3000, 200
    corp = O.synth_bm25_corpus(n, vocab)
    off = corp["term_offsets"]
    with oi.GpuIndex(n_docs=n, dim=64, max_k=5) as ix:
        long_t = int(np.argmax(np.diff(off)))
        p = int(off[long_t]) + 3
        bad = corp["doc_ids"].copy()
        bad[p] = n + 7
        with pytest.raises(oi.OiError) as e:
            ix.load_bm25(off, bad, corp["tfs"], corp["doc_len"])
        assert e.value.status == 1 and "outside the shard" in str(e.value) and str(p) in str(e.value)
        bad = corp["doc_ids"].copy()
        bad[p] = bad[p - 1]
        with pytest.raises(oi.OiError) as e:
            ix.load_bm25(off, bad, corp["tfs"], corp["doc_len"])
        assert e.value.status == 1 and "ascending" in str(e.value)
        # the first posting of a list may be smaller than the last posting of the previous list
        ix.load_bm25(off, corp["doc_ids"], corp["tfs"], corp["doc_len"])
        ix.bm25_finalize()
        ids, _ = ix.search_bm25([[long_t]], 5)
        assert ids[0][0] != oi.NO_DOC


def test_bm25_before_finalize_is_a_state_error(oi):
    corp = O.synth_bm25_corpus(500, 100)
    with oi.GpuIndex(n_docs=500, dim=64, max_k=5) as ix:
        with pytest.raises(oi.OiError) as e:
            ix.search_bm25([[1, 2]], 5)
        assert e.value.status == 5
        ix.load_bm25(corp["term_offsets"], corp["doc_ids"], corp["tfs"], corp["doc_len"])
        with pytest.raises(oi.OiError) as e:
            ix.search_bm25([[1, 2]], 5)
        assert e.value.status == 5


def test_bm25_global_statistics_make_shards_equal_the_whole(oi):
    """SPEC §5: a sharded index scored with GLOBAL N / df / avgdl reproduces the unsharded lists."""
    n, vocab, k, G = 30000, 4000, 50, 3
    corp = O.synth_bm25_corpus(n, vocab)
    w = _weights(corp, n)
    qs = O.synth_query_terms(6, 8, corp["cdf"])
    want_ids, want_sc = _oracle_lists(corp, w, qs, n, k)
    gdf = np.diff(corp["term_offsets"]).astype(np.uint32)
    avgdl = O.bm25_avgdl(corp["doc_len"])
    per = (n + G - 1) // G
    all_ids, all_sc = [], []
    for r in range(G):
        lo, hi = r * per, min(n, (r + 1) * per)
        with oi.GpuIndex(n_docs=hi - lo, dim=64, max_k=k, max_batch=6, doc_base=lo) as ix:
            ix.synth_bm25(O.SEED, vocab, corp["cdf"])
            ix.bm25_finalize(avgdl=avgdl, n_docs_global=n, global_df=gdf)
            ids, sc = ix.search_bm25(qs, k)
        all_ids.append(ids)
        all_sc.append(sc)
    for j in range(6):
        cand = [(O.key(float(s), int(d)), int(d), s) for r in range(G) for d, s in zip(all_ids[r][j], all_sc[r][j]) if d != oi.NO_DOC]
        cand.sort(key=lambda t: -t[0])
        got = [c[1] for c in cand[:k]]
        assert got == [int(x) for x in want_ids[j] if x != oi.NO_DOC]
        assert [c[2].view(np.uint32) for c in cand[:k]] == [s.view(np.uint32) for s, d in zip(want_sc[j], want_ids[j]) if d != oi.NO_DOC]


@pytest.mark.parametrize("k", [10, 100])
def test_hybrid_matches_oracle(oi, k):
    n, dim, vocab, nq = 30000, 384, 3000, 6
    rows = O.synth_rows_f32(n, dim)
    corp = O.synth_bm25_corpus(n, vocab)
    w = _weights(corp, n)
    qv = np.concatenate([O.synth_rows_f32(3, dim, stream=1), O.synth_planted_queries(3, dim, n)[0]])
    qt = O.synth_query_terms(nq, 8, corp["cdf"])
    with oi.GpuIndex(n_docs=n, dim=dim, max_k=k, max_batch=nq) as ix:
        ix.synth_embeddings(O.SEED)
        ix.synth_bm25(O.SEED, vocab, corp["cdf"])
        ix.bm25_finalize()
        cos_ids, cos_sc = ix.search_cosine(qv, k)
        bm_ids, bm_sc = ix.search_bm25(qt, k)
        ids, rrf, rc, rb = ix.search_hybrid(qv, qt, k)
    wb_ids, wb_sc = _oracle_lists(corp, w, qt, n, k)
    assert np.array_equal(bm_ids, wb_ids)
    for j in range(nq):
        allsc = O.cosine_scores_f32(rows, qv[j])
        wi, ws, _ = O.topk_f64(allsc, k)
        swaps = assert_ranked_close(cos_ids[j], cos_sc[j], wi, ws, allsc, 1e-5)
        # RRF is bit-exact GIVEN the two input lists (SPEC §4): fuse the GPU's own lists with the oracle
        oi_ids, oi_val, oi_rc, oi_rb, m = O.rrf(cos_ids[j], bm_ids[j], k)
        assert np.array_equal(ids[j], oi_ids)
        assert np.array_equal(rrf[j].view(np.uint32), oi_val.view(np.uint32))
        assert np.array_equal(rc[j], oi_rc) and np.array_equal(rb[j], oi_rb)
        if swaps == 0:  # and end to end when the cosine list had no tie-band swap
            e_ids, e_val, _, _, _ = O.rrf(wi, wb_ids[j], k)
            assert np.array_equal(ids[j], e_ids) and np.array_equal(rrf[j].view(np.uint32), e_val.view(np.uint32))


def test_host_buffer_copy_paths_agree(oi):
    """small host-buffer calls go through one pinned staging block each way, large ones copy from / to the caller's
    buffers directly: same results either way (cosine, BM25, hybrid; ragged term lists)"""
    n, dim, vocab, k, nq = 20000, 128, 2000, 25, 9
    corp = O.synth_bm25_corpus(n, vocab)
    qv = O.synth_rows_f32(nq, dim, stream=1)
    qt = [list(map(int, t[: 1 + j % 8])) for j, t in enumerate(O.synth_query_terms(nq, 8, corp["cdf"]))]
    qt[4] = []
    with oi.GpuIndex(n_docs=n, dim=dim, max_k=k, max_batch=nq) as ix:
        ix.synth_embeddings(O.SEED)
        ix.load_bm25(corp["term_offsets"], corp["doc_ids"], corp["tfs"], corp["doc_len"])
        ix.bm25_finalize()
        res = []
        for off in (0, 1):
            ix.set_option("no_pinned_staging", off)
            res.append((ix.search_cosine(qv, k), ix.search_bm25(qt, k), ix.search_hybrid(qv, qt, k)))
    for a, b in zip(res[0], res[1]):
        for x, y in zip(a, b):
            assert np.array_equal(x.view(np.uint32), y.view(np.uint32))


def test_rrf_short_and_disjoint_lists(oi):
    """BM25 list shorter than k (padding) and a query with no lexical match at all."""
    n, dim, vocab, k = 4000, 128, 50000, 30
    corp = O.synth_bm25_corpus(n, vocab)
    qv = O.synth_rows_f32(2, dim, stream=1)
    rare = int(np.nonzero(np.diff(corp["term_offsets"]) == 1)[0][0])
    qt = [[rare], [vocab + 9]]
    with oi.GpuIndex(n_docs=n, dim=dim, max_k=k, max_batch=2) as ix:
        ix.synth_embeddings(O.SEED)
        ix.load_bm25(corp["term_offsets"], corp["doc_ids"], corp["tfs"], corp["doc_len"])
        ix.bm25_finalize()
        cos_ids, _ = ix.search_cosine(qv, k)
        bm_ids, _ = ix.search_bm25(qt, k)
        ids, rrf, rc, rb = ix.search_hybrid(qv, qt, k)
    assert (bm_ids[0] != oi.NO_DOC).sum() == 1 and np.all(bm_ids[1] == oi.NO_DOC)
    for j in range(2):
        e_ids, e_val, e_rc, e_rb, _ = O.rrf(cos_ids[j], bm_ids[j], k)
        assert np.array_equal(ids[j], e_ids) and np.array_equal(rrf[j].view(np.uint32), e_val.view(np.uint32))
        assert np.array_equal(rc[j], e_rc) and np.array_equal(rb[j], e_rb)
    assert np.all(rb[1] == 0)


def test_full_size_config3_properties(oi):
    """BASELINE config 3 scale (10M docs, 1M-term Zipf vocabulary, 8-term queries) through size-independent
    properties: lists are sorted and duplicate-free, every reported score equals the oracle's f32 sum of the
    folded weights read back from the device (bit for bit), and the result equals the merge of the results of
    two half-corpus shards scored with GLOBAL statistics (checksum of checksums)."""
    n, vocab, k, nq = 10_000_000, 1_000_000, 100, 32
    cdf = O.zipf_cdf(vocab)
    qs = np.concatenate([O.synth_query_terms(nq // 2, 8, cdf), O.synth_query_terms(nq // 2, 8, cdf, uniform=True)])
    with oi.GpuIndex(n_docs=n, dim=64, max_k=k, max_batch=nq) as ix:
        ix.synth_bm25(O.SEED, vocab, cdf)
        df, sdl, npost = ix.bm25_local_stats()
        ix.bm25_finalize()
        ids, sc = ix.search_bm25(qs, k)
        got = ix.read_bm25(npost)
    avgdl = np.float32(np.float64(sdl) / n)
    for j in range(nq):
        valid = ids[j] != oi.NO_DOC
        assert np.all(np.diff(sc[j][valid]) <= 0) and len(set(ids[j][valid])) == int(valid.sum())
        # ties broken by ascending doc id
        same = np.diff(sc[j][valid]) == 0
        assert np.all(np.diff(ids[j][valid].astype(np.int64))[same] > 0)
    # recompute a sample of reported scores from the device's own CSR + weights in SPEC order (ascending term id)
    to, di, w = got["term_offsets"], got["doc_ids"], got["weights"]
    for j in (0, 5, nq // 2, nq - 1):
        terms = sorted(set(int(t) for t in qs[j]))
        for i in (0, 1, k // 2, k - 1):
            d = ids[j][i]
            if d == oi.NO_DOC:
                continue
            s = np.float32(0.0)
            for t in terms:
                lo, hi = int(to[t]), int(to[t + 1])
                p = lo + int(np.searchsorted(di[lo:hi], d))
                if p < hi and di[p] == d:
                    s = np.float32(s + w[p])
            assert s.view(np.uint32) == sc[j][i].view(np.uint32), (j, i, d)
    # two half shards with global statistics reproduce the unsharded lists exactly
    halves = []
    for base in (0, n // 2):
        with oi.GpuIndex(n_docs=n // 2, dim=64, max_k=k, max_batch=nq, doc_base=base) as ix:
            ix.synth_bm25(O.SEED, vocab, cdf)
            ix.bm25_finalize(avgdl=float(avgdl), n_docs_global=n, global_df=df)
            halves.append(ix.search_bm25(qs, k))
    for j in range(nq):
        cat_ids = np.concatenate([halves[0][0][j], halves[1][0][j]])
        cat_sc = np.concatenate([halves[0][1][j], halves[1][1][j]])
        keep = cat_ids != oi.NO_DOC
        order = sorted(np.nonzero(keep)[0], key=lambda i: (-float(cat_sc[i]), int(cat_ids[i])))[:k]
        m = len(order)
        assert np.array_equal(cat_ids[order], ids[j][:m]) and np.array_equal(cat_sc[order].view(np.uint32), sc[j][:m].view(np.uint32))
        assert np.all(ids[j][m:] == oi.NO_DOC)


def test_gpu_matches_the_published_formulas_on_hand_sized_input(oi):
    """the GPU path against the published definitions directly (50-digit decimals / exact fractions from
    tests/test_oracle_known_answers.py), without the oracle in between: Lucene-idf BM25, RRF with k = 60"""
    from decimal import Decimal
    from fractions import Fraction
    from test_oracle_known_answers import _bm25_decimal, _csr
    docs = [[0, 1, 1, 2], [0, 2, 2, 2, 3, 4, 4], [0, 1, 5], [0, 0, 0, 3], [0, 4, 5, 5, 5, 5, 1, 2, 3], [0, 3]]
    vocab, k = 7, 4
    off, di, tf, dl = _csr(docs, vocab)
    rows = np.zeros((6, 64), dtype=np.float32)
    for d in range(6):
        rows[d, d] = 1.0  # orthonormal rows: cosine scores are exactly 0 or the query component
    q = np.zeros((1, 64), dtype=np.float32)
    q[0, :6] = [0.1, 0.2, 0.3, 0.4, 0.5, 0.6]
    q /= np.linalg.norm(q)
    with oi.GpuIndex(n_docs=6, dim=64, max_k=k, max_batch=1) as ix:
        ix.load_embeddings(rows)
        ix.load_bm25(off, di, tf, dl)
        ix.bm25_finalize()
        for query in ([1, 2], [0], [5, 5, 3], [4, 6, 9], [0, 1, 2, 3, 4, 5]):
            ids, sc = ix.search_bm25([query], k)
            want = _bm25_decimal(docs, vocab, query)
            exact = sorted([(-x, d) for d, x in enumerate(want) if x > 0])[:k]
            assert [int(i) for i in ids[0][:len(exact)]] == [d for _, d in exact], query
            for i, s in zip(ids[0][:len(exact)], sc[0]):
                assert abs(Decimal(float(s)) - want[int(i)]) <= Decimal("2e-6") * max(want[int(i)], Decimal(1))
            assert np.all(ids[0][len(exact):] == oi.NO_DOC)
        cos_ids, cos_sc = ix.search_cosine(q, k)
        assert [int(i) for i in cos_ids[0]] == [5, 4, 3, 2] and np.allclose(cos_sc[0], q[0, [5, 4, 3, 2]], atol=1e-7)
        h_ids, h_rrf, h_rc, h_rb = ix.search_hybrid(q, [[5, 5, 3]], k)
        bm_ids, _ = ix.search_bm25([[5, 5, 3]], k)
        fused = {}
        for lst in ([int(i) for i in cos_ids[0]], [int(i) for i in bm_ids[0] if i != oi.NO_DOC]):
            for r, d in enumerate(lst, 1):
                fused[d] = fused.get(d, Fraction(0)) + Fraction(1, 60 + r)
        order = sorted(fused, key=lambda d: (-fused[d], d))[:k]
        assert [int(i) for i in h_ids[0][:len(order)]] == order
        for d, v in zip(h_ids[0][:len(order)], h_rrf[0]):
            assert abs(Fraction(float(v)) - fused[int(d)]) < Fraction(1, 10 ** 8)

