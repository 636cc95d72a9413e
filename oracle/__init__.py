"""ctypes front end of the CPU oracle (oracle/oracle.c).  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module, and only as the checker.  The product package (openintel_b200) never does.

Parity status: BM25 / cosine / top-k / RRF are "parity unpinned" (the reference has no such
code); tokenizer / lexicon / engine summary are pinned to the reference's goldens.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboi_oracle.so")
NO_DOC = 0xFFFFFFFF
SEED = 20261018


def build(force=False):
    """Compile oracle/*.c with gcc (oracle/Makefile; make rebuilds what is older than its sources)."""
    if force or not os.path.exists(_SO) or any(
            os.path.getmtime(os.path.join(_HERE, f)) > os.path.getmtime(_SO)
            for f in os.listdir(_HERE) if f.endswith((".c", ".h"))):
        if os.path.exists("/usr/bin/gcc"):
            subprocess.check_call(["make", "-C", _HERE, "-s"] + (["-B"] if force else []))
    return _SO


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.oio_hash64.restype = C.c_uint64
        L.oio_hash64.argtypes = [C.c_uint64] * 4
        L.oio_key.restype = C.c_uint64
        L.oio_key.argtypes = [C.c_float, C.c_uint32]
        L.oio_csr_count.restype = C.c_uint64
        L.oio_bm25_avgdl.restype = C.c_float
        L.oio_crowding.restype = C.c_double
        L.oio_f32_to_bf16.restype = C.c_uint16
        L.oio_f32_to_bf16.argtypes = [C.c_float]
        L.oio_bf16_to_f32.restype = C.c_float
        L.oio_bf16_to_f32.argtypes = [C.c_uint16]
        _lib = L
    return _lib


u64, u32, i32 = C.c_uint64, C.c_uint32, C.c_int


def hash64(seed, stream, row, col):
    return lib().oio_hash64(seed, stream, row, col)


def key(score, doc):
    return lib().oio_key(C.c_float(score), doc)


def synth_rows_f32(n, dim, seed=SEED, stream=0, first=0):
    out = np.empty((n, dim), dtype=np.float32)
    lib().oio_synth_rows_f32(u64(seed), u64(stream), u64(first), u64(n), u32(dim), _p(out, C.c_float))
    return out


def synth_rows_bf16(n, dim, seed=SEED, stream=0, first=0):
    out = np.empty((n, dim), dtype=np.uint16)
    lib().oio_synth_rows_bf16(u64(seed), u64(stream), u64(first), u64(n), u32(dim), _p(out, C.c_uint16))
    return out


def synth_planted_queries(nq, dim, n_docs, seed=SEED, first=0):
    out = np.empty((nq, dim), dtype=np.float32)
    tgt = np.empty(nq, dtype=np.uint32)
    lib().oio_synth_planted_queries_f32(u64(seed), u64(first), u64(nq), u32(dim), u64(n_docs),
                                        _p(out, C.c_float), _p(tgt, C.c_uint32))
    return out, tgt


def bf16_to_f32(a):
    return (a.astype(np.uint32) << 16).view(np.float32)


def f32_to_bf16(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    u = a.view(np.uint32).astype(np.uint64)
    u = u + 0x7FFF + ((u >> 16) & 1)
    return (u >> 16).astype(np.uint16)


def synth_doc_lens(n_docs, seed=SEED, first=0):
    out = np.empty(n_docs, dtype=np.uint32)
    lib().oio_synth_doc_lens(u64(seed), u64(first), u64(n_docs), _p(out, C.c_uint32))
    return out


def zipf_cdf(vocab):
    out = np.empty(vocab, dtype=np.float64)
    lib().oio_zipf_cdf(u32(vocab), _p(out, C.c_double))
    return out


def synth_tokens(doc_len, cdf, seed=SEED, first=0):
    n = len(doc_len)
    off = np.zeros(n + 1, dtype=np.uint64)
    np.cumsum(doc_len, out=off[1:])
    toks = np.empty(int(off[-1]), dtype=np.uint32)
    lib().oio_synth_tokens(u64(seed), u64(first), u64(n), _p(doc_len, C.c_uint32), _p(off, C.c_uint64),
                           _p(cdf, C.c_double), u32(len(cdf)), _p(toks, C.c_uint32))
    return off, toks


def synth_query_terms(nq, terms_per_query, cdf, uniform=False, seed=SEED, first=0):
    out = np.empty((nq, terms_per_query), dtype=np.uint32)
    lib().oio_synth_query_terms(u64(seed), u64(6 if uniform else 5), u64(first), u64(nq),
                                u32(terms_per_query), _p(cdf, C.c_double), u32(len(cdf)),
                                _p(out, C.c_uint32))
    return out


def build_csr(tok_off, tokens, vocab):
    n = len(tok_off) - 1
    counts = np.empty(vocab, dtype=np.uint64)
    total = lib().oio_csr_count(u64(n), _p(tok_off, C.c_uint64), _p(tokens, C.c_uint32), u32(vocab),
                                _p(counts, C.c_uint64))
    term_off = np.empty(vocab + 1, dtype=np.uint64)
    doc_ids = np.empty(total, dtype=np.uint32)
    tfs = np.empty(total, dtype=np.uint32)
    lib().oio_csr_fill(u64(n), _p(tok_off, C.c_uint64), _p(tokens, C.c_uint32), u32(vocab),
                       _p(counts, C.c_uint64), _p(term_off, C.c_uint64), _p(doc_ids, C.c_uint32),
                       _p(tfs, C.c_uint32))
    return term_off, doc_ids, tfs


def synth_bm25_corpus(n_docs, vocab, seed=SEED, first=0):
    """-> dict(doc_len, term_offsets, doc_ids, tfs, cdf) of docs [first, first+n_docs)."""
    cdf = zipf_cdf(vocab)
    dl = synth_doc_lens(n_docs, seed, first)
    off, toks = synth_tokens(dl, cdf, seed, first)
    term_off, doc_ids, tfs = build_csr(off, toks, vocab)
    return dict(doc_len=dl, term_offsets=term_off, doc_ids=doc_ids, tfs=tfs, cdf=cdf)


def bm25_idf(n_docs_global, df):
    df = np.ascontiguousarray(df, dtype=np.uint32)
    out = np.empty(len(df), dtype=np.float32)
    lib().oio_bm25_idf(u64(n_docs_global), _p(df, C.c_uint32), u32(len(df)), _p(out, C.c_float))
    return out


def bm25_avgdl(doc_len):
    return float(lib().oio_bm25_avgdl(_p(doc_len, C.c_uint32), u64(len(doc_len))))


def bm25_weights(term_off, doc_ids, tfs, doc_len, idf, k1=1.2, b=0.75, avgdl=None):
    if avgdl is None:
        avgdl = bm25_avgdl(doc_len)
    w = np.empty(len(doc_ids), dtype=np.float32)
    lib().oio_bm25_weights(_p(term_off, C.c_uint64), _p(doc_ids, C.c_uint32), _p(tfs, C.c_uint32),
                           _p(doc_len, C.c_uint32), _p(idf, C.c_float), u32(len(term_off) - 1),
                           C.c_float(k1), C.c_float(b), C.c_float(avgdl), _p(w, C.c_float))
    return w


def bm25_score_dense(term_off, doc_ids, w, q_terms, n_docs):
    q_terms = np.ascontiguousarray(q_terms, dtype=np.uint32)
    out = np.empty(n_docs, dtype=np.float32)
    lib().oio_bm25_score_dense(_p(term_off, C.c_uint64), _p(doc_ids, C.c_uint32), _p(w, C.c_float),
                               u32(len(term_off) - 1), _p(q_terms, C.c_uint32), u32(len(q_terms)),
                               u64(n_docs), _p(out, C.c_float))
    return out


def cosine_scores_f32(rows, q):
    rows = np.ascontiguousarray(rows, dtype=np.float32)
    q = np.ascontiguousarray(q, dtype=np.float32)
    out = np.empty(rows.shape[0], dtype=np.float64)
    lib().oio_cosine_scores_f32(_p(rows, C.c_float), u64(rows.shape[0]), u32(rows.shape[1]),
                                _p(q, C.c_float), _p(out, C.c_double))
    return out


def cosine_scores_bf16(rows_bf16, q):
    rows = np.ascontiguousarray(rows_bf16, dtype=np.uint16)
    q = np.ascontiguousarray(q, dtype=np.float32)
    out = np.empty(rows.shape[0], dtype=np.float64)
    lib().oio_cosine_scores_bf16(_p(rows, C.c_uint16), u64(rows.shape[0]), u32(rows.shape[1]),
                                 _p(q, C.c_float), _p(out, C.c_double))
    return out


def topk_f64(scores, k, doc_base=0):
    scores = np.ascontiguousarray(scores, dtype=np.float64)
    ids = np.empty(k, dtype=np.uint32)
    sc = np.empty(k, dtype=np.float64)
    m = lib().oio_topk_f64(_p(scores, C.c_double), u64(len(scores)), u32(k), u32(doc_base),
                           _p(ids, C.c_uint32), _p(sc, C.c_double))
    return ids, sc, m


def topk_f32(scores, k, only_positive=False, doc_base=0):
    scores = np.ascontiguousarray(scores, dtype=np.float32)
    ids = np.empty(k, dtype=np.uint32)
    sc = np.empty(k, dtype=np.float32)
    m = lib().oio_topk_f32(_p(scores, C.c_float), u64(len(scores)), u32(k), i32(int(only_positive)),
                           u32(doc_base), _p(ids, C.c_uint32), _p(sc, C.c_float))
    return ids, sc, m


def cosine_topk_f32_fast(rows, q, k, n_threads=None):
    """The timed CPU baseline (multi-threaded f32 scan + heap top-k)."""
    if n_threads is None:
        n_threads = max_threads()
    ids = np.empty(k, dtype=np.uint32)
    sc = np.empty(k, dtype=np.float32)
    lib().oio_cosine_topk_f32_fast(_p(rows, C.c_float), u64(rows.shape[0]), u32(rows.shape[1]),
                                   _p(q, C.c_float), u32(k), i32(n_threads), _p(ids, C.c_uint32),
                                   _p(sc, C.c_float))
    return ids, sc


def max_threads():
    return int(lib().oio_max_threads())


def host_threads():
    """Threads the timed / full-size legs use: every core this process may run on.  NOT omp_get_max_threads():
    torch.distributed.run exports OMP_NUM_THREADS=1, which made the round-1 CPU arm a 1-thread run at N >= 2."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:  # pragma: no cover
        return max(1, os.cpu_count() or 1)


# ---- full-size drivers (oracle_scale.c): oracle.c's own functions, chunked and spread over the host threads ----
def scale_synth_rows(n, dim, bf16=False, seed=SEED, stream=0, first=0, n_threads=None):
    out = np.empty((n, dim), dtype=np.uint16 if bf16 else np.float32)
    lib().oio_scale_synth_rows(u64(seed), u64(stream), u64(first), u64(n), u32(dim), i32(int(bf16)),
                               i32(n_threads or host_threads()), C.c_void_p(out.ctypes.data))
    return out


def scale_cosine_topk(n_rows, dim, q, k, bf16=False, seed=SEED, first=0, n_threads=None):
    """exact top-k (double accumulation) of each query over synthetic rows [first, first + n_rows), regenerated in
    chunks -> (ids [nq][k] global row numbers, scores f64 [nq][k])"""
    q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, dim)
    nq = q.shape[0]
    ids = np.empty((nq, k), dtype=np.uint32)
    sc = np.empty((nq, k), dtype=np.float64)
    lib().oio_scale_cosine_topk(u64(seed), u64(first), u64(n_rows), u32(dim), i32(int(bf16)), _p(q, C.c_float), u32(nq), u32(k),
                                i32(n_threads or host_threads()), _p(ids, C.c_uint32), _p(sc, C.c_double))
    return ids, sc


def scale_bm25_mini_index(n_docs, vocab, terms, cdf=None, seed=SEED, first=0, n_threads=None):
    """The CSR restricted to `terms` (any iterable of term ids) of synthetic documents [first, first + n_docs), built
    by regenerating every document's tokens (oio_synth_tokens) -- the oracle's own index of the touched terms, at
    sizes where the whole CSR would take minutes.  -> dict(terms (ascending), term_offsets, doc_ids, tfs, doc_len,
    sum_doc_len); mini term id i stands for terms[i], so ascending order is preserved (SPEC §3 summation order)."""
    L = lib()
    L.oio_scale_mini_csr_build.restype = C.c_void_p
    L.oio_scale_mini_csr_size.restype = C.c_uint64
    L.oio_scale_mini_csr_size.argtypes = [C.c_void_p]
    if cdf is None:
        cdf = zipf_cdf(vocab)
    terms = np.unique(np.asarray(list(terms), dtype=np.uint32))
    assert len(terms) <= 64
    obj = L.oio_scale_mini_csr_build(u64(seed), u64(first), u64(n_docs), _p(cdf, C.c_double), u32(vocab), _p(terms, C.c_uint32),
                                     u32(len(terms)), i32(n_threads or host_threads()))
    assert obj, "oio_scale_mini_csr_build failed"
    n = int(L.oio_scale_mini_csr_size(C.c_void_p(obj)))
    term_off = np.empty(len(terms) + 1, dtype=np.uint64)
    d = np.empty(max(n, 1), dtype=np.uint32)
    tf = np.empty(max(n, 1), dtype=np.uint32)
    sdl = C.c_uint64()
    L.oio_scale_mini_csr_take(C.c_void_p(obj), _p(term_off, C.c_uint64), _p(d, C.c_uint32), _p(tf, C.c_uint32), C.byref(sdl))
    return dict(terms=terms, term_offsets=term_off, doc_ids=d[:n], tfs=tf[:n],
                doc_len=synth_doc_lens(n_docs, seed, first), sum_doc_len=int(sdl.value), cdf=cdf)


def scale_bm25_topk(mini, q_terms, k, n_docs_global=None, avgdl=None, df_global=None, doc_base=0):
    """BM25 top-k of ONE query from a mini index (scale_bm25_mini_index): oracle.c's idf / weights / dense scoring /
    top-k on the touched terms only.  df_global: df of mini['terms'] over the whole corpus (default: this index)."""
    terms = mini["terms"]
    n_docs = len(mini["doc_len"])
    df = np.diff(mini["term_offsets"]).astype(np.uint32) if df_global is None else np.asarray(df_global, dtype=np.uint32)
    idf = bm25_idf(n_docs_global or n_docs, df)
    if avgdl is None:
        avgdl = float(np.float32(np.float64(mini["sum_doc_len"]) / np.float64(n_docs)))
    w = bm25_weights(mini["term_offsets"], mini["doc_ids"], mini["tfs"], mini["doc_len"], idf, avgdl=avgdl)
    pos = np.searchsorted(terms, np.asarray(q_terms, dtype=np.uint32))
    known = [int(p) for p, t in zip(pos, q_terms) if p < len(terms) and terms[p] == t]
    s = bm25_score_dense(mini["term_offsets"], mini["doc_ids"], w, np.asarray(known, dtype=np.uint32), n_docs)
    return topk_f32(s, k, only_positive=True, doc_base=doc_base)


def hybrid_batch_fast(rows_bf16, q, term_off, doc_ids, w, q_terms, k, rrf_k=60, n_threads=None):
    """The timed CPU baseline of one hybrid batch (oracle_fast.c) over a corpus slice."""
    rows = np.ascontiguousarray(rows_bf16, dtype=np.uint16)
    q = np.ascontiguousarray(q, dtype=np.float32).reshape(-1, rows.shape[1])
    qt = np.ascontiguousarray(q_terms, dtype=np.uint32)
    nq = q.shape[0]
    assert qt.shape[0] == nq
    ids = np.empty((nq, k), dtype=np.uint32)
    rrf_ = np.empty((nq, k), dtype=np.float32)
    rc = np.empty((nq, k), dtype=np.uint32)
    rb = np.empty((nq, k), dtype=np.uint32)
    r = lib().oio_hybrid_batch_fast(_p(rows, C.c_uint16), u64(rows.shape[0]), u32(rows.shape[1]), _p(q, C.c_float), u32(nq),
                                    _p(term_off, C.c_uint64), _p(doc_ids, C.c_uint32), _p(w, C.c_float), u32(len(term_off) - 1),
                                    _p(qt, C.c_uint32), u32(qt.shape[1]), u32(k), u32(rrf_k), i32(n_threads or host_threads()),
                                    _p(ids, C.c_uint32), _p(rrf_, C.c_float), _p(rc, C.c_uint32), _p(rb, C.c_uint32))
    assert r == 0
    return ids, rrf_, rc, rb


def rrf(ids_cos, ids_bm25, k, rrf_k=60):
    a = np.ascontiguousarray(ids_cos, dtype=np.uint32)
    b = np.ascontiguousarray(ids_bm25, dtype=np.uint32)
    ids = np.empty(k, dtype=np.uint32)
    val = np.empty(k, dtype=np.float32)
    rc = np.empty(k, dtype=np.uint32)
    rb = np.empty(k, dtype=np.uint32)
    m = lib().oio_rrf(_p(a, C.c_uint32), u32(len(a)), _p(b, C.c_uint32), u32(len(b)), u32(k), u32(rrf_k),
                      _p(ids, C.c_uint32), _p(val, C.c_float), _p(rc, C.c_uint32), _p(rb, C.c_uint32))
    return ids, val, rc, rb, m


def tokenize(text):
    raw = text.encode("utf-8") if isinstance(text, str) else bytes(text)
    n = len(raw)
    buf = (C.c_uint8 * max(n, 1)).from_buffer_copy(raw or b"\0")
    norm = (C.c_uint8 * (n + 1))()
    st = (C.c_uint32 * (n + 1))()
    ln = (C.c_uint32 * (n + 1))()
    m = lib().oio_tokenize(buf, C.c_size_t(n), norm, st, ln, u32(n + 1))
    nb = bytes(norm)
    return [nb[st[i]:st[i] + ln[i]].decode("ascii") for i in range(m)]


def lexicon_score(text):
    raw = text.encode("utf-8")
    buf = (C.c_uint8 * max(len(raw), 1)).from_buffer_copy(raw or b"\0")
    pol, spec, bull, bear = C.c_double(), C.c_int(), C.c_uint32(), C.c_uint32()
    lib().oio_lexicon_score(buf, C.c_size_t(len(raw)), C.byref(pol), C.byref(spec), C.byref(bull), C.byref(bear))
    return pol.value, bool(spec.value), bull.value, bear.value


class _Summary(C.Structure):
    _fields_ = [("total", C.c_uint64), ("bullish", C.c_uint64), ("bearish", C.c_uint64),
                ("neutral", C.c_uint64), ("net_sentiment", C.c_double),
                ("speculation_index", C.c_double), ("bull_bear_ratio", C.c_double)]


def social_summary(polarity, speculative, threshold=0.2):
    p = np.ascontiguousarray(polarity, dtype=np.float64)
    s = np.ascontiguousarray(speculative, dtype=np.int32)
    out = _Summary()
    lib().oio_social_summary(_p(p, C.c_double), _p(s, C.c_int), u64(len(p)), C.c_double(threshold), C.byref(out))
    return {f: getattr(out, f) for f, _ in _Summary._fields_}


def crowding(total, spec_index, rvol=None, iv=None, w=(0.5, 0.3, 0.2), rvol_cap=3.0):
    return lib().oio_crowding(u64(total), C.c_double(spec_index), i32(rvol is not None),
                              C.c_double(rvol or 0.0), i32(iv is not None), C.c_double(iv or 0.0),
                              C.c_double(w[0]), C.c_double(w[1]), C.c_double(w[2]), C.c_double(rvol_cap))


ALIGNMENT = ("ConfirmingBullish", "ConfirmingBearish", "Diverging", "Quiet")


def alignment(has_market, total, net, pct_change, min_sample=10, net_thr=0.05, price_thr=1.0):
    return ALIGNMENT[lib().oio_alignment(i32(int(has_market)), u64(total), C.c_double(net),
                                         C.c_double(pct_change), u64(min_sample), C.c_double(net_thr),
                                         C.c_double(price_thr))]


# ---- synthetic posts of BASELINE.json configs[0] (test / tool input, not product code) -----------
def synth_posts(n_docs, vocab, seed=SEED):
    """10k-post style synthetic corpus: token ids from the SPEC §9 Zipf generator,
    rendered as words ("w<id>", with a few upper-case / punctuation variants so the tokenizer has work to do) in
    dicts that carry the reference's SocialPost fields.  -> (posts, zipf cdf)"""
    cdf = zipf_cdf(vocab)
    dl = synth_doc_lens(n_docs, seed)
    off, toks = synth_tokens(dl, cdf, seed)
    posts = []
    for d in range(n_docs):
        words = []
        for j, t in enumerate(toks[int(off[d]):int(off[d + 1])]):
            w = "w%d" % t
            if (d + j) % 7 == 0:
                w = w.upper()
            if (d + j) % 5 == 0:
                w = "$" + w + ","
            words.append(w)
        posts.append(dict(id="post-%d" % d, source=("reddit", "bluesky")[d % 2], author="user%d" % (d % 97), text=" ".join(words),
                          created_at="2026-10-18T00:00:%02dZ" % (d % 60), engagement=d % 1000))
    return posts, cdf
