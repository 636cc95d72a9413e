"""GPU parity: cosine scan + fused top-k through the C ABI vs the CPU oracle (SPEC §2)."""
import numpy as np
import pytest

import oracle as O
from gpu_util import assert_ranked_close

pytestmark = pytest.mark.gpu

F32_TOL, BF16_TOL = 1e-5, 2e-3


@pytest.fixture(scope="module")
def oi():
    import __graft_entry__ as ge
    import openintel_b200
    openintel_b200.load_library()
    return openintel_b200


def test_device_synth_is_bit_identical_to_oracle(oi):
    for dim, dtype in ((384, oi.DTYPE_F32), (768, oi.DTYPE_BF16), (64, oi.DTYPE_F32)):
        n = 3000
        with oi.GpuIndex(n_docs=n, dim=dim, dtype=dtype, doc_base=1000) as ix:
            ix.synth_embeddings(O.SEED)
            got = ix.read_embeddings(0, n)
        want = (O.synth_rows_f32 if dtype == oi.DTYPE_F32 else O.synth_rows_bf16)(n, dim, first=1000)
        assert np.array_equal(got, want)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n,dim,k", [(50000, 384, 100), (20000, 768, 10), (4097, 128, 1), (9000, 1024, 100),
                                     (3000, 2048, 7), (70000, 64, 1000)])
def test_cosine_f32_parity(oi, variant, n, dim, k):
    rows = O.synth_rows_f32(n, dim)
    qs = np.concatenate([O.synth_rows_f32(2, dim, stream=1), O.synth_planted_queries(2, dim, n)[0]])
    with oi.GpuIndex(n_docs=n, dim=dim, max_k=k, max_batch=4) as ix:
        ix.load_embeddings(rows)
        ix.set_option("cosine_variant", variant)
        ix.set_option("cosine_multi_query", 0)  # the single-query kernels; the multi-query scan has its own test
        ids, sc = ix.search_cosine(qs, k)
    for j in range(4):
        allsc = O.cosine_scores_f32(rows, qs[j])
        wi, ws, _ = O.topk_f64(allsc, k)
        assert_ranked_close(ids[j], sc[j], wi, ws, allsc, F32_TOL)


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("n,dim,k", [(40000, 768, 100), (10000, 384, 10), (5000, 1024, 50)])
def test_cosine_bf16_parity(oi, variant, n, dim, k):
    rows = O.synth_rows_bf16(n, dim)
    qs = O.synth_rows_f32(3, dim, stream=1)
    with oi.GpuIndex(n_docs=n, dim=dim, dtype=oi.DTYPE_BF16, max_k=k, max_batch=3, doc_base=7) as ix:
        ix.load_embeddings(rows)
        ix.set_option("cosine_variant", variant)
        ix.set_option("cosine_multi_query", 0)
        ids, sc = ix.search_cosine(qs, k)
    for j in range(3):
        allsc = O.cosine_scores_bf16(rows, qs[j])
        wi, ws, _ = O.topk_f64(allsc, k, doc_base=7)
        assert_ranked_close(ids[j], sc[j], wi, ws, allsc, BF16_TOL, doc_base=7)
        # measured error is far inside the budget
        assert np.max(np.abs(sc[j] - ws)) < 1e-5


@pytest.mark.parametrize("mode", [1, 2])
@pytest.mark.parametrize("n,dim,bf16,k,nq", [(50000, 384, False, 100, 9), (30011, 256, False, 10, 4), (4097, 64, False, 1, 2),
                                             (70000, 128, False, 1000, 5), (40000, 768, True, 100, 3), (9000, 128, True, 7, 6),
                                             (3, 64, False, 5, 7)])
def test_cosine_multi_query_scan_parity(oi, mode, n, dim, bf16, k, nq):
    """calls with several queries share one matrix pass per group of 4 (rows of at most 1536 bytes): every query's
    list must match the oracle and the single-query kernel; the last group may be partial, shards may be ragged"""
    rows = O.synth_rows_bf16(n, dim) if bf16 else O.synth_rows_f32(n, dim)
    n_pl = min(2, nq)
    qs = np.concatenate([O.synth_rows_f32(nq - n_pl, dim, stream=1), O.synth_planted_queries(n_pl, dim, n)[0]]) if n >= 100 \
        else O.synth_rows_f32(nq, dim, stream=1)
    with oi.GpuIndex(n_docs=n, dim=dim, dtype=oi.DTYPE_BF16 if bf16 else oi.DTYPE_F32, max_k=k, max_batch=nq, doc_base=5) as ix:
        ix.load_embeddings(rows)
        ix.set_option("cosine_gemm_min_batch", 0)  # bf16 batches would otherwise take the tensor-core path
        ix.set_option("cosine_multi_query", mode)  # 1 = direct loads, 2 = bulk-copy pipeline (f32 rows)
        l0 = ix.launch_count()
        ids, sc = ix.search_cosine(qs, k)
        multi_launches = ix.launch_count() - l0
        ix.set_option("cosine_multi_query", 0)
        ids1, sc1 = ix.search_cosine(qs, k)
    assert multi_launches == (nq + 3) // 4 + 1  # one scan per group of 4 + the unpack kernel
    tol = BF16_TOL if bf16 else F32_TOL
    for j in range(nq):
        allsc = O.cosine_scores_bf16(rows, qs[j]) if bf16 else O.cosine_scores_f32(rows, qs[j])
        wi, ws, _ = O.topk_f64(allsc, k, doc_base=5)
        kk = min(k, n)
        assert_ranked_close(ids[j][:kk], sc[j][:kk], wi[:kk], ws[:kk], allsc, tol, doc_base=5)
        assert np.all(ids[j][kk:] == oi.NO_DOC) and np.all(sc[j][kk:] == 0)
        # against the single-query kernel: same documents unless two scores tie within the f32 summation noise
        if not np.array_equal(ids[j], ids1[j]):
            assert sorted(ids[j].tolist()) == sorted(ids1[j].tolist()) or np.max(np.abs(sc[j] - sc1[j])) < 1e-6
        assert np.allclose(sc[j], sc1[j], rtol=0, atol=2e-6)


@pytest.mark.parametrize("variant", [0, 1])
def test_edge_cases(oi, variant):
    dim = 64
    # fewer docs than k -> padded with (NO_DOC, 0); exact ties broken by ascending doc id
    rows = O.synth_rows_f32(5, dim)
    rows[3] = rows[1]
    with oi.GpuIndex(n_docs=5, dim=dim, max_k=8) as ix:
        ix.load_embeddings(rows)
        ix.set_option("cosine_variant", variant)
        ids, sc = ix.search_cosine(rows[1:2], 8)
        assert list(ids[0][:2]) == [1, 3] and sc[0][0] == sc[0][1]
        assert list(ids[0][5:]) == [oi.NO_DOC] * 3 and list(sc[0][5:]) == [0.0] * 3
        assert sorted(ids[0][:5]) == [0, 1, 2, 3, 4]
        # nq = 0 is a no-op; k > max_k and unloaded state are errors
        ix.search_cosine(np.zeros((0, dim), np.float32), 3)
        with pytest.raises(oi.OiError) as e:
            ix.search_cosine(rows[:1], 9)
        assert e.value.status == 1
    with oi.GpuIndex(n_docs=100, dim=dim, max_k=8) as ix:
        with pytest.raises(oi.OiError) as e:
            ix.search_cosine(rows[:1], 3)
        assert e.value.status == 5
    # all-identical rows: every score ties, order is by doc id
    rows = np.tile(O.synth_rows_f32(1, dim), (5000, 1))
    with oi.GpuIndex(n_docs=5000, dim=dim, max_k=100, doc_base=10) as ix:
        ix.load_embeddings(rows)
        ix.set_option("cosine_variant", variant)
        ids, sc = ix.search_cosine(rows[:1], 100)
        assert list(ids[0]) == list(range(10, 110))


@pytest.mark.parametrize("variant", [0, 1])
def test_full_size_config2_properties(oi, variant):
    """BASELINE config 2 (1M x 384 f32, top-100) through size-independent properties: planted
    targets are found at rank 1, the list is sorted, ids are unique, and the result equals the
    merge of the results of two half-size shards (checksum of checksums)."""
    n, dim, k = 1_000_000, 384, 100
    q, tgt = O.synth_planted_queries(2, dim, n)
    q = np.concatenate([q, O.synth_rows_f32(1, dim, stream=1)])
    with oi.GpuIndex(n_docs=n, dim=dim, max_k=k, max_batch=3) as ix:
        ix.synth_embeddings(O.SEED)
        ix.set_option("cosine_variant", variant)
        ids, sc = ix.search_cosine(q, k)
    assert ids[0][0] == tgt[0] and ids[1][0] == tgt[1] and sc[0][0] > 0.85
    for j in range(3):
        assert np.all(np.diff(sc[j]) <= 0) and len(set(ids[j])) == k
    halves = []
    for base in (0, n // 2):
        with oi.GpuIndex(n_docs=n // 2, dim=dim, max_k=k, max_batch=3, doc_base=base) as ix:
            ix.synth_embeddings(O.SEED)
            ix.set_option("cosine_variant", variant)
            halves.append(ix.search_cosine(q, k))
    for j in range(3):
        cat_ids = np.concatenate([halves[0][0][j], halves[1][0][j]])
        cat_sc = np.concatenate([halves[0][1][j], halves[1][1][j]])
        order = sorted(range(2 * k), key=lambda i: (-cat_sc[i], cat_ids[i]))[:k]
        assert np.array_equal(cat_ids[order], ids[j]) and np.array_equal(cat_sc[order], sc[j])
    # spot-check against the oracle on the rows that matter: recompute the reported docs' scores
    for j in range(3):
        for i in (0, 1, 50, 99):
            row = O.synth_rows_f32(1, dim, first=int(ids[j][i]))
            want = float(O.cosine_scores_f32(row, q[j])[0])
            assert abs(want - sc[j][i]) <= F32_TOL * max(abs(want), 1e-2)


def test_empty_shard_returns_padding(oi):
    """a shard without documents (more ranks than documents) answers every query with padding"""
    dim, k = 64, 5
    q = O.synth_rows_f32(5, dim, stream=1)
    with oi.GpuIndex(n_docs=0, dim=dim, max_k=k, max_batch=5) as ix:
        ix.load_embeddings(np.zeros((0, dim), np.float32))
        for nq in (1, 5):
            ids, sc = ix.search_cosine(q[:nq], k)
            assert np.all(ids == oi.NO_DOC) and np.all(sc == 0)
        ix.load_bm25(np.zeros(4, np.uint64), np.zeros(0, np.uint32), np.zeros(0, np.uint32), np.zeros(0, np.uint32))
        ix.bm25_finalize(avgdl=10.0, n_docs_global=100, global_df=np.array([3, 0, 7], np.uint32))
        ids, sc = ix.search_bm25([[0, 2], [1]], k)
        assert np.all(ids == oi.NO_DOC) and np.all(sc == 0)
        ids, rrf, rc, rb = ix.search_hybrid(q[:2], [[0, 2], [1]], k)
        assert np.all(ids == oi.NO_DOC) and np.all(rrf == 0) and np.all(rc == 0) and np.all(rb == 0)
