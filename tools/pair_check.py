#!/usr/bin/env python3
"""tools/pair_check.py — first contact with the CTA-pair (cta_group::2) GEMM: every raw score of a small corpus against
the oracle, pair kernel vs single-CTA kernel, then top-k parity at a larger size.  Run under `timeout` on the GPU box."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import openintel_b200 as oi
import oracle as O

for n, dim, nq in ((4097, 768, 130), (200, 128, 256), (3000, 384, 200)):
    rows = O.synth_rows_bf16(n, dim)
    qs = O.synth_rows_f32(nq, dim, stream=1)
    with oi.GpuIndex(n_docs=n, dim=dim, dtype=oi.DTYPE_BF16, max_k=8, max_batch=nq) as ix:
        ix.load_embeddings(rows)
        got = {}
        for pair in (0, 1):
            ix.set_option("cosine_gemm_pair", pair)
            got[pair] = ix.debug_cosine_gemm_scores(qs)
            print("n=%d dim=%d nq=%d pair=%d: scores computed" % (n, dim, nq, pair), flush=True)
    worst = 0.0
    for j in range(0, nq, max(1, nq // 8)):
        want = O.cosine_scores_bf16(rows, qs[j])
        worst = max(worst, float(np.max(np.abs(got[1][j] - want))))
    print("   pair vs oracle max err %.3g; pair == single-CTA bitwise: %s" % (worst, np.array_equal(got[0], got[1])), flush=True)
    assert worst < 2e-5
n, dim, k, nq = 300000, 768, 100, 256
qs = O.synth_rows_f32(nq, dim, stream=1)
with oi.GpuIndex(n_docs=n, dim=dim, dtype=oi.DTYPE_BF16, max_k=k, max_batch=nq) as ix:
    ix.synth_embeddings(O.SEED)
    out = {}
    for pair in (0, 1):
        ix.set_option("cosine_gemm_pair", pair)
        out[pair] = ix.search_cosine(qs, k)
    print("top-k lists equal:", np.array_equal(out[0][0], out[1][0]), np.array_equal(out[0][1], out[1][1]))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])
print("pair check ok")
