"""Helpers shared by the GPU parity tests: tie-band comparison of ranked lists (SPEC §2)."""
import numpy as np


def assert_ranked_close(ids, scores, want_ids, want_scores, all_scores, rel, floor=1e-2, doc_base=0):
    """ids/scores: GPU list; want_*: oracle list (f64 scores); all_scores: oracle score of every doc
    (f64, local index).  Outside tie bands the lists must be identical (SPEC §2)."""
    ids = np.asarray(ids).astype(np.int64)
    want_ids = np.asarray(want_ids).astype(np.int64)
    k = len(ids)
    tol = rel * np.maximum(np.abs(want_scores), floor)
    # scores agree rank by rank (the sorted score sequences must match within tolerance)
    assert np.all(np.abs(np.asarray(scores, dtype=np.float64) - want_scores) <= tol), \
        "score mismatch: max abs diff %g" % np.max(np.abs(scores - want_scores))
    # each GPU doc's own oracle score agrees with the score the GPU reported for it
    own = all_scores[ids - doc_base]
    assert np.all(np.abs(own - np.asarray(scores, dtype=np.float64)) <= rel * np.maximum(np.abs(own), floor))
    # position-wise identity outside tie bands
    for i in range(k):
        if ids[i] != want_ids[i]:
            assert abs(all_scores[ids[i] - doc_base] - want_scores[i]) <= tol[i], \
                "rank %d: doc %d (%.9g) vs oracle doc %d (%.9g)" % (i, ids[i], all_scores[ids[i] - doc_base], want_ids[i], want_scores[i])
    assert len(set(ids.tolist())) == k, "duplicate doc ids in a ranked list"
    return int(np.sum(ids != want_ids))
