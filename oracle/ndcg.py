"""Oracle restatement of the reference DCG / NDCG (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/util/ndcg.py:9-45 (dcg_at_k) and :48-78 (ndcg_at_k).
``np.asfarray`` (removed in NumPy 2) is spelled ``np.asarray(..., float64)``,
which is what it did in the reference's pinned NumPy 1.21.5.
"""
import numpy as np


def dcg_at_k(r, k, method=0):
    # util/ndcg.py:37-45
    r = np.asarray(r, dtype=np.float64)[:k]
    if r.size:
        if method == 0:
            return r[0] + np.sum(r[1:] / np.log2(np.arange(2, r.size + 1)))
        elif method == 1:
            return np.sum(r / np.log2(np.arange(2, r.size + 2)))
        else:
            raise ValueError('method must be 0 or 1.')
    return 0.


def ndcg_at_k(r, k, method=0):
    # util/ndcg.py:75-78
    dcg_max = dcg_at_k(sorted(r, reverse=True), k, method)
    if not dcg_max:
        return 0.
    return dcg_at_k(r, k, method) / dcg_max


def ndcg_from_hits(hits, n_pos, k, n_total):
    """NDCG@k of a binary relevance list known only through its first
    ``len(hits)`` entries (``hits`` = relevance of the top ranks, in rank order)
    and the total number of ones ``n_pos`` among ``n_total`` items.

    Equal to ``ndcg_at_k(full_list, k)`` whenever ``len(hits) >= min(k, n_total)``:
    dcg_at_k only reads the first k entries (util/ndcg.py:37) and the ideal
    ordering is ``min(n_pos, k)`` ones followed by zeros (util/ndcg.py:75).
    """
    kk = min(k, n_total)
    assert len(hits) >= kk
    ideal = np.zeros(kk, dtype=np.float64)
    ideal[:min(n_pos, kk)] = 1.0
    dcg_max = dcg_at_k(ideal, k)
    if not dcg_max:
        return 0.
    return dcg_at_k(np.asarray(hits[:kk], dtype=np.float64), k) / dcg_max
