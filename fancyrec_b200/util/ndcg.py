"""DCG / NDCG with the reference's exact float64 arithmetic (util/ndcg.py:9-45, 48-78).

Host-side (NB-length inputs at most); the device side only produces the integer relevance bits.
np.sum is used on the same expressions as the reference so that its pairwise summation order --
which matters in the last bit at k = 50 -- is reproduced.
"""
import numpy as np


def dcg_at_k(r, k, method=0):
    r = np.asarray(r, dtype=np.float64)[:k]
    if not r.size:
        return 0.
    if method == 0:
        discounts = np.log2(np.arange(2, r.size + 1))
        return r[0] + np.sum(r[1:] / discounts)
    if method == 1:
        discounts = np.log2(np.arange(2, r.size + 2))
        return np.sum(r / discounts)
    raise ValueError('method must be 0 or 1.')


def ndcg_at_k(r, k, method=0):
    best = dcg_at_k(sorted(r, reverse=True), k, method)
    if not best:
        return 0.
    return dcg_at_k(r, k, method) / best


def ndcg_from_hits(hits, n_pos, k, n_total):
    """ndcg_at_k of a 0/1 list of length n_total with n_pos ones, given only its first
    min(k, n_total) entries: the ideal list is min(n_pos, k) ones (util/ndcg.py:75)."""
    depth = min(k, n_total)
    ideal = np.zeros(depth, dtype=np.float64)
    ideal[:min(int(n_pos), depth)] = 1.0
    best = dcg_at_k(ideal, k)
    if not best:
        return 0.
    return dcg_at_k(np.asarray(hits[:depth], dtype=np.float64), k) / best
