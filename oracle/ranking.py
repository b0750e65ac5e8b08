"""Oracle restatement of the reference ranking + metric loop (TEST INFRASTRUCTURE ONLY).

Follows /root/reference/evaluator.py:
  * l2norm / cal_sim                      evaluator.py:14-29
  * per-brand sort, AUC, NDCG, ranks      evaluator.py:103-127
  * recall / MedR / MeanR / averages      evaluator.py:129-143

Tie-break (the stated deterministic order): (score descending, post index
ascending).  It IS the reference order wherever the reference uses Python's
stable ``sorted(..., reverse=True)`` (evaluator.py:109 -> MedR, MeanR, NDCG) and
it is the deterministic refinement of ``np.argsort(-d)`` (evaluator.py:124 ->
recall@k), whose order under ties is unspecified.

Two implementations:
  rank_metrics_loop  -- literal pure-Python transcription, small inputs only;
  rank_stats / rank_metrics_vec -- vectorised NumPy, proven equal to the loop
                        in tests/test_oracle_ranking.py, usable up to ~1e8 pairs.
"""
import numpy as np

from .ndcg import ndcg_at_k, ndcg_from_hits


# ----------------------------------------------------------------------------
# scores
# ----------------------------------------------------------------------------
def l2norm(x):
    """evaluator.py:14-19 -- row / sqrt(sum(row**2)); no epsilon (zero row -> NaN)."""
    x = np.asarray(x, dtype=np.float32)
    norm = np.sqrt(np.sum(np.power(x, 2), axis=1, keepdims=True, dtype=np.float32))
    return (x / norm).astype(np.float32)


def cal_sim(im, s):
    """evaluator.py:23-29 -- cosine similarity of every (brand, post) pair, fp32."""
    return l2norm(im) @ l2norm(s).T


# ----------------------------------------------------------------------------
# literal loop (evaluator.py:103-143)
# ----------------------------------------------------------------------------
def rank_metrics_loop(scores, brands):
    """Pure-Python transcription of evaluator.py:103-143 for metric == 'auc'.

    ``scores`` [NB, NP] float32, ``brands`` [NP] int.  Returns the 8-tuple
    (MedR, MeanR, AUC, NDCG@10, NDCG@50, r1, r5, r10).  The one deviation:
    ``np.argsort(-d)`` at evaluator.py:124 is taken with kind='stable'
    (the stated tie-break)."""
    scores = np.asarray(scores)
    brands = np.asarray(brands)
    nb, npost = scores.shape
    brand_list = list(range(nb))
    queries = []
    ranks = np.zeros(nb)
    for b in range(nb):
        predictions = [(scores[b, j], int(brands[j])) for j in range(npost)]
        s_predictions = sorted(predictions, key=lambda x: x[0], reverse=True)
        pos = [v[0] for v in s_predictions if brand_list[b] == v[-1]]
        neg = [v[0] for v in s_predictions if brand_list[b] != v[-1]]
        total = np.sum([len([el for el in neg if e > el]) for e in pos])
        if len(pos) != 0:
            rank_of_first_pos = list(zip(*s_predictions))[-1].index(brand_list[b])
            rel = [1 if brand_list[b] == v[-1] else 0 for v in s_predictions]
            queries.append((rank_of_first_pos,
                            float(total) / (len(pos) * len(neg)),
                            ndcg_at_k(rel, 10),
                            ndcg_at_k(rel, 50)))
            inds = np.argsort(-scores[b], kind='stable')
            brand_idx = brands[inds]
            ranks[b] = np.where(brand_idx == b)[0][0]
    return _aggregate_queries(queries, ranks)


def _aggregate_queries(queries, ranks):
    # evaluator.py:129-143
    r1 = 100.0 * len(np.where(ranks < 1)[0]) / len(ranks)
    r5 = 100.0 * len(np.where(ranks < 5)[0]) / len(ranks)
    r10 = 100.0 * len(np.where(ranks < 10)[0]) / len(ranks)
    queries = list(zip(*queries))
    return (np.floor(np.median(queries[0])),
            np.floor(np.mean(queries[0])),
            np.average(queries[1]),
            np.average(queries[2]),
            np.average(queries[3]),
            r1, r5, r10)


# ----------------------------------------------------------------------------
# vectorised restatement with the integer intermediates exposed
# ----------------------------------------------------------------------------
def order_desc(row):
    """Permutation that sorts one score row by (score desc, index asc)."""
    return np.argsort(-np.asarray(row), kind='stable')


def topk_indices(scores, k):
    """[NB, min(k, NP)] int64 post indices in (score desc, index asc) order."""
    scores = np.asarray(scores)
    k = min(k, scores.shape[1])
    return np.stack([order_desc(scores[b])[:k] for b in range(scores.shape[0])])


def rank_stats(scores, brands, hit_depth=50):
    """Integer statistics per brand, everything the 8-tuple is a function of.

    Returns dict of arrays (length NB):
      n_pos       number of posts labelled b
      first_rank  0-based rank of the best positive under the stated order
                  (-1 if n_pos == 0)                       evaluator.py:116,122-127
      auc_num     sum over positives e of #{negatives el : e > el}  evaluator.py:111-113
      hits        [NB, min(hit_depth, NP)] uint8, 1 where the post at that rank is
                  labelled b                               evaluator.py:119-120
    """
    scores = np.asarray(scores)
    brands = np.asarray(brands).astype(np.int64)
    nb, npost = scores.shape
    depth = min(hit_depth, npost)
    n_pos = np.zeros(nb, dtype=np.int64)
    first_rank = np.full(nb, -1, dtype=np.int64)
    auc_num = np.zeros(nb, dtype=np.int64)
    hits = np.zeros((nb, depth), dtype=np.uint8)
    for b in range(nb):
        row = scores[b]
        order = order_desc(row)
        rel = brands[order] == b
        n_pos[b] = int(rel.sum())
        hits[b] = rel[:depth]
        if n_pos[b]:
            first_rank[b] = int(np.argmax(rel))
            is_pos = brands == b
            neg_sorted = np.sort(row[~is_pos])
            # strict: ties contribute 0 (evaluator.py:113 "e > el")
            auc_num[b] = int(np.searchsorted(neg_sorted, row[is_pos], side='left').sum())
    return dict(n_pos=n_pos, first_rank=first_rank, auc_num=auc_num, hits=hits)


def aggregate(stats, n_posts):
    """evaluator.py:105,115-143 applied to the integer statistics.

    Brands without positives are skipped for MedR/MeanR/AUC/NDCG and keep
    ranks[b] = 0, so they count as hits in recall@k (evaluator.py:105,129-131)."""
    n_pos = stats['n_pos']
    nb = len(n_pos)
    ranks = np.zeros(nb)
    queries = []
    for b in range(nb):
        if n_pos[b] != 0:
            n_neg = n_posts - int(n_pos[b])
            hits = stats['hits'][b]
            queries.append((int(stats['first_rank'][b]),
                            float(np.int64(stats['auc_num'][b])) / (int(n_pos[b]) * n_neg),
                            ndcg_from_hits(hits, int(n_pos[b]), 10, n_posts),
                            ndcg_from_hits(hits, int(n_pos[b]), 50, n_posts)))
            ranks[b] = stats['first_rank'][b]
    return _aggregate_queries(queries, ranks)


def rank_metrics_vec(scores, brands):
    return aggregate(rank_stats(scores, brands), np.asarray(scores).shape[1])


# ----------------------------------------------------------------------------
# candidate-list merge (multi-GPU exchange step)
# ----------------------------------------------------------------------------
def merge_topk(score_lists, index_lists, k):
    """Merge per-shard top-k lists ([G][NB, k_g] scores / global indices) into the
    global top-k under (score desc, index asc).  Entries with index < 0 are padding."""
    s = np.concatenate(score_lists, axis=1)
    i = np.concatenate(index_lists, axis=1).astype(np.int64)
    nb = s.shape[0]
    out_s = np.full((nb, k), -np.inf, dtype=np.float32)
    out_i = np.full((nb, k), -1, dtype=np.int64)
    for b in range(nb):
        valid = i[b] >= 0
        sb, ib = s[b][valid], i[b][valid]
        order = np.lexsort((ib, -sb.astype(np.float64)))[:k]
        out_s[b, :len(order)] = sb[order]
        out_i[b, :len(order)] = ib[order]
    return out_s, out_i
