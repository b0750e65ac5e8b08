"""Host orchestration of the score -> top-k -> rank-statistics pipeline (one GPU, or one shard of a
post-sharded job -- see sharded.py).  All device work is done by libfrx_b200.so through ops.py;
the float64 metric values are computed here from the integer statistics with the same NumPy calls
the reference makes (evaluator.py:129-143), so the returned 8-tuple is bit-identical to the
reference's for identical scores.
"""
import numpy as np
import torch

from . import ops
from .util.ndcg import ndcg_from_hits

HIT_DEPTH = 50          # evaluator.py:120 reads NDCG@50 -> 50 relevance bits per brand
MIN_TOPK = 64
DENSE_BUDGET_BYTES = 2 << 30


def to_operand(x_f32, final_norm=True):
    """fp32 [N, D] embeddings -> L2-normalised bf16 operand [N, round_up(D, 64)] (zero padded)."""
    return ops.finalize_posts(x_f32, final_norm=final_norm, want_f32=False, want_bf16=True)[1]


def device_rank_statistics(brand_bf16, post_bf16, labels_i32, d, k=MIN_TOPK, want_auc=True, index_base=0,
                           workspace=None):
    """Runs the fused score+top-k kernel and the statistic kernels on ONE shard.

    Returns a dict of DEVICE tensors:
      topk_scores/topk_index [NB, k], n_pos [NB] i32, best_score/best_index [NB],
      hit_mask [NB] (u64 bits in an int64), first_in_list [NB] i32,
      before_first [NB] i64 (valid where computed: all brands when want_auc, else only the brands
      whose first positive fell outside the list), auc_num [NB] i64 (want_auc only).
    """
    nb = brand_bf16.shape[0]
    dev = post_bf16.device
    k = max(int(k), min(MIN_TOPK, 1024))
    res = ops.score_topk(brand_bf16, post_bf16, k, d=d, labels=labels_i32, index_base=index_base,
                         workspace=workspace)
    n_pos, best_score, best_index = ops.label_stats(labels_i32, res["pos_score"], nb, index_base)
    hit_mask, first_in_list = ops.rank_from_topk(res["index"], labels_i32, index_base)
    out = dict(topk_scores=res["scores"], topk_index=res["index"], n_pos=n_pos, best_score=best_score,
               best_index=best_index, hit_mask=hit_mask, first_in_list=first_in_list, workspace=res["workspace"],
               pos_score=res["pos_score"])
    before_first = torch.zeros(nb, dtype=torch.int64, device=dev)
    if want_auc:
        auc_num = torch.zeros(nb, dtype=torch.int64, device=dev)
        seg_ptr, pos_sorted = ops.group_positives(labels_i32, res["pos_score"], n_pos)
        n_posts = post_bf16.shape[0]
        rows = max(1, min(nb, DENSE_BUDGET_BYTES // (4 * n_posts)))
        if rows >= 128:
            rows = rows // 128 * 128
        dense = torch.empty((min(rows, nb), n_posts), dtype=torch.float32, device=dev)
        for r0 in range(0, nb, rows):
            r1 = min(nb, r0 + rows)
            tile = dense[:r1 - r0]
            ops.score_dense(brand_bf16[r0:r1], post_bf16, d=d, out=tile)
            ops.auc_rows(tile, r0, labels_i32, seg_ptr, pos_sorted, best_score, best_index, auc_num, before_first,
                         index_base)
        out["auc_num"] = auc_num
        out["before_first_valid"] = torch.ones(nb, dtype=torch.bool, device=dev)
    else:
        missing = (first_in_list < 0) & (n_pos > 0)
        if bool(missing.any().item()):
            thr_index = torch.where(missing, best_index, torch.full_like(best_index, -1))
            ops.score_count(brand_bf16, post_bf16, best_score, thr_index, d=d, index_base=index_base,
                            out=before_first)
        out["before_first_valid"] = missing
    out["before_first"] = before_first
    return out


def host_statistics(dev_stats, n_posts, want_auc=True):
    """Device statistics -> the integer per-brand arrays the metrics are functions of (NumPy, host)."""
    n_pos = dev_stats["n_pos"].cpu().numpy().astype(np.int64)
    first_in_list = dev_stats["first_in_list"].cpu().numpy().astype(np.int64)
    before = dev_stats["before_first"].cpu().numpy().astype(np.int64)
    valid = dev_stats["before_first_valid"].cpu().numpy()
    first_rank = np.where(valid, before, first_in_list)
    first_rank = np.where(n_pos > 0, first_rank, -1)
    mask = dev_stats["hit_mask"].cpu().numpy().view(np.uint64)
    depth = min(HIT_DEPTH, n_posts)
    hits = ((mask[:, None] >> np.arange(depth, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(np.uint8)
    st = dict(n_pos=n_pos, first_rank=first_rank, hits=hits)
    if want_auc:
        st["auc_num"] = dev_stats["auc_num"].cpu().numpy().astype(np.int64)
    return st


def aggregate(stats, n_posts, want_auc=True):
    """evaluator.py:105,115-143 on the integer statistics.  Returns
    (MedR, MeanR, AUC, NDCG@10, NDCG@50, r1, r5, r10); AUC is NaN when want_auc is False."""
    n_pos = stats["n_pos"]
    nb = len(n_pos)
    ranks = np.zeros(nb)                      # brands without positives keep 0 -> count as recall hits
    first, aucs, n10, n50 = [], [], [], []
    for b in range(nb):
        if n_pos[b] == 0:
            continue
        ranks[b] = stats["first_rank"][b]
        first.append(int(stats["first_rank"][b]))
        if want_auc:
            aucs.append(float(np.int64(stats["auc_num"][b])) / (int(n_pos[b]) * (n_posts - int(n_pos[b]))))
        n10.append(ndcg_from_hits(stats["hits"][b], n_pos[b], 10, n_posts))
        n50.append(ndcg_from_hits(stats["hits"][b], n_pos[b], 50, n_posts))
    if not first:
        raise IndexError("no brand has a positive post (the reference fails the same way, evaluator.py:132-134)")
    r1 = 100.0 * len(np.where(ranks < 1)[0]) / len(ranks)
    r5 = 100.0 * len(np.where(ranks < 5)[0]) / len(ranks)
    r10 = 100.0 * len(np.where(ranks < 10)[0]) / len(ranks)
    return (np.floor(np.median(tuple(first))), np.floor(np.mean(tuple(first))),
            np.average(tuple(aucs)) if want_auc else np.float64("nan"),
            np.average(tuple(n10)), np.average(tuple(n50)), r1, r5, r10)


def rank_posts(brand_f32, post_f32, labels, k=MIN_TOPK, want_auc=True):
    """Single-GPU convenience: fp32 brand [NB, D] / post [NP, D] embeddings + labels -> (8-tuple, stats)."""
    labels_i32 = labels.to(torch.int32).contiguous()
    d = post_f32.shape[1]
    brand_bf16 = to_operand(brand_f32.contiguous().float())
    post_bf16 = to_operand(post_f32.contiguous().float())
    dev_stats = device_rank_statistics(brand_bf16, post_bf16, labels_i32, d, k=k, want_auc=want_auc)
    stats = host_statistics(dev_stats, post_f32.shape[0], want_auc)
    return aggregate(stats, post_f32.shape[0], want_auc), stats, dev_stats
