"""Host orchestration of the score -> top-k -> rank-statistics pipeline (one GPU, or one shard of a
post-sharded job -- see sharded.py).  All device work is done by libfrx_b200.so through ops.py;
the float64 metric values are computed here from the integer statistics with the same NumPy calls
the reference makes (evaluator.py:129-143), so the returned 8-tuple is bit-identical to the
reference's for identical scores.
"""
import numpy as np
import torch

from . import ops

HIT_DEPTH = 50          # evaluator.py:120 reads NDCG@50 -> 50 relevance bits per brand
MIN_TOPK = 64
DENSE_BUDGET_BYTES = 2 << 30          # row-chunked AUC sweep: bytes of one dense score tile
FUSED_DENSE_BUDGET_BYTES = 8 << 30    # AUC fast path: the fused top-k pass also writes the [NB, NP] fp32 scores when they fit


# Score precision (north_star: 1e-3 relative for bf16 inputs, 1e-5 for tf32):
#   "bf16"    bf16 operands, fp32 accumulation, full tensor rate               |err| <= 1e-3   (default)
#   "tf32"    fp32 operands read as tf32 (one pass, half rate)                 |err| <= 2e-4
#   "tf32x3"  3xTF32 split operands (K = 3D, ~6x the bf16 time), fp32-grade    |err| <= 1e-5
PRECISION = "bf16"
BRAND_SIDE, POST_SIDE = 0, 1


def contraction_depth(d, precision=None):
    """K extent of the tensor-core contraction for embedding size d (3d for the K-concatenated 3xTF32 operands)."""
    return 3 * d if (precision or PRECISION) == "tf32x3" else d


def to_operand(x_f32, final_norm=True, precision=None, side=POST_SIDE):
    """fp32 [N, D] embeddings -> L2-normalised score operand: bf16 [N, round_up(D, 64)] (zero padded), fp32
    [N, D] for the tf32 path, or fp32 [N, 3D] (`side`-specific hi/lo layout) for the 3xTF32 path (D % 4 == 0)."""
    precision = precision or PRECISION
    if precision in ("tf32", "tf32x3"):
        if x_f32.shape[1] % 4:
            raise ValueError("tf32 operands need D % 4 == 0")
        unit = ops.finalize_posts(x_f32, final_norm=final_norm, want_f32=True, want_bf16=False)[0]
        return ops.split_tf32x3(unit, side) if precision == "tf32x3" else unit
    if precision != "bf16":
        raise ValueError("unknown precision %r" % (precision,))
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
    # exact AUC needs every score once more (against the positives' scores, which this pass produces): when the matrix fits
    # the budget the fused pass writes it from its own accumulators, else the sweep re-contracts row chunks (auc_sweep)
    fused_dense = want_auc and dense_fits(nb, post_bf16.shape[0])
    res = ops.score_topk(brand_bf16, post_bf16, k, d=d, labels=labels_i32, index_base=index_base,
                         workspace=workspace, dense=fused_dense)
    n_pos, best_score, best_index = ops.label_stats(labels_i32, res["pos_score"], nb, index_base)
    hit_mask, first_in_list = ops.rank_from_topk(res["index"], labels_i32, index_base)
    out = dict(topk_scores=res["scores"], topk_index=res["index"], n_pos=n_pos, best_score=best_score,
               best_index=best_index, hit_mask=hit_mask, first_in_list=first_in_list, workspace=res["workspace"],
               pos_score=res["pos_score"])
    before_first = torch.zeros(nb, dtype=torch.int64, device=dev)
    if want_auc:
        seg_ptr, pos_sorted = ops.group_positives(labels_i32, res["pos_score"], n_pos)
        out["auc_num"] = auc_sweep(ops, brand_bf16, post_bf16, d, labels_i32, seg_ptr, pos_sorted, best_score,
                                   best_index, index_base, before_first, dense=res.get("dense"))
    else:
        # enqueued unconditionally: frx_score_count skips every 128-brand tile without a missing first positive
        thr_index = ops.missing_thresholds(n_pos, first_in_list, best_index)
        ops.score_count(brand_bf16, post_bf16, best_score, thr_index, d=d, index_base=index_base, out=before_first)
    out["before_first"] = before_first
    return out


def dense_fits(nb, n_posts):
    return 4 * nb * n_posts <= FUSED_DENSE_BUDGET_BYTES


def auc_sweep(kernels, brand_op, post_op, d, labels_i32, seg_ptr, pos_sorted, best_score, best_index, index_base,
              before_first, dense=None):
    """Exact AUC numerators (evaluator.py:111-113) of the posts in `post_op` against the sorted positives in
    (seg_ptr, pos_sorted) -- which may be the positives of the WHOLE job when `post_op` is one shard of it.
    `dense` = the [NB, NP] scores the fused top-k pass wrote on its way: one streaming pass over them.  Without it
    (matrix above FUSED_DENSE_BUDGET_BYTES) row chunks of at most DENSE_BUDGET_BYTES are re-contracted and swept, never
    the whole matrix.  Also accumulates into `before_first` the number of these posts that precede each brand's best
    positive.  Returns auc_num [NB] int64."""
    nb, n_posts = brand_op.shape[0], post_op.shape[0]
    auc_num = torch.zeros(nb, dtype=torch.int64, device=post_op.device)
    if dense is not None:
        kernels.auc_rows(dense, 0, labels_i32, seg_ptr, pos_sorted, best_score, best_index, auc_num, before_first,
                         index_base)
        return auc_num
    rows = max(1, min(nb, DENSE_BUDGET_BYTES // (4 * max(n_posts, 1))))
    if rows >= 128:
        rows = rows // 128 * 128
    dense = torch.empty((min(rows, nb), n_posts), dtype=torch.float32, device=post_op.device)
    for r0 in range(0, nb, rows):
        r1 = min(nb, r0 + rows)
        tile = dense[:r1 - r0]
        kernels.score_dense(brand_op[r0:r1], post_op, d=d, out=tile)
        kernels.auc_rows(tile, r0, labels_i32, seg_ptr, pos_sorted, best_score, best_index, auc_num, before_first,
                         index_base)
    return auc_num


def pack_statistics(dev_stats, want_auc=True, kernels=ops):
    """Everything the host needs, packed by one kernel into one int64 [5 | 6, NB] DEVICE tensor, so that the
    device->host read of an evaluation is ONE copy (`kernels`: provider of the device steps, as in sharded.py)."""
    return kernels.pack_rank_stats(dev_stats["n_pos"], dev_stats["first_in_list"], dev_stats["before_first"],
                                   dev_stats["hit_mask"], dev_stats["auc_num"] if want_auc else None,
                                   all_valid=want_auc)


def unpack_statistics(packed, n_posts, want_auc=True):
    """The packed block (NumPy, host) -> the integer per-brand arrays the metrics are functions of."""
    n_pos, first_in_list, before, valid, mask = packed[0], packed[1], packed[2], packed[3] != 0, packed[4]
    first_rank = np.where(valid, before, first_in_list)
    first_rank = np.where(n_pos > 0, first_rank, -1)
    depth = min(HIT_DEPTH, n_posts)
    # bit r of the 64-bit mask = relevance of rank r: unpack little-endian bytes, keep the first `depth` ranks
    hits = np.unpackbits(np.ascontiguousarray(mask).view(np.uint8).reshape(-1, 8), axis=1, bitorder='little')[:, :depth]
    st = dict(n_pos=n_pos.copy(), first_rank=first_rank, hits=hits)
    if want_auc:
        st["auc_num"] = packed[5].copy()
    return st


def host_statistics(dev_stats, n_posts, want_auc=True, kernels=ops):
    """Device statistics -> host integer arrays (blocking device->host copy; pipeline.py has the asynchronous form)."""
    return unpack_statistics(pack_statistics(dev_stats, want_auc, kernels).cpu().numpy(), n_posts, want_auc)


_IDEAL_CACHE = {}


def _ideal_table(depth):
    """(ideal DCG for m = 0..depth ones followed by zeros, discounts log2(2..depth), their reciprocals), each ideal value
    computed with the reference's expression (util/ndcg.py:37-42 on sorted(r, reverse=True)); cached per `depth`."""
    tab = _IDEAL_CACHE.get(depth)
    if tab is None:
        disc = np.log2(np.arange(2, depth + 1)) if depth > 1 else None
        tab = np.zeros(depth + 1, dtype=np.float64)
        for m in range(1, depth + 1):
            ideal = np.zeros(depth, dtype=np.float64)
            ideal[:m] = 1.0
            tab[m] = ideal[0] + (np.sum(ideal[1:] / disc) if depth > 1 else 0.0)
        inv = 1.0 / disc if depth > 1 else None
        _IDEAL_CACHE[depth] = (tab, disc, inv)
        tab = _IDEAL_CACHE[depth]
    return tab


def _ndcg_rows(hits_u8, n_pos, k, n_posts):
    """ndcg_at_k (util/ndcg.py:48-78) for every row of a 0/1 `hits` matrix at once.  A hit contributes
    1.0 / log2(i + 1), which is bit for bit the reference's r[i] / log2(i + 1) for r[i] = 1.0 (a miss contributes 0.0),
    and the row reductions run over the contiguous last axis of a fresh array, so NumPy applies to each row the same
    pairwise summation it applies to the reference's 1-D np.sum -- results are bit-identical (tests/test_abi.py)."""
    depth = min(k, n_posts, hits_u8.shape[1])
    table, disc, inv = _ideal_table(depth)
    dcg = hits_u8[:, 0].astype(np.float64)
    if depth > 1:
        dcg = dcg + np.sum(np.multiply(hits_u8[:, 1:depth], inv), axis=1)
    best = table[np.minimum(n_pos, depth)]
    if best.all():                       # every row has a positive (the caller filters on n_pos > 0): ideal DCG > 0
        return dcg / best
    out = np.zeros(len(n_pos), dtype=np.float64)
    nz = best != 0
    out[nz] = dcg[nz] / best[nz]
    return out


def aggregate(stats, n_posts, want_auc=True):
    """evaluator.py:105,115-143 on the integer statistics.  Returns
    (MedR, MeanR, AUC, NDCG@10, NDCG@50, r1, r5, r10); AUC is NaN when want_auc is False."""
    n_pos = np.asarray(stats["n_pos"], dtype=np.int64)
    nb = len(n_pos)
    has = n_pos > 0
    if not has.any():
        raise IndexError("no brand has a positive post (the reference fails the same way, evaluator.py:132-134)")
    first_rank = np.asarray(stats["first_rank"], dtype=np.int64)
    ranks = np.where(has, first_rank, 0).astype(np.float64)   # brands without positives keep 0 -> recall hits
    first = first_rank[has]
    r1 = 100.0 * len(np.where(ranks < 1)[0]) / len(ranks)
    r5 = 100.0 * len(np.where(ranks < 5)[0]) / len(ranks)
    r10 = 100.0 * len(np.where(ranks < 10)[0]) / len(ranks)
    if want_auc:
        num = np.asarray(stats["auc_num"], dtype=np.int64)[has].astype(np.float64)
        den = (n_pos[has] * (n_posts - n_pos[has])).astype(np.float64)   # exact below 2^53, as float(int)/int is
        if not den.all():        # a brand owns every post: float(num) / (len(pos) * 0), evaluator.py:117
            raise ZeroDivisionError("float division by zero")
        auc = np.average(num / den)
    else:
        auc = np.float64("nan")
    hits = np.asarray(stats["hits"])
    if not has.all():
        hits = hits[has]
    n10 = _ndcg_rows(hits, n_pos[has], 10, n_posts)
    n50 = _ndcg_rows(hits, n_pos[has], 50, n_posts)
    return (np.floor(np.median(first)), np.floor(np.mean(first)), auc,
            np.average(n10), np.average(n50), r1, r5, r10)


def rank_posts(brand_f32, post_f32, labels, k=MIN_TOPK, want_auc=True):
    """Single-GPU convenience: fp32 brand [NB, D] / post [NP, D] embeddings + labels -> (8-tuple, stats)."""
    labels_i32 = labels.to(torch.int32).contiguous()
    d = contraction_depth(post_f32.shape[1])
    brand_bf16 = to_operand(brand_f32.contiguous().float(), side=BRAND_SIDE)
    post_bf16 = to_operand(post_f32.contiguous().float(), side=POST_SIDE)
    dev_stats = device_rank_statistics(brand_bf16, post_bf16, labels_i32, d, k=k, want_auc=want_auc)
    stats = host_statistics(dev_stats, post_f32.shape[0], want_auc)
    return aggregate(stats, post_f32.shape[0], want_auc), stats, dev_stats
