"""Drop-in for the reference's evaluator.py (same names, positional signatures, return types).

  l2norm            evaluator.py:14-19      -> frx_finalize_posts (row L2 norm, one HBM pass)
  cal_sim           evaluator.py:23-29      -> L2-normalise to bf16 operands + tcgen05 dense score tile
  random_sim        evaluator.py:33-34
  encode_data       evaluator.py:38-81      (host loop over the loader; same outputs, same quirks)
  test_post_ranking evaluator.py:85-143     -> brand embed kernel, fused score + top-k GEMM, rank-statistic
                                               kernels; float64 aggregation on the host as the reference does

The reference copies the whole [NB, NP] score matrix to the host (evaluator.py:96) and ranks it in
Python; here the matrix never leaves the tensor-core epilogue.
"""
import time

import numpy as np
import torch

from . import ops, ranking
from .util.constant import device


class AverageMeter(object):
    """Running average (util/util.py AverageMeter contract: val, avg, sum, count, update)."""

    def __init__(self):
        self.val = self.avg = self.sum = self.count = 0

    def update(self, val, n=1):
        self.val = val
        self.sum += val * n
        self.count += n
        self.avg = self.sum / (.0001 + self.count)


def l2norm(X):
    """L2-normalize the rows of X (no epsilon: a zero row gives NaN, like the reference)."""
    return ops.finalize_posts(X.contiguous().float(), final_norm=True, want_f32=True, want_bf16=False)[0]


def cal_sim(im, s):
    """Cosine similarity between every (brand, post) pair -> [M, N] fp32 on the device
    (default: bf16 operands, fp32 accumulation, |error| <= 1e-3 on the cosine scale; ranking.PRECISION =
    "tf32x3" gives |error| <= 1e-5 on the tf32 tensor-core path)."""
    a = ranking.to_operand(im.contiguous().float(), side=ranking.BRAND_SIDE)
    b = ranking.to_operand(s.contiguous().float(), side=ranking.POST_SIDE)
    return ops.score_dense(a, b, d=ranking.contraction_depth(im.shape[1]))


def random_sim(num_brands, num_test_posts):
    return np.random.rand(num_brands, num_test_posts)


def encode_data(model, data_loader, log_step=10, logging=print):
    """Encode every post `data_loader` yields.  Returns (brands [NP] on device in LOADER order,
    post_embs [NP, common_embedding_size] on device in DATASET order) -- the reference's contract,
    including that asymmetry (evaluator.py:62 vs :69)."""
    batch_time = AverageMeter()
    model.brand_encoding.eval()
    if model.opt.single_modal_text:
        model.text_encoding.eval()
    elif model.opt.single_modal_visual:
        model.vid_encoding.eval()
    else:
        model.vid_encoding.eval()
        model.text_encoding.eval()
        model.fusion_encoding.eval()
    end = time.time()
    brands = torch.tensor([], dtype=torch.int).to(device)
    post_embs = torch.zeros((len(data_loader.dataset), model.opt.common_embedding_size)).to(device)
    with torch.no_grad():
        for i, (brand_ids, videos, captions, idxs, cap_ids, vid_ids) in enumerate(data_loader):
            brand_ids = brand_ids.to(device)
            brands = torch.cat((brands, brand_ids), 0)
            _, post_emb = model(brand_ids, videos, captions)
            post_embs[np.array(idxs)] = post_emb
            batch_time.update(time.time() - end)
            end = time.time()
            if i % log_step == 0:
                logging('Process: [{0:2d}/{1:2d}]\t'
                        'Time {batch_time.val:.3f} ({batch_time.avg:.3f})\t'.format(
                            i, len(data_loader), batch_time=batch_time))
            del videos, captions
    return brands, post_embs


def brand_matrix(model, brand_num):
    """evaluator.py:89-94 without the [NB, A, D] intermediate: mean over aspects of W[b, a] * E[a, :]."""
    enc = model.brand_encoding.eval()
    w = enc.brand_embeddings.weight.detach().float().contiguous()
    e = enc.aspects_embeddings.detach().float().contiguous()
    return ops.brand_embed(w, e, nb=brand_num)


def test_post_ranking(brand_num, metric, model, post_embs, brands):
    """Returns (MedR, MeanR, AUC, NDCG@10, NDCG@50, r1, r5, r10) for metric == 'auc', else None
    (evaluator.py:103).  Ranking order: (score descending, post index ascending)."""
    if metric != 'auc':
        return None
    brand = brand_matrix(model, brand_num)
    result, _, _ = ranking.rank_posts(brand, post_embs, brands, want_auc=True)
    return result


test_post_ranking.__test__ = False   # not a pytest test


def test_post_ranking_sharded(brand_num, metric, model, post_embs_local, brands_local, n_posts_total=None, group=None):
    """test_post_ranking for a job whose posts are sharded over the ranks of a torch.distributed group (one process
    per GPU; rank r holds the contiguous post range sharded.shard_bounds(n_posts_total, world, r)).  Returns on every
    rank the SAME 8-tuple the single-GPU call returns for the concatenated posts, exact AUC included: one packed
    all-gather of candidate lists / label statistics / labels / positives' scores, then two NB-length all-reduces."""
    if metric != 'auc':
        return None
    import torch.distributed as dist
    from . import sharded
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if n_posts_total is None:
        n = torch.tensor([post_embs_local.shape[0]], dtype=torch.int64, device=post_embs_local.device)
        if world > 1:
            dist.all_reduce(n, group=group)
        n_posts_total = int(n.item())
    brand = brand_matrix(model, brand_num)
    d = ranking.contraction_depth(post_embs_local.shape[1])
    brand_op = ranking.to_operand(brand, side=ranking.BRAND_SIDE)
    post_op = ranking.to_operand(post_embs_local.contiguous().float(), side=ranking.POST_SIDE)
    st = sharded.sharded_rank_statistics(brand_op, post_op, brands_local.to(torch.int32).contiguous(), d,
                                         ranking.MIN_TOPK, n_posts_total, group=group, want_auc=True)
    stats = ranking.host_statistics(st, n_posts_total, want_auc=True)
    return ranking.aggregate(stats, n_posts_total, want_auc=True)


test_post_ranking_sharded.__test__ = False
