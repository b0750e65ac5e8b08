"""The evaluation-side entry points of the reference's trainer.py, behind the same names.

  validate         trainer.py:398-416   encode_data -> test_post_ranking -> the eight printed lines ->
                                        (rsum, AUC, NDCG@10, NDCG@50, MedR, MeanR, r1, r5, r10)
  save_checkpoint  trainer.py:419-424   checkpoint-selection rule (keep when within 1 % of the best, copy when best)

The training loop itself (optimisers, schedulers, TensorBoard, argparse) is control plane and stays with the
reference; it only needs these two names plus fancyrec_b200.loss / loss_ctrs to run on the B200 kernels.
"""
import logging
import shutil

import torch

from . import evaluator
from .evaluator import test_post_ranking


def validate(opt, val_loader, model):
    brands, post_embs = evaluator.encode_data(model, val_loader, opt.log_step, logging.info)
    MedR, MeanR, AUC, NDCG_10, NDCG_50, r1, r5, r10 = test_post_ranking(opt.brand_num, opt.metric, model, post_embs,
                                                                        brands)
    print('MedR:', MedR)
    print('MeanR:', MeanR)
    print('AUC[0-1]:', AUC)
    print('NDCG@10[0-1]:', NDCG_10)
    print('NDCG@50[0-1]:', NDCG_50)
    print('recall@1:', r1)
    print('recall@5:', r5)
    print('recall@10:', r10)
    rsum = 0.0
    rsum += ((AUC + NDCG_10 + NDCG_50) * 100 + r1 + r5 + r10)
    return rsum, AUC, NDCG_10, NDCG_50, MedR, MeanR, r1, r5, r10


def save_checkpoint(state, sum, best_rsum, filename='checkpoint.pth.tar', prefix='', best_epoch=None):
    if best_epoch is None or sum > best_rsum * 0.99:
        torch.save(state, prefix + filename)
    if sum > best_rsum:
        shutil.copyfile(prefix + filename, prefix + 'model_best.pth.tar')
    return max(sum, best_rsum)
