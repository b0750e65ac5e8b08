"""Evaluation tail of the reference's tester.py (tester.py:102-113) as a function: the CLI around it
(argparse, checkpoint and vocabulary files, data loaders) is control plane and stays with the reference,
which can call this with the objects it has built.

    model.load_state_dict(checkpoint['model'])  ->  evaluate(options, model, data_loader['test'], log_step)
"""
import logging

from . import evaluator
from .evaluator import test_post_ranking


def evaluate(options, model, data_loader, log_step=10):
    """encode_data + test_post_ranking + the four printed lines of tester.py:110-113.  Returns the 8-tuple
    (MedR, MeanR, AUC, NDCG@10, NDCG@50, r1, r5, r10)."""
    brands, post_embs = evaluator.encode_data(model, data_loader, log_step, logging.info)
    ranking_metrics = test_post_ranking(options.brand_num, options.metric, model, post_embs, brands)
    print('AUC[0-1]:', ranking_metrics[2])
    print('NDCG@10[0-1]:', ranking_metrics[3])
    print('NDCG@50[0-1]:', ranking_metrics[4])
    print('recall@1:', ranking_metrics[5])
    return ranking_metrics
