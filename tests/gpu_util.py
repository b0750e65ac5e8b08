"""Helpers shared by the -m gpu tests (call the product through ops.py -> C ABI)."""
import numpy as np
import torch


def dev():
    return torch.device("cuda:0")


def to_dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev())


def operand(x_np, final_norm=True):
    from fancyrec_b200 import ranking
    return ranking.to_operand(to_dev(x_np.astype(np.float32)), final_norm=final_norm)
