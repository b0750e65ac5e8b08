"""Drop-in for the reference's loss.py.

  TripletLoss   loss.py:67-143   forward + backward fused on the device (frx_triplet_fwd_bwd): the B x B
                                 tile, rank weights, hinge, same-brand mask and dS are produced without the B
                                 GEMV launches, four sorts and B^2 host loop of the reference.
  LabLoss / cosine_sim / order_sim / euclidean_sim / l2norm: same formulas on torch device ops
                                 (not on the hot path north_star names).
As in the reference, `max_violation`, `measure` and `loss_fun` are accepted and do not change the
result (loss.py:85 vs :87-143), and `direction != 'all'` raises TypeError (loss.py:131-132).
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import ops
from .util.constant import device


def l2norm(X):
    norm = torch.pow(X, 2).sum(dim=1, keepdim=True).sqrt()
    return torch.div(X, norm)


def cosine_sim(im, s):
    return l2norm(im).mm(l2norm(s).t())


def order_sim(im, s):
    YmX = (s.unsqueeze(1).expand(s.size(0), im.size(0), s.size(1))
           - im.unsqueeze(0).expand(s.size(0), im.size(0), s.size(1)))
    return -YmX.clamp(min=0).pow(2).sum(2).sqrt().t()


def euclidean_sim(im, s):
    YmX = (s.unsqueeze(1).expand(s.size(0), im.size(0), s.size(1))
           - im.unsqueeze(0).expand(s.size(0), im.size(0), s.size(1)))
    return -YmX.pow(2).sum(2).t()


class LabLoss(nn.Module):
    def __init__(self):
        super(LabLoss, self).__init__()

    def forward(self, brand_embs):
        s = cosine_sim(brand_embs, brand_embs)
        eye = torch.eye(s.size(0), device=s.device) > .5
        s = s.masked_fill(eye, 0)
        return (torch.sum(torch.exp(s)) - s.size(0)) / s.size(0)


class _TripletFn(Function):
    @staticmethod
    def forward(ctx, brand_ids, brand_emb, post_emb, margin, mean_style):
        loss, d_brand, d_post = ops.triplet_fwd_bwd(brand_ids, brand_emb, post_emb, margin, mean_style, True)
        ctx.save_for_backward(d_brand, d_post)
        return loss.reshape(())

    @staticmethod
    def backward(ctx, grad_out):
        d_brand, d_post = ctx.saved_tensors
        return None, d_brand * grad_out, d_post * grad_out, None, None


class TripletLoss(nn.Module):
    """triplet ranking loss (rank-weighted hinge on the in-batch similarity tile)"""

    def __init__(self, margin=0, measure='cosine', max_violation=False, cost_style='sum', direction='all',
                 loss_fun='mrl'):
        super(TripletLoss, self).__init__()
        self.margin = margin
        self.cost_style = cost_style
        self.direction = direction
        self.loss_fun = loss_fun
        self.sim = {'order': order_sim, 'euclidean': euclidean_sim}.get(measure, cosine_sim)
        self.max_violation = max_violation

    def forward(self, brand_ids, brand_emb, post_emb):
        if self.direction != 'all':
            # loss.py:131-132 multiplies the missing direction's None by a tensor
            raise TypeError("unsupported operand type(s) for *: 'Tensor' and 'NoneType' "
                            "(direction=%r; the reference only works with 'all')" % (self.direction,))
        ids = torch.as_tensor(brand_ids).to(brand_emb.device, torch.int64).contiguous()
        return _TripletFn.apply(ids, brand_emb.contiguous().float(), post_emb.contiguous().float(),
                                float(self.margin), 0 if self.cost_style == 'sum' else 1)
