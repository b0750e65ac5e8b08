"""Drop-in for the reference's loss.py.

  TripletLoss   loss.py:67-143   forward + backward fused on the device (frx_triplet_fwd_bwd): the B x B
                                 tile, rank weights, hinge, same-brand mask and dS are produced without the B
                                 GEMV launches, four sorts and B^2 host loop of the reference.
  LabLoss       loss.py:55-63    forward + backward fused on the device (frx_lab_fwd_bwd).
  cosine_sim / order_sim / euclidean_sim / l2norm: same formulas on torch device ops (helpers the reference's
                                 TripletLoss stores but never calls, loss.py:78-85).
As in the reference, `max_violation`, `measure` and `loss_fun` are accepted and do not change the
result (loss.py:85 vs :87-143), and `direction != 'all'` raises TypeError (loss.py:131-132).
"""
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

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


class _LabFn(Function):
    @staticmethod
    def forward(ctx, brand_embs):
        # the gradient tile / GEMM run only when autograd will ask for them (not under no_grad / validation)
        want = ctx.needs_input_grad[0]
        loss, d_brand = ops.lab_fwd_bwd(brand_embs, want)
        if want:
            ctx.save_for_backward(d_brand)
        return loss.reshape(())

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        d_brand, = ctx.saved_tensors
        return d_brand * grad_out


class LabLoss(nn.Module):
    """loss.py:55-63 fused on the device (frx_lab_fwd_bwd): cosine Gram tile on the tensor cores (3xTF32),
    exp / masked sum and the gradient tile in one row kernel."""

    def __init__(self):
        super(LabLoss, self).__init__()

    def forward(self, brand_embs):
        return _LabFn.apply(brand_embs.contiguous().float())


class _TripletFn(Function):
    @staticmethod
    def forward(ctx, brand_ids, brand_emb, post_emb, margin, mean_style):
        want = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]     # False under no_grad: forward kernels only
        loss, d_brand, d_post = ops.triplet_fwd_bwd(brand_ids, brand_emb, post_emb, margin, mean_style, want)
        if want:
            ctx.save_for_backward(d_brand, d_post)
        return loss.reshape(())

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        d_brand, d_post = ctx.saved_tensors
        return None, d_brand * grad_out, d_post * grad_out, None, None


class _VseppFn(Function):
    @staticmethod
    def forward(ctx, brand_ids, brand_emb, post_emb, margin, mean_style):
        want = ctx.needs_input_grad[1] or ctx.needs_input_grad[2]
        loss, d_brand, d_post = ops.vsepp_fwd_bwd(brand_ids, brand_emb, post_emb, margin, mean_style, want)
        if want:
            ctx.save_for_backward(d_brand, d_post)
        return loss.reshape(())

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        d_brand, d_post = ctx.saved_tensors
        return None, d_brand * grad_out, d_post * grad_out, None, None


class TripletLoss(nn.Module):
    """triplet ranking loss (rank-weighted hinge on the in-batch similarity tile)"""

    def __init__(self, margin=0, measure='cosine', max_violation=False, cost_style='sum', direction='all',
                 loss_fun='mrl', hardest_negative=False):
        """The reference's constructor (loss.py:79-85): `max_violation`, `measure` and `loss_fun` are stored and never
        read by forward -- so they are here.  `hardest_negative=True` is a NEW opt-in keyword (not in the reference): the
        VSE++ hinge over the hardest in-batch negative of every row and column (frx_vsepp_fwd_bwd), no rank weights."""
        super(TripletLoss, self).__init__()
        self.hardest_negative = bool(hardest_negative)
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
        fn = _VseppFn if self.hardest_negative else _TripletFn
        return fn.apply(ids, brand_emb.contiguous().float(), post_emb.contiguous().float(),
                        float(self.margin), 0 if self.cost_style == 'sum' else 1)
