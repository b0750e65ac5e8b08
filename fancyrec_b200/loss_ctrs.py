"""Drop-in for the reference's loss_ctrs.py.

  ContrastiveLoss             loss_ctrs.py:120-214  forward + backward on the device
                              (frx_contrastive_fwd_bwd); queue / pointer semantics, including the
                              positive mask read AFTER the pointer moves (loss_ctrs.py:149-159) and the
                              errors for Q % B != 0, are the reference's.
  CrossCLR_onlyIntraModality  loss_ctrs.py:28-117   forward + backward on the device (frx_crossclr_fwd_bwd).
"""
import numpy as np
import torch
import torch.nn.functional as F
from torch import nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import ops
from .util.constant import device


def l2norm(X):
    norm = torch.pow(X, 2).sum(dim=1, keepdim=True).sqrt()
    return torch.div(X, norm)


def cosine_sim(im, s):
    return l2norm(im).mm(l2norm(s).t())


class _CrossCLRFn(Function):
    @staticmethod
    def forward(ctx, brand, post, temperature, negative_w, mean_style):
        want = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]     # False under no_grad: forward kernels only
        loss, d_brand, d_post = ops.crossclr_fwd_bwd(brand, post, temperature, negative_w, mean_style, want)
        if want:
            ctx.save_for_backward(d_brand, d_post)
        return loss.reshape(())

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        d_brand, d_post = ctx.saved_tensors
        return d_brand * grad_out, d_post * grad_out, None, None, None


class CrossCLR_onlyIntraModality(nn.Module):
    """loss_ctrs.py:28-117 fused on the device (frx_crossclr_fwd_bwd): rank weights by counting instead of four
    sorts, the four Gram tiles and the two gradient GEMMs on the tensor cores (3xTF32), both soft-maxes in one row
    kernel."""

    def __init__(self, temperature=0.03, negative_weight=0.8, logger=None, cost_style='sum'):
        super(CrossCLR_onlyIntraModality, self).__init__()
        self.logit_scale = nn.Parameter(torch.ones([]))
        self.criterion = torch.nn.CrossEntropyLoss(reduction='none')
        self.temperature = temperature
        self.logger = logger
        self.cost_style = cost_style
        self.negative_w = negative_weight

    def compute_loss(self, logits, mask):
        return - torch.log((F.softmax(logits, dim=1) * mask).sum(1))

    def forward(self, brand, post):
        return _CrossCLRFn.apply(brand.contiguous().float(), post.contiguous().float(), float(self.temperature),
                                 float(self.negative_w), 0 if self.cost_style == 'sum' else 1)


class _ContrastiveFn(Function):
    @staticmethod
    def forward(ctx, brand, post, keys, mask_col0, no_intra, temperature, negative_w, mean_style):
        want = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]     # False under no_grad: forward kernels only
        loss, d_brand, d_post = ops.contrastive_fwd_bwd(brand, post, keys, mask_col0, no_intra, temperature,
                                                        negative_w, mean_style, want)
        if want:
            ctx.save_for_backward(d_brand, d_post)
        return loss.reshape(())

    @staticmethod
    @once_differentiable
    def backward(ctx, grad_out):
        d_brand, d_post = ctx.saved_tensors
        return d_brand * grad_out, d_post * grad_out, None, None, None, None, None, None


class ContrastiveLoss(nn.Module):
    def __init__(self, opt, temperature=0.03, negative_weight=0.8):
        super(ContrastiveLoss, self).__init__()
        self.opt = opt
        self.logit_scale = nn.Parameter(torch.ones([]))
        self.criterion = torch.nn.CrossEntropyLoss(reduction='none')
        self.temperature = temperature
        self.cost_style = opt.cost_style
        self.negative_w = negative_weight
        self.register_buffer("queue", torch.zeros(opt.queue_size, opt.common_embedding_size))
        self.register_buffer("queue_ptr", torch.zeros(1, dtype=torch.long))

    @torch.no_grad()
    def _dequeue_and_enqueue(self, post):
        batch_size = post.shape[0]
        ptr = int(self.queue_ptr)
        self.queue[ptr: ptr + batch_size] = post        # RuntimeError when the slice is short, as in the reference
        self.queue_ptr[0] = (ptr + batch_size) % self.opt.queue_size

    def forward(self, brand, post):
        brand = brand.contiguous().float()
        post = post.contiguous().float()
        b = brand.shape[0]
        if self.opt.no_queue or self.opt.no_intra:
            keys, n_cols = None, b
        else:
            self._dequeue_and_enqueue(ops.normalize_rows(post.detach()))
            keys, n_cols = self.queue, self.queue.shape[0]
        ptr = int(self.queue_ptr[0])
        if ptr + b > n_cols:
            # loss_ctrs.py:155-158: mask[i][ptr + i] runs off the row
            raise IndexError("index %d is out of bounds for dimension 0 with size %d" % (n_cols, n_cols))
        return _ContrastiveFn.apply(brand, post, keys, ptr, bool(self.opt.no_intra), float(self.temperature),
                                    float(self.negative_w), 0 if self.cost_style == 'sum' else 1)
