"""Hot-path pieces of the reference's model.py behind the same names.

  l2norm        model.py:39-44
  L1Penalty     model.py:389-402   (identity forward; backward adds 1e-4 * sign(input))
  BrandAspects  model.py:406-428   (same parameters / state_dict keys; .forward keeps the [B, A, D]
                                    contract, .embed() is the fused (W.E)/A kernel used at eval time)
  MFC           model.py:59-83     Linear + ReLU (+ dropout): eval path = ONE tensor-core GEMM with the bias / ReLU
                                    epilogue (ops.linear -> frx_linear, 3xTF32 split inside the kernel)
  PrjHeadFusionEncoder model.py:463-491  concat -> fc1 -> BatchNorm1d -> ReLU -> fc2: eval path = concat kernel + two
                                    GEMMs, the BatchNorm folded into the first one's scale / shift epilogue
  FancyRec      model.py:538-649   shell only: brand side + post finalisation are ours, the learned
                                    visual / text / fusion encoders (out of scope) are
                                    injected by the caller, e.g. the reference's own classes.
Same parameter names / state_dict keys as the reference classes, so reference checkpoints load unchanged.  Training
(autograd) runs the reference's torch formulas; the fused kernels serve the no-grad path that encode_data takes.
"""
import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function

from . import ops
from .util.constant import device


def l2norm(X):
    """L2-normalize the rows of X.  Differentiable inputs stay on the autograd path (torch ops, the
    reference formula); detached CUDA tensors take the one-pass kernel."""
    if X.requires_grad or not X.is_cuda:
        if not X.is_cuda:
            raise RuntimeError("fancyrec_b200.model.l2norm needs a CUDA tensor (no CPU fallback)")
        norm = torch.pow(X, 2).sum(dim=1, keepdim=True).sqrt()
        return torch.div(X, norm)
    return ops.finalize_posts(X.contiguous().float(), final_norm=True, want_f32=True, want_bf16=False)[0]


def xavier_init_fc(fc, bias=True):
    """Xavier initialization for the fully connected layer (model.py:47-54)."""
    r = np.sqrt(6.) / np.sqrt(fc.in_features + fc.out_features)
    fc.weight.data.uniform_(-r, r)
    if bias:
        fc.bias.data.fill_(0)


def _fused_ok(module, x):
    """The kernels serve inference: eval mode, no autograd graph wanted, CUDA fp32 input."""
    return (not module.training) and (not torch.is_grad_enabled() or not x.requires_grad) and x.is_cuda and \
        x.dtype == torch.float32 and x.dim() == 2


def fold_batchnorm(bn):
    """Eval-mode BatchNorm1d as a per-column (scale, shift): y = x * scale + shift."""
    scale = (bn.weight if bn.affine else torch.ones_like(bn.running_var)) / torch.sqrt(bn.running_var + bn.eps)
    shift = (bn.bias if bn.affine else torch.zeros_like(bn.running_mean)) - bn.running_mean * scale
    return scale.detach().float().contiguous(), shift.detach().float().contiguous()


class MFC(nn.Module):
    """Multi Fully Connected Layers (model.py:59-83): fc1 -> ReLU -> dropout."""

    def __init__(self, fc_layers, dropout):
        super(MFC, self).__init__()
        self.fc1 = nn.Linear(fc_layers[0], fc_layers[1])
        self.dropout = nn.Dropout(p=dropout)
        self.relu = nn.ReLU()
        self.init_weights()

    def init_weights(self):
        xavier_init_fc(self.fc1)

    def forward(self, inputs):
        if _fused_ok(self, inputs):      # eval: dropout is the identity; bias + ReLU live in the GEMM epilogue
            return ops.linear(inputs.contiguous(), self.fc1.weight.detach(), bias=self.fc1.bias.detach(), relu=True)
        return self.dropout(self.relu(self.fc1(inputs)))


class PrjHeadFusionEncoder(nn.Module):
    """Non-linear projection head over the concatenated branches (model.py:463-491)."""

    def __init__(self, opt):
        super(PrjHeadFusionEncoder, self).__init__()
        self.opt = opt
        self.common_embedding_size = opt.common_embedding_size
        self.visual_mapping_size = opt.visual_mapping_size[1]
        self.text_mapping_size = opt.text_mapping_size[1]
        self.fc1 = nn.Linear(self.text_mapping_size + self.visual_mapping_size, 512, bias=False)
        self.fc2 = nn.Linear(512, self.common_embedding_size, bias=True)
        self.projection_head = nn.Sequential(self.fc1, nn.BatchNorm1d(512), nn.ReLU(), self.fc2)
        self.init_weights()

    def forward(self, visual_embs, text_embs):
        if _fused_ok(self, visual_embs) and _fused_ok(self, text_embs):
            # concat in one pass of the finalisation kernel (no norms), then fc1 (+ folded BatchNorm + ReLU) and fc2 (+ bias)
            fusion_vt = ops.finalize_posts(visual_embs.contiguous(), text_embs.contiguous(), final_norm=False,
                                           want_f32=True, want_bf16=False)[0]
            if self.opt.prj_head_output:
                return fusion_vt
            scale, shift = fold_batchnorm(self.projection_head[1])
            hidden = ops.linear(fusion_vt, self.fc1.weight.detach(), bias=shift, col_scale=scale, relu=True)
            return ops.linear(hidden, self.fc2.weight.detach(), bias=self.fc2.bias.detach())
        fusion_vt = torch.cat((visual_embs, text_embs), 1)
        if self.opt.prj_head_output:
            return fusion_vt
        return self.projection_head(fusion_vt)

    def init_weights(self):
        xavier_init_fc(self.fc1, bias=False)
        xavier_init_fc(self.fc2)


class L1Penalty(Function):
    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        return input.clone()

    @staticmethod
    def backward(ctx, grad_output):
        input, = ctx.saved_tensors
        return input.sign().mul(0.0001) + grad_output


class _BrandTrainFn(Function):
    """Dropout(0.5) on the [B, A, D] products + mean over aspects + L1Penalty, fused (frx_brand_train_fwd / _bwd): the
    keep bits are a counter-based hash of (seed, b, a, d), regenerated in the backward instead of stored."""

    @staticmethod
    def forward(ctx, w_rows, aspects, seed):
        w_rows, aspects = w_rows.contiguous().float(), aspects.contiguous().float()
        ctx.save_for_backward(w_rows, aspects)
        ctx.seed = seed
        return ops.brand_train_fwd(w_rows, aspects, seed)

    @staticmethod
    def backward(ctx, grad_out):
        w_rows, aspects = ctx.saved_tensors
        d_w, d_e = ops.brand_train_bwd(grad_out.contiguous().float(), w_rows, aspects, ctx.seed)
        return d_w, d_e, None


class BrandAspects(nn.Module):
    def __init__(self, opt):
        super(BrandAspects, self).__init__()
        self.brand_num = opt.brand_num
        self.common_embedding_size = opt.common_embedding_size
        self.num_aspects = opt.brand_aspect
        self.brand_embeddings = nn.Embedding(self.brand_num + 1, self.num_aspects)
        self.aspects_embeddings = nn.Parameter(
            torch.randn(self.num_aspects, self.common_embedding_size), requires_grad=True)
        self.dropout = nn.Dropout()

    def forward(self, brand_list):
        """[B] ids -> [B, A, D] weighted aspects (reference contract; training path with dropout)."""
        w = L1Penalty.apply(self.brand_embeddings(brand_list))
        return self.dropout(w.unsqueeze(2) * self.aspects_embeddings.unsqueeze(0))

    def embed_train(self, brand_list, seed=None):
        """Training-time [B, D] brand embedding = forward(brand_list).permute(1, 0, 2).mean(0) of the reference
        (model.py:594) without the [B, A, D] intermediate: dropout(0.5) on the products and the L1Penalty gradient are
        fused into the kernels.  `seed` fixes the dropout mask (default: drawn from torch's CPU generator)."""
        if self.dropout.p != 0.5:
            raise ValueError("the fused path implements nn.Dropout()'s default p = 0.5 (model.py:417)")
        if seed is None:
            seed = int(torch.empty((), dtype=torch.int64).random_().item())
        return _BrandTrainFn.apply(self.brand_embeddings(brand_list), self.aspects_embeddings, seed)

    def embed(self, brand_list):
        """Eval-time fast path: mean over aspects -> [B, D] without materialising [B, A, D]."""
        return ops.brand_embed(self.brand_embeddings.weight.detach().float().contiguous(),
                               self.aspects_embeddings.detach().float().contiguous(),
                               brand_ids=brand_list.to(self.aspects_embeddings.device, torch.int64).contiguous())


class FancyRec(nn.Module):
    """Same attribute / method surface as the reference wrapper (model.py:538-649)."""

    def __init__(self, opt, vid_encoding=None, text_encoding=None, fusion_encoding=None):
        super(FancyRec, self).__init__()
        self.opt = opt
        self.brand_encoding = BrandAspects(opt)
        params1 = list(self.brand_encoding.parameters())
        self.vid_encoding = vid_encoding
        self.text_encoding = text_encoding
        self.fusion_encoding = fusion_encoding
        self.text_net = getattr(opt, 'text_net', None)
        self.fusion_style = getattr(opt, 'fusion_style', None)
        for enc in (vid_encoding, text_encoding, fusion_encoding):
            if enc is not None:
                params1 += list(enc.parameters())
        self.params1 = params1
        self.Eiters = 0

    def forward(self, brand_ids, videos, captions):
        brand_embs = self.embed_brand(brand_ids)
        if self.opt.single_modal_visual:
            post_embs = self.embed_vis(videos)
        elif self.opt.single_modal_text:
            post_embs = self.embed_txt(captions)
        else:
            post_embs = self.fusion_encoding(self.embed_vis(videos), self.embed_txt(captions))
        return brand_embs, post_embs

    def embed_brand(self, brand_ids, volatile=True):
        brand_ids = brand_ids.to(device)
        if not self.brand_encoding.training:
            if not torch.is_grad_enabled():
                return self.brand_encoding.embed(brand_ids)
            return self.brand_encoding(brand_ids).permute((1, 0, 2)).mean(0)      # eval with autograd: no dropout, torch path
        return self.brand_encoding.embed_train(brand_ids)                          # training: fused, no [B, A, D] tensor

    def embed_vis(self, vis_data, volatile=True):
        frames, mean_origin, video_lengths, vidoes_mask = vis_data
        data = (frames.to(device), mean_origin.to(device), video_lengths, vidoes_mask.to(device))
        return self.vid_encoding(data)

    def embed_txt(self, text_data, volatile=True):
        moved = tuple(t.to(device) if isinstance(t, torch.Tensor) else t for t in text_data)
        return self.text_encoding(moved)

    def state_dict(self):
        return [self.vid_encoding.state_dict(), self.text_encoding.state_dict(),
                self.brand_encoding.state_dict(), self.fusion_encoding.state_dict()]

    def load_state_dict(self, state_dict):
        self.vid_encoding.load_state_dict(state_dict[0])
        self.text_encoding.load_state_dict(state_dict[1])
        self.brand_encoding.load_state_dict(state_dict[2])
        self.fusion_encoding.load_state_dict(state_dict[3])
