"""Hot-path pieces of the reference's model.py behind the same names.

  l2norm        model.py:39-44
  L1Penalty     model.py:389-402   (identity forward; backward adds 1e-4 * sign(input))
  BrandAspects  model.py:406-428   (same parameters / state_dict keys; .forward keeps the [B, A, D]
                                    contract, .embed() is the fused (W.E)/A kernel used at eval time)
  FancyRec      model.py:538-649   shell only: brand side + post finalisation are ours, the learned
                                    visual / text / fusion encoders (out of scope) are
                                    injected by the caller, e.g. the reference's own classes.
"""
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


class L1Penalty(Function):
    @staticmethod
    def forward(ctx, input):
        ctx.save_for_backward(input)
        return input.clone()

    @staticmethod
    def backward(ctx, grad_output):
        input, = ctx.saved_tensors
        return input.sign().mul(0.0001) + grad_output


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
        if not self.brand_encoding.training and not torch.is_grad_enabled():
            return self.brand_encoding.embed(brand_ids)
        return self.brand_encoding(brand_ids).permute((1, 0, 2)).mean(0)

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
