"""A4, training path: the fused dropout + mean-over-aspects + L1Penalty brand embedding (frx_brand_train_fwd / _bwd,
BrandAspects.embed_train) against the reference formula (model.py:419-428, :594, L1Penalty :389-402) evaluated by torch
autograd under the SAME keep mask (frx_brand_dropout_mask writes the hash-generated mask out), plus the statistics of
the mask and the memory claim (no [B, A, D] tensor)."""
import types

import numpy as np
import pytest
import torch

from tests.gpu_util import dev

pytestmark = pytest.mark.gpu


def _reference(w_rows, aspects, mask):
    """forward(brand_list).permute(1, 0, 2).mean(0) with nn.Dropout(0.5) replaced by the given keep mask."""
    from fancyrec_b200.model import L1Penalty
    w = L1Penalty.apply(w_rows)
    prod = w.unsqueeze(2) * aspects.unsqueeze(0)                    # [B, A, D]  (model.py:422-426)
    return (prod * mask * 2.0).permute(1, 0, 2).mean(0)             # dropout keep -> x 1/(1-p); model.py:594


@pytest.mark.parametrize("b,a,d", [(5, 7, 40), (17, 33, 1100), (24, 50, 2048), (64, 200, 96)])
def test_fused_training_embedding_matches_reference_under_the_same_mask(b, a, d):
    from fancyrec_b200 import model, ops
    torch.manual_seed(b * 100 + a)
    opt = types.SimpleNamespace(brand_num=b + 3, common_embedding_size=d, brand_aspect=a)
    enc = model.BrandAspects(opt).to(dev()).train()
    ids = torch.randint(0, b + 4, (b,), device=dev())
    ids[1] = ids[0]                                                  # a repeated brand: gradients must accumulate
    seed = 123456789 + d
    out = enc.embed_train(ids, seed=seed)
    g = torch.randn_like(out)
    out.backward(g)
    got_dw, got_de = enc.brand_embeddings.weight.grad.clone(), enc.aspects_embeddings.grad.clone()
    enc.zero_grad()
    mask = ops.brand_dropout_mask(b, a, d, seed, dev()).float()
    ref = _reference(enc.brand_embeddings(ids), enc.aspects_embeddings, mask)
    ref.backward(g)
    scale = ref.abs().max().item()
    assert torch.allclose(out, ref, rtol=0, atol=2e-6 * max(scale, 1.0))
    assert torch.allclose(got_dw, enc.brand_embeddings.weight.grad, rtol=1e-5, atol=2e-6)
    assert torch.allclose(got_de, enc.aspects_embeddings.grad, rtol=1e-5, atol=2e-6)
    # the L1 term is there: rows of the table that the batch touched carry +-1e-4 even where g would cancel
    touched = torch.unique(ids)
    assert bool((enc.brand_embeddings.weight.grad[touched].abs() > 0).any())


def test_mask_is_fair_and_seeded():
    from fancyrec_b200 import ops
    m1 = ops.brand_dropout_mask(16, 64, 2048, 42, dev()).float()
    m2 = ops.brand_dropout_mask(16, 64, 2048, 42, dev()).float()
    m3 = ops.brand_dropout_mask(16, 64, 2048, 43, dev()).float()
    assert torch.equal(m1, m2) and not torch.equal(m1, m3)
    assert abs(m1.mean().item() - 0.5) < 2e-3                        # 2 M bits: sigma = 3.5e-4
    assert abs((m1 * m3).mean().item() - 0.25) < 3e-3                # different seeds: independent
    for dim in (0, 1, 2):                                            # no biased row, aspect or column: within 5.5 sigma
        others = tuple(x for x in (0, 1, 2) if x != dim)
        mean = m1.mean(dim=others)
        sigma = 0.5 / (m1.numel() / m1.shape[dim]) ** 0.5
        assert 0.5 - 5.5 * sigma < mean.min().item() and mean.max().item() < 0.5 + 5.5 * sigma


def test_training_embed_brand_goes_through_the_fused_path_and_stays_small():
    """FancyRec.embed_brand in training mode at the config-3 size (B = 512, A = 2000, D = 3072): the reference
    materialises [B, A, D] = 12.6 GB (+ the dropout mask and the autograd copies); the fused path stays below 1 GB."""
    from fancyrec_b200 import model
    b, a, d = 512, 2000, 3072
    opt = types.SimpleNamespace(brand_num=51, common_embedding_size=d, brand_aspect=a, single_modal_text=False,
                                single_modal_visual=False)
    mdl = model.FancyRec(opt).to(dev()).train()
    ids = torch.randint(0, 52, (b,), device=dev())
    torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    out = mdl.embed_brand(ids)
    out.square().mean().backward()
    torch.cuda.synchronize()
    peak = torch.cuda.max_memory_allocated() - base
    assert out.shape == (b, d) and peak < 1 << 30, peak
    # expectation over the mask = the eval-mode embedding: the batch mean of (train - eval) is ~0 relative to the spread
    mdl.eval()
    with torch.no_grad():
        ev = mdl.embed_brand(ids)
    rel = ((out.detach() - ev).mean() / ev.std()).abs().item()
    assert rel < 0.02, rel
    g_w = mdl.brand_encoding.brand_embeddings.weight.grad
    assert g_w is not None and mdl.brand_encoding.aspects_embeddings.grad is not None and bool(torch.isfinite(g_w).all())
