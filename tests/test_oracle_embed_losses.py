"""Pins oracle/embed.py and oracle/losses.py to reference outputs (tests/golden/*.npz)."""
import os

import numpy as np
import pytest

from oracle import embed, losses, synth

RTOL = 2e-6   # fp32 accumulation-order tolerance (oracle accumulates in fp64, torch in fp32)


def test_finalize_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "finalize.npz"))
    frames, row_ptr = synth.frames_csr(707, 40, 96, 1, 9)
    text = np.abs(synth.gaussian(708, 40, 24, 0.3))
    pooled = embed.mean_pool_csr(frames, row_ptr)
    np.testing.assert_allclose(pooled, g["pooled"], rtol=RTOL, atol=1e-7)
    idx = np.arange(frames.shape[0])
    np.testing.assert_array_equal(embed.mean_pool_gather(frames, idx, row_ptr), pooled)
    np.testing.assert_allclose(embed.finalize_posts(pooled, text, True, True, False), g["cat_branchnorm"],
                               rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(embed.finalize_posts(pooled, text, True, True, True), g["final_branchnorm"],
                               rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(embed.finalize_posts(pooled, text, False, False, True), g["final_raw"],
                               rtol=RTOL, atol=1e-7)
    w = synth.gaussian(709, 9, 30)
    e = synth.gaussian(710, 30, 20)
    np.testing.assert_allclose(embed.brand_embed(w, e, g["brand_ids"]), g["brand_emb"], rtol=1e-5, atol=1e-6)


def test_zero_row_gives_nan_like_reference():
    out = embed.finalize_posts(np.zeros((1, 4), np.float32), None, True, True, True)
    assert np.isnan(out).all()


@pytest.mark.parametrize("style", ["sum", "mean"])
def test_triplet_golden(golden_dir, style):
    g = np.load(os.path.join(golden_dir, "losses.npz"))
    ids, brand, post = synth.loss_inputs()
    loss, db, dp, aux = losses.triplet_loss(ids, brand, post, margin=0.2, cost_style=style)
    for mv in (0, 1):   # max_violation is a no-op in the reference (loss.py:85 vs :87-143)
        key = "triplet_%s_mv%d" % (style, mv)
        np.testing.assert_allclose(loss, g[key + "_loss"], rtol=1e-5)
        np.testing.assert_allclose(db, g[key + "_dbrand"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(dp, g[key + "_dpost"], rtol=1e-4, atol=1e-6)
    assert str(g["triplet_dir_p2b"]) == "TypeError" and str(g["triplet_dir_b2p"]) == "TypeError"


@pytest.mark.parametrize("style", ["sum", "mean"])
@pytest.mark.parametrize("mode", ["queue", "no_queue", "no_intra"])
def test_contrastive_golden(golden_dir, style, mode):
    g = np.load(os.path.join(golden_dir, "losses.npz"))
    b, d = 24, 32
    queue, ptr = np.zeros((2 * b, d)), 0
    for step in range(3):
        _, brand, post = synth.loss_inputs(seed=600 + step)
        loss, db, dp, queue, ptr = losses.contrastive_loss(
            brand, post, queue=queue, queue_ptr=ptr, cost_style=style,
            no_queue=(mode == "no_queue"), no_intra=(mode == "no_intra"))
        key = "ctr_%s_%s_s%d" % (style, mode, step)
        np.testing.assert_allclose(loss, g[key + "_loss"], rtol=2e-5)
        np.testing.assert_allclose(db, g[key + "_dbrand"], rtol=2e-3, atol=2e-5)
        np.testing.assert_allclose(dp, g[key + "_dpost"], rtol=2e-3, atol=2e-5)
        np.testing.assert_allclose(queue, g[key + "_queue"], rtol=1e-6, atol=1e-7)
        assert ptr == int(g[key + "_ptr"][0])


def test_contrastive_bad_queue_size_raises(golden_dir):
    g = np.load(os.path.join(golden_dir, "losses.npz"))
    want = [str(x) for x in g["ctr_bad_queue_errors"]]
    _, brand, post = synth.loss_inputs()
    b, d = brand.shape
    queue, ptr = np.zeros((2 * b + 4, d)), 0
    got = []
    for step in range(3):
        try:
            _, _, _, queue, ptr = losses.contrastive_loss(brand, post, queue=queue, queue_ptr=ptr)
            got.append("ok")
        except (IndexError, RuntimeError) as ex:
            got.append(type(ex).__name__)
            queue, ptr = getattr(ex, "state", (queue, ptr))   # the reference mutates before raising
    assert got == want, (got, want)


def test_crossclr_lab_golden(golden_dir):
    g = np.load(os.path.join(golden_dir, "losses.npz"))
    _, brand, post = synth.loss_inputs()
    for style in ("sum", "mean"):
        np.testing.assert_allclose(losses.crossclr_loss(brand, post, cost_style=style),
                                   g["crossclr_%s_loss" % style], rtol=2e-5)
    np.testing.assert_allclose(losses.lab_loss(brand), g["lab_loss"], rtol=1e-5)


def test_vsepp_restatement_matches_torch_autograd():
    """oracle.losses.vsepp_loss (the opt-in hardest-negative hinge; not a reference function) against torch autograd of
    `cost_p.max(1) + cost_b.max(0)` on the reference's tile and same-brand mask, fp64."""
    import torch
    from oracle import losses as oloss
    rs = np.random.RandomState(3)
    b, d = 48, 24
    ids = rs.randint(0, 9, b)
    brand = rs.standard_normal((b, d)) * 0.3
    post = rs.standard_normal((b, d)) * 0.3 + brand * 0.5
    for style in ("sum", "mean"):
        wl, wdb, wdp, _ = oloss.vsepp_loss(ids, brand, post, 0.2, style)
        tb = torch.from_numpy(brand).requires_grad_()
        tp = torch.from_numpy(post).requires_grad_()
        s = tp @ tb.t()
        dg = s.diag()
        same = torch.from_numpy(ids[:, None] == ids[None, :])
        cp = (0.2 + s - dg[:, None]).clamp(min=0).masked_fill(same, 0)
        cb = (0.2 + s - dg[None, :]).clamp(min=0).masked_fill(same, 0)
        ref = cp.max(1)[0].sum() + cb.max(0)[0].sum()
        if style == "mean":
            ref = ref / b
        ref.backward()
        np.testing.assert_allclose(wl, ref.item(), rtol=1e-12)
        np.testing.assert_allclose(wdb, tb.grad.numpy(), atol=1e-12)
        np.testing.assert_allclose(wdp, tp.grad.numpy(), atol=1e-12)
