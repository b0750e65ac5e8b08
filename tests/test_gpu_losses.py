"""GPU parity: TripletLoss / ContrastiveLoss forward + backward against the reference's autograd
(tests/golden/losses.npz) and the fp64 oracle.  fp32 kernels: rtol 2e-4 on loss, 2e-3 on gradients
(sum over B^2 hinge terms / softmax over B + Q logits, different summation order)."""
import os

import numpy as np
import pytest
import torch

from oracle import losses as oloss
from oracle import synth
from tests.gpu_util import dev, to_dev

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("style", ["sum", "mean"])
def test_triplet_golden(golden_dir, style):
    from fancyrec_b200 import loss as floss
    g = np.load(os.path.join(golden_dir, "losses.npz"))
    ids, brand, post = synth.loss_inputs()
    for mv in (False, True):
        bt = to_dev(brand).requires_grad_()
        pt = to_dev(post).requires_grad_()
        crit = floss.TripletLoss(margin=0.2, max_violation=mv, cost_style=style)
        val = crit(torch.from_numpy(ids), bt, pt)
        (val * 1.0).backward()
        key = "triplet_%s_mv%d" % (style, int(mv))
        np.testing.assert_allclose(val.item(), g[key + "_loss"], rtol=2e-5)
        np.testing.assert_allclose(bt.grad.cpu().numpy(), g[key + "_dbrand"], rtol=1e-4, atol=1e-6)
        np.testing.assert_allclose(pt.grad.cpu().numpy(), g[key + "_dpost"], rtol=1e-4, atol=1e-6)
    for direction in ("p2b", "b2p"):
        with pytest.raises(TypeError):
            floss.TripletLoss(margin=0.2, direction=direction)(torch.from_numpy(ids), to_dev(brand), to_dev(post))


@pytest.mark.parametrize("b,d", [(512, 1024), (512, 3072), (100, 96), (257, 300)])
def test_triplet_vs_oracle_config3(b, d):
    """Values and decisions are tested SEPARATELY.  (1) the tile: |S_device - S_fp64| <= 1e-5 of its scale (3xTF32);
    (2) everything downstream of the tile -- integer rank weights, hinge active set, dS, both gradient GEMMs, the loss --
    against the fp64 oracle evaluated on the device's OWN tile: 1e-5 of the gradient scale; (3) the decisions taken on the
    device tile differ from those taken on the fp64 tile only where a hinge argument is within rounding of zero
    (a handful of the B^2 terms), and the end-to-end loss agrees to 2e-5."""
    from fancyrec_b200 import ops
    rs = np.random.RandomState(b + d)
    ids = rs.randint(0, 51, b).astype(np.int64)
    brand = (rs.standard_normal((b, d)) * 0.05).astype(np.float32)
    post = (rs.standard_normal((b, d)) * 0.05).astype(np.float32) + brand * 0.5
    aligned = d % 4 == 0
    s_dev = ops.linear(to_dev(post), to_dev(brand)).cpu().numpy() if aligned else None      # the tile the loss kernel sees
    for style in (0, 1):
        loss, db, dp = ops.triplet_fwd_bwd(to_dev(ids), to_dev(brand), to_dev(post), 0.2, style)
        name = 'mean' if style else 'sum'
        wl, wdb, wdp, aux = oloss.triplet_loss(ids, brand, post, 0.2, name)
        np.testing.assert_allclose(loss.item(), wl, rtol=2e-5)
        if s_dev is not None:
            assert np.abs(s_dev - aux["s"]).max() <= 1e-5 * np.abs(aux["s"]).max()
            ol, odb, odp, oaux = oloss.triplet_loss(ids, brand, post, 0.2, name, s_override=s_dev)
            np.testing.assert_allclose(loss.item(), ol, rtol=1e-5)
            np.testing.assert_allclose(db.cpu().numpy(), odb, rtol=0, atol=1e-5 * np.abs(odb).max())
            np.testing.assert_allclose(dp.cpu().numpy(), odp, rtol=0, atol=1e-5 * np.abs(odp).max())
            flips = int((oaux["active_p"] != aux["active_p"]).sum() + (oaux["active_b"] != aux["active_b"]).sum())
            assert flips <= max(4, b * b // 20000), flips
            assert np.array_equal(oaux["rank_p"], aux["rank_p"]) or np.abs(oaux["rank_p"] - aux["rank_p"]).max() < 0.5
        else:                                         # odd width: fp32 FMA fallback, summation-order tolerance
            np.testing.assert_allclose(db.cpu().numpy(), wdb, rtol=2e-3, atol=2e-3 * np.abs(wdb).max())
            np.testing.assert_allclose(dp.cpu().numpy(), wdp, rtol=2e-3, atol=2e-3 * np.abs(wdp).max())
        # forward-only call (no_grad path): same loss, no gradient work
        loss_f, none_b, none_p = ops.triplet_fwd_bwd(to_dev(ids), to_dev(brand), to_dev(post), 0.2, style, want_grad=False)
        assert none_b is None and none_p is None and loss_f.item() == loss.item()


def test_triplet_upstream_gradient_scales():
    from fancyrec_b200 import loss as floss
    ids, brand, post = synth.loss_inputs()
    bt = to_dev(brand).requires_grad_()
    val = floss.TripletLoss(margin=0.2)(torch.from_numpy(ids), bt, to_dev(post))
    (val * 3.0).backward()
    g3 = bt.grad.clone()
    bt.grad = None
    floss.TripletLoss(margin=0.2)(torch.from_numpy(ids), bt, to_dev(post)).backward()
    torch.testing.assert_close(g3, bt.grad * 3.0)


@pytest.mark.parametrize("style", ["sum", "mean"])
@pytest.mark.parametrize("mode", ["queue", "no_queue", "no_intra"])
def test_contrastive_golden(golden_dir, style, mode):
    import types
    from fancyrec_b200 import loss_ctrs as fctrs
    g = np.load(os.path.join(golden_dir, "losses.npz"))
    b, d = 24, 32
    opt = types.SimpleNamespace(cost_style=style, queue_size=2 * b, common_embedding_size=d,
                                no_queue=(mode == "no_queue"), no_intra=(mode == "no_intra"))
    crit = fctrs.ContrastiveLoss(opt).to(dev())
    for step in range(3):
        _, brand, post = synth.loss_inputs(seed=600 + step)
        bt = to_dev(brand).requires_grad_()
        pt = to_dev(post).requires_grad_()
        val = crit(bt, pt)
        val.backward()
        key = "ctr_%s_%s_s%d" % (style, mode, step)
        np.testing.assert_allclose(val.item(), g[key + "_loss"], rtol=1e-4)
        np.testing.assert_allclose(bt.grad.cpu().numpy(), g[key + "_dbrand"], rtol=5e-3, atol=5e-5)
        np.testing.assert_allclose(pt.grad.cpu().numpy(), g[key + "_dpost"], rtol=5e-3, atol=5e-5)
        np.testing.assert_allclose(crit.queue.cpu().numpy(), g[key + "_queue"], rtol=1e-5, atol=1e-7)
        assert int(crit.queue_ptr[0]) == int(g[key + "_ptr"][0])


def test_contrastive_bad_queue_size_errors_like_reference(golden_dir):
    import types
    from fancyrec_b200 import loss_ctrs as fctrs
    g = np.load(os.path.join(golden_dir, "losses.npz"))
    want = [str(x) for x in g["ctr_bad_queue_errors"]]
    _, brand, post = synth.loss_inputs()
    b, d = brand.shape
    opt = types.SimpleNamespace(cost_style="sum", queue_size=2 * b + 4, common_embedding_size=d, no_queue=False,
                                no_intra=False)
    crit = fctrs.ContrastiveLoss(opt).to(dev())
    got = []
    for step in range(3):
        try:
            crit(to_dev(brand), to_dev(post))
            got.append("ok")
        except (IndexError, RuntimeError) as ex:
            got.append(type(ex).__name__)
    assert got == want


def test_contrastive_vs_oracle_config3():
    from fancyrec_b200 import ops
    rs = np.random.RandomState(5)
    b, d, q = 512, 1024, 5120
    brand = rs.standard_normal((b, d)).astype(np.float32)
    post = (rs.standard_normal((b, d)) + brand).astype(np.float32)
    queue = rs.standard_normal((q, d)).astype(np.float32)
    queue /= np.linalg.norm(queue, axis=1, keepdims=True)
    wl, wdb, wdp, nq, nptr = oloss.contrastive_loss(brand, post, queue=queue, queue_ptr=1024, cost_style='mean')
    loss, db, dp = ops.contrastive_fwd_bwd(to_dev(brand), to_dev(post), to_dev(nq.astype(np.float32)), nptr, False,
                                           0.03, 0.8, 1)
    np.testing.assert_allclose(loss.item(), wl, rtol=2e-4)
    np.testing.assert_allclose(db.cpu().numpy(), wdb, rtol=5e-3, atol=5e-3 * np.abs(wdb).max())
    np.testing.assert_allclose(dp.cpu().numpy(), wdp, rtol=5e-3, atol=5e-3 * np.abs(wdp).max())


def test_crossclr_lab_golden(golden_dir):
    from fancyrec_b200 import loss as floss
    from fancyrec_b200 import loss_ctrs as fctrs
    g = np.load(os.path.join(golden_dir, "losses.npz"))
    _, brand, post = synth.loss_inputs()
    for style in ("sum", "mean"):
        bt = to_dev(brand).requires_grad_()
        pt = to_dev(post).requires_grad_()
        val = fctrs.CrossCLR_onlyIntraModality(cost_style=style).to(dev())(bt, pt)
        val.backward()
        np.testing.assert_allclose(val.item(), g["crossclr_%s_loss" % style], rtol=1e-4)
        np.testing.assert_allclose(bt.grad.cpu().numpy(), g["crossclr_%s_dbrand" % style], rtol=5e-3, atol=5e-5)
        np.testing.assert_allclose(pt.grad.cpu().numpy(), g["crossclr_%s_dpost" % style], rtol=5e-3, atol=5e-5)
    bt = to_dev(brand).requires_grad_()
    val = floss.LabLoss()(bt)
    val.backward()
    np.testing.assert_allclose(val.item(), g["lab_loss"], rtol=1e-5)
    np.testing.assert_allclose(bt.grad.cpu().numpy(), g["lab_dbrand"], rtol=1e-3, atol=1e-6)


def _torch_crossclr(brand, post, t=0.03, w=0.8, style='sum'):
    """Plain torch fp32 restatement of loss_ctrs.py:52-117 (rank weights by counting, ties to the smaller index)."""
    import torch.nn.functional as F
    b = brand.shape[0]
    with torch.no_grad():
        sc = post @ brand.t()
        idx = torch.arange(b, device=sc.device)
        dg = sc.diag()
        lower = idx[None, :] < idx[:, None]
        pos_r = (sc > dg[:, None]).sum(1) + ((sc == dg[:, None]) & lower).sum(1)
        pos_c = (sc > dg[None, :]).sum(0) + ((sc == dg[None, :]) & lower.t()).sum(0)
        rank_p = 1 / (b - (pos_r + 1).float() + 1) + 1
        rank_b = 1 / (b - (pos_c + 1).float() + 1) + 1
    bn, pn = F.normalize(brand, dim=1), F.normalize(post, dim=1)
    off = 1 - torch.eye(b, device=brand.device)
    bl = torch.cat([bn @ pn.t() / t, w * (bn @ bn.t() / t) * off], 1)
    pl = torch.cat([pn @ bn.t() / t, w * (pn @ pn.t() / t) * off], 1)
    lb = rank_b * -torch.log(F.softmax(bl, 1).diagonal())
    lp = rank_p * -torch.log(F.softmax(pl, 1).diagonal())
    return (lb.sum() + lp.sum()) / 2 if style == 'sum' else (lb.mean() + lp.mean()) / 2


@pytest.mark.parametrize("b,d,style", [(512, 1024, 'sum'), (512, 3072, 'mean'), (96, 200, 'sum'), (33, 50, 'mean')])
def test_crossclr_config3_vs_oracle_and_torch(b, d, style):
    """Fused CrossCLR at the config-3 batch: value vs the fp64 oracle, gradients vs torch fp32 autograd of the
    same formula (tolerances as for the other loss tiles: value rtol 2e-4, gradients 2e-3 of the largest entry)."""
    from fancyrec_b200 import loss_ctrs as fctrs
    rs = np.random.RandomState(b + d)
    brand = rs.standard_normal((b, d)).astype(np.float32)
    post = (rs.standard_normal((b, d)) + 0.1 * brand).astype(np.float32)     # weakly aligned: the loss stays O(1)
    bt, pt = to_dev(brand).requires_grad_(), to_dev(post).requires_grad_()
    val = fctrs.CrossCLR_onlyIntraModality(cost_style=style).to(dev())(bt, pt)
    val.backward()
    np.testing.assert_allclose(val.item(), oloss.crossclr_loss(brand, post, cost_style=style), rtol=2e-4)
    br, pr = to_dev(brand).requires_grad_(), to_dev(post).requires_grad_()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref = _torch_crossclr(br, pr, style=style)
        ref.backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    np.testing.assert_allclose(val.item(), ref.item(), rtol=2e-4)
    for got, want in ((bt.grad, br.grad), (pt.grad, pr.grad)):
        want = want.cpu().numpy()
        np.testing.assert_allclose(got.cpu().numpy(), want, rtol=2e-3, atol=2e-3 * np.abs(want).max())


@pytest.mark.parametrize("b,d", [(512, 1024), (52, 2048), (33, 50)])
def test_lab_loss_vs_oracle_and_torch(b, d):
    from fancyrec_b200 import loss as floss
    rs = np.random.RandomState(b * 3 + d)
    brand = rs.standard_normal((b, d)).astype(np.float32)
    bt = to_dev(brand).requires_grad_()
    val = floss.LabLoss()(bt)
    val.backward()
    np.testing.assert_allclose(val.item(), oloss.lab_loss(brand), rtol=1e-5)
    br = to_dev(brand).requires_grad_()
    n = br / br.pow(2).sum(1, keepdim=True).sqrt()
    prev = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        sref = (n @ n.t()).masked_fill(torch.eye(b, device=br.device) > .5, 0)
        ref = (torch.exp(sref).sum() - b) / b
        ref.backward()
    finally:
        torch.backends.cuda.matmul.allow_tf32 = prev
    want = br.grad.cpu().numpy()
    np.testing.assert_allclose(bt.grad.cpu().numpy(), want, rtol=2e-3, atol=2e-3 * np.abs(want).max())


@pytest.mark.parametrize("b,d,style", [(512, 3072, "sum"), (512, 1024, "mean"), (100, 96, "sum"), (33, 64, "mean")])
def test_hardest_negative_hinge_opt_in(b, d, style):
    """The opt-in VSE++ mode (`TripletLoss(hardest_negative=True)`, NOT the reference's forward): the selected negatives and
    dS against the fp64 restatement evaluated on the device's own tile, the loss to 2e-5, and a 5-line torch autograd
    restatement as a second witness."""
    from fancyrec_b200 import loss as floss
    from fancyrec_b200 import ops
    rs = np.random.RandomState(7 * b + d)
    ids = rs.randint(0, 23, b).astype(np.int64)
    brand = (rs.standard_normal((b, d)) * 0.05).astype(np.float32)
    post = (rs.standard_normal((b, d)) * 0.05).astype(np.float32) + brand * 0.5
    crit = floss.TripletLoss(margin=0.2, cost_style=style, hardest_negative=True)
    bt, pt = to_dev(brand).requires_grad_(), to_dev(post).requires_grad_()
    val = crit(torch.from_numpy(ids), bt, pt)
    val.backward()
    # fp64 restatement on the tile the device used (aligned shapes: the same gemm3x product)
    s_dev = ops.linear(to_dev(post), to_dev(brand)).cpu().numpy() if (d % 4 == 0 and b % 4 == 0 and b >= 32) else None
    wl, wdb, wdp, aux = oloss.vsepp_loss(ids, brand, post, 0.2, style, s_override=s_dev)
    np.testing.assert_allclose(val.item(), wl, rtol=2e-5)
    if s_dev is not None:
        np.testing.assert_allclose(bt.grad.cpu().numpy(), wdb, rtol=0, atol=1e-5 * np.abs(wdb).max())
        np.testing.assert_allclose(pt.grad.cpu().numpy(), wdp, rtol=0, atol=1e-5 * np.abs(wdp).max())
    # torch autograd witness (CPU, fp64)
    tb = torch.from_numpy(brand).double().requires_grad_()
    tp = torch.from_numpy(post).double().requires_grad_()
    s = tp @ tb.t()
    dg = s.diag()
    same = torch.from_numpy(ids[:, None] == ids[None, :])
    cost_p = (0.2 + s - dg[:, None]).clamp(min=0).masked_fill(same, 0)
    cost_b = (0.2 + s - dg[None, :]).clamp(min=0).masked_fill(same, 0)
    ref = cost_p.max(1)[0].sum() + cost_b.max(0)[0].sum()
    if style == "mean":
        ref = ref / b
    ref.backward()
    np.testing.assert_allclose(val.item(), ref.item(), rtol=1e-4)
    np.testing.assert_allclose(bt.grad.cpu().numpy(), tb.grad.numpy(), rtol=0, atol=2e-3 * tb.grad.abs().max().item())
    # the reference's own forward is untouched by the new keyword's default
    plain = floss.TripletLoss(margin=0.2, cost_style=style, max_violation=True)
    assert plain.hardest_negative is False
