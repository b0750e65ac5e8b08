"""The in-kernel-split 3xTF32 GEMM (csrc/gemm3x.cu) through its public faces: `ops.linear` (K-major operands, bias /
scale / ReLU epilogue, K split for small outputs, ragged tiles) against float64 torch, and -- for the transposed
(MN-major) operand paths -- `ops.matmul3x` in all four layout combinations, whole and ragged 32-row groups (and the
TripletLoss gradients in tests/test_gpu_losses.py)."""
import numpy as np
import pytest
import torch

from tests.gpu_util import dev, to_dev

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("m,n,k", [(128, 128, 32), (512, 512, 3072), (37, 200, 100), (300, 130, 2048), (1024, 3072, 1024),
                                   (5, 7, 8), (129, 257, 36)])
@pytest.mark.parametrize("epilogue", ["plain", "bias_relu", "scale_bias"])
def test_linear_matches_float64(m, n, k, epilogue):
    from fancyrec_b200 import ops
    g = torch.Generator(device=dev()).manual_seed(m * 31 + n * 7 + k)
    x = torch.randn((m, k), generator=g, device=dev())
    w = torch.randn((n, k), generator=g, device=dev()) / k ** 0.5
    bias = torch.randn(n, generator=g, device=dev()) if epilogue != "plain" else None
    scale = (torch.rand(n, generator=g, device=dev()) + 0.5) if epilogue == "scale_bias" else None
    got = ops.linear(x, w, bias=bias, col_scale=scale, relu=(epilogue == "bias_relu"))
    want = x.double() @ w.double().t()
    if scale is not None:
        want = want * scale.double()
    if bias is not None:
        want = want + bias.double()
    if epilogue == "bias_relu":
        want = want.clamp(min=0)
    # fp32-grade: 3xTF32 products (2^-22 relative each) + fp32 accumulation over k terms
    err = (got.double() - want).abs().max().item()
    ref = (x.double().abs() @ w.double().abs().t()).max().item()
    assert err <= 4e-6 * ref + 1e-6, (err, ref)      # observed <= 2.3e-6


def test_linear_strided_rows_and_reuse():
    """Row pitches larger than the row (views into wider buffers), repeated launches give identical bits."""
    from fancyrec_b200 import ops
    g = torch.Generator(device=dev()).manual_seed(9)
    xb = torch.randn((200, 640), generator=g, device=dev())
    wb = torch.randn((96, 1024), generator=g, device=dev())
    x, w = xb[:, :512], wb[:, 256:768]
    out = torch.zeros((200, 128), device=dev())
    a = ops.linear(x, w, out=out[:, :96]).clone()
    b = ops.linear(x, w, out=out[:, :96]).clone()
    assert torch.equal(a, b) and bool((out[:, 96:] == 0).all())
    want = x.double() @ w.double().t()
    assert (a.double() - want).abs().max().item() <= 2e-6 * (x.double().abs() @ w.double().abs().t()).max().item()


@pytest.mark.parametrize("m,n,k", [(512, 3072, 512), (128, 128, 32), (96, 160, 72), (100, 36, 52), (260, 300, 516), (512, 5120, 3072),
                                   (512, 3072, 5120)])
@pytest.mark.parametrize("at,bt", [(False, False), (False, True), (True, False), (True, True)])
def test_matmul3x_operand_layouts(m, n, k, at, bt):
    """Every operand layout of the gradient products (loss.py:87-143): K-major, transposed with whole 32-row groups (fed to the
    tensor core as MN-major tiles) and transposed with ragged row counts (transposed in shared memory) -- 4e-6 of scale.
    The last two shapes (ContrastiveLoss's queue products: 160 and 96 tiles on 148 SMs) and the first run on the stream
    mapping (equal unit ranges per SM, shared tiles fixed up), the small ones on one CTA per tile."""
    from fancyrec_b200 import ops
    rs = np.random.RandomState(m + 3 * n + 7 * k + 2 * at + bt)
    a = rs.standard_normal((m, k)).astype(np.float32)
    b = rs.standard_normal((n, k)).astype(np.float32)
    want = a.astype(np.float64) @ b.astype(np.float64).T * 0.5
    ad = to_dev(np.ascontiguousarray(a.T) if at else a)
    bd = to_dev(np.ascontiguousarray(b.T) if bt else b)
    got = ops.matmul3x(ad, bd, a_transposed=at, b_transposed=bt, alpha=0.5).cpu().numpy()
    ref = (np.abs(a).astype(np.float64) @ np.abs(b).astype(np.float64).T).max() * 0.5
    err = np.abs(got - want).max()
    assert err <= 4e-6 * ref + 1e-6, (err, ref)
