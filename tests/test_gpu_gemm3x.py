"""The in-kernel-split 3xTF32 GEMM (csrc/gemm3x.cu) through its public faces: `ops.linear` (K-major operands, bias /
scale / ReLU epilogue, K split for small outputs, ragged tiles) against float64 torch, and -- for the transposed
(MN-major) operand paths -- the TripletLoss gradients in tests/test_gpu_losses.py."""
import numpy as np
import pytest
import torch

from tests.gpu_util import dev

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
