"""One MFC-sized Linear layer (512 x 3072 x 3072) on gemm3x, a few launches -- the target of an `ncu --set full` capture."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fancyrec_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((512, 3072), generator=g, device=dev)
w = torch.randn((3072, 3072), generator=g, device=dev) / 3072 ** 0.5
for _ in range(4):
    y = ops.linear(x, w)
torch.cuda.synchronize()
print(float(y.abs().max()))
