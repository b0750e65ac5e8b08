"""Device time of one MFC-sized Linear layer (512 x 3072 x 3072) on gemm3x, events around 20 launches, best of 5.
Used for DESIGN.md 4.5's breakdown with a scratch build whose kernel honoured FRX_G3_DBG (bit 0: converters skip the split,
bit 1: the issuer skips the MMAs -- results are then garbage, only the time means something); the shipped kernel has no such
knob, FRX_G3_STREAM=0 (one CTA per tile instead of the stream mapping) is the only switch left."""
import os, sys, torch
sys.path.insert(0, "/root/repo")
from fancyrec_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((512, 3072), generator=g, device=dev)
w = torch.randn((3072, 3072), generator=g, device=dev) / 3072 ** 0.5
out = torch.empty((512, 3072), device=dev)
for _ in range(5): ops.linear(x, w, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
best = 1e9
for rep in range(5):
    e0.record()
    for _ in range(20): ops.linear(x, w, out=out)
    e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 20 * 1e3)
print("FRX_G3_DBG=%s FRX_G3_STREAM=%s linear 512x3072x3072: %.1f us" % (os.environ.get("FRX_G3_DBG"), os.environ.get("FRX_G3_STREAM"), best))
