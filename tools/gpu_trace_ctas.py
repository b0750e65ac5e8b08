"""Per-CTA placement and progress of the fused score kernel (build/libfrx_trace.so): SM id, start time, time at tile 50 of
the first item, end time -- to see how the CTAs that share a post range (one split) are placed and how far they drift."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from fancyrec_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "build", "libfrx_trace.so")
from fancyrec_b200 import ops, ranking
lib = _lib.load()
lib.frx_debug_set_cta_trace.argtypes = [ctypes.c_void_p]
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
nb, n, d, k = 1000, 1000000, 3072, 100
a = ranking.to_operand(torch.randn((nb, d), generator=g, device=dev))
b = ranking.to_operand(torch.randn((n, d), generator=g, device=dev))
lab = (torch.randperm(n, generator=g, device=dev) % nb).to(torch.int32)
for pairs in (0, 1):
    lib.frx_set_cta_pairs(pairs)
    ws = None
    for _ in range(2):
        r = ops.score_topk(a, b, k, d=d, labels=lab, workspace=ws); ws = r["workspace"]
    tr = torch.zeros(148 * 4, dtype=torch.int64, device=dev)
    lib.frx_debug_set_cta_trace(tr.data_ptr())
    ops.score_topk(a, b, k, d=d, labels=lab, workspace=ws)
    torch.cuda.synchronize()
    lib.frx_debug_set_cta_trace(None)
    t = tr.cpu().numpy().reshape(148, 4)
    t0 = t[:, 1].min()
    print("pairs=%d  kernel span %.1f us ; start skew %.1f us ; end skew %.1f us" % (
        pairs, (t[:, 3].max() - t0) / 1e3, (t[:, 1].max() - t0) / 1e3, (t[:, 3].max() - t[:, 3].min()) / 1e3))
    print("  smid of blocks 0..31:", t[:32, 0].tolist())
    grp = 8
    sk = [(t[i:i + grp, 2].max() - t[i:i + grp, 2].min()) / 1e3 for i in range(0, 144, grp)]
    print("  skew at tile 50 inside each group of 8 consecutive blocks (one split), us: mean %.1f max %.1f" % (np.mean(sk), np.max(sk)))
    print("  time from start to tile 50 per block (us): min %.1f mean %.1f max %.1f" % (
        ((t[:, 2] - t[:, 1]) / 1e3).min(), ((t[:, 2] - t[:, 1]) / 1e3).mean(), ((t[:, 2] - t[:, 1]) / 1e3).max()))
    print("  smid parity of the 8 blocks of splits 0..3:", [[int(x) for x in t[i:i + 8, 0]] for i in range(0, 32, 8)])
