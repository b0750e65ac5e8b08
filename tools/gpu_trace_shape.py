"""Pipeline trace (see gpu_trace_pipeline.py) of the first 256 tiles of CTA 0 at an arbitrary shape: nb,n,d,k."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
from fancyrec_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "build", "libfrx_trace.so")
from fancyrec_b200 import ops, ranking
lib = _lib.load()
lib.frx_debug_set_trace.argtypes = [ctypes.c_void_p]
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
nb, n, d, k = [int(v) for v in sys.argv[1].split(",")]
a = ranking.to_operand(torch.randn((nb, d), generator=g, device=dev))
b = torch.empty((n, ops.round_up(d, 64)), dtype=torch.bfloat16, device=dev)
for lo in range(0, n, 250000):
    b[lo:lo + 250000] = ranking.to_operand(torch.randn((min(250000, n - lo), d), generator=g, device=dev))
lab = (torch.randperm(n, generator=g, device=dev) % nb).to(torch.int32)
trace = torch.zeros(256 * 4, dtype=torch.int64, device=dev)
ws = None
for _ in range(2):
    r = ops.score_topk(a, b, k, d=d, labels=lab, workspace=ws); ws = r["workspace"]
lib.frx_debug_set_trace(trace.data_ptr())
ops.score_topk(a, b, k, d=d, labels=lab, workspace=ws)
torch.cuda.synchronize()
lib.frx_debug_set_trace(None)
t = trace.cpu().numpy().reshape(256, 4)
nt = int((t[:, 0] > 0).sum())
kb = (d + 63) // 64
print("shape %s: %d traced tiles of CTA 0's first item ; ideal MMA cycles per tile %d" % (sys.argv[1], nt, kb * 4 * 128))
for lo, hi in ((0, 8), (8, 32), (32, 96), (96, 200), (200, 255)):
    hi = min(hi, nt - 1)
    if hi <= lo: break
    per = (t[hi, 0] - t[lo, 0]) / float(hi - lo)
    print("tiles %3d..%3d: period %7.0f  epilogue %7.0f (max %7.0f)  mma issue span %7.0f  wait for accumulator stage %7.0f" % (
        lo, hi, per, (t[lo:hi, 3] - t[lo:hi, 2]).mean(), (t[lo:hi, 3] - t[lo:hi, 2]).max(),
        (t[lo:hi, 1] - t[lo:hi, 0]).mean(), (t[lo + 1:hi + 1, 0] - t[lo:hi, 1]).mean()))
