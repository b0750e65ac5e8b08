"""Warp-per-row finalisation kernels: bandwidth vs blocks per SM (probe build with FRX_FIN_GRID)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fancyrec_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "build", "libfrx_probe.so")
from fancyrec_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
def timeit(fn, reps=10):
    fn(); torch.cuda.synchronize()
    b, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return b.elapsed_time(e) / reps
grid = os.environ.get("FRX_FIN_GRID")
np5, f, d = 51200, 32, 2048
frames = torch.randn((np5 * f, d), generator=g, device=dev).abs_()
rp = (torch.arange(np5 + 1, device=dev) * f).to(torch.int64)
t = timeit(lambda: ops.finalize_posts(frames, row_ptr=rp, final_norm=True))
print("grid %s: config-5 rows (32 x 2048 pooled): %.3f ms %.0f GB/s" % (grid, t, (np5 * f * d * 4 + np5 * d * 2) / t / 1e6))
del frames
n = 1000000
x = torch.randn((n, 1024), generator=g, device=dev)
t = timeit(lambda: ops.finalize_posts(x, final_norm=True))
print("grid %s: 1024-d rows: %.3f ms %.0f GB/s" % (grid, t, n * 1024 * 6 / t / 1e6))
