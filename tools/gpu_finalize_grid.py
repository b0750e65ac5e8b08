"""Finalisation bandwidth vs blocks per SM (probe build with FRX_FIN_GRID) for three row shapes."""
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
n = 1000000
x2048 = torch.randn((n, 2048), generator=g, device=dev)
t = timeit(lambda: ops.finalize_posts(x2048, final_norm=True))
print("grid %s: 2048-d rows -> bf16: %.3f ms %.0f GB/s" % (os.environ.get("FRX_FIN_GRID"), t, n * 2048 * 6 / t / 1e6))
del x2048
np_ = 200000
frames = torch.randn((np_ * 4, 2048), generator=g, device=dev)
text = torch.randn((np_, 1024), generator=g, device=dev)
rp = (torch.arange(np_ + 1, device=dev) * 4).to(torch.int64)
t = timeit(lambda: ops.finalize_posts(frames, text, row_ptr=rp, visual_norm=True, text_norm=True, final_norm=True))
gb = (np_ * 4 * 2048 * 4 + np_ * 1024 * 4 + np_ * 3072 * 2) / 1e9
print("grid %s: pooled 4 x 2048 + 1024 text -> 3072 bf16: %.3f ms %.0f GB/s" % (os.environ.get("FRX_FIN_GRID"), t, gb / t * 1e3))
