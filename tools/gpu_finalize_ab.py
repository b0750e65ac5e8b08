"""A/B of the config-2 finalisation pass between two builds of the library (argv[1] = path of the alternative .so)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fancyrec_b200 import _lib
if len(sys.argv) > 1:
    _lib.LIB_PATH = os.path.join(ROOT, sys.argv[1])
from fancyrec_b200 import ops
dev = torch.device("cuda:0")
n, dv, dt = 1000000, 2048, 1024
g = torch.Generator(device=dev).manual_seed(1)
visual = torch.randn((n, dv), generator=g, device=dev)
text = torch.randn((n, dt), generator=g, device=dev)
def timeit(fn, reps=20):
    fn(); torch.cuda.synchronize()
    b, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.record()
    for _ in range(reps): fn()
    e.record(); torch.cuda.synchronize()
    return b.elapsed_time(e) / reps
gb = (n * (dv + dt) * 4 + n * 3072 * 2) / 1e9
t = timeit(lambda: ops.finalize_posts(visual, text, visual_norm=True, text_norm=True, final_norm=True))
print("%s: finalize config-2 rows %.3f ms  %.0f GB/s" % (os.path.basename(_lib.LIB_PATH), t, gb / t * 1e3))
