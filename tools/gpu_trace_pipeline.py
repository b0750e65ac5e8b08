"""Per-tile clock64 timeline of the score kernel's MMA / epilogue pipeline on CTA 0 (needs build/libfrx_trace.so,
built with -DFRX_TRACE).  Columns per tile: MMA start (accumulator stage free), MMA last commit issued, epilogue start
(accumulator ready), epilogue end."""
import os, sys, ctypes
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from fancyrec_b200 import _lib
_lib.LIB_PATH = os.path.join(ROOT, "build", "libfrx_trace.so")
from fancyrec_b200 import ops, ranking
lib = _lib.load()
lib.frx_debug_set_trace.argtypes = [ctypes.c_void_p]
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
nb, n = 1000, 1000000
for d in (1024, 3072):
    a = ranking.to_operand(torch.randn((nb, d), generator=g, device=dev))
    b = ranking.to_operand(torch.randn((n, d), generator=g, device=dev))
    lab = (torch.randperm(n, generator=g, device=dev) % nb).to(torch.int32)
    trace = torch.zeros(256 * 4, dtype=torch.int64, device=dev)
    for _ in range(2): ops.score_topk(a, b, 100, d=d, labels=lab)
    lib.frx_debug_set_trace(trace.data_ptr())
    ops.score_topk(a, b, 100, d=d, labels=lab)
    torch.cuda.synchronize()
    lib.frx_debug_set_trace(None)
    t = trace.cpu().numpy().reshape(256, 4)
    t = t[:100]
    base = t[0, 0]
    print("D=%d  (cycles relative to tile 0 MMA start; per tile: mma_start mma_issued epi_start epi_end | mma_dur epi_dur gap_to_next_mma)" % d)
    for i in range(40, 52):
        r = t[i] - base
        print("tile %3d: %9d %9d %9d %9d | issue %6d  epi %6d  period %6d  epi_start-mma_issued %6d" % (
            i, r[0], r[1], r[2], r[3], r[1] - r[0], r[3] - r[2], t[i + 1, 0] - t[i, 0], r[2] - r[1]))
    per = (t[90, 0] - t[30, 0]) / 60.0
    print("mean period tiles 30..90: %.0f cycles ; mean epilogue %.0f ; mean mma issue span %.0f" % (
        per, (t[30:90, 3] - t[30:90, 2]).mean(), (t[30:90, 1] - t[30:90, 0]).mean()))
