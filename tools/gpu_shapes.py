"""Fused score + top-k kernel at the per-GPU shapes of configs 4 and 5 (run on the GPU box).

The per-row behaviour of the epilogue (pass rate of the seeded threshold, candidate appends) depends on
posts-per-shard and k, not on the number of brand rows, so each config is timed with its true posts-per-GPU and k
on ONE wave-filling slab of brand rows (1184 = 8 x 148 ... rounded to m-tiles) instead of all 10 k / 5 k brands.
"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fancyrec_b200 import _lib, ops, ranking

dev = torch.device("cuda:0")
lib = _lib.load()
g = torch.Generator(device=dev).manual_seed(7)


def unit_rows(n, d):
    out = torch.empty((n, ops.round_up(d, 64)), dtype=torch.bfloat16, device=dev)
    step = 250000
    for lo in range(0, n, step):
        hi = min(n, lo + step)
        x = torch.randn((hi - lo, d), generator=g, device=dev)
        out[lo:hi] = ops.finalize_posts(x, final_norm=True)[1]
    return out


def run(name, nb, n, d, k, reps=5):
    a = unit_rows(nb, d)
    b = unit_rows(n, d)
    lab = (torch.randperm(n, generator=g, device=dev) % nb).to(torch.int32)
    ws = None
    for _ in range(2):
        res = ops.score_topk(a, b, k, d=d, labels=lab, workspace=ws); ws = res["workspace"]
    torch.cuda.synchronize()
    lib.frx_probe_enable(1)
    beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    beg.record()
    for _ in range(reps):
        ops.score_topk(a, b, k, d=d, labels=lab, workspace=ws)
    end.record(); torch.cuda.synchronize()
    buf = np.zeros(64, dtype=np.float32)
    m = lib.frx_probe_read(buf.ctypes.data, 64)
    lib.frx_probe_enable(0)
    total = beg.elapsed_time(end) / reps
    kern = float(buf[:m].mean())
    fl = 2.0 * nb * n * d
    print("%-58s call %.3f ms (%.0f TFLOP/s) ; main kernel %.3f ms (%.0f TFLOP/s, %.3e pairs/s)" %
          (name, total, fl / total / 1e9, kern, fl / kern / 1e9, nb * n / kern * 1e3), flush=True)
    del a, b, lab, ws, res
    torch.cuda.empty_cache()


which = sys.argv[1:] or ["c2", "c4", "c5", "d1024"]
if "c2" in which:
    run("C2  1000 x 1M      D=3072 k=100", 1000, 1000000, 3072, 100)
if "c4" in which:
    run("C4/8 1152 of 10k brands x 2.5M-post shard D=3072 k=1000", 1152, 2500000, 3072, 1000, reps=3)
if "c5" in which:
    run("C5/8 1152 of 5k brands x 625k-post shard D=2048 k=100", 1152, 625000, 2048, 100)
    run("C5   1152 of 5k brands x 5M posts D=2048 k=100", 1152, 5000000, 2048, 100, reps=2)
if "d1024" in which:
    run("insCar-size 1000 x 1M D=1024 k=100", 1000, 1000000, 1024, 100)
