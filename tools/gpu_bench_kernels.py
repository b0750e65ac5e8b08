"""Micro-benchmarks of the individual kernels at the C2 shapes (run on the GPU box)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fancyrec_b200 import ops, ranking

dev = torch.device("cuda:0")
n, dv, dt, nb = 1000000, 2048, 1024, 1000
g = torch.Generator(device=dev).manual_seed(1)
visual = torch.randn((n, dv), generator=g, device=dev)
text = torch.randn((n, dt), generator=g, device=dev)
w = torch.randn((nb + 1, 2000), generator=g, device=dev)
e = torch.randn((2000, dv + dt), generator=g, device=dev)

def timeit(fn, reps=5):
    fn(); torch.cuda.synchronize()
    b, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.record()
    for _ in range(reps): fn()
    en.record(); torch.cuda.synchronize()
    return b.elapsed_time(en) / reps

gb = (n * (dv + dt) * 4 + n * 3072 * 2) / 1e9
t = timeit(lambda: ops.finalize_posts(visual, text, visual_norm=True, text_norm=True, final_norm=True))
print("finalize branch-norm+concat+norm -> bf16: %.3f ms  %.0f GB/s" % (t, gb / t * 1e3))
t = timeit(lambda: ops.finalize_posts(visual, text, final_norm=True))
print("finalize concat+norm -> bf16:            %.3f ms  %.0f GB/s" % (t, gb / t * 1e3))
post = torch.cat([visual[:, :1024], text], 1)[:, :2048].contiguous()
gb2 = (n * 2048 * 4 + n * 2048 * 2) / 1e9
t = timeit(lambda: ops.finalize_posts(post, final_norm=True))
print("finalize 2048-d norm -> bf16:            %.3f ms  %.0f GB/s" % (t, gb2 / t * 1e3))
# C5-shaped pooling: 32 frames x 2048 per post, 100k posts (26 GB would not fit next to the rest)
np5 = 60000
frames = torch.randn((np5 * 32, dv), generator=g, device=dev).abs_()
row_ptr = (torch.arange(np5 + 1, device=dev) * 32).to(torch.int64)
gb5 = (np5 * 32 * dv * 4 + np5 * dv * 2) / 1e9
t = timeit(lambda: ops.finalize_posts(frames, row_ptr=row_ptr, final_norm=True))
print("finalize C5 pool 32x2048 -> bf16:        %.3f ms  %.0f GB/s  %.3e posts/s" % (t, gb5 / t * 1e3, np5 / t * 1e3))
t = timeit(lambda: ops.brand_embed(w, e, nb=nb))
print("brand_embed 1000x2000x3072:              %.3f ms  %.1f TFLOP/s" % (t, 2 * nb * 2000 * 3072 / t / 1e9))
del frames
a = ranking.to_operand(ops.brand_embed(w, e, nb=nb))
b = ops.finalize_posts(visual, text, visual_norm=True, text_norm=True, final_norm=True)[1]
lab = (torch.randperm(n, generator=g, device=dev) % nb).to(torch.int32)
import bench as _bench
def sustained(name, fn, flops, secs=2.0):
    fn(); torch.cuda.synchronize()
    smp = _bench.ClockSampler(0); smp.start()
    t0 = time.time(); n = 0
    b, en = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    b.record()
    while time.time() - t0 < secs:
        for _ in range(10): fn()
        n += 10
        torch.cuda.synchronize()
    en.record(); torch.cuda.synchronize()
    ms = b.elapsed_time(en) / n
    c = smp.stop()
    print("%-42s %.3f ms  %.0f TFLOP/s  sm_mhz(median)=%s reasons=%s" % (name, ms, flops / ms / 1e9, c["sm_mhz"], c["reasons"]))
F = 2 * nb * n * 3072
thr_s = torch.zeros(nb, device=dev); thr_i = torch.zeros(nb, dtype=torch.int32, device=dev)
sustained("sustained score_count", lambda: ops.score_count(a, b, thr_s, thr_i, d=3072), F)
sustained("sustained score_topk k=100 labels", lambda: ops.score_topk(a, b, 100, d=3072, labels=lab), F)
sustained("sustained score_topk k=100 no labels", lambda: ops.score_topk(a, b, 100, d=3072), F)
am = a.float() ; bm = b[:, :3072]
sustained("sustained torch.matmul bf16 (cuBLAS) 1000x1Mx3072", lambda: torch.matmul(a, b.t()), F)
for k in (64, 100, 1000):
    t = timeit(lambda: ops.score_topk(a, b, k, d=3072, labels=lab))
    print("score_topk k=%4d (kernel+merge):         %.3f ms  %.0f TFLOP/s" % (k, t, 2 * nb * n * 3072 / t / 1e9))
t = timeit(lambda: ops.score_count(a, b, torch.zeros(nb, device=dev), torch.zeros(nb, dtype=torch.int32, device=dev), d=3072))
print("score_count:                             %.3f ms  %.0f TFLOP/s" % (t, 2 * nb * n * 3072 / t / 1e9))
d1 = 1024
a1 = ranking.to_operand(torch.randn((nb, d1), generator=g, device=dev))
b1 = ranking.to_operand(torch.randn((n, d1), generator=g, device=dev))
t = timeit(lambda: ops.score_topk(a1, b1, 100, d=d1, labels=lab))
print("score_topk D=1024 k=100:                 %.3f ms  %.0f TFLOP/s" % (t, 2 * nb * n * d1 / t / 1e9))
