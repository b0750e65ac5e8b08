"""Config-4 shard (10 000 x 2.5 M x 3072): fused top-1000 vs the COUNT epilogue vs cuBLAS (bf16 out), back to back."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fancyrec_b200 import _lib, ops, ranking
dev = torch.device("cuda:0")
lib = _lib.load()
nb, n, d, k = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "10000,2500000,3072,1000".split(","))]
planted = len(sys.argv) > 2 and sys.argv[2] == "planted"
g = torch.Generator(device=dev).manual_seed(3)
brand = torch.randn((nb, d), generator=g, device=dev)
a = ranking.to_operand(brand)
bn = brand / brand.norm(dim=1, keepdim=True)
lab = (torch.randperm(n, generator=g, device=dev) % nb).to(torch.int32)
b = torch.empty((n, ops.round_up(d, 64)), dtype=torch.bfloat16, device=dev)
for lo in range(0, n, 250000):
    hi = min(n, lo + 250000)
    x = torch.randn((hi - lo, d), generator=g, device=dev)
    if planted:
        x += 0.05 * (d ** 0.5) * bn[lab[lo:hi].long()]
    b[lo:hi] = ranking.to_operand(x)
fl = 2.0 * nb * n * d
def bench(name, fn, reps=6):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); ts.append((e0, e1))
    torch.cuda.synchronize()
    ms = [x.elapsed_time(y) for x, y in ts]
    print("%-34s %s ms -> %.0f TFLOP/s (last)" % (name, " ".join("%.1f" % m for m in ms), fl / ms[-1] / 1e9), flush=True)
ws = torch.empty(lib.frx_score_topk_workspace_bytes(nb, n, d, k), dtype=torch.uint8, device=dev)
thr_s = torch.zeros(nb, device=dev); thr_i = torch.zeros(nb, dtype=torch.int32, device=dev)
cnt = torch.zeros(nb, dtype=torch.int64, device=dev)
bench("score_topk k=%d (sample+main+merge)" % k, lambda: ops.score_topk(a, b, k, d=d, labels=lab, workspace=ws))
bench("score_topk k=%d no labels" % k, lambda: ops.score_topk(a, b, k, d=d, workspace=ws))
bench("score_count (trivial epilogue)", lambda: ops.score_count(a, b, thr_s, thr_i, d=d, out=cnt))
out = torch.empty((nb, n), dtype=torch.bfloat16, device=dev)
bench("torch.matmul bf16 out (cuBLAS)", lambda: torch.matmul(a, b.t(), out=out))
bench("score_topk again", lambda: ops.score_topk(a, b, k, d=d, labels=lab, workspace=ws))
