"""Where the time of one config-4 shard evaluation goes (single GPU): per-phase CUDA-event and host timings."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fancyrec_b200 import ops, ranking

dev = torch.device("cuda:0")
nb, n, d, k = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "10000,2500000,3072,1000".split(","))]
g = torch.Generator(device=dev).manual_seed(3)
brand = torch.randn((nb, d), generator=g, device=dev)
a = ranking.to_operand(brand)
bn = brand / brand.norm(dim=1, keepdim=True)
lab = (torch.randperm(n, generator=g, device=dev) % nb).to(torch.int32)
b = torch.empty((n, ops.round_up(d, 64)), dtype=torch.bfloat16, device=dev)
for lo in range(0, n, 250000):
    hi = min(n, lo + 250000)
    b[lo:hi] = ranking.to_operand(torch.randn((hi - lo, d), generator=g, device=dev) + 0.05 * (d ** 0.5) * bn[lab[lo:hi].long()])
ws = None
def run(trace):
    global ws
    marks = []
    def mark(name):
        if trace:
            e = torch.cuda.Event(enable_timing=True); e.record(); marks.append((name, e, time.perf_counter()))
    mark("start")
    res = ops.score_topk(a, b, k, d=d, labels=lab, workspace=ws); ws = res["workspace"]
    mark("score_topk")
    n_pos, bs, bi = ops.label_stats(lab, res["pos_score"], nb, 0)
    mark("label_stats")
    hit, first = ops.rank_from_topk(res["index"], lab, 0)
    mark("rank_from_topk")
    before = torch.zeros(nb, dtype=torch.int64, device=dev)
    ops.score_count(a, b, bs, ops.missing_thresholds(n_pos, first, bi), d=d, out=before)
    mark("count")
    st = dict(n_pos=n_pos, first_in_list=first, before_first=before, hit_mask=hit)
    t0 = time.perf_counter(); hs = ranking.host_statistics(st, n, False); t1 = time.perf_counter()
    out = ranking.aggregate(hs, n, False); t2 = time.perf_counter()
    mark("end")
    if trace:
        torch.cuda.synchronize()
        for (nm, e, t), (_, e0, t0_) in zip(marks[1:], marks[:-1]):
            print("%-16s gpu %8.3f ms   host %8.3f ms" % (nm, e0.elapsed_time(e), (t - t0_) * 1e3))
        print("host_statistics %.3f ms ; aggregate %.3f ms" % ((t1 - t0) * 1e3, (t2 - t1) * 1e3))
    return out
for _ in range(2): run(False)
torch.cuda.synchronize()
run(True)
t0 = time.perf_counter()
for _ in range(3): run(False)
torch.cuda.synchronize()
print("avg evaluation %.2f ms" % ((time.perf_counter() - t0) / 3 * 1e3))
