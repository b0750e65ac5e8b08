"""Phase timeline of one bench step (CUDA events + host clocks)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import bench
from fancyrec_b200 import ops, ranking, sharded
dev = torch.device("cuda:0")
torch.cuda.set_device(0)
cfg = dict(bench.CFG)
nb, n = 1000, 1000000
w, e, labels, visual, text = bench.make_workload(dev, 0, nb, n, cfg)
ws = None
def step(trace):
    global ws
    ev = []
    def mark(name):
        if trace:
            x = torch.cuda.Event(enable_timing=True); x.record(); ev.append((name, x, time.perf_counter()))
    mark("start")
    brand = ops.brand_embed(w, e, nb=nb); brand_op = ops.finalize_posts(brand, final_norm=True)[1]
    mark("brand")
    post_op = ops.finalize_posts(visual, text, visual_norm=True, text_norm=True, final_norm=True)[1]
    mark("finalize")
    res = ops.score_topk(brand_op, post_op, 100, d=3072, labels=labels, workspace=ws); ws = res["workspace"]
    mark("score_topk")
    n_pos, bs, bi = ops.label_stats(labels, res["pos_score"], nb, 0)
    hit, first = ops.rank_from_topk(res["index"], labels, 0)
    mark("stats")
    before = torch.zeros(nb, dtype=torch.int64, device=dev)
    thr_index = ops.missing_thresholds(n_pos, first, bi)
    ops.score_count(brand_op, post_op, bs, thr_index, d=3072, out=before)
    mark("count")
    t0 = t1 = time.perf_counter()
    st = dict(n_pos=n_pos, first_in_list=first, before_first=before, hit_mask=hit)
    t2 = time.perf_counter(); hs = ranking.host_statistics(st, n, False); t3 = time.perf_counter()
    out = ranking.aggregate(hs, n, False); t4 = time.perf_counter()
    mark("end")
    if trace:
        torch.cuda.synchronize()
        base_ev, base_t = ev[0][1], ev[0][2]
        for name, x, t in ev[1:]:
            print("%-12s gpu %.3f ms   host %.3f ms" % (name, base_ev.elapsed_time(x), (t - base_t) * 1e3))
        print("item() wait %.3f ms ; host_statistics %.3f ms ; aggregate %.3f ms" % ((t1 - t0) * 1e3, (t3 - t2) * 1e3, (t4 - t3) * 1e3))
    return out
for _ in range(3): step(False)
torch.cuda.synchronize()
step(True)
t0 = time.perf_counter()
for _ in range(20): step(False)
torch.cuda.synchronize()
print("avg step %.3f ms" % ((time.perf_counter() - t0) / 20 * 1e3))
