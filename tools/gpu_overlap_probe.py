"""Does the post finalisation really run UNDER the contraction?  Event timelines on a B200:
  1. each kernel alone (finalise of 1 M x 3072 rows; fused score + top-k call of 1 000 x 1 M);
  2. both at once on two streams (contraction first, finalisation of the next batch right behind it);
  3. the pipelined evaluation (pipeline.EvalPipeline), per-step phases on both streams + host submit times.
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, numpy as np
import bench
from fancyrec_b200 import ops, ranking, sharded, pipeline

dev = torch.device("cuda:0")
torch.cuda.set_device(0)
cfg = dict(bench.CFG)
nb, n, d, k = 1000, 1000000, 3072, 100
w, e, labels, visual, text = bench.make_workload(dev, 0, nb, n, cfg)
brand_op = ops.finalize_posts(ops.brand_embed(w, e, nb=nb), final_norm=True)[1]
post_a = ops.finalize_posts(visual, text, visual_norm=True, text_norm=True, final_norm=True)[1]
post_b = torch.empty_like(post_a)
ws = ops.score_topk(brand_op, post_a, k, d=d, labels=labels)["workspace"]
side = torch.cuda.Stream(dev)
main = torch.cuda.current_stream(dev)


def ev():
    return torch.cuda.Event(enable_timing=True)


BPS = 0


def fin(out):
    ops.finalize_posts(visual, text, visual_norm=True, text_norm=True, final_norm=True, out_bf16=out, blocks_per_sm=BPS)


def gemm():
    return ops.score_topk(brand_op, post_a, k, d=d, labels=labels, workspace=ws)


for _ in range(3):
    fin(post_b); gemm()
torch.cuda.synchronize()
for BPS in (0, 1, 2):
    a, b, c = ev(), ev(), ev()
    a.record(); fin(post_b); b.record(); gemm(); c.record()
    torch.cuda.synchronize()
    print("alone (finalise blocks/SM bound %d): finalise" % BPS + " %.3f ms   score_topk call %.3f ms   sum %.3f" % (a.elapsed_time(b), b.elapsed_time(c), a.elapsed_time(c)))
for order, BPS in (("gemm first", 0), ("finalise first", 0), ("gemm first", 1), ("finalise first", 1), ("gemm first", 2),
                   ("finalise first", 2), ("finalise first", 3)):
    for rep in range(3):
        t0, g0, g1, f0, f1 = ev(), ev(), ev(), ev(), ev()
        torch.cuda.synchronize()
        t0.record(main)
        side.wait_event(t0)
        if order == "gemm first":
            g0.record(main); gemm(); g1.record(main)
            with torch.cuda.stream(side):
                f0.record(side); fin(post_b); f1.record(side)
        else:
            with torch.cuda.stream(side):
                f0.record(side); fin(post_b); f1.record(side)
            g0.record(main); gemm(); g1.record(main)
        torch.cuda.synchronize()
        print("concurrent (%s, finalise blocks/SM bound %d): gemm [%.3f, %.3f] = %.3f ms   finalise [%.3f, %.3f] = %.3f ms   makespan %.3f"
              % (order, BPS, t0.elapsed_time(g0), t0.elapsed_time(g1), g0.elapsed_time(g1), t0.elapsed_time(f0), t0.elapsed_time(f1),
                 f0.elapsed_time(f1), max(t0.elapsed_time(g1), t0.elapsed_time(f1))))

BPS = 0
# ---- pipelined evaluation, per-step phases
for overlap in (False, True):
    pipe = pipeline.EvalPipeline(dev, nb, n, cfg["dv"], cfg["dt"], k=k, overlap=overlap)
    inputs = (w, e, visual, text, labels)
    for _ in range(3):
        pipe.result(pipe.submit(*inputs))
    torch.cuda.synchronize()
    marks = []
    t_ref = ev(); t_ref.record()
    h_ref = time.perf_counter()
    prev = None
    steps = 8
    for t in range(steps):
        h0 = time.perf_counter()
        s0 = ev(); s0.record()
        tk = pipe.submit(*inputs)
        s1 = ev(); s1.record()
        h1 = time.perf_counter()
        if prev is not None:
            pipe.result(prev)
        h2 = time.perf_counter()
        prev = tk
        slot = tk % pipe.depth
        marks.append((s0, s1, h0 - h_ref, h1 - h_ref, h2 - h_ref))
    pipe.result(prev)
    end = ev(); end.record()
    torch.cuda.synchronize()
    print("pipeline overlap=%s: total %.3f ms for %d steps = %.3f ms/step" % (overlap, t_ref.elapsed_time(end), steps, t_ref.elapsed_time(end) / steps))
    for t, (s0, s1, h0, h1, h2) in enumerate(marks):
        print("  step %d: main-stream [%.3f, %.3f] ms   host submit [%.3f, %.3f] collect-prev done %.3f"
              % (t, t_ref.elapsed_time(s0), t_ref.elapsed_time(s1), h0 * 1e3, h1 * 1e3, h2 * 1e3))
    del pipe
    torch.cuda.empty_cache()
