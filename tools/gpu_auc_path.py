"""One AUC-inclusive evaluation at config-2 size (for an ncu launch list / event breakdown of its kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from fancyrec_b200 import ops, ranking
dev = torch.device("cuda:0")
cfg = dict(bench.CFG)
nb, n = 1000, 1000000
w, e, labels, visual, text = bench.make_workload(dev, 0, nb, n, cfg)
brand = ops.brand_embed(w, e, nb=nb)
posts = ops.finalize_posts(visual, text, visual_norm=True, text_norm=True, final_norm=False, want_f32=True, want_bf16=False)[0]
del visual, text
reps = int(os.environ.get("REPS", "3"))
for _ in range(reps):
    res = ranking.rank_posts(brand, posts, labels, k=100, want_auc=True)[0]
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
res = ranking.rank_posts(brand, posts, labels, k=100, want_auc=True)[0]
b.record(); torch.cuda.synchronize()
print("rank_posts(want_auc=True) = evaluator.test_post_ranking body: %.3f ms" % a.elapsed_time(b), tuple(float(x) for x in res))
