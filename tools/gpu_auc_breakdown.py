"""Kernel list of ONE evaluator.test_post_ranking call (exact AUC included) at config-2 size; run under
ncu --metrics gpu__time_duration.sum to see where the AUC sweep spends its time."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fancyrec_b200 import evaluator, model as fmodel
dev = torch.device("cuda:0")
nb, n, d = 1000, 1000000, 3072
g = torch.Generator(device=dev).manual_seed(2)
opt = types.SimpleNamespace(brand_num=nb, common_embedding_size=d, brand_aspect=2000)
mdl = types.SimpleNamespace(brand_encoding=fmodel.BrandAspects(opt).to(dev), opt=opt)
brand = evaluator.brand_matrix(mdl, nb)
lab = (torch.randperm(n, generator=g, device=dev) % nb)
post = torch.randn((n, d), generator=g, device=dev)
post += 0.05 * (d ** 0.5) * (brand / brand.norm(dim=1, keepdim=True))[lab]
evaluator.test_post_ranking(nb, 'auc', mdl, post, lab)
torch.cuda.synchronize()
torch.cuda.profiler.start()
import time
t0 = time.perf_counter()
res = evaluator.test_post_ranking(nb, 'auc', mdl, post, lab)
torch.cuda.synchronize()
print("test_post_ranking (AUC incl.): %.2f ms  %s" % ((time.perf_counter() - t0) * 1e3, [round(float(x), 4) for x in res]))
torch.cuda.profiler.stop()
