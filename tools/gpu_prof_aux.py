"""Small driver for ncu captures of the HBM-bound kernels at scale: C5 pooling and the AUC row sweep."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fancyrec_b200 import ops, ranking
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
n5, f, dv = 60000, 32, 2048
frames = torch.randn((n5 * f, dv), generator=g, device=dev).abs_()
rp = (torch.arange(n5 + 1, device=dev) * f).to(torch.int64)
for _ in range(2): ops.finalize_posts(frames, row_ptr=rp, final_norm=True)
del frames
nb, n, d = 512, 1000000, 1024
brand = torch.randn((nb, d), generator=g, device=dev)
post = torch.randn((n, d), generator=g, device=dev)
lab = (torch.randperm(n, generator=g, device=dev) % nb).to(torch.int32)
a, b = ranking.to_operand(brand), ranking.to_operand(post)
st = ranking.device_rank_statistics(a, b, lab, d, k=64, want_auc=True)
torch.cuda.synchronize()
print("ok", int(st["auc_num"].sum()))
