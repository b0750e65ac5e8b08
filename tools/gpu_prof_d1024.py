import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fancyrec_b200 import ops, ranking
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
nb, n, d = 1000, 1000000, int(os.environ.get("D", "1024"))
a = ranking.to_operand(torch.randn((nb, d), generator=g, device=dev))
b = ranking.to_operand(torch.randn((n, d), generator=g, device=dev))
lab = (torch.randperm(n, generator=g, device=dev) % nb).to(torch.int32)
for _ in range(3):
    r = ops.score_topk(a, b, 100, d=d, labels=lab)
torch.cuda.synchronize()
print("ok")
