"""Multicast-cluster variant of the score kernel: bit-identity with the single-CTA kernel and timing at config 2."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fancyrec_b200 import _lib, ops, ranking
lib = _lib.load()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(3)
def problem(nb, n, d):
    brand = ranking.to_operand(torch.randn((nb, d), generator=g, device=dev))
    post = ranking.to_operand(torch.randn((n, d), generator=g, device=dev))
    labels = (torch.randperm(n, generator=g, device=dev) % nb).to(torch.int32)
    return brand, post, labels
cl = int(sys.argv[1])
for nb, n, d, k in [(300, 70001, 128, 100), (1000, 300000, 256, 64), (130, 5000, 96, 10)]:
    brand, post, labels = problem(nb, n, d)
    lib.frx_set_cluster(0)
    ref = ops.score_topk(brand, post, k, d=d, labels=labels)
    refd = ops.score_dense(brand, post, d=d)
    lib.frx_set_cluster(cl)
    got = ops.score_topk(brand, post, k, d=d, labels=labels, dense=True)
    torch.cuda.synchronize()
    ok = torch.equal(got["index"], ref["index"]) and torch.equal(got["scores"], ref["scores"]) and \
        torch.equal(got["pos_score"], ref["pos_score"]) and torch.equal(got["dense"], refd)
    cnt = ops.score_count(brand, post, ref["scores"][:, k // 2].contiguous(), ref["index"][:, k // 2].contiguous(), d=d)
    ok = ok and bool((cnt == k // 2).all())
    print("cluster %d  %d x %d x %d: %s" % (cl, nb, n, d, "bit-identical" if ok else "MISMATCH"))
if len(sys.argv) > 2:
    nb, n, d, k = 1000, 1000000, 3072, 100
    brand = ranking.to_operand(torch.randn((nb, d), generator=g, device=dev))
    post = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
    for lo in range(0, n, 65536):
        post[lo:lo + 65536] = ranking.to_operand(torch.randn((min(65536, n - lo), d), generator=g, device=dev))
    labels = (torch.randperm(n, generator=g, device=dev) % nb).to(torch.int32)
    for c in (0, cl):
        lib.frx_set_cluster(c)
        ws = ops.score_topk(brand, post, k, d=d, labels=labels)["workspace"]
        a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(); a.record()
        for _ in range(10):
            ops.score_topk(brand, post, k, d=d, labels=labels, workspace=ws)
        e.record(); torch.cuda.synchronize()
        print("cluster %d: fused score + top-k call at config 2: %.3f ms" % (c, a.elapsed_time(e) / 10))
