"""First-contact diagnostics for the tcgen05 score kernel (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fancyrec_b200 import ops, ranking

dev = torch.device("cuda:0")
print(torch.cuda.get_device_name(0), flush=True)
for (nb, npost, d) in [(128, 256, 64), (128, 256, 128), (100, 300, 200), (300, 5000, 1024)]:
    g = torch.Generator(device="cpu").manual_seed(nb + npost)
    a = torch.randn(nb, d, generator=g).to(dev)
    b = torch.randn(npost, d, generator=g).to(dev)
    A = ranking.to_operand(a); B = ranking.to_operand(b)
    ref = (A[:, :d].float() @ B[:, :d].float().t())
    t0 = time.time()
    out = ops.score_dense(A, B, d=d)
    torch.cuda.synchronize()
    err = (out - ref).abs()
    print("dense nb=%d np=%d d=%d: max err %.3e mean err %.3e (%.1f ms)" % (nb, npost, d, err.max().item(), err.mean().item(), (time.time()-t0)*1e3), flush=True)
    if err.max().item() > 1e-3:
        bad = (err > 1e-3)
        print("  bad fraction %.4f; bad rows %s ; bad cols %s" % (bad.float().mean().item(), bad.any(1).nonzero().flatten()[:16].tolist(), bad.any(0).nonzero().flatten()[:16].tolist()))
        print("  out[0,:8]", out[0, :8].tolist()); print("  ref[0,:8]", ref[0, :8].tolist())
    res = ops.score_topk(A, B, 10, d=d)
    torch.cuda.synchronize()
    want = torch.topk(out, min(10, npost), dim=1)
    ok = torch.equal(res["scores"][:, :min(10, npost)], want.values)
    print("  topk scores equal torch.topk on our tile:", ok, flush=True)
print("done")
