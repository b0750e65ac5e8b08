"""Config 4 at its true per-GPU size: 10 000 brands x 2.5 M posts (one of 8 shards), D = 3072, fused top-1000.
One timed call (1.536e17 flop, about two minutes), then an exactness check of 128 brand rows against the dense tile."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from fancyrec_b200 import _lib, ops, ranking

dev = torch.device("cuda:0")
lib = _lib.load()
nb, n, d, k = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "10000,2500000,3072,1000".split(","))]
g = torch.Generator(device=dev).manual_seed(11)
a = ranking.to_operand(torch.randn((nb, d), generator=g, device=dev))
b = torch.empty((n, ops.round_up(d, 64)), dtype=torch.bfloat16, device=dev)
for lo in range(0, n, 250000):
    b[lo:lo + 250000] = ranking.to_operand(torch.randn((min(250000, n - lo), d), generator=g, device=dev))
lab = (torch.randperm(n, generator=g, device=dev) % nb).to(torch.int32)
need = lib.frx_score_topk_workspace_bytes(nb, n, d, k)
print("workspace %.2f GB, operands %.2f GB" % (need / 1e9, (a.numel() + b.numel()) * 2 / 1e9), flush=True)
ws = torch.empty(need, dtype=torch.uint8, device=dev)
ops.score_topk(a[:256], b[:300000], k, d=d, labels=None, workspace=ws)          # warm the kernels / tensor maps
torch.cuda.synchronize()
lib.frx_probe_enable(1)
t0 = time.perf_counter()
res = ops.score_topk(a, b, k, d=d, labels=lab, index_base=5000000, workspace=ws)
torch.cuda.synchronize()
t = time.perf_counter() - t0
buf = np.zeros(8, dtype=np.float32)
m = lib.frx_probe_read(buf.ctypes.data, 8)
lib.frx_probe_enable(0)
fl = 2.0 * nb * n * d
print("C4 shard %d x %d D=%d k=%d: call %.2f s (%.0f TFLOP/s, %.3e pairs/s) ; main kernel %.2f s (%.0f TFLOP/s)"
      % (nb, n, d, k, t, fl / t / 1e12, nb * n / t, buf[m - 1] / 1e3, fl / buf[m - 1] / 1e9), flush=True)
rows = torch.arange(0, nb, max(1, nb // 128), device=dev)[:128]
dense = ops.score_dense(a[rows].contiguous(), b, d=d)
want = torch.topk(dense, k, dim=1).values
got = res["scores"][rows]
idx = res["index"][rows].long() - 5000000
ok = torch.equal(got, want) and torch.equal(torch.gather(dense, 1, idx), got)
ps = res["pos_score"]
sel = torch.isin(lab, rows.to(torch.int32))
pos_ok = torch.equal(ps[sel], dense[torch.searchsorted(rows, lab[sel].long()), sel.nonzero().squeeze(1)])
print("exact top-%d scores on %d sampled brand rows: %s ; positives' scores exact: %s" % (k, len(rows), ok, pos_ok))
assert ok and pos_ok
