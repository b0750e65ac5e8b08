"""Config 5 on ONE GPU, device side: 5 M video posts x 32 frames x 2048-d fp32 (1.31 TB of frame rows, streamed through
one resident 13.4 GB chunk that is re-permuted between chunks so the posts differ) -> mean-pool + L2 norm -> bf16
operand (20.5 GB) -> 5 000 brands, fused top-100 -> recall@1/5/10, MedR, MeanR, NDCG@10/50.
Times the finalisation passes and the evaluation separately (the frame rows cannot be resident; feeding them is I/O)."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fancyrec_b200 import ops, ranking

dev = torch.device("cuda:0")
nb, n, f, d = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "5000,5000000,32,2048".split(","))]
chunk = 51200
g = torch.Generator(device=dev).manual_seed(5)
brand = torch.randn((nb, d), generator=g, device=dev)
a = ranking.to_operand(brand)
frames = torch.randn((chunk * f, d), generator=g, device=dev).abs_()
rp = (torch.arange(chunk + 1, device=dev) * f).to(torch.int64)
lab = (torch.randperm(n, generator=g, device=dev) % nb).to(torch.int32)
post_op = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
fin_ms = 0.0
lo = 0
while lo < n:
    m = min(chunk, n - lo)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    rc = ops._lib.load().frx_finalize_posts(frames.data_ptr(), rp.data_ptr(), 0, 0, m, d, 0, 4, 0,
                                             post_op[lo:lo + m].data_ptr(), d, torch.cuda.current_stream().cuda_stream)
    e1.record()
    assert rc == 0
    torch.cuda.synchronize()
    fin_ms += e0.elapsed_time(e1)
    lo += m
    frames = frames.roll(shifts=7919, dims=0)            # next chunk: different frames pool together (not timed)
gb = (n * f * d * 4 + n * d * 2) / 1e9
print("finalise: %d posts x %d frames x %d in %d chunks: %.1f ms of kernel time -> %.3e posts/s, %.0f GB/s (%.2f TB streamed)"
      % (n, f, d, (n + chunk - 1) // chunk, fin_ms, n / fin_ms * 1e3, gb / fin_ms * 1e3, gb / 1e3), flush=True)
del frames
ws = None
def evaluate():
    global ws
    st = ranking.device_rank_statistics(a, post_op, lab, d, k=100, want_auc=False, workspace=ws)
    ws = st["workspace"]
    return ranking.aggregate(ranking.host_statistics(st, n, False), n, False), st
evaluate(); torch.cuda.synchronize()
t0 = time.perf_counter()
res, st = evaluate()
torch.cuda.synchronize()
t = time.perf_counter() - t0
missing = int(((st["first_in_list"] < 0) & (st["n_pos"] > 0)).sum())
print("evaluate: %d brands x %d posts, top-100 + recall/MedR/MeanR/NDCG sweep: %.1f ms -> %.3e pairs/s (%d brands needed the count pass)"
      % (nb, n, t * 1e3, nb * n / t, missing))
print("config 5 device total: %.1f ms ; result MedR %.0f MeanR %.0f NDCG@10 %.4f NDCG@50 %.4f r@1 %.2f r@5 %.2f r@10 %.2f"
      % (fin_ms + t * 1e3, res[0], res[1], res[3], res[4], res[5], res[6], res[7]))
