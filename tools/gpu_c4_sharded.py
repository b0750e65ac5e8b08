"""Config 4 end to end: 10 000 brands x (2.5 M posts per GPU) sharded over the GPUs of one box, D = 3072, fused top-1000
per shard + ONE NCCL all-gather of the candidate lists + merge + global rank statistics.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
      tools/gpu_c4_sharded.py [nb,posts_per_gpu,d,k]
"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from fancyrec_b200 import ops, ranking, sharded

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    os.environ.setdefault("NCCL_DEBUG", "WARN")
    dist.init_process_group("nccl", device_id=dev)
nb, n_local, d, k = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "10000,2500000,3072,1000".split(","))]
n_total = n_local * world
g = torch.Generator(device=dev).manual_seed(3)                   # brands identical on every rank
brand = torch.randn((nb, d), generator=g, device=dev)
a = ranking.to_operand(brand)
bn = brand / brand.norm(dim=1, keepdim=True)
g = torch.Generator(device=dev).manual_seed(100 + rank)
lab = (torch.randperm(n_local, generator=g, device=dev) % nb).to(torch.int32)
b = torch.empty((n_local, ops.round_up(d, 64)), dtype=torch.bfloat16, device=dev)
for lo in range(0, n_local, 250000):
    hi = min(n_local, lo + 250000)
    x = torch.randn((hi - lo, d), generator=g, device=dev) + 0.05 * (d ** 0.5) * bn[lab[lo:hi].long()]   # planted signal
    b[lo:hi] = ranking.to_operand(x)
del x
ws = None
def step():
    global ws
    st = sharded.sharded_rank_statistics(a, b, lab, d, k, n_total, workspace=ws)
    ws = st["workspace"]
    hs = ranking.host_statistics(st, n_total, want_auc=False)
    return ranking.aggregate(hs, n_total, want_auc=False), st
for _ in range(2):
    res, st = step()
def sync():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
sync()
beg, end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
beg.record()
for _ in range(reps):
    res, st = step()
end.record()
sync()
ms = torch.tensor([beg.elapsed_time(end) / reps], device=dev, dtype=torch.float64)
if world > 1:
    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
ms = float(ms.item())
if rank == 0:
    pairs = float(nb) * n_total
    print("C4 %d brands x %d posts (%d per GPU) on %d GPU(s), D=%d, top-%d: %.1f ms per evaluation -> %.3e pairs/s "
          "(%.0f TFLOP/s aggregate) ; brands needing the count pass: %d ; MedR %.0f NDCG@10 %.4f r@1 %.1f"
          % (nb, n_total, n_local, world, d, k, ms, pairs / ms * 1e3, 2 * pairs * d / ms / 1e9, int(((st["first_in_list"] < 0) & (st["n_pos"] > 0)).sum()),
             res[0], res[3], res[5]), flush=True)
if world > 1:
    dist.destroy_process_group()
