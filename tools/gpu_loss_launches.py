"""Config-3 losses (B = 512, D = 3072) a few times: for an ncu launch list and an event timing of each loss."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import torch
from fancyrec_b200 import loss, loss_ctrs
dev = torch.device("cuda:0")
b, d = 512, int(os.environ.get("D", "3072"))
g = torch.Generator(device=dev).manual_seed(1)
ids = torch.randint(0, 51, (b,), generator=g, device=dev)
brand = torch.randn((b, d), generator=g, device=dev, requires_grad=True)
post = torch.randn((b, d), generator=g, device=dev, requires_grad=True)
trip = loss.TripletLoss(margin=0.2, cost_style="sum").to(dev)
opt = SimpleNamespace(cost_style="sum", queue_size=5120, common_embedding_size=d, no_queue=False, no_intra=False)
con = loss_ctrs.ContrastiveLoss(opt).to(dev)
reps = int(os.environ.get("REPS", "20"))


def timed(name, fn):
    for _ in range(3):
        fn()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        fn()
    e.record()
    torch.cuda.synchronize()
    print("%-28s %.1f us" % (name, a.elapsed_time(e) / reps * 1e3))


def t_fb():
    brand.grad = None; post.grad = None
    trip(ids, brand, post).backward()


def c_fb():
    brand.grad = None; post.grad = None
    con(brand, post).backward()


def t_f():
    with torch.no_grad():
        trip(ids, brand, post)


timed("triplet fwd+bwd", t_fb)
timed("triplet fwd (no_grad)", t_f)
timed("contrastive fwd+bwd", c_fb)


# ---- device time without host gaps: the C entry point captured into a CUDA graph and replayed
from fancyrec_b200 import ops


def graph_us(fn, reps=50):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        fn()
    for _ in range(3):
        g.replay()
    a, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(reps):
        g.replay()
    e.record()
    torch.cuda.synchronize()
    return a.elapsed_time(e) / reps * 1e3


bd, pd = brand.detach(), post.detach()
W = torch.randn((3072, d), device=dev) / d ** 0.5
print("triplet fwd+bwd  device (graph replay) %.1f us" % graph_us(lambda: ops.triplet_fwd_bwd(ids, bd, pd, 0.2, 0, True)))
print("triplet fwd      device (graph replay) %.1f us" % graph_us(lambda: ops.triplet_fwd_bwd(ids, bd, pd, 0.2, 0, False)))
keys = con.queue
print("contrastive fwd+bwd device (graph replay) %.1f us" % graph_us(
    lambda: ops.contrastive_fwd_bwd(bd, pd, keys, 0, False, 0.03, 0.8, 0, True)))
print("linear 512x3072x3072 (MFC-sized) device %.1f us" % graph_us(lambda: ops.linear(pd, W)))
