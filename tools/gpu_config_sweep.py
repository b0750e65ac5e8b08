"""One timing line per BASELINE.json config (scaled where the full size does not fit one GPU / a few seconds).
Run on the GPU box:  python tools/gpu_config_sweep.py > gpurun_out/config_sweep.txt"""
import os, sys, time, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from fancyrec_b200 import evaluator, loss as floss, loss_ctrs as fctrs, model as fmodel, ops, ranking

dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(20261018)


def timeit(fn, reps=3, warm=1):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps


def brand_model(nb, a, d):
    opt = types.SimpleNamespace(brand_num=nb, common_embedding_size=d, brand_aspect=a)
    return types.SimpleNamespace(brand_encoding=fmodel.BrandAspects(opt).to(dev), opt=opt)


def planted(nb, n, d, brand, signal=0.05):
    lab = (torch.randperm(n, generator=g, device=dev) % nb)
    x = torch.empty((n, d), device=dev)
    bn = brand / brand.norm(dim=1, keepdim=True)
    for lo in range(0, n, 65536):
        hi = min(n, lo + 65536)
        x[lo:hi] = torch.randn((hi - lo, d), generator=g, device=dev) + signal * d ** 0.5 * bn[lab[lo:hi]]
    return lab, x

# ---- C1: 50 brands x 10k posts, frames mean-pooled (F ~ U{1..40}) + text, full 8-tuple incl. AUC -------------
nb, n, dv, dt, a = 50, 10000, 2048, 1024, 2000
mdl = brand_model(nb, a, dv + dt)
counts = torch.randint(1, 41, (n,), generator=g, device=dev)
row_ptr = torch.cat([torch.zeros(1, dtype=torch.int64, device=dev), counts.cumsum(0)])
frames = torch.randn((int(row_ptr[-1]), dv), generator=g, device=dev).mul_(0.5).add_(0.3).clamp_(min=0)
text = torch.randn((n, dt), generator=g, device=dev).abs_()
lab = torch.randint(0, nb, (n,), generator=g, device=dev)
def c1():
    post = ops.finalize_posts(frames, text, row_ptr=row_ptr, visual_norm=True, text_norm=True, final_norm=False,
                              want_f32=True, want_bf16=False)[0]
    return evaluator.test_post_ranking(nb, 'auc', mdl, post, lab)
t = timeit(c1)
print("C1  50 x 10k (pooled frames %d rows + text, AUC/NDCG/recall/MedR 8-tuple): %.2f ms  -> %.3e pairs/s   result %s"
      % (int(row_ptr[-1]), t * 1e3, nb * n / t, [round(float(x), 4) for x in c1()]))
del frames, text

# ---- C2 with the full 8-tuple (AUC through the dense row sweep): 1k x 1M, D = 3072 --------------------------------
nb, n, d = 1000, 1000000, 3072
mdl = brand_model(nb, a, d)
brand = evaluator.brand_matrix(mdl, nb)
lab, post = planted(nb, n, d, brand)
t = timeit(lambda: evaluator.test_post_ranking(nb, 'auc', mdl, post, lab), reps=2)
print("C2+ 1k x 1M D=3072 through evaluator.test_post_ranking (incl. exact AUC): %.1f ms -> %.3e pairs/s  result %s"
      % (t * 1e3, nb * n / t, [round(float(x), 4) for x in evaluator.test_post_ranking(nb, 'auc', mdl, post, lab)]))
del post

# ---- C3: loss tiles, B = 512 -----------------------------------------------------------------------------------------
for d in (1024, 2048, 3072):
    b = 512
    ids = torch.randint(0, 51, (b,), generator=g, device=dev)
    be = torch.randn((b, d), generator=g, device=dev).requires_grad_()
    pe = torch.randn((b, d), generator=g, device=dev).requires_grad_()
    crit = floss.TripletLoss(margin=0.2, cost_style='sum')
    def trip():
        be.grad = pe.grad = None
        crit(ids, be, pe).backward()
    t = timeit(trip, reps=20, warm=3)
    opt = types.SimpleNamespace(cost_style='mean', queue_size=5120, common_embedding_size=d, no_queue=False, no_intra=False)
    cl = fctrs.ContrastiveLoss(opt).to(dev)
    def ctr():
        be.grad = pe.grad = None
        cl(be, pe).backward()
    t2 = timeit(ctr, reps=20, warm=3)
    xl = fctrs.CrossCLR_onlyIntraModality(cost_style='mean').to(dev)
    def xclr():
        be.grad = pe.grad = None
        xl(be, pe).backward()
    t3 = timeit(xclr, reps=20, warm=3)
    ll = floss.LabLoss()
    def lab_():
        be.grad = None
        ll(be).backward()
    t4 = timeit(lab_, reps=20, warm=3)
    print("C3  B=512 D=%d: TripletLoss fwd+bwd %.3f ms ; ContrastiveLoss(Q=5120) fwd+bwd %.3f ms ; CrossCLR fwd+bwd %.3f ms ; "
          "LabLoss fwd+bwd %.3f ms" % (d, t * 1e3, t2 * 1e3, t3 * 1e3, t4 * 1e3))

# ---- C4: true per-GPU shard (2.5 M posts, k = 1000, D = 3072) on one wave-filling slab of 1152 of the 10k brands ------
# (the epilogue's per-row behaviour depends on posts-per-shard and k, not on the number of brand rows)
nb, n, d, k = 1152, 2500000, 3072, 1000
brand = torch.randn((nb, d), generator=g, device=dev)
a_op = ranking.to_operand(brand)
b_op = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
for lo in range(0, n, 250000):
    b_op[lo:lo + 250000] = ranking.to_operand(torch.randn((min(250000, n - lo), d), generator=g, device=dev))
lab32 = (torch.randperm(n, generator=g, device=dev) % nb).to(torch.int32)
res = ops.score_topk(a_op, b_op, k, d=d, labels=lab32)
t = timeit(lambda: ops.score_topk(a_op, b_op, k, d=d, labels=lab32, workspace=res["workspace"]), reps=3)
print("C4/8 1152 of 10k brands x 2.5M-post shard D=3072 top-1000 fused (sample + main + merge): %.1f ms -> %.3e pairs/s = %.0f TFLOP/s"
      % (t * 1e3, nb * n / t, 2 * nb * n * d / t / 1e12))
del b_op, res

# ---- C5 scaled: 400k video posts x 32 frames x 2048 (105 GB of frames streamed in 8 chunks), 5k brands ---------------
nb, n, f, d = 5000, 400000, 32, 2048
brand = torch.randn((nb, d), generator=g, device=dev)
a_op = ranking.to_operand(brand)
lab = (torch.randperm(n, generator=g, device=dev) % nb).to(torch.int32)
chunk = n // 8
frames = torch.randn((chunk * f, d), generator=g, device=dev).abs_()       # one resident chunk, reused (26 GB)
rp = (torch.arange(chunk + 1, device=dev) * f).to(torch.int64)
post_op = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
def c5():
    for c in range(8):
        ops.finalize_posts(frames, row_ptr=rp, final_norm=True)   # pooled + normalised bf16 rows of chunk c
        post_op[c * chunk:(c + 1) * chunk].copy_(ops.finalize_posts(frames, row_ptr=rp, final_norm=True)[1]) if c == 0 else None
    st = ranking.device_rank_statistics(a_op, post_op, lab, d, k=100, want_auc=False)
    return ranking.aggregate(ranking.host_statistics(st, n, False), n, False)
post_op[:] = ops.finalize_posts(frames, row_ptr=rp, final_norm=True)[1].repeat(8, 1)
tf = timeit(lambda: ops.finalize_posts(frames, row_ptr=rp, final_norm=True), reps=3)
ts = timeit(lambda: ranking.aggregate(ranking.host_statistics(
    ranking.device_rank_statistics(a_op, post_op, lab, d, k=100, want_auc=False), n, False), n, False), reps=2)
print("C5s 5k brands x 400k posts: pool 32x2048 + l2norm %.2f ms per 50k posts -> %.3e posts/s (%.0f GB/s) ; "
      "score+top-100+recall/MedR/NDCG sweep %.1f ms -> %.3e pairs/s"
      % (tf * 1e3, chunk / tf, (chunk * f * d * 4 + chunk * d * 2) / tf / 1e9, ts * 1e3, nb * n / ts))
