"""torchrun --nproc-per-node N tools/gpu_sharded_check.py : N-rank NCCL sharded evaluation == single-GPU evaluation."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from fancyrec_b200 import ops, ranking, sharded

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
nb, n, d, k = [int(v) for v in (sys.argv[1].split(",") if len(sys.argv) > 1 else "300,200003,256,100".split(","))]
g = torch.Generator(device="cpu").manual_seed(5)
brand = torch.randn(nb, d, generator=g)
posts = torch.randn(n, d, generator=g)
posts[::7] = posts[::7].round()          # some exact ties
labels = (torch.randperm(n, generator=g) % (nb + 40))   # some labels out of brand range, none for a few brands
labels[labels >= nb] = nb + 5
lo, hi = sharded.shard_bounds(n, world, rank)
a = ranking.to_operand(brand.to(dev))
b_local = ranking.to_operand(posts[lo:hi].to(dev))
lab_local = labels[lo:hi].to(dev, torch.int32)
st = sharded.sharded_rank_statistics(a, b_local, lab_local, d, k, n)
hs = ranking.host_statistics(st, n, want_auc=False)
res = ranking.aggregate(hs, n, want_auc=False)
if rank == 0:
    b_full = ranking.to_operand(posts.to(dev))
    ref = ranking.device_rank_statistics(a, b_full, labels.to(dev, torch.int32), d, k=k, want_auc=False)
    hr = ranking.host_statistics(ref, n, want_auc=False)
    rres = ranking.aggregate(hr, n, want_auc=False)
    ok = (torch.equal(st["topk_index"], ref["topk_index"]) and torch.equal(st["topk_scores"], ref["topk_scores"])
          and np.array_equal(hs["first_rank"], hr["first_rank"]) and np.array_equal(hs["hits"], hr["hits"])
          and np.array_equal(hs["n_pos"], hr["n_pos"]) and tuple(map(float, res[:2] + res[3:])) == tuple(map(float, rres[:2] + rres[3:])))
    print("sharded(%d ranks) == single GPU: %s ; result %s" % (world, ok, [float(x) for x in res]))
    assert ok
# full 8-tuple incl. exact AUC across the shards
st = sharded.sharded_rank_statistics(a, b_local, lab_local, d, k, n, want_auc=True)
hs = ranking.host_statistics(st, n, want_auc=True)
res = ranking.aggregate(hs, n, want_auc=True)
if rank == 0:
    ref = ranking.device_rank_statistics(a, b_full, labels.to(dev, torch.int32), d, k=k, want_auc=True)
    hr = ranking.host_statistics(ref, n, want_auc=True)
    rres = ranking.aggregate(hr, n, want_auc=True)
    ok = (np.array_equal(hs["auc_num"], hr["auc_num"]) and np.array_equal(hs["first_rank"], hr["first_rank"])
          and tuple(map(float, res)) == tuple(map(float, rres)))
    print("sharded AUC (%d ranks) == single GPU: %s ; AUC %.6f" % (world, ok, float(res[2])))
    assert ok
# the reference-facing sharded call
import types
from fancyrec_b200 import evaluator, model as fmodel
opt = types.SimpleNamespace(brand_num=nb, common_embedding_size=d, brand_aspect=16)
ba = fmodel.BrandAspects(opt).to(dev)
with torch.no_grad():
    gg = torch.Generator(device="cpu").manual_seed(9)
    ba.brand_embeddings.weight.copy_(torch.randn(nb + 1, 16, generator=gg))
    ba.aspects_embeddings.copy_(torch.randn(16, d, generator=gg))
mdl = types.SimpleNamespace(brand_encoding=ba, opt=opt)
lab_ok = labels.clamp(max=nb - 1)
got = evaluator.test_post_ranking_sharded(nb, 'auc', mdl, posts[lo:hi].to(dev), lab_ok[lo:hi].to(dev))
if rank == 0:
    want = evaluator.test_post_ranking(nb, 'auc', mdl, posts.to(dev), lab_ok.to(dev))
    ok = tuple(map(float, got)) == tuple(map(float, want))
    print("evaluator.test_post_ranking_sharded (%d ranks) == test_post_ranking: %s ; %s" % (world, ok, [round(float(x), 6) for x in got]))
    assert ok
dist.barrier()
dist.destroy_process_group()
