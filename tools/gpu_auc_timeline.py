"""Timeline of the AUC-inclusive path (evaluator.test_post_ranking) at config-2 scale."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fancyrec_b200 import ops, ranking
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(3)
nb, n, d = 1000, 1000000, 3072
brand = torch.randn((nb, d), generator=g, device=dev)
lab = (torch.randperm(n, generator=g, device=dev) % nb).to(torch.int32)
post = torch.empty((n, d), device=dev)
bn = brand / brand.norm(dim=1, keepdim=True)
for lo in range(0, n, 65536):
    hi = min(n, lo + 65536)
    post[lo:hi] = torch.randn((hi - lo, d), generator=g, device=dev) + 0.05 * d ** 0.5 * bn[lab[lo:hi].long()]
a_op, b_op = ranking.to_operand(brand), ranking.to_operand(post)
def ev():
    x = torch.cuda.Event(enable_timing=True); x.record(); return x
def run(trace):
    marks = [("start", ev())]
    res = ops.score_topk(a_op, b_op, 64, d=d, labels=lab); marks.append(("score_topk", ev()))
    n_pos, bs, bi = ops.label_stats(lab, res["pos_score"], nb, 0); marks.append(("label_stats", ev()))
    seg_ptr, pos_sorted = ops.group_positives(lab, res["pos_score"], n_pos); marks.append(("group_positives", ev()))
    auc = torch.zeros(nb, dtype=torch.int64, device=dev); before = torch.zeros(nb, dtype=torch.int64, device=dev)
    rows = 512
    dense = torch.empty((rows, n), dtype=torch.float32, device=dev)
    for r0 in range(0, nb, rows):
        r1 = min(nb, r0 + rows)
        ops.score_dense(a_op[r0:r1], b_op, d=d, out=dense[:r1 - r0]); marks.append(("dense %d" % r0, ev()))
        ops.auc_rows(dense[:r1 - r0], r0, lab, seg_ptr, pos_sorted, bs, bi, auc, before, 0); marks.append(("auc_rows %d" % r0, ev()))
    torch.cuda.synchronize()
    if trace:
        for (na, e0), (nb_, e1) in zip(marks[:-1], marks[1:]):
            print("%-18s %.3f ms" % (nb_, e0.elapsed_time(e1)))
run(False); run(True)
