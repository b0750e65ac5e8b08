"""GPU parity for the round-2 evaluation paths:

  * the fused top-k pass that also writes the score matrix (same accumulators as frx_score_dense, bit for bit);
  * the exact-AUC sweep (`frx_auc_rows`, monotone bucket table) against the oracle on adversarial score layouts:
    heavy ties, positives 1 ulp apart with a tied negative, all positives equal, clustered positives with an outlier,
    more positives than one shared-memory chunk, NaN scores;
  * the fused route and the row-chunked route give identical numerators; the public drop-in call at BASELINE.json's
    config-2 size (1 000 x 1 000 000, D = 3072) against independent torch arithmetic on sampled brands;
  * EvalPipeline (side-stream finalisation, asynchronous D2H, deferred host aggregation) == the synchronous call.
"""
import types

import numpy as np
import pytest
import torch

from oracle import ranking as oref
from oracle import synth
from tests.gpu_util import dev, to_dev

pytestmark = pytest.mark.gpu


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("nb,n,d,k", [(5, 300, 64, 7), (130, 1000, 96, 64), (257, 70001, 128, 100), (1000, 300000, 64, 100)])
def test_topk_pass_writes_the_dense_tile_bit_for_bit(nb, n, d, k):
    from fancyrec_b200 import ops, ranking
    g = torch.Generator(device=dev()).manual_seed(nb * 7 + n)
    brand_op = ranking.to_operand(torch.randn((nb, d), generator=g, device=dev()))
    post_op = ranking.to_operand(torch.randn((n, d), generator=g, device=dev()))
    labels = (torch.randperm(n, generator=g, device=dev()) % nb).to(torch.int32)
    res = ops.score_topk(brand_op, post_op, k, d=d, labels=labels, dense=True)
    plain = ops.score_topk(brand_op, post_op, k, d=d, labels=labels)
    dense = ops.score_dense(brand_op, post_op, d=d)
    assert torch.equal(res["dense"], dense)
    assert torch.equal(res["index"], plain["index"]) and torch.equal(res["scores"], plain["scores"])
    assert torch.equal(res["pos_score"], plain["pos_score"])
    kk = min(k, n)
    assert np.array_equal(res["index"].cpu().numpy()[:, :kk], oref.topk_indices(dense.cpu().numpy(), kk))


# ---------------------------------------------------------------------------------------------
def _auc_device(scores_np, labels_np, nb, index_base=0):
    """frx_group_positives + frx_auc_rows on a given score matrix -> (auc_num, before_first)."""
    from fancyrec_b200 import ops
    scores = to_dev(scores_np.astype(np.float32))
    labels = to_dev(labels_np.astype(np.int32))
    n = scores.shape[1]
    lab_l = labels.long()
    pos_score = torch.full((n,), float("nan"), device=dev())
    ok = (lab_l >= 0) & (lab_l < nb)
    cols = ok.nonzero().flatten()
    pos_score[cols] = scores[lab_l[cols], cols]
    n_pos, best_s, best_i = ops.label_stats(labels, pos_score, nb, index_base)
    seg_ptr, pos_sorted = ops.group_positives(labels, pos_score, n_pos)
    auc = torch.zeros(nb, dtype=torch.int64, device=dev())
    before = torch.zeros(nb, dtype=torch.int64, device=dev())
    ops.auc_rows(scores, 0, labels, seg_ptr, pos_sorted, best_s, best_i, auc, before, index_base)
    return auc.cpu().numpy(), before.cpu().numpy()


def _auc_oracle(scores, labels, nb):
    """evaluator.py:111-113 literally: sum over positives e of #{negatives el : e > el} (NaN compares False)."""
    out = np.zeros(nb, dtype=np.int64)
    for b in range(nb):
        pos = scores[b][labels == b]
        neg = scores[b][labels != b]
        neg = np.sort(neg[~np.isnan(neg)])
        out[b] = int(np.searchsorted(neg, pos[~np.isnan(pos)], side="left").sum())
    return out


def _first_rank_oracle(scores, labels, nb):
    out = np.zeros(nb, dtype=np.int64)
    for b in range(nb):
        if (labels == b).any():
            order = oref.order_desc(scores[b])
            out[b] = int(np.argmax(labels[order] == b))
    return out


def _check_auc(scores, labels, nb):
    got, before = _auc_device(scores, labels, nb)
    assert np.array_equal(got, _auc_oracle(scores.astype(np.float32), labels, nb))
    has = np.bincount(labels[(labels >= 0) & (labels < nb)], minlength=nb) > 0
    if not np.isnan(scores).any():
        assert np.array_equal(before[has], _first_rank_oracle(scores.astype(np.float32), labels, nb)[has])


def test_auc_rows_random_and_ragged():
    rs = np.random.RandomState(1)
    for nb, n in [(3, 40), (17, 5000), (64, 70003)]:
        scores = rs.standard_normal((nb, n)).astype(np.float32) * 0.02
        labels = rs.randint(0, nb + 2, size=n)            # labels nb, nb+1: posts of brands outside the table
        labels[labels == 1] = 0                           # brand 1 has no positive
        _check_auc(scores, labels, nb)


def test_auc_rows_heavy_ties_lattice():
    rs = np.random.RandomState(2)
    nb, n = 9, 20000
    scores = (rs.randint(-40, 41, size=(nb, n)) / 1024.0).astype(np.float32)      # 81 distinct values
    labels = rs.randint(0, nb, size=n)
    _check_auc(scores, labels, nb)


def test_auc_rows_positives_one_ulp_apart_with_tied_negatives():
    """Near-tied top positives and negatives tied with them (the round-1 table could over-count here)."""
    rs = np.random.RandomState(3)
    nb, n = 4, 9000
    scores = rs.standard_normal((nb, n)).astype(np.float32) * 0.01
    labels = rs.randint(0, nb, size=n)
    top = np.float32(0.75)
    ladder = top + np.arange(-4, 5) * np.spacing(top)       # 9 consecutive floats around 0.75
    for b in range(nb):
        pos = np.where(labels == b)[0]
        neg = np.where(labels != b)[0]
        scores[b, pos[:9]] = ladder
        scores[b, neg[:27]] = np.tile(ladder, 3)             # negatives exactly tied with every rung
        scores[b, neg[27:30]] = [np.nextafter(ladder[-1], np.float32(2)), ladder[0], np.nextafter(ladder[0], np.float32(-2))]
    _check_auc(scores, labels, nb)


def test_auc_rows_all_positives_equal_and_single_positive():
    rs = np.random.RandomState(4)
    nb, n = 5, 3000
    scores = rs.standard_normal((nb, n)).astype(np.float32)
    labels = rs.randint(0, nb, size=n)
    scores[0, labels == 0] = np.float32(0.125)                # all positives of brand 0 tie
    scores[0, np.where(labels != 0)[0][:50]] = np.float32(0.125)
    only = np.where(labels == 1)[0]
    labels[only[1:]] = 2                                       # brand 1 keeps a single positive
    _check_auc(scores, labels, nb)


def test_auc_rows_clustered_positives_with_outlier_and_many_positives():
    """Thousands of positives inside one bucket (binary-search path) next to an outlier that stretches the bucket range,
    and a brand with more positives than one shared-memory chunk (4 000)."""
    rs = np.random.RandomState(5)
    nb, n = 3, 40000
    scores = rs.standard_normal((nb, n)).astype(np.float32) * 0.05
    labels = np.where(rs.rand(n) < 0.5, 0, rs.randint(1, nb, size=n))     # brand 0 owns ~20 000 posts (5 chunks)
    p1 = np.where(labels == 1)[0]
    scores[1, p1] = np.float32(0.3) + rs.randint(0, 2000, size=len(p1)).astype(np.float32) * np.spacing(np.float32(0.3))
    scores[1, p1[0]] = np.float32(0.9)                        # outlier: every other positive shares one bucket
    n1 = np.where(labels != 1)[0]
    scores[1, n1[:4000]] = np.float32(0.3) + rs.randint(-50, 2050, size=4000).astype(np.float32) * np.spacing(np.float32(0.3))
    _check_auc(scores, labels, nb)


def test_auc_rows_nan_scores_count_as_the_reference_counts_them():
    rs = np.random.RandomState(6)
    nb, n = 3, 2000
    scores = rs.standard_normal((nb, n)).astype(np.float32)
    labels = rs.randint(0, nb, size=n)
    for b in range(nb):
        scores[b, np.where(labels != b)[0][:25]] = np.nan     # `e > nan` is False: these negatives add nothing
    got, _ = _auc_device(scores, labels, nb)
    assert np.array_equal(got, _auc_oracle(scores, labels, nb))


# ---------------------------------------------------------------------------------------------
def test_fused_and_chunked_auc_routes_agree(monkeypatch):
    from fancyrec_b200 import ranking
    g = torch.Generator(device=dev()).manual_seed(77)
    nb, n, d = 300, 50000, 192
    brand = torch.randn((nb, d), generator=g, device=dev())
    labels = (torch.randperm(n, generator=g, device=dev()) % (nb + 3)).to(torch.int32)
    posts = torch.randn((n, d), generator=g, device=dev()) + 0.3 * brand[labels.long() % nb]
    fused = ranking.rank_posts(brand, posts, labels, want_auc=True)
    monkeypatch.setattr(ranking, "FUSED_DENSE_BUDGET_BYTES", 0)
    monkeypatch.setattr(ranking, "DENSE_BUDGET_BYTES", 4 * n * 128)        # 3 row chunks
    chunked = ranking.rank_posts(brand, posts, labels, want_auc=True)
    assert np.array_equal(fused[1]["auc_num"], chunked[1]["auc_num"])
    assert np.array_equal(fused[1]["first_rank"], chunked[1]["first_rank"])
    assert tuple(map(float, fused[0])) == tuple(map(float, chunked[0]))
    ours = ranking.to_operand(brand), ranking.to_operand(posts)
    from fancyrec_b200 import ops
    dense = ops.score_dense(ours[0], ours[1], d=d).cpu().numpy()
    want = oref.rank_stats(dense, labels.cpu().numpy())
    assert np.array_equal(fused[1]["auc_num"], want["auc_num"])
    assert np.array_equal(fused[1]["first_rank"], want["first_rank"])


def test_public_call_with_auc_at_config2_size():
    """evaluator.test_post_ranking (AUC always on, evaluator.py:103) at 1 000 brands x 1 000 000 posts, D = 3072: the
    AUC numerators of sampled brands equal independent torch arithmetic on the dense tile, and the 8-tuple equals the
    aggregation of those statistics."""
    from fancyrec_b200 import evaluator, model, ops, ranking
    free, _ = torch.cuda.mem_get_info()
    if free < 60 << 30:
        pytest.skip("needs ~60 GB of free HBM")
    g = torch.Generator(device=dev()).manual_seed(20261018)
    nb, n, d, a = 1000, 1000000, 3072, 64
    opt = types.SimpleNamespace(brand_num=nb, common_embedding_size=d, brand_aspect=a)
    enc = model.BrandAspects(opt).to(dev())
    mdl = types.SimpleNamespace(brand_encoding=enc, opt=opt)
    brand = evaluator.brand_matrix(mdl, nb)
    labels = (torch.randperm(n, generator=g, device=dev()) % nb).to(torch.int64)
    posts = torch.empty((n, d), device=dev())
    bn = brand / brand.norm(dim=1, keepdim=True)
    for lo in range(0, n, 65536):
        hi = min(n, lo + 65536)
        posts[lo:hi] = torch.randn((hi - lo, d), generator=g, device=dev()) + 0.04 * d ** 0.5 * bn[labels[lo:hi]]
    result = evaluator.test_post_ranking(nb, "auc", mdl, posts, labels)
    res2, stats, dev_stats = ranking.rank_posts(brand, posts, labels, want_auc=True)
    assert tuple(map(float, result)) == tuple(map(float, res2))
    assert 0.5 < float(result[2]) < 1.0
    brand_op, post_op = ranking.to_operand(brand), ranking.to_operand(posts)
    del posts
    rows = [0, 1, 127, 128, 511, 640, 998, 999]
    lab32 = labels.to(torch.int32)
    for b in rows:
        row = ops.score_dense(brand_op[b:b + 1].contiguous(), post_op, d=d)[0]
        is_pos = lab32 == b
        neg_sorted = torch.sort(row[~is_pos]).values
        num = int(torch.searchsorted(neg_sorted, row[is_pos].contiguous(), right=False).sum())
        assert num == int(stats["auc_num"][b]), b
        order = torch.sort(row, descending=True, stable=True).indices
        assert int(is_pos[order].nonzero()[0]) == int(stats["first_rank"][b]), b


# ---------------------------------------------------------------------------------------------
@pytest.mark.parametrize("want_auc", [False, True])
@pytest.mark.parametrize("overlap", [False, True])
def test_pipeline_equals_synchronous_call(want_auc, overlap):
    from fancyrec_b200 import ops, pipeline, ranking
    g = torch.Generator(device=dev()).manual_seed(5 + want_auc)
    nb, n, dv, dt, a, k = 300, 90000, 1024, 512, 40, 100
    w = torch.randn((nb + 1, a), generator=g, device=dev())
    e = torch.randn((a, dv + dt), generator=g, device=dev())
    batches = []
    for t in range(5):
        labels = (torch.randperm(n, generator=g, device=dev()) % nb).to(torch.int32)
        visual = torch.randn((n, dv), generator=g, device=dev())
        text = torch.randn((n, dt), generator=g, device=dev())
        batches.append((visual, text, labels))
    pipe = pipeline.EvalPipeline(dev(), nb, n, dv, dt, k=k, want_auc=want_auc, overlap=overlap)
    tickets = [pipe.submit(w, e, v, t, lab) for (v, t, lab) in batches]       # slots are recycled: depth 2, five batches
    got = [pipe.result(tk) for tk in tickets]
    brand_op = ops.finalize_posts(ops.brand_embed(w, e, nb=nb), final_norm=True)[1]
    for (v, t, lab), res in zip(batches, got):
        post_op = ops.finalize_posts(v, t, visual_norm=True, text_norm=True, final_norm=True)[1]     # the same single pass
        st = ranking.device_rank_statistics(brand_op, post_op, lab, dv + dt, k=k, want_auc=want_auc)
        want = ranking.aggregate(ranking.host_statistics(st, n, want_auc), n, want_auc)
        a_, b_ = tuple(map(float, res)), tuple(map(float, want))
        assert all((x == y) or (x != x and y != y) for x, y in zip(a_, b_)), (a_, b_)
