"""Full BASELINE.json config-2 size (1 000 brands x 1 000 000 posts, D = 3072) through size-independent properties:
the oracle cannot rank 1e9 pairs in seconds, but these invariants pin the fused top-k exactly:

  * sortedness / uniqueness of every list under (score desc, index asc);
  * rank consistency: the count pass (independent epilogue mode) must report exactly k-1 posts preceding the k-th entry
    and exactly r posts preceding the entry at rank r -- with sortedness this proves the list IS the exact top-k;
  * the listed scores equal the dense tile bit for bit (sampled brand rows);
  * shard / merge equivalence: top-k of two halves merged == top-k of the whole (the multi-GPU exchange, on one GPU);
  * idempotence: a second run returns identical lists.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def c2():
    from fancyrec_b200 import ops, ranking
    dev = torch.device("cuda:0")
    free, _ = torch.cuda.mem_get_info()
    if free < 40 << 30:
        pytest.skip("needs ~40 GB of free HBM")
    g = torch.Generator(device=dev).manual_seed(20261018)
    nb, n, d, k = 1000, 1000000, 3072, 100
    brand = torch.randn((nb, d), generator=g, device=dev)
    labels = (torch.randperm(n, generator=g, device=dev) % nb).to(torch.int32)
    bn = brand / brand.norm(dim=1, keepdim=True)
    post_op = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
    for lo in range(0, n, 65536):                       # operand generated chunk-wise (fp32 chunk -> normalised bf16)
        hi = min(n, lo + 65536)
        x = torch.randn((hi - lo, d), generator=g, device=dev) + 0.05 * d ** 0.5 * bn[labels[lo:hi].long()]
        x[::1000] = x[0]                                # exact duplicates -> exact ties across the index range
        post_op[lo:hi] = ranking.to_operand(x)
    brand_op = ranking.to_operand(brand)
    res = ops.score_topk(brand_op, post_op, k, d=d, labels=labels)
    torch.cuda.synchronize()
    return dict(nb=nb, n=n, d=d, k=k, brand_op=brand_op, post_op=post_op, labels=labels, res=res)


def test_lists_sorted_unique_in_range(c2):
    s, i = c2["res"]["scores"], c2["res"]["index"].long()
    assert bool(((i >= 0) & (i < c2["n"])).all())
    ds = s[:, 1:] - s[:, :-1]
    assert bool((ds <= 0).all())                                          # scores non-increasing
    tie = ds == 0
    assert bool((i[:, 1:][tie] > i[:, :-1][tie]).all())                    # ties: index ascending
    assert int(tie.sum()) > 0                                              # the fixture does contain exact ties
    srt = torch.sort(i, dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())                         # no post twice in a list


@pytest.mark.parametrize("rank", [99, 50, 0])
def test_rank_consistency_via_count_pass(c2, rank):
    from fancyrec_b200 import ops
    res = c2["res"]
    thr_s = res["scores"][:, rank].contiguous()
    thr_i = res["index"][:, rank].contiguous()
    cnt = ops.score_count(c2["brand_op"], c2["post_op"], thr_s, thr_i, d=c2["d"])
    assert bool((cnt == rank).all()), cnt[:8]


def test_scores_equal_dense_tile_and_positive_scores(c2):
    from fancyrec_b200 import ops
    rows = slice(384, 512)
    dense = ops.score_dense(c2["brand_op"][rows], c2["post_op"], d=c2["d"])
    got = c2["res"]["scores"][rows]
    want = torch.gather(dense, 1, c2["res"]["index"][rows].long())
    assert torch.equal(got, want)
    # independent check of the selection on these rows: torch.topk values of the dense tile
    assert torch.equal(got, torch.topk(dense, c2["k"], dim=1).values)
    lab = c2["labels"].long()
    sel = (lab >= 384) & (lab < 512)
    cols = sel.nonzero().flatten()
    assert torch.equal(c2["res"]["pos_score"][cols], dense[lab[cols] - 384, cols])


def test_shard_merge_equals_whole_and_idempotent(c2):
    from fancyrec_b200 import ops
    half = c2["n"] // 2 + 12345
    a = ops.score_topk(c2["brand_op"], c2["post_op"][:half], c2["k"], d=c2["d"])
    b = ops.score_topk(c2["brand_op"], c2["post_op"][half:], c2["k"], d=c2["d"], index_base=half)
    ms, mi = ops.topk_merge(torch.stack([a["scores"], b["scores"]]), torch.stack([a["index"], b["index"]]), c2["k"])
    assert torch.equal(mi, c2["res"]["index"]) and torch.equal(ms, c2["res"]["scores"])
    again = ops.score_topk(c2["brand_op"], c2["post_op"], c2["k"], d=c2["d"], labels=c2["labels"])
    assert torch.equal(again["index"], c2["res"]["index"]) and torch.equal(again["scores"], c2["res"]["scores"])
    assert torch.equal(again["pos_score"], c2["res"]["pos_score"])


def test_metrics_pipeline_matches_properties(c2):
    """First-positive ranks from the list == count-pass ranks of the best positive (two independent routes)."""
    from fancyrec_b200 import ops
    res, lab = c2["res"], c2["labels"]
    n_pos, best_s, best_i = ops.label_stats(lab, res["pos_score"], c2["nb"])
    hit, first = ops.rank_from_topk(res["index"], lab)
    cnt = ops.score_count(c2["brand_op"], c2["post_op"], best_s, best_i, d=c2["d"])
    inside = first >= 0
    assert bool((n_pos == c2["n"] // c2["nb"]).all())
    assert bool((cnt[inside] == first[inside].long()).all())
    assert bool((cnt[~inside] >= c2["k"]).all())


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[3], one GPU's shard of the 8-GPU job: ALL 10 000 brands x 2 500 000 posts, D = 3072, k = 1000
# ---------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def c4():
    from fancyrec_b200 import ops, ranking
    dev = torch.device("cuda:0")
    torch.cuda.empty_cache()
    free, _ = torch.cuda.mem_get_info()
    if free < 60 << 30:
        pytest.skip("needs ~60 GB of free HBM")
    g = torch.Generator(device=dev).manual_seed(20261018 + 4)
    nb, n, d, k, base = 10000, 2500000, 3072, 1000, 5000000        # the shard of rank 2: global indices start at 5 M
    brand = torch.randn((nb, d), generator=g, device=dev)
    bn = brand / brand.norm(dim=1, keepdim=True)
    labels = ((torch.arange(base, base + n, device=dev) * 7919) % nb).to(torch.int32)
    post_op = torch.empty((n, d), dtype=torch.bfloat16, device=dev)
    for lo in range(0, n, 65536):
        hi = min(n, lo + 65536)
        x = torch.randn((hi - lo, d), generator=g, device=dev) + 0.05 * d ** 0.5 * bn[labels[lo:hi].long()]
        x[::997] = x[0]                                 # exact duplicates -> exact ties across the index range
        post_op[lo:hi] = ranking.to_operand(x)
    brand_op = ranking.to_operand(brand)
    res = ops.score_topk(brand_op, post_op, k, d=d, labels=labels, index_base=base)
    torch.cuda.synchronize()
    yield dict(nb=nb, n=n, d=d, k=k, base=base, brand_op=brand_op, post_op=post_op, labels=labels, res=res)
    torch.cuda.empty_cache()


def test_c4_shard_lists_sorted_unique_in_range(c4):
    s, i = c4["res"]["scores"], c4["res"]["index"].long()
    assert bool(((i >= c4["base"]) & (i < c4["base"] + c4["n"])).all())
    ds = s[:, 1:] - s[:, :-1]
    assert bool((ds <= 0).all())
    tie = ds == 0
    assert bool((i[:, 1:][tie] > i[:, :-1][tie]).all())
    assert int(tie.sum()) > 0
    srt = torch.sort(i, dim=1).values
    assert bool((srt[:, 1:] != srt[:, :-1]).all())


@pytest.mark.parametrize("rank", [999, 500, 0])
def test_c4_shard_rank_consistency_via_count_pass(c4, rank):
    """Independent epilogue mode: exactly `rank` posts precede the entry at rank `rank`, for all 10 000 brands."""
    from fancyrec_b200 import ops
    res = c4["res"]
    cnt = ops.score_count(c4["brand_op"], c4["post_op"], res["scores"][:, rank].contiguous(),
                          res["index"][:, rank].contiguous(), d=c4["d"], index_base=c4["base"])
    assert bool((cnt == rank).all()), cnt[:8]


def test_c4_shard_slab_equals_dense_tile_and_torch_topk(c4):
    from fancyrec_b200 import ops
    for r0 in (0, 9872):                                 # the first and the last (ragged: 10 000 = 78 * 128 + 16) m-tile
        rows = slice(r0, min(r0 + 128, c4["nb"]))
        dense = ops.score_dense(c4["brand_op"][rows].contiguous(), c4["post_op"], d=c4["d"])
        got_s, got_i = c4["res"]["scores"][rows], c4["res"]["index"][rows].long() - c4["base"]
        assert torch.equal(got_s, torch.gather(dense, 1, got_i))
        assert torch.equal(got_s, torch.topk(dense, c4["k"], dim=1).values)
        lab = c4["labels"].long()
        cols = ((lab >= rows.start) & (lab < rows.stop)).nonzero().flatten()
        assert torch.equal(c4["res"]["pos_score"][cols], dense[lab[cols] - rows.start, cols])
        del dense


def test_c4_shard_statistics_and_merge_of_two_half_shards(c4):
    """The exchange step on one GPU: two half shards merged == the shard; first-positive ranks from the list agree with
    the count pass of the best positive."""
    from fancyrec_b200 import ops
    half = c4["n"] // 2 + 4321
    a = ops.score_topk(c4["brand_op"], c4["post_op"][:half], c4["k"], d=c4["d"], index_base=c4["base"])
    b = ops.score_topk(c4["brand_op"], c4["post_op"][half:], c4["k"], d=c4["d"], index_base=c4["base"] + half)
    ms, mi = ops.topk_merge(torch.stack([a["scores"], b["scores"]]), torch.stack([a["index"], b["index"]]), c4["k"])
    assert torch.equal(mi, c4["res"]["index"]) and torch.equal(ms, c4["res"]["scores"])
    del a, b
    res, lab = c4["res"], c4["labels"]
    n_pos, best_s, best_i = ops.label_stats(lab, res["pos_score"], c4["nb"], c4["base"])
    assert bool((n_pos == c4["n"] // c4["nb"]).all())
    hit, first = ops.rank_from_topk(res["index"], lab, c4["base"])
    cnt = ops.score_count(c4["brand_op"], c4["post_op"], best_s, best_i, d=c4["d"], index_base=c4["base"])
    inside = first >= 0
    assert bool((cnt[inside] == first[inside].long()).all())
    assert bool((cnt[~inside] >= c4["k"]).all())


# ---------------------------------------------------------------------------------------------
# BASELINE.json configs[4] shape: 32 frames x 2048-d per post pooled from feature.bin-layout rows, then the full sweep
# ---------------------------------------------------------------------------------------------
def test_c5_pooled_posts_through_rank_posts_match_the_oracle():
    """50 000 video posts x 32 frames x 2048 fp32 (13 GB of frame rows) -> mean-pool + L2 norm on the device ->
    rank_posts against 1 152 brands: the 8-tuple (recall@1/5/10, MedR, MeanR, NDCG@10/50, AUC) and every integer
    statistic equal the oracle applied to our own scores, bit for bit; pooled rows match torch within fp32 tolerance."""
    from fancyrec_b200 import ops, ranking
    from oracle import ranking as oref
    dev = torch.device("cuda:0")
    torch.cuda.empty_cache()
    g = torch.Generator(device=dev).manual_seed(20261018 + 5)
    n, f, dv, nb = 50000, 32, 2048, 1152
    frames = torch.empty((n * f, dv), device=dev)
    brand = torch.randn((nb, dv), generator=g, device=dev)
    bn = brand / brand.norm(dim=1, keepdim=True)
    labels = (torch.randperm(n, generator=g, device=dev) % (nb + 5)).to(torch.int32)     # 5 label values without a brand row
    for lo in range(0, n, 2048):
        hi = min(n, lo + 2048)
        x = torch.relu(torch.randn(((hi - lo) * f, dv), generator=g, device=dev) * 0.5 + 0.3)
        x += 0.08 * bn[(labels[lo:hi].long() % nb).repeat_interleave(f)].abs()
        frames[lo * f:hi * f] = x
    row_ptr = torch.arange(n + 1, device=dev, dtype=torch.int64) * f
    pooled = ops.finalize_posts(frames, row_ptr=row_ptr, final_norm=True, want_f32=True, want_bf16=False)[0]
    want = frames[:64 * f].view(64, f, dv).double().mean(1)
    want = (want / want.norm(dim=1, keepdim=True)).float()
    assert torch.allclose(pooled[:64], want, rtol=4e-6, atol=1e-7)
    del frames
    result, stats, dev_stats = ranking.rank_posts(brand, pooled, labels, k=100, want_auc=True)
    ours = ops.score_dense(ranking.to_operand(brand), ranking.to_operand(pooled), d=dv).cpu().numpy()
    lab = labels.cpu().numpy()
    ost = oref.rank_stats(ours, lab)
    assert np.array_equal(stats["n_pos"], ost["n_pos"])
    assert np.array_equal(stats["first_rank"], ost["first_rank"])
    assert np.array_equal(stats["auc_num"], ost["auc_num"])
    assert np.array_equal(stats["hits"], ost["hits"])
    assert tuple(map(float, result)) == tuple(map(float, oref.aggregate(ost, n)))
    assert np.array_equal(dev_stats["topk_index"].cpu().numpy(), oref.topk_indices(ours, 100))
