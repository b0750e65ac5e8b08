"""GPU parity: fused score + top-k GEMM, dense tile, count pass, merge -- through the C ABI.

Protocol: (2) scores vs fp32 cal_sim within 1e-3 on the cosine scale (bf16 operands,
fp32 accumulation); (3) rankings / indices bit-exact against the oracle applied to OUR score tile;
(4) exact-lattice inputs: bit-exact against the pure fp32 reference scores.
"""
import numpy as np
import pytest
import torch

from oracle import ranking as oref
from oracle import synth
from tests.gpu_util import dev, operand, score_atol, to_dev

pytestmark = pytest.mark.gpu



def _dense_and_topk(brand, posts, k, labels=None, index_base=0):
    from fancyrec_b200 import ops
    a, b = operand(brand), operand(posts)
    d = brand.shape[1]
    lab = to_dev(labels.astype(np.int32)) if labels is not None else None
    res = ops.score_topk(a, b, k, d=d, labels=lab, index_base=index_base)
    dense = ops.score_dense(a, b, d=d)
    torch.cuda.synchronize()
    return res, dense.cpu().numpy(), a, b


@pytest.mark.parametrize("nb,npost,d,k", [
    (5, 257, 48, 10), (1, 37, 64, 64), (50, 10000, 1024, 64), (130, 3000, 200, 100),
    (300, 70001, 256, 100), (64, 5000, 3072, 1000), (129, 513, 72, 1), (1000, 20000, 128, 100),
    (50, 10000, 320, 64),     # 80 candidate lists x 64: merge smem lands between 32 and 48 KB
])
def test_topk_matches_oracle_on_our_scores(nb, npost, d, k):
    rs = np.random.RandomState(nb * 7 + npost)
    brand = rs.standard_normal((nb, d)).astype(np.float32)
    posts = rs.standard_normal((npost, d)).astype(np.float32)
    res, dense, _, _ = _dense_and_topk(brand, posts, k)
    ref = oref.cal_sim(brand, posts)
    assert np.abs(dense - ref).max() <= score_atol(d)
    kk = min(k, npost)
    want_idx = oref.topk_indices(dense, k)
    got_idx = res["index"].cpu().numpy()
    got_s = res["scores"].cpu().numpy()
    assert np.array_equal(got_idx[:, :kk], want_idx)
    assert np.array_equal(got_s[:, :kk], np.take_along_axis(dense, want_idx, 1))
    if kk < k:   # padding
        assert (got_idx[:, kk:] == -1).all() and np.isneginf(got_s[:, kk:]).all()


@pytest.mark.parametrize("name", ["lattice", "lattice_small_k"])
def test_lattice_scores_bit_exact_vs_fp32_reference(golden_dir, name):
    import os
    g = np.load(os.path.join(golden_dir, "ranking_%s.npz" % name))
    nb, lab, w, e, posts = synth.ranking_inputs(name)
    res, dense, _, _ = _dense_and_topk(w[:nb], posts, 64, labels=lab)
    assert np.array_equal(dense, g["scores"])            # identical in bf16 and fp32, any order
    kk = min(64, posts.shape[0])
    assert np.array_equal(res["index"].cpu().numpy()[:, :kk], oref.topk_indices(g["scores"], 64))
    ps = res["pos_score"].cpu().numpy()
    assert np.array_equal(ps, g["scores"][lab, np.arange(len(lab))])


@pytest.mark.parametrize("kind,k", [("gauss", 100), ("lattice", 64), ("gauss", 1000)])
def test_sample_seeded_thresholds_large_np(kind, k):
    """n_posts >= 262144 takes the sample pass (strided TMA view) that seeds the per-row thresholds;
    the result must still be the exact top-k."""
    nb, npost, d = 40, 300000, 64
    rs = np.random.RandomState(k)
    if kind == "lattice":
        brand = synth.lattice(11, nb, d, nnz=16)
        posts = synth.lattice(12, 3000, d, nnz=16)[rs.randint(0, 3000, npost)]    # massive exact ties
    else:
        brand = rs.standard_normal((nb, d)).astype(np.float32)
        posts = rs.standard_normal((npost, d)).astype(np.float32)
    lab = synth.labels(13, npost, nb)
    res, dense, _, _ = _dense_and_topk(brand, posts, k, labels=lab, index_base=7)
    want = oref.topk_indices(dense, k)
    assert np.array_equal(res["index"].cpu().numpy(), want + 7)
    assert np.array_equal(res["scores"].cpu().numpy(), np.take_along_axis(dense, want, 1))
    assert np.array_equal(res["pos_score"].cpu().numpy(), dense[lab, np.arange(npost)])


@pytest.mark.parametrize("nb,npost,d,k", [(130, 3000, 1024, 100), (40, 300000, 64, 64), (7, 513, 52, 10)])
def test_tf32_path(nb, npost, d, k):
    """fp32 operands through tcgen05.mma.kind::tf32: tighter score tolerance, same exact ranking contract."""
    from fancyrec_b200 import ops, ranking
    rs = np.random.RandomState(nb + d)
    brand = rs.standard_normal((nb, d)).astype(np.float32)
    posts = rs.standard_normal((npost, d)).astype(np.float32)
    lab = synth.labels(3, npost, nb)
    a = ranking.to_operand(to_dev(brand), precision="tf32")
    b = ranking.to_operand(to_dev(posts), precision="tf32")
    assert a.dtype == torch.float32
    dense = ops.score_dense(a, b, d=d).cpu().numpy()
    ref = oref.cal_sim(brand, posts)
    # tf32 keeps 10 mantissa bits (operands truncated by the tensor core): <= 2 * 2^-10 worst case on the
    # cosine scale; stated tolerance 2e-4 at D >= 1024 (observed 1.1e-4 at D = 1024: per-product relative
    # error ~4e-4 averaged over D terms).  north_star's 1e-5 is not reachable by a single tf32 pass: that is
    # what precision "tf32x3" is for (next test).
    err = float(np.abs(dense - ref).max())
    print("tf32 max |score - fp32 cal_sim| at D=%d: %.3e" % (d, err))
    assert err <= (2e-4 if d >= 1024 else 2.0 ** -9)
    res = ops.score_topk(a, b, k, d=d, labels=to_dev(lab.astype(np.int32)), index_base=5)
    want = oref.topk_indices(dense, k)
    assert np.array_equal(res["index"].cpu().numpy(), want + 5)
    assert np.array_equal(res["scores"].cpu().numpy(), np.take_along_axis(dense, want, 1))
    assert np.array_equal(res["pos_score"].cpu().numpy(), dense[lab, np.arange(npost)])
    tj = rs.randint(0, npost, nb)
    cnt = ops.score_count(a, b, to_dev(dense[np.arange(nb), tj].copy()), to_dev((tj + 5).astype(np.int32)), d=d,
                          index_base=5).cpu().numpy()
    for r in range(0, nb, max(1, nb // 8)):
        assert cnt[r] == int(np.where(oref.order_desc(dense[r]) == tj[r])[0][0])


@pytest.mark.parametrize("nb,npost,d,k", [(130, 3000, 1024, 100), (64, 2000, 3072, 64), (40, 300000, 64, 64),
                                          (7, 513, 52, 10)])
def test_tf32x3_path_meets_1e5(nb, npost, d, k):
    """3xTF32 split operands (K = 3D) through the same tf32 kernels: scores within north_star's 1e-5 of
    fp32 cal_sim (relative, on the cosine scale |s| <= 1: |ours - ref| <= 1e-5), rankings exact on our scores."""
    from fancyrec_b200 import ops, ranking
    rs = np.random.RandomState(nb + d + 1)
    brand = rs.standard_normal((nb, d)).astype(np.float32)
    posts = rs.standard_normal((npost, d)).astype(np.float32)
    lab = synth.labels(3, npost, nb)
    a = ranking.to_operand(to_dev(brand), precision="tf32x3", side=ranking.BRAND_SIDE)
    b = ranking.to_operand(to_dev(posts), precision="tf32x3", side=ranking.POST_SIDE)
    assert a.shape == (nb, 3 * d) and b.shape == (npost, 3 * d) and a.dtype == torch.float32
    kd = ranking.contraction_depth(d, "tf32x3")
    dense = ops.score_dense(a, b, d=kd).cpu().numpy()
    ref = oref.cal_sim(brand, posts)                       # fp64-accumulated cosine, the oracle of evaluator.py:23-29
    err = float(np.abs(dense - ref).max())
    print("tf32x3 max |score - cal_sim| at D=%d: %.3e" % (d, err))
    assert err <= 1e-5
    res = ops.score_topk(a, b, k, d=kd, labels=to_dev(lab.astype(np.int32)), index_base=5)
    want = oref.topk_indices(dense, k)
    assert np.array_equal(res["index"].cpu().numpy(), want + 5)
    assert np.array_equal(res["scores"].cpu().numpy(), np.take_along_axis(dense, want, 1))
    assert np.array_equal(res["pos_score"].cpu().numpy(), dense[lab, np.arange(npost)])


def test_tf32x3_through_public_api(monkeypatch):
    """ranking.PRECISION = 'tf32x3' switches cal_sim / rank_posts to the fp32-grade path."""
    from fancyrec_b200 import evaluator, ranking
    monkeypatch.setattr(ranking, "PRECISION", "tf32x3")
    rs = np.random.RandomState(5)
    nb, npost, d = 50, 4000, 1024
    brand = rs.standard_normal((nb, d)).astype(np.float32)
    posts = rs.standard_normal((npost, d)).astype(np.float32)
    lab = synth.labels(9, npost, nb)
    ours = evaluator.cal_sim(to_dev(brand), to_dev(posts)).cpu().numpy()
    assert np.abs(ours - oref.cal_sim(brand, posts)).max() <= 1e-5
    result, stats, _ = ranking.rank_posts(to_dev(brand), to_dev(posts), to_dev(lab), want_auc=True)
    want = oref.rank_metrics_vec(ours, lab)
    assert tuple(map(float, result)) == tuple(map(float, want))
    assert np.array_equal(stats["auc_num"], oref.rank_stats(ours, lab)["auc_num"])


@pytest.mark.parametrize("kind", ["ascending", "descending", "constant", "two_level", "negative", "sparse_nonneg"])
@pytest.mark.parametrize("k", [64, 1000])
def test_adversarial_orders_large_np(kind, k):
    """Worst cases for the threshold machinery (sample-seeded thresholds + the shared per-row candidate histogram):
    scores that keep improving along the post axis (every tile beats everything before it), scores that only get
    worse, one constant score (every post ties), two score levels, all-negative scores, and sparse non-negative
    rows whose sample threshold is exactly 0.  The result must still be the exact top-k of our own score tile."""
    nb, npost, d = 33, 400000, 64
    rs = np.random.RandomState(len(kind) + k)
    brand = np.zeros((nb, d), np.float32)
    brand[:, 0] = 1.0
    brand[:, 1:] = 0.05 * rs.standard_normal((nb, d - 1))
    posts = 0.05 * rs.standard_normal((npost, d)).astype(np.float32)
    ramp = np.linspace(0.1, 3.0, npost, dtype=np.float32)
    if kind == "ascending":
        posts[:, 0] = ramp
    elif kind == "descending":
        posts[:, 0] = ramp[::-1]
    elif kind == "constant":
        posts[:] = 0.0
        posts[:, 0] = 1.0
        brand[:, 1:] = 0.0
    elif kind == "two_level":
        posts[:] = 0.0
        posts[:, 0] = 1.0
        posts[::1000, 1] = 1.0                                       # every 1000th post scores lower
        brand[:, 1:] = 0.0
    elif kind == "negative":
        posts[:, 0] = -ramp
    else:
        posts = np.abs(posts) * (rs.random_sample((npost, d)) < 0.02)    # most dot products are exactly 0
        posts[:, 1] += 1e-3                                              # no zero rows
        brand = np.abs(brand)
    lab = synth.labels(17, npost, nb)
    res, dense, _, _ = _dense_and_topk(brand, posts.astype(np.float32), k, labels=lab, index_base=3)
    want = oref.topk_indices(dense, k)
    assert np.array_equal(res["index"].cpu().numpy(), want + 3)
    assert np.array_equal(res["scores"].cpu().numpy(), np.take_along_axis(dense, want, 1))
    assert np.array_equal(res["pos_score"].cpu().numpy(), dense[lab, np.arange(npost)])


@pytest.fixture
def cta_pairs():
    """Run the body on the CTA-pair (tcgen05 cta_group::2) variant of the score kernel."""
    from fancyrec_b200 import _lib
    lib = _lib.load()
    prev = lib.frx_set_cta_pairs(1)
    yield
    lib.frx_set_cta_pairs(prev)


@pytest.mark.parametrize("nb,npost,d,k", [
    (5, 257, 48, 10), (130, 3000, 200, 100), (300, 70001, 256, 100), (64, 5000, 3072, 1000), (129, 513, 72, 1),
    (1000, 20000, 128, 100), (385, 300000, 64, 64),      # odd m-tile counts leave the second CTA of a pair without rows
])
def test_cta_pair_variant_bit_identical(cta_pairs, nb, npost, d, k):
    """The cta_group::2 variant must give bit for bit what the single-CTA kernel gives: dense tile, fused top-k
    (incl. the sample-seeded / histogram-refined path at 300 k posts), positives' scores and the count pass."""
    from fancyrec_b200 import _lib, ops
    lib = _lib.load()
    rs = np.random.RandomState(nb + npost + 1)
    brand = rs.standard_normal((nb, d)).astype(np.float32)
    posts = rs.standard_normal((npost, d)).astype(np.float32)
    lab = synth.labels(5, npost, nb)
    res, dense, a, b = _dense_and_topk(brand, posts, k, labels=lab, index_base=9)
    want = oref.topk_indices(dense, k)
    kk = min(k, npost)
    assert np.array_equal(res["index"].cpu().numpy()[:, :kk], want + 9)
    assert np.array_equal(res["scores"].cpu().numpy()[:, :kk], np.take_along_axis(dense, want, 1))
    assert np.array_equal(res["pos_score"].cpu().numpy(), dense[lab, np.arange(npost)])
    tj = rs.randint(0, npost, nb)
    tidx = (tj + 9).astype(np.int32)
    tidx[::3] = -1
    cnt = ops.score_count(a, b, to_dev(dense[np.arange(nb), tj].copy()), to_dev(tidx), d=d, index_base=9).cpu().numpy()
    lib.frx_set_cta_pairs(0)                              # single-CTA kernel on the same operands
    dense1 = ops.score_dense(a, b, d=d).cpu().numpy()
    res1 = ops.score_topk(a, b, k, d=d, labels=to_dev(lab.astype(np.int32)), index_base=9)
    cnt1 = ops.score_count(a, b, to_dev(dense[np.arange(nb), tj].copy()), to_dev(tidx), d=d, index_base=9).cpu().numpy()
    lib.frx_set_cta_pairs(1)
    assert np.array_equal(dense, dense1)
    assert torch.equal(res["index"], res1["index"]) and torch.equal(res["scores"], res1["scores"])
    assert np.array_equal(cnt, cnt1)


def test_cta_pair_variant_tf32_and_loss_tiles(cta_pairs):
    """tf32 operands and the K-split dense path (the B x B loss tile) on CTA pairs."""
    from fancyrec_b200 import loss as floss, ops, ranking
    rs = np.random.RandomState(2)
    nb, npost, d = 70, 3000, 256
    brand = rs.standard_normal((nb, d)).astype(np.float32)
    posts = rs.standard_normal((npost, d)).astype(np.float32)
    a = ranking.to_operand(to_dev(brand), precision="tf32x3", side=ranking.BRAND_SIDE)
    b = ranking.to_operand(to_dev(posts), precision="tf32x3", side=ranking.POST_SIDE)
    dense = ops.score_dense(a, b, d=3 * d).cpu().numpy()
    assert np.abs(dense - oref.cal_sim(brand, posts)).max() <= 1e-5
    ids = to_dev(rs.randint(0, 20, 512).astype(np.int64))
    be, pe = to_dev(rs.standard_normal((512, 1024)).astype(np.float32)), to_dev(rs.standard_normal((512, 1024)).astype(np.float32))
    crit = floss.TripletLoss(margin=0.2)
    pair_val = crit(ids, be, pe).item()
    from fancyrec_b200 import _lib
    _lib.load().frx_set_cta_pairs(0)
    single_val = crit(ids, be, pe).item()
    _lib.load().frx_set_cta_pairs(1)
    assert pair_val == single_val


@pytest.mark.parametrize("pairs", [0, 1])
def test_random_shape_sweep(pairs):
    """40 seeded random problems (ragged sizes around the tile boundaries, k around n_posts, quantised scores with
    heavy ties every third case): exact top-k, positives' scores and count pass against the oracle on our tile, on
    both kernel variants."""
    from fancyrec_b200 import _lib, ops
    lib = _lib.load()
    prev = lib.frx_set_cta_pairs(pairs)
    try:
        rs = np.random.RandomState(1234 + pairs)
        for case in range(40):
            nb = int(rs.choice([1, 2, 31, 127, 128, 129, 200, 255, 256, 257, 300]))
            npost = int(rs.choice([1, 3, 63, 255, 256, 257, 511, 513, 1000, 4099]))
            d = int(rs.choice([4, 8, 60, 64, 68, 128, 200]))
            k = int(rs.choice([1, 2, 10, 64, 100, 333, 1024]))
            if case % 3 == 0:
                brand = rs.randint(-2, 3, size=(nb, d)).astype(np.float32)
                posts = rs.randint(-2, 3, size=(npost, d)).astype(np.float32)
                brand[np.abs(brand).sum(1) == 0, 0] = 1.0
                posts[np.abs(posts).sum(1) == 0, 0] = 1.0
            else:
                brand = rs.standard_normal((nb, d)).astype(np.float32)
                posts = rs.standard_normal((npost, d)).astype(np.float32)
            lab = rs.randint(0, nb, npost).astype(np.int64)
            base = int(rs.choice([0, 1, 77777]))
            res, dense, a, b = _dense_and_topk(brand, posts, k, labels=lab, index_base=base)
            want = oref.topk_indices(dense, k)
            kk = min(k, npost)
            msg = "case %d: nb=%d npost=%d d=%d k=%d" % (case, nb, npost, d, k)
            assert np.array_equal(res["index"].cpu().numpy()[:, :kk], want + base), msg
            assert np.array_equal(res["scores"].cpu().numpy()[:, :kk], np.take_along_axis(dense, want, 1)), msg
            assert np.array_equal(res["pos_score"].cpu().numpy(), dense[lab, np.arange(npost)]), msg
            tj = rs.randint(0, npost, nb)
            cnt = ops.score_count(a, b, to_dev(dense[np.arange(nb), tj].copy()), to_dev((tj + base).astype(np.int32)),
                                  d=d, index_base=base).cpu().numpy()
            r = int(rs.randint(0, nb))
            assert cnt[r] == int(np.where(oref.order_desc(dense[r]) == tj[r])[0][0]), msg
    finally:
        lib.frx_set_cta_pairs(prev)


def test_heavy_ties_and_index_base():
    """Quantised scores (few distinct values) -> the tie-break carries the whole ranking."""
    rs = np.random.RandomState(3)
    nb, npost, d = 40, 9000, 64
    brand = synth.lattice(1, nb, d, nnz=16)
    posts = synth.lattice(2, npost, d, nnz=16)
    lab = synth.labels(5, npost, nb)
    base = 1234567
    res, dense, _, _ = _dense_and_topk(brand, posts, 100, labels=lab, index_base=base)
    assert len(np.unique(dense)) <= 33
    want = oref.topk_indices(dense, 100)
    assert np.array_equal(res["index"].cpu().numpy(), want + base)
    assert np.array_equal(res["pos_score"].cpu().numpy(), dense[lab, np.arange(npost)])


def test_adversarial_increasing_scores():
    """Scores increase with the post index: every element passes the running threshold."""
    nb, npost, d = 3, 6000, 64
    brand = np.zeros((nb, d), np.float32)
    brand[:, 0] = 1.0
    posts = np.zeros((npost, d), np.float32)
    posts[:, 0] = np.linspace(0.1, 1.0, npost)
    posts[:, 1] = np.sqrt(1.0 - posts[:, 0] ** 2)
    res, dense, _, _ = _dense_and_topk(brand, posts, 50)
    assert np.array_equal(res["index"].cpu().numpy(), oref.topk_indices(dense, 50))


def test_labels_out_of_range_give_nan_pos_score():
    rs = np.random.RandomState(9)
    brand = rs.standard_normal((4, 64)).astype(np.float32)
    posts = rs.standard_normal((300, 64)).astype(np.float32)
    lab = rs.randint(0, 4, 300)
    lab[[3, 77]] = 9
    lab[5] = -1
    res, dense, _, _ = _dense_and_topk(brand, posts, 8, labels=lab)
    ps = res["pos_score"].cpu().numpy()
    bad = (lab < 0) | (lab >= 4)
    assert np.isnan(ps[bad]).all()
    assert np.array_equal(ps[~bad], dense[lab[~bad], np.arange(300)[~bad]])


def test_count_pass_matches_oracle():
    from fancyrec_b200 import ops
    rs = np.random.RandomState(21)
    nb, npost, d = 70, 12000, 128
    brand = synth.lattice(7, nb, d, nnz=16)      # ties matter for the (score, index) comparison
    posts = synth.lattice(8, npost, d, nnz=16)
    a, b = operand(brand), operand(posts)
    dense = ops.score_dense(a, b, d=d).cpu().numpy()
    tj = rs.randint(0, npost, nb)
    ts = dense[np.arange(nb), tj].copy()
    tidx = tj.astype(np.int32)
    tidx[::9] = -1                                # skipped rows
    base = 100
    out = ops.score_count(a, b, to_dev(ts), to_dev(np.where(tidx >= 0, tidx + base, -1).astype(np.int32)), d=d,
                          index_base=base)
    got = out.cpu().numpy()
    for r in range(nb):
        if tidx[r] < 0:
            assert got[r] == 0
            continue
        order = oref.order_desc(dense[r])
        assert got[r] == int(np.where(order == tj[r])[0][0])


@pytest.mark.parametrize("need", ["none", "one_tile", "scattered"])
def test_count_pass_skips_tiles_without_thresholds(need):
    """frx_score_count skips every 128-brand tile whose rows all have thr_index < 0 (and returns at once when no row
    has a threshold): counts of the rows that DO have one are unchanged, all others stay zero."""
    from fancyrec_b200 import ops
    rs = np.random.RandomState(33)
    nb, npost, d = 400, 9000, 64                  # 4 m-tiles
    brand = rs.standard_normal((nb, d)).astype(np.float32)
    posts = rs.standard_normal((npost, d)).astype(np.float32)
    a, b = operand(brand), operand(posts)
    dense = ops.score_dense(a, b, d=d).cpu().numpy()
    tj = rs.randint(0, npost, nb)
    tidx = np.full(nb, -1, np.int32)
    rows = {"none": [], "one_tile": [130, 200, 255], "scattered": [0, 127, 128, 399]}[need]
    tidx[rows] = tj[rows]
    out = ops.score_count(a, b, to_dev(dense[np.arange(nb), tj].copy()), to_dev(tidx), d=d).cpu().numpy()
    for r in range(nb):
        want = int(np.where(oref.order_desc(dense[r]) == tj[r])[0][0]) if r in rows else 0
        assert out[r] == want


def test_missing_thresholds_and_packed_stats():
    from fancyrec_b200 import ops
    n_pos = to_dev(np.array([3, 0, 5, 1], np.int32))
    first = to_dev(np.array([-1, -1, 7, -1], np.int32))
    best = to_dev(np.array([11, -1, 22, 33], np.int32))
    assert ops.missing_thresholds(n_pos, first, best).cpu().tolist() == [11, -1, -1, 33]
    before = to_dev(np.array([100, 0, 0, 2 ** 40], np.int64))
    mask = to_dev(np.array([0, 0, 1 << 7, -1], np.int64))
    packed = ops.pack_rank_stats(n_pos, first, before, mask).cpu().numpy()
    assert packed.tolist() == [[3, 0, 5, 1], [-1, -1, 7, -1], [100, 0, 0, 2 ** 40], [1, 0, 0, 1], [0, 0, 128, -1]]
    auc = to_dev(np.array([9, 8, 7, 6], np.int64))
    packed = ops.pack_rank_stats(n_pos, first, before, mask, auc, all_valid=True).cpu().numpy()
    assert packed[3].tolist() == [1, 1, 1, 1] and packed[5].tolist() == [9, 8, 7, 6]


@pytest.mark.parametrize("nb,npost", [(300, 200000), (2048, 70000), (3000, 200000), (50, 999)])
def test_label_stats_both_paths(nb, npost):
    """n_pos and the best positive per brand (ties -> smaller index), shared-memory-privatised kernel (nb <= 2048,
    many posts) and the plain one; labels outside [0, nb) are ignored; brands without positives give (-inf, -1)."""
    from fancyrec_b200 import ops
    rs = np.random.RandomState(nb + npost)
    lab = rs.randint(-2, nb + 3, npost).astype(np.int32)
    lab[lab == 7] = 8                                           # brand 7 has no positive
    score = (rs.randint(-20, 21, npost) / 16.0).astype(np.float32)   # heavy ties
    base = 1000
    n_pos, bs, bi = ops.label_stats(to_dev(lab), to_dev(score), nb, base)
    n_pos, bs, bi = n_pos.cpu().numpy(), bs.cpu().numpy(), bi.cpu().numpy()
    ok = (lab >= 0) & (lab < nb)
    assert np.array_equal(n_pos, np.bincount(lab[ok], minlength=nb))
    order = np.lexsort((np.arange(npost), -score))              # score desc, index asc
    first = {}
    for j in order:
        if ok[j] and lab[j] not in first:
            first[lab[j]] = j
    for b in range(nb):
        if b in first:
            assert bi[b] == first[b] + base and bs[b] == score[first[b]]
        else:
            assert bi[b] == -1 and np.isneginf(bs[b])


def test_topk_merge_matches_oracle():
    from fancyrec_b200 import ops
    rs = np.random.RandomState(31)
    g, nb, k = 4, 33, 100
    scores = (rs.randint(-50, 51, size=(nb, 4000)) / 64.0).astype(np.float32)
    bounds = [0, 1000, 1900, 3100, 4000]
    sl, il = [], []
    for a, b in zip(bounds[:-1], bounds[1:]):
        idx = oref.topk_indices(scores[:, a:b], k) + a
        il.append(idx.astype(np.int32))
        sl.append(np.take_along_axis(scores, idx, 1))
    il[2][:, -7:] = -1                       # padding entries are ignored
    ws, wi = oref.merge_topk(sl, il, k)
    gs, gi = ops.topk_merge(to_dev(np.stack(sl)), to_dev(np.stack(il)), k)
    assert np.array_equal(gi.cpu().numpy(), wi.astype(np.int32))
    assert np.array_equal(gs.cpu().numpy(), ws)


def test_bad_arguments_raise():
    from fancyrec_b200 import _lib, ops
    a = torch.zeros((4, 64), dtype=torch.bfloat16, device=dev())
    with pytest.raises(_lib.FrxError):
        ops.score_topk(a, a, 2000)                       # k > 1024
    with pytest.raises(_lib.FrxError):
        ops.score_topk(a.cpu(), a, 4)                    # no CPU path


@pytest.mark.parametrize("cl", [2, 4, 8])
@pytest.mark.parametrize("nb,npost,d,k", [(300, 70001, 128, 100), (1000, 300000, 256, 64), (130, 5000, 96, 10), (5, 300, 64, 7)])
def test_multicast_cluster_variant_bit_identical(cl, nb, npost, d, k):
    """Clusters of 2 / 4 / 8 CTAs that share every post tile through TMA multicast (frx_set_cluster): fused top-k with the
    score matrix written on the way, positives' scores and the count pass are bit for bit those of the single-CTA kernel
    (m-tile counts that are not a multiple of the cluster size included: 300 brands = 3 m-tiles, 5 brands = 1)."""
    from fancyrec_b200 import _lib, ops, ranking
    lib = _lib.load()
    g = torch.Generator(device=dev()).manual_seed(nb + npost + cl)
    brand = ranking.to_operand(torch.randn((nb, d), generator=g, device=dev()))
    post = ranking.to_operand(torch.randn((npost, d), generator=g, device=dev()))
    labels = (torch.randperm(npost, generator=g, device=dev()) % nb).to(torch.int32)
    prev = lib.frx_set_cluster(0)
    try:
        ref = ops.score_topk(brand, post, k, d=d, labels=labels, index_base=3)
        refd = ops.score_dense(brand, post, d=d)
        kk = min(k, npost) - 1
        thr_s, thr_i = ref["scores"][:, kk // 2].contiguous(), ref["index"][:, kk // 2].contiguous()
        refc = ops.score_count(brand, post, thr_s, thr_i, d=d, index_base=3)
        lib.frx_set_cluster(cl)
        got = ops.score_topk(brand, post, k, d=d, labels=labels, index_base=3, dense=True)
        gotc = ops.score_count(brand, post, thr_s, thr_i, d=d, index_base=3)
        gotd = ops.score_dense(brand, post, d=d)
    finally:
        lib.frx_set_cluster(prev if prev > 1 else 0)
    assert torch.equal(got["index"], ref["index"]) and torch.equal(got["scores"], ref["scores"])
    assert torch.equal(got["pos_score"], ref["pos_score"])
    assert torch.equal(got["dense"], refd) and torch.equal(gotd, refd)
    assert torch.equal(gotc, refc) and bool((refc == kk // 2).all())
