"""GPU parity: finalisation, brand embedding, rank statistics, the evaluator drop-in API."""
import os
import types

import numpy as np
import pytest
import torch

from oracle import embed as oembed
from oracle import ranking as oref
from oracle import synth
from tests.gpu_util import dev, score_atol, to_dev

pytestmark = pytest.mark.gpu

RTOL = 4e-6   # fp32 accumulation-order tolerance (kernel and torch sum in fp32, oracle in fp64)


def _bf16_to_f32(t):
    return t.float().cpu().numpy()


@pytest.mark.parametrize("vn,tn,fn", [(True, True, True), (False, False, True), (True, False, False), (False, True, True)])
def test_finalize_csr_matches_oracle_and_golden(golden_dir, vn, tn, fn):
    from fancyrec_b200 import ops
    frames, row_ptr = synth.frames_csr(707, 40, 96, 1, 9)
    text = np.abs(synth.gaussian(708, 40, 24, 0.3))
    out32, out16 = ops.finalize_posts(to_dev(frames), to_dev(text), row_ptr=to_dev(row_ptr), visual_norm=vn,
                                      text_norm=tn, final_norm=fn, want_f32=True, want_bf16=True)
    want = oembed.finalize_posts(oembed.mean_pool_csr(frames, row_ptr), text, vn, tn, fn)
    got = out32.cpu().numpy()
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-7)
    b16 = _bf16_to_f32(out16)
    assert b16.shape == (40, 128) and (b16[:, 120:] == 0).all()
    np.testing.assert_allclose(b16[:, :120], got, rtol=2 ** -8, atol=1e-30)   # bf16 round-to-nearest
    g = np.load(os.path.join(golden_dir, "finalize.npz"))
    if vn and tn and fn:
        np.testing.assert_allclose(got, g["final_branchnorm"], rtol=RTOL, atol=1e-7)
    if not vn and not tn and fn:
        np.testing.assert_allclose(got, g["final_raw"], rtol=RTOL, atol=1e-7)


def test_finalize_gather_unaligned_and_edge_cases():
    from fancyrec_b200 import ops
    # non-contiguous frame rows (row_idx), odd dims -> scalar path, F = 1 and F = 33 posts
    rs = np.random.RandomState(4)
    dv, dt = 50, 7
    frames = np.maximum(rs.standard_normal((200, dv)), 0).astype(np.float32)
    counts = np.array([1, 33, 2, 5, 1, 17], dtype=np.int64)
    row_ptr = np.concatenate([[0], np.cumsum(counts)])
    row_idx = rs.permutation(200)[:row_ptr[-1]].astype(np.int32)
    text = rs.standard_normal((6, dt)).astype(np.float32)
    out32, _ = ops.finalize_posts(to_dev(frames), to_dev(text), row_ptr=to_dev(row_ptr), row_idx=to_dev(row_idx),
                                  visual_norm=True, text_norm=True, final_norm=True, want_f32=True, want_bf16=False)
    want = oembed.finalize_posts(oembed.mean_pool_gather(frames, row_idx, row_ptr), text, True, True, True)
    np.testing.assert_allclose(out32.cpu().numpy(), want, rtol=RTOL, atol=1e-7)
    # zero row -> NaN like the reference (no epsilon); empty input is a no-op
    z = ops.finalize_posts(torch.zeros((2, 64), device=dev()), want_f32=True, want_bf16=False)[0]
    assert torch.isnan(z).all()
    e32, e16 = ops.finalize_posts(torch.zeros((0, 64), device=dev()), want_f32=True, want_bf16=True)
    assert e32.shape == (0, 64) and e16.shape == (0, 64)


def test_finalize_video_config_shape():
    """C5-shaped rows: 32 frames x 2048-d mean-pool + L2 norm (small NP)."""
    from fancyrec_b200 import ops
    rs = np.random.RandomState(6)
    n, f, dv = 64, 32, 2048
    frames = np.maximum(rs.standard_normal((n * f, dv)) * 0.5 + 0.3, 0).astype(np.float32)
    row_ptr = (np.arange(n + 1) * f).astype(np.int64)
    out32, out16 = ops.finalize_posts(to_dev(frames), row_ptr=to_dev(row_ptr), final_norm=True, want_f32=True)
    want = oembed.finalize_posts(oembed.mean_pool_csr(frames, row_ptr), None, False, False, True)
    np.testing.assert_allclose(out32.cpu().numpy(), want, rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(_bf16_to_f32(out16), want, rtol=2 ** -8, atol=1e-30)


def test_brand_embed_matches_oracle_and_golden(golden_dir):
    from fancyrec_b200 import ops
    g = np.load(os.path.join(golden_dir, "finalize.npz"))
    w = synth.gaussian(709, 9, 30)
    e = synth.gaussian(710, 30, 20)
    got = ops.brand_embed(to_dev(w), to_dev(e), brand_ids=to_dev(g["brand_ids"])).cpu().numpy()
    np.testing.assert_allclose(got, g["brand_emb"], rtol=1e-5, atol=1e-6)
    w2 = synth.gaussian(1, 131, 2000)
    e2 = synth.gaussian(2, 2000, 200)
    want2 = oembed.brand_embed(w2, e2, np.arange(130))
    for tc in (True, False):      # 3xTF32 tensor-core GEMM and the fp32 CUDA-core GEMM: both fp32-grade
        got2 = ops.brand_embed(to_dev(w2), to_dev(e2), nb=130, tensor_cores=tc).cpu().numpy()
        np.testing.assert_allclose(got2, want2, rtol=2e-4, atol=2e-6)
        assert np.abs(got2 - want2).max() <= 2e-6 * np.abs(want2).max() + 1e-7
    ids = np.array([5, 5, 130, 0, 77] * 60, dtype=np.int64)
    got3 = ops.brand_embed(to_dev(w2), to_dev(e2), brand_ids=to_dev(ids)).cpu().numpy()
    np.testing.assert_allclose(got3, oembed.brand_embed(w2, e2, ids), rtol=2e-4, atol=2e-6)


def _fake_model(nb, w, e):
    from fancyrec_b200 import model as fmodel
    opt = types.SimpleNamespace(brand_num=nb, common_embedding_size=e.shape[1], brand_aspect=e.shape[0])
    ba = fmodel.BrandAspects(opt)
    with torch.no_grad():
        ba.brand_embeddings.weight.copy_(torch.from_numpy(w))
        ba.aspects_embeddings.copy_(torch.from_numpy(e))
    return types.SimpleNamespace(brand_encoding=ba.to(dev()), opt=opt)


@pytest.mark.parametrize("name", list(synth.RANKING_CASES))
def test_test_post_ranking_dropin(golden_dir, name):
    """The reference-facing call.  Lattice fixtures: bit-exact against the UNMODIFIED reference run
    (stable tie-break).  Real-valued fixtures: bit-exact against the oracle on our own score tile, and
    close to the fp32 reference result."""
    from fancyrec_b200 import evaluator
    g = np.load(os.path.join(golden_dir, "ranking_%s.npz" % name))
    nb, lab, w, e, posts = synth.ranking_inputs(name)
    mdl = _fake_model(nb, w, e)
    post_t, lab_t = to_dev(posts), to_dev(lab)
    got = evaluator.test_post_ranking(nb, 'auc', mdl, post_t, lab_t)
    assert evaluator.test_post_ranking(nb, 'recall', mdl, post_t, lab_t) is None
    got = tuple(float(x) for x in got)
    assert isinstance(evaluator.test_post_ranking(nb, 'auc', mdl, post_t, lab_t)[0], np.float64)
    brand = evaluator.brand_matrix(mdl, nb)
    np.testing.assert_allclose(brand.cpu().numpy(), g["brand"], rtol=1e-5, atol=1e-6)
    ours = evaluator.cal_sim(brand, post_t).cpu().numpy()
    assert np.abs(ours - g["scores"]).max() <= score_atol(posts.shape[1])
    assert got == tuple(float(x) for x in oref.rank_metrics_vec(ours, lab))
    if name.startswith("lattice"):
        assert np.array_equal(ours, g["scores"])
        assert got == tuple(g["result_stable"])
        assert got[:5] == tuple(g["result"])[:5]
    else:
        ref = tuple(g["result"])
        assert abs(got[2] - ref[2]) < 5e-3 and abs(got[3] - ref[3]) < 0.05


@pytest.mark.parametrize("want_auc", [True, False])
def test_rank_statistics_bit_exact_medium(want_auc):
    """1k-post-per-brand scale, positives beyond the top-k list for most brands (pure noise) ->
    exercises the count pass (want_auc False) and the dense AUC sweep (True)."""
    from fancyrec_b200 import ops, ranking
    rs = np.random.RandomState(77)
    nb, npost, d = 37, 30000, 256
    brand = rs.standard_normal((nb, d)).astype(np.float32)
    lab = synth.labels(78, npost, nb, empty_brands=(3, 20))
    posts = synth.planted_posts(79, brand, lab, signal=0.0)   # pure noise: first positives land deep
    result, stats, dev_stats = ranking.rank_posts(to_dev(brand), to_dev(posts), to_dev(lab), want_auc=want_auc)
    ours = ops.score_dense(ranking.to_operand(to_dev(brand)), ranking.to_operand(to_dev(posts)), d=d).cpu().numpy()
    ost = oref.rank_stats(ours, lab)
    assert np.array_equal(stats["n_pos"], ost["n_pos"])
    assert np.array_equal(stats["first_rank"], ost["first_rank"])
    assert np.array_equal(stats["hits"], ost["hits"])
    want = oref.aggregate(ost, npost)
    if want_auc:
        assert np.array_equal(stats["auc_num"], ost["auc_num"])
        assert tuple(map(float, result)) == tuple(map(float, want))
    else:
        assert np.isnan(result[2])
        assert tuple(map(float, result[:2] + result[3:])) == tuple(map(float, want[:2] + want[3:]))
    assert (ost["first_rank"][ost["n_pos"] > 0] >= 64).any()   # the list alone would not have been enough


def test_all_brands_empty_raises_like_reference():
    from fancyrec_b200 import ranking
    rs = np.random.RandomState(1)
    brand = rs.standard_normal((3, 64)).astype(np.float32)
    posts = rs.standard_normal((100, 64)).astype(np.float32)
    with pytest.raises(IndexError):
        ranking.rank_posts(to_dev(brand), to_dev(posts), to_dev(np.full(100, 7)), want_auc=True)


def test_l2norm_and_cal_sim_api():
    from fancyrec_b200 import evaluator, model
    rs = np.random.RandomState(2)
    x = rs.standard_normal((33, 100)).astype(np.float32)
    y = rs.standard_normal((7, 100)).astype(np.float32)
    np.testing.assert_allclose(evaluator.l2norm(to_dev(x)).cpu().numpy(), oref.l2norm(x), rtol=RTOL, atol=1e-7)
    np.testing.assert_allclose(model.l2norm(to_dev(x)).cpu().numpy(), oref.l2norm(x), rtol=RTOL, atol=1e-7)
    s = evaluator.cal_sim(to_dev(y), to_dev(x))
    assert s.shape == (7, 33) and s.dtype == torch.float32 and s.is_cuda
    assert np.abs(s.cpu().numpy() - oref.cal_sim(y, x)).max() <= score_atol(100)
    assert evaluator.random_sim(3, 5).shape == (3, 5)


def test_encode_data_contract():
    """encode_data with a fake loader/model: dataset-order scatter of post rows, loader-order labels."""
    from fancyrec_b200 import evaluator

    class _Enc(torch.nn.Module):
        def forward(self, x):
            return x

    class _Model:
        def __init__(self):
            self.opt = types.SimpleNamespace(single_modal_text=False, single_modal_visual=False,
                                             common_embedding_size=8)
            self.brand_encoding = self.vid_encoding = self.text_encoding = self.fusion_encoding = _Enc()

        def __call__(self, brand_ids, videos, captions):
            return None, videos.to(dev())

    data = torch.arange(10 * 8, dtype=torch.float32).reshape(10, 8)

    class _Loader:
        dataset = list(range(10))

        def __iter__(self):
            for idxs in ([3, 1, 2], [0, 9, 8, 7], [4, 6, 5]):
                yield torch.tensor(idxs) % 3, data[idxs], None, idxs, None, None

        def __len__(self):
            return 3

    logs = []
    brands, embs = evaluator.encode_data(_Model(), _Loader(), log_step=1, logging=logs.append)
    assert torch.equal(embs.cpu(), data) and embs.is_cuda
    assert brands.cpu().tolist() == [0, 1, 2, 0, 0, 2, 1, 1, 0, 2] and len(logs) == 3


def test_bigfile_ingest_to_device(tmp_path):
    """feature.bin -> ImageBigFile.read_csr -> ONE finalize pass on the device == the reference's per-frame
    read_one + torch.mean + l2norm path (util/imgbigfile.py:19-57, util/data_provider.py:40, evaluator.py:14-19)."""
    from fancyrec_b200 import ops
    from fancyrec_b200.util.imgbigfile import ImageBigFile
    rs = np.random.RandomState(17)
    n_rows, dims = 300, 128
    feats = np.maximum(rs.standard_normal((n_rows, dims)), 0).astype(np.float32)
    names = ["v%d_frame_%d_cls%d" % (i // 7, i % 7, i % 5) for i in range(n_rows)]
    perm = rs.permutation(n_rows)                          # frames of one video are scattered in the file
    feats_file, names_file = feats[perm], [names[i] for i in perm]
    feats_file.tofile(os.path.join(str(tmp_path), "feature.bin"))
    open(os.path.join(str(tmp_path), "id.txt"), "w", encoding="utf8").write("#".join(names_file))
    open(os.path.join(str(tmp_path), "shape.txt"), "w").write("%d %d" % (n_rows, dims))
    bf = ImageBigFile(str(tmp_path))
    posts = [[n for n in names if n.startswith("v%d_" % v)] for v in range(n_rows // 7 + 1)]
    posts = [p for p in posts if p]
    row_idx, row_ptr = bf.read_csr(posts)
    out = ops.finalize_posts(to_dev(np.array(bf.matrix)), row_ptr=to_dev(row_ptr), row_idx=to_dev(row_idx),
                             final_norm=True, want_f32=True, want_bf16=False)[0].cpu().numpy()
    want = []
    for frames in posts:                                   # the reference way, frame by frame
        vecs = np.array([bf.read_one(f) for f in frames], dtype=np.float32)
        m = vecs.astype(np.float64).mean(0)
        want.append(m / np.sqrt((m * m).sum()))
    np.testing.assert_allclose(out, np.array(want), rtol=RTOL, atol=1e-7)


@pytest.mark.parametrize("chunk", [7, 64, 100000])
def test_ingest_from_host_bit_identical_to_device_path(tmp_path, chunk):
    """ingest.finalize_from_host (memmap / NumPy / pinned host features, chunked + double-buffered H2D) gives
    bit for bit what ops.finalize_posts gives on device-resident inputs: pooled + scattered rows (feature.bin),
    pooled + contiguous rows, and un-pooled visual + text rows."""
    from fancyrec_b200 import ingest, ops
    from fancyrec_b200.util.imgbigfile import ImageBigFile
    rs = np.random.RandomState(31)
    n_rows, dims = 900, 2048
    feats = np.maximum(rs.standard_normal((n_rows, dims)), 0).astype(np.float32)
    names = ["v%d_frame_%d" % (i // 9, i % 9) for i in range(n_rows)]
    perm = rs.permutation(n_rows)
    feats[perm].tofile(os.path.join(str(tmp_path), "feature.bin"))
    open(os.path.join(str(tmp_path), "id.txt"), "w", encoding="utf8").write("#".join(names[i] for i in perm))
    open(os.path.join(str(tmp_path), "shape.txt"), "w").write("%d %d" % (n_rows, dims))
    bf = ImageBigFile(str(tmp_path))
    posts = [[n for n in names[v * 9:(v + 1) * 9]][:1 + (v % 9)] for v in range(n_rows // 9)]     # 1..9 frames
    row_idx, row_ptr = bf.read_csr(posts)
    text = rs.standard_normal((len(posts), 300)).astype(np.float32)
    kw = dict(visual_norm=True, text_norm=True, final_norm=True, want_f32=True, want_bf16=True)
    want = ops.finalize_posts(to_dev(np.array(bf.matrix)), to_dev(text), row_ptr=to_dev(row_ptr),
                              row_idx=to_dev(row_idx), **kw)
    got = ingest.finalize_from_host(bf.matrix, text, row_ptr=row_ptr, row_idx=row_idx, chunk_posts=chunk, **kw)
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
    # contiguous rows per post (no gather list), pinned source
    gathered = torch.from_numpy(np.array(bf.matrix)[row_idx]).pin_memory()
    got = ingest.finalize_from_host(gathered, torch.from_numpy(text), row_ptr=row_ptr, chunk_posts=chunk, **kw)
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
    # one row per post + text
    vis = rs.standard_normal((len(posts), 1024)).astype(np.float32)
    want = ops.finalize_posts(to_dev(vis), to_dev(text), **kw)
    got = ingest.finalize_from_host(torch.from_numpy(vis).pin_memory(), torch.from_numpy(text).pin_memory(),
                                    chunk_posts=chunk, **kw)
    assert torch.equal(got[0], want[0]) and torch.equal(got[1], want[1])
    got = ingest.finalize_from_host(vis, None, chunk_posts=chunk, final_norm=True, want_bf16=True)
    assert torch.equal(got[1], ops.finalize_posts(to_dev(vis), final_norm=True)[1]) and got[0] is None


def test_masked_mean_pool_matches_reference_loop():
    """the encoders' per-sample `torch.mean(seq[:len], 0)` loops (model.py:105-114) as one
    pass of the pooling kernel."""
    from fancyrec_b200 import ops
    rs = np.random.RandomState(23)
    b, t, d = 17, 64, 256
    x = rs.standard_normal((b, t, d)).astype(np.float32)
    lengths = rs.randint(1, t + 1, b)
    lengths[0], lengths[1] = t, 1
    got = ops.masked_mean_pool(to_dev(x), lengths.tolist()).cpu().numpy()
    want = np.stack([x[i, :lengths[i]].astype(np.float64).mean(0) for i in range(b)])
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=1e-6)
    gotn = ops.masked_mean_pool(to_dev(x), to_dev(lengths), l2norm=True).cpu().numpy()
    np.testing.assert_allclose(gotn, want / np.linalg.norm(want, axis=1, keepdims=True), rtol=RTOL, atol=1e-6)


def test_masked_softmax_pool_matches_reference_loop():
    """MultiHeadSelfAttention.forward's per-sample soft-max loop + weighted mean (model.py:105-114)
    as one pass; compared with the loop itself restated in fp64."""
    from fancyrec_b200 import ops
    rs = np.random.RandomState(29)
    for b, t, d in ((17, 64, 2048), (5, 7, 50), (3, 300, 128)):
        x = rs.standard_normal((b, t, d)).astype(np.float32)
        att = (3 * rs.standard_normal((b, t, 1))).astype(np.float32)
        lengths = rs.randint(1, t + 1, b)
        lengths[0], lengths[1], lengths[2] = t, 1, 0               # full, single step, empty (-> zero row)
        got = ops.masked_softmax_pool(to_dev(x), to_dev(att), lengths.tolist()).cpu().numpy()
        want = np.zeros((b, d))
        for i in range(b):
            if lengths[i] == 0:
                continue
            a = att[i, :lengths[i], 0].astype(np.float64)
            w = np.exp(a - a.max()); w /= w.sum()
            weight = np.zeros(t); weight[:lengths[i]] = w
            want[i] = (weight[:, None] * x[i].astype(np.float64)).mean(0)
        np.testing.assert_allclose(got, want, rtol=2e-5, atol=2e-7)   # expf + fp32 weighted sums vs fp64


def test_validate_and_tester_flow(golden_dir, capsys, tmp_path):
    """trainer.validate / tester.evaluate on the exact-lattice fixture through a fake loader: the 8 metrics are
    bit-identical to the UNMODIFIED reference run stored in the golden file, rsum follows trainer.py:413-415, the
    printed lines are the reference's, and save_checkpoint keeps the reference's selection rule."""
    from fancyrec_b200 import tester, trainer
    g = np.load(os.path.join(golden_dir, "ranking_lattice.npz"))
    nb, lab, w, e, posts = synth.ranking_inputs("lattice")
    mdl = _fake_model(nb, w, e)
    d = posts.shape[1]
    mdl.opt = types.SimpleNamespace(single_modal_text=False, single_modal_visual=False, common_embedding_size=d,
                                    brand_num=nb, metric='auc', log_step=100)

    class _Enc(torch.nn.Module):
        def forward(self, x):
            return x
    mdl.vid_encoding = mdl.text_encoding = mdl.fusion_encoding = _Enc()

    class _Callable:
        """model(brand_ids, videos, captions) -> (None, post embeddings); attribute access falls through to mdl."""
        def __call__(self, brand_ids, videos, captions):
            return None, videos.to(dev())

        def __getattr__(self, name):
            return getattr(mdl, name)

    n = posts.shape[0]
    posts_t, lab_t = torch.from_numpy(posts), torch.from_numpy(lab)

    class _Loader:
        dataset = list(range(n))

        def __iter__(self):
            for lo in range(0, n, 128):
                idxs = list(range(lo, min(n, lo + 128)))
                yield lab_t[idxs], posts_t[idxs], None, idxs, None, None

        def __len__(self):
            return (n + 127) // 128

    out = trainer.validate(mdl.opt, _Loader(), _Callable())
    printed = capsys.readouterr().out.splitlines()
    medr, meanr, auc, n10, n50, r1, r5, r10 = tuple(g["result_stable"])
    assert tuple(float(x) for x in out[1:]) == (auc, n10, n50, medr, meanr, r1, r5, r10)
    assert float(out[0]) == float((np.float64(auc) + np.float64(n10) + np.float64(n50)) * 100 + r1 + r5 + r10)
    assert [l.split(':')[0] for l in printed] == ['MedR', 'MeanR', 'AUC[0-1]', 'NDCG@10[0-1]', 'NDCG@50[0-1]',
                                                  'recall@1', 'recall@5', 'recall@10']
    res = tester.evaluate(mdl.opt, _Callable(), _Loader(), log_step=100)
    assert tuple(float(x) for x in res) == tuple(g["result_stable"])
    assert len(capsys.readouterr().out.splitlines()) == 4
    # checkpoint selection (trainer.py:419-424)
    prefix = str(tmp_path) + os.sep
    best = trainer.save_checkpoint({"x": 1}, 10.0, 0.0, prefix=prefix, best_epoch=None)
    assert best == 10.0 and os.path.exists(prefix + 'model_best.pth.tar')
    os.remove(prefix + 'checkpoint.pth.tar')
    assert trainer.save_checkpoint({"x": 2}, 9.5, best, prefix=prefix, best_epoch=1) == 10.0
    assert not os.path.exists(prefix + 'checkpoint.pth.tar')          # more than 1 % below the best: not saved
    assert trainer.save_checkpoint({"x": 3}, 9.95, best, prefix=prefix, best_epoch=1) == 10.0
    assert os.path.exists(prefix + 'checkpoint.pth.tar')
