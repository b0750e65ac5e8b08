"""ImageBigFile mirror vs the reference's own outputs on a toy directory (tests/golden/bigfile.json)."""
import json
import os

import numpy as np


def _make_dir(tmp_path, g):
    rs = np.random.RandomState(g["feats_seed"])
    feats = rs.standard_normal((7, 5)).astype(np.float32)
    feats.tofile(os.path.join(tmp_path, "feature.bin"))
    open(os.path.join(tmp_path, "id.txt"), "w", encoding="utf8").write("#".join(g["names"]))
    open(os.path.join(tmp_path, "shape.txt"), "w").write("7 5")
    return feats


def test_reference_api_semantics(tmp_path, golden_dir):
    from fancyrec_b200.util.imgbigfile import ImageBigFile
    g = json.load(open(os.path.join(golden_dir, "bigfile.json")))
    feats = _make_dir(str(tmp_path), g)
    bf = ImageBigFile(str(tmp_path))
    names, vecs = bf.read(g["request"])
    assert names == g["read_names"] and vecs == g["read_vecs"]          # dedup, unknown dropped, row order
    assert isinstance(vecs[0], list) and isinstance(vecs[0][0], float)
    assert bf.read_one("img9_cls0") == g["read_one"]
    ni, vi = bf.read([5, 0, 5], isname=False)
    assert ni == g["read_idx_names"] and vi == g["read_idx_vecs"]
    assert [list(x) for x in bf.read(["nope"])] == g["empty"]
    assert bf.shape() == g["shape"]
    assert bf.ndims == 5 and bf.nr_of_images == 7 and bf.name2index["a"] == 5


def test_bulk_csr_path(tmp_path, golden_dir):
    from fancyrec_b200.util.imgbigfile import ImageBigFile
    from oracle import embed
    g = json.load(open(os.path.join(golden_dir, "bigfile.json")))
    feats = _make_dir(str(tmp_path), g)
    bf = ImageBigFile(str(tmp_path))
    posts = [["v1_frame_1_cls2", "v1_frame_0_cls2"], ["img9_cls0"], ["z", "a", "b"]]
    row_idx, row_ptr = bf.read_csr(posts)
    assert row_idx.dtype == np.int32 and row_ptr.dtype == np.int64
    assert row_idx.tolist() == [1, 0, 2, 6, 5, 4] and row_ptr.tolist() == [0, 2, 3, 6]
    pooled = embed.mean_pool_gather(np.asarray(bf.matrix), row_idx, row_ptr)
    np.testing.assert_allclose(pooled[0], feats[[0, 1]].mean(0), rtol=1e-6)
    nm, mat = bf.read_matrix(["b", "a"])
    assert nm == ["b", "a"] and np.array_equal(mat, feats[[4, 5]])
