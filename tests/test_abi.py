"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol that
include/frx.h declares, and refuses to compute without an sm_100 GPU (no CPU fallback)."""
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "frx.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(frx_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as entry
    entry.build()
    from fancyrec_b200 import _lib
    lib = _lib.load()
    declared = _declared_symbols()
    assert len(declared) >= 19
    for name in declared:
        assert hasattr(lib, name), name
        assert name in _lib.SIGNATURES, "no ctypes prototype for %s" % name
    assert sorted(_lib.SIGNATURES) == declared
    assert lib.frx_abi_version() == 1


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the no-GPU behaviour")
def test_no_cpu_fallback():
    from fancyrec_b200 import _lib, evaluator, ops
    lib = _lib.load()
    assert lib.frx_device_check(0) == -2
    assert b"no CPU fallback" in lib.frx_last_error()
    x = torch.zeros((4, 64))
    with pytest.raises(_lib.FrxError):
        ops.finalize_posts(x)
    with pytest.raises(_lib.FrxError):
        evaluator.l2norm(x)
    with pytest.raises(_lib.FrxError):
        evaluator.cal_sim(x, x)


def test_workspace_queries_need_no_gpu():
    from fancyrec_b200 import _lib
    lib = _lib.load()
    assert lib.frx_score_topk_workspace_bytes(1000, 1000000, 3072, 100) > 0
    assert lib.frx_score_topk_workspace_bytes(1000, 1000000, 3072, 5000) == 0     # k out of range
    assert lib.frx_triplet_workspace_bytes(512, 1024) >= 2 * 512 * 512 * 4
    assert lib.frx_contrastive_workspace_bytes(512, 1024, 5120) >= 512 * 5120 * 4


def test_host_mirrors_match_reference_goldens(golden_dir):
    import json
    from fancyrec_b200.util import metric, ndcg
    g = json.load(open(os.path.join(golden_dir, "ndcg_metric.json")))
    for lst, k, m, want in g["dcg"]:
        assert ndcg.dcg_at_k(lst, k, m) == want
    for lst, k, m, want in g["ndcg"]:
        assert ndcg.ndcg_at_k(lst, k, m) == want
        if m == 0 and set(lst) <= {0, 1}:
            assert ndcg.ndcg_from_hits(lst[:50], sum(lst), k, len(lst)) == want
    for name, rows in g["scorers"].items():
        for lst, nm, ln, sc in rows:
            s = metric.getScorer(name)
            assert (s.name(), s.getLength(lst), s.score(list(lst))) == (nm, ln, sc)
    ms = metric.MetricScorer(10)
    assert [ms.name(), ms.score([3, 2, 3, 0, 1, 2]), ms.getLength([3, 2, 3, 0, 1, 2])] == g["metric_scorer_main"]
    with pytest.raises(ZeroDivisionError):
        metric.NDCGScorer(10).score([0, 0, 0])
    with pytest.raises(ValueError):
        ndcg.dcg_at_k([1, 2], 2, method=3)


def test_aggregate_matches_oracle_on_host():
    """Host aggregation (evaluator.py:129-143) from integer statistics == oracle, bit for bit."""
    import numpy as np
    from fancyrec_b200 import ranking
    from oracle import ranking as oref
    from oracle import synth
    rs = np.random.RandomState(5)
    scores = (rs.randint(-6, 7, size=(7, 150)) / 8.0).astype(np.float32)
    lab = synth.labels(3, 150, 7, empty_brands=(2,))
    st = oref.rank_stats(scores, lab)
    got = ranking.aggregate(dict(n_pos=st["n_pos"], first_rank=st["first_rank"], hits=st["hits"],
                                 auc_num=st["auc_num"]), 150, True)
    assert tuple(map(float, got)) == tuple(map(float, oref.rank_metrics_loop(scores, lab)))


def test_vectorised_aggregate_is_bit_identical_to_per_brand_reference_calls():
    """Row-wise np.sum over [NB, 50] == the reference's per-brand 1-D np.sum (pairwise order), and the
    Python-int AUC division == float64 division, over many random relevance patterns."""
    import numpy as np
    from fancyrec_b200 import ranking
    from oracle import ranking as oref
    rs = np.random.RandomState(11)
    for trial in range(20):
        nb, n_posts = 200, 100000
        n_pos = rs.randint(0, 300, nb)
        n_pos[rs.randint(0, nb, 5)] = 0
        hits = (rs.rand(nb, 50) < rs.rand(nb, 1)).astype(np.uint8)
        hits[n_pos == 0] = 0
        first = np.where(n_pos > 0, rs.randint(0, 5000, nb), -1)
        auc_num = (rs.rand(nb) * n_pos * (n_posts - n_pos)).astype(np.int64)
        st = dict(n_pos=n_pos.astype(np.int64), first_rank=first.astype(np.int64), hits=hits, auc_num=auc_num)
        got = ranking.aggregate(st, n_posts, True)
        want = oref.aggregate(st, n_posts)
        assert tuple(map(float, got)) == tuple(map(float, want))


def test_header_is_plain_c_and_links(tmp_path):
    """include/frx.h is the drop-in boundary: it must compile as C99 (no C++ / torch types in the signatures) and a plain
    C program must link against libfrx_b200.so and call it (no compute without a GPU: only the version query)."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        import pytest
        pytest.skip("gcc not available")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "c_check.c"
    src.write_text('#include "frx.h"\n#include <stdio.h>\n'
                   'int main(void) {\n'
                   '  size_t ws = frx_score_topk_workspace_bytes(1000, 1000000, 3072, 100);\n'
                   '  printf("%d %d\\n", frx_abi_version(), ws > 0);\n'
                   '  return frx_device_check(0) == 0 ? 0 : 3;   /* 3 on a box without an sm_100 GPU */\n}\n')
    exe = tmp_path / "c_check"
    libdir = os.path.join(root, "fancyrec_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", os.path.join(root, "include"),
                    str(src), "-o", str(exe), "-L", libdir, "-l:libfrx_b200.so", "-Wl,-rpath," + libdir], check=True)
    p = subprocess.run([str(exe)], stdout=subprocess.PIPE, text=True)
    assert p.stdout.split() == ["1", "1"]
    import torch
    assert p.returncode == (0 if torch.cuda.is_available() else 3)
