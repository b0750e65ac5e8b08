#!/usr/bin/env python
"""Generate tests/golden/*.npz|json by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference mounted):

    python tests/golden/make_golden.py

Shims: ``np.asfarray`` was removed in NumPy 2 and is restored with
the NumPy-1.21 semantics; bytecode writing is off because the tree is read-only.
Inputs come from oracle/synth.py (bit-stable RandomState streams), so the golden
files store seeds + reference OUTPUTS.  Nothing at test/bench time reads
/root/reference; only these committed files travel.
"""
import json
import os
import sys
import tempfile
import types

sys.dont_write_bytecode = True
import numpy as np

np.asfarray = lambda a, dtype=np.float64: np.asarray(a, dtype=dtype)  # noqa: E731

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("FANCYREC_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
sys.path.insert(0, REF)

import torch  # noqa: E402

torch.manual_seed(0)
torch.set_num_threads(1)

import evaluator as ref_eval  # noqa: E402
import loss as ref_loss  # noqa: E402
import loss_ctrs as ref_ctrs  # noqa: E402
import model as ref_model  # noqa: E402
from util import metric as ref_metric  # noqa: E402
from util import ndcg as ref_ndcg  # noqa: E402
from util.imgbigfile import ImageBigFile  # noqa: E402

from oracle import synth  # noqa: E402
from oracle.synth import RANKING_CASES, loss_inputs, ranking_inputs  # noqa: E402


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path), "bytes")


# ----------------------------------------------------------------------------
# 1. NDCG doctests + util/metric scorers
# ----------------------------------------------------------------------------
def gen_ndcg_metric():
    r = [3, 2, 3, 0, 0, 1, 2, 2, 3, 0]
    out = {
        "dcg": [[r, 1, 0, ref_ndcg.dcg_at_k(r, 1)], [r, 1, 1, ref_ndcg.dcg_at_k(r, 1, method=1)],
                [r, 2, 0, ref_ndcg.dcg_at_k(r, 2)], [r, 2, 1, ref_ndcg.dcg_at_k(r, 2, method=1)],
                [r, 10, 0, ref_ndcg.dcg_at_k(r, 10)], [r, 11, 0, ref_ndcg.dcg_at_k(r, 11)],
                [[], 3, 0, ref_ndcg.dcg_at_k([], 3)]],
        "ndcg": [[r, 1, 0, ref_ndcg.ndcg_at_k(r, 1)], [[2, 1, 2, 0], 4, 0, ref_ndcg.ndcg_at_k([2, 1, 2, 0], 4)],
                 [[2, 1, 2, 0], 4, 1, ref_ndcg.ndcg_at_k([2, 1, 2, 0], 4, method=1)],
                 [[0], 1, 0, ref_ndcg.ndcg_at_k([0], 1)], [[1], 2, 0, ref_ndcg.ndcg_at_k([1], 2)]],
        "scorers": {},
    }
    rs = np.random.RandomState(7)
    lists = [[1, 1, 0, 0, 0], [3, 2, 3, 0, 1, 2], [0, 0, 1], rs.randint(0, 2, 60).tolist(),
             rs.randint(0, 4, 25).tolist()]
    for k in (10, 50):
        for lst in lists[3:]:
            out["ndcg"].append([lst, k, 0, ref_ndcg.ndcg_at_k(lst, k)])
    names = ["P@1", "P@5", "P", "AP", "AP@2", "AP@10", "NDCG", "NDCG@10", "NDCG@3", "RR", "DCG@5", "DCG@25"]
    for name in names:
        rows = []
        for lst in lists:
            sc = ref_metric.getScorer(name)
            rows.append([lst, sc.name(), sc.getLength(lst), sc.score(list(lst))])
        out["scorers"][name] = rows
    ms = ref_metric.MetricScorer(10)
    out["metric_scorer_main"] = [ms.name(), ms.score([3, 2, 3, 0, 1, 2]), ms.getLength([3, 2, 3, 0, 1, 2])]
    with open(os.path.join(HERE, "ndcg_metric.json"), "w") as f:
        json.dump(out, f)
    print("wrote ndcg_metric.json")


# ----------------------------------------------------------------------------
# 2. test_post_ranking through the reference (BrandAspects -> cal_sim -> loops)
# ----------------------------------------------------------------------------
class _Opt(types.SimpleNamespace):
    pass


def _brand_model(nb, a, d, w, e):
    opt = _Opt(brand_num=nb, common_embedding_size=d, brand_aspect=a)
    ba = ref_model.BrandAspects(opt)
    with torch.no_grad():
        ba.brand_embeddings.weight.copy_(torch.from_numpy(w))
        ba.aspects_embeddings.copy_(torch.from_numpy(e))
    return types.SimpleNamespace(brand_encoding=ba, opt=opt)


def gen_ranking():
    for name in RANKING_CASES:
        nb, lab, w, e, posts = ranking_inputs(name)
        a, d = e.shape
        mdl = _brand_model(nb, a, d, w, e)
        with torch.no_grad():
            aspects = mdl.brand_encoding.eval()(torch.arange(nb))
            brand = aspects.permute((1, 0, 2)).mean(0)          # evaluator.py:93-94
            scores = ref_eval.cal_sim(brand, torch.from_numpy(posts)).numpy().copy()
            res = ref_eval.test_post_ranking(nb, 'auc', mdl, torch.from_numpy(posts), torch.from_numpy(lab))
            # evaluator.py:124 uses np.argsort(-d), whose order under ties is unspecified (not stable on
            # this AVX-512 host).  Second run with the stated tie-break: argsort forced to kind='stable'.
            orig_argsort = np.argsort
            np.argsort = lambda a, *args, **kw: orig_argsort(a, kind='stable')
            try:
                res_stable = ref_eval.test_post_ranking(nb, 'auc', mdl, torch.from_numpy(posts),
                                                        torch.from_numpy(lab))
            finally:
                np.argsort = orig_argsort
        none_res = ref_eval.test_post_ranking(nb, 'recall', mdl, torch.from_numpy(posts), torch.from_numpy(lab))
        assert none_res is None
        save("ranking_%s.npz" % name, result=np.array([float(x) for x in res], dtype=np.float64),
             result_stable=np.array([float(x) for x in res_stable], dtype=np.float64),
             scores=scores, brand=brand.numpy())


# ----------------------------------------------------------------------------
# 3. losses (value + autograd gradients)
# ----------------------------------------------------------------------------
def gen_losses():
    ids, brand, post = loss_inputs()
    out = {}
    for style in ("sum", "mean"):
        for mv in (False, True):
            bt = torch.from_numpy(brand).requires_grad_()
            pt = torch.from_numpy(post).requires_grad_()
            crit = ref_loss.TripletLoss(margin=0.2, max_violation=mv, cost_style=style)
            val = crit(torch.from_numpy(ids), bt, pt)
            val.backward()
            key = "triplet_%s_mv%d" % (style, int(mv))
            out[key + "_loss"] = val.detach().numpy()
            out[key + "_dbrand"] = bt.grad.numpy().copy()
            out[key + "_dpost"] = pt.grad.numpy().copy()
    # contrastive: two consecutive steps so the queue pointer moves and wraps (Q = 2B)
    b, d = brand.shape
    for style in ("sum", "mean"):
        for mode in ("queue", "no_queue", "no_intra"):
            opt = _Opt(cost_style=style, queue_size=2 * b, common_embedding_size=d,
                       no_queue=(mode == "no_queue"), no_intra=(mode == "no_intra"))
            crit = ref_ctrs.ContrastiveLoss(opt)
            for step in range(3):
                ids2, brand2, post2 = loss_inputs(seed=600 + step)
                bt = torch.from_numpy(brand2).requires_grad_()
                pt = torch.from_numpy(post2).requires_grad_()
                val = crit(bt, pt)
                val.backward()
                key = "ctr_%s_%s_s%d" % (style, mode, step)
                out[key + "_loss"] = val.detach().numpy()
                out[key + "_dbrand"] = bt.grad.numpy().copy()
                out[key + "_dpost"] = pt.grad.numpy().copy()
                out[key + "_queue"] = crit.queue.numpy().copy()
                out[key + "_ptr"] = crit.queue_ptr.numpy().copy()
    # the reference raises when Q % B != 0: record the exception type
    opt = _Opt(cost_style="sum", queue_size=2 * b + 4, common_embedding_size=d, no_queue=False, no_intra=False)
    crit = ref_ctrs.ContrastiveLoss(opt)
    errs = []
    for step in range(3):
        try:
            crit(torch.from_numpy(brand), torch.from_numpy(post))
            errs.append("ok")
        except Exception as ex:  # noqa: BLE001
            errs.append(type(ex).__name__)
    out["ctr_bad_queue_errors"] = np.array(errs)
    for style in ("sum", "mean"):
        bt = torch.from_numpy(brand).requires_grad_()
        pt = torch.from_numpy(post).requires_grad_()
        val = ref_ctrs.CrossCLR_onlyIntraModality(cost_style=style)(bt, pt)
        val.backward()
        out["crossclr_%s_loss" % style] = val.detach().numpy()
        out["crossclr_%s_dbrand" % style] = bt.grad.numpy().copy()
        out["crossclr_%s_dpost" % style] = pt.grad.numpy().copy()
    bt = torch.from_numpy(brand).requires_grad_()
    val = ref_loss.LabLoss()(bt)
    val.backward()
    out["lab_loss"] = val.detach().numpy()
    out["lab_dbrand"] = bt.grad.numpy().copy()
    # direction != 'all' crashes in the reference: record
    for direction in ("p2b", "b2p"):
        try:
            ref_loss.TripletLoss(margin=0.2, direction=direction)(torch.from_numpy(ids), torch.from_numpy(brand),
                                                                  torch.from_numpy(post))
            out["triplet_dir_%s" % direction] = np.array("ok")
        except Exception as ex:  # noqa: BLE001
            out["triplet_dir_%s" % direction] = np.array(type(ex).__name__)
    save("losses.npz", **out)


# ----------------------------------------------------------------------------
# 4. post finalisation + brand embedding (torch ops the reference uses)
# ----------------------------------------------------------------------------
def gen_finalize():
    frames, row_ptr = synth.frames_csr(707, 40, 96, 1, 9)
    text = np.abs(synth.gaussian(708, 40, 24, 0.3))
    pooled = torch.stack([torch.mean(torch.from_numpy(frames[row_ptr[p]:row_ptr[p + 1]]), 0)
                          for p in range(40)])                  # data_provider.py:40
    v = ref_model.l2norm(pooled)                                # model.py:207-208
    t = ref_model.l2norm(torch.from_numpy(text))                # model.py:301-302
    cat_nn = torch.cat((v, t), 1)                               # model.py:483
    cat_raw = torch.cat((pooled, torch.from_numpy(text)), 1)
    w = synth.gaussian(709, 9, 30)
    e = synth.gaussian(710, 30, 20)
    mdl = _brand_model(8, 30, 20, w, e)
    ids = torch.tensor([3, 0, 7, 7, 8, 1])
    with torch.no_grad():
        emb = mdl.brand_encoding.eval()(ids).permute((1, 0, 2)).mean(0)   # model.py:593-594
    save("finalize.npz", pooled=pooled.numpy(), cat_branchnorm=cat_nn.numpy(),
         final_branchnorm=ref_eval.l2norm(cat_nn).numpy(), final_raw=ref_eval.l2norm(cat_raw).numpy(),
         brand_emb=emb.numpy(), brand_ids=ids.numpy())


# ----------------------------------------------------------------------------
# 5. bigfile reader semantics (dedup, ascending row order, unknown names dropped)
# ----------------------------------------------------------------------------
def gen_bigfile():
    rs = np.random.RandomState(808)
    feats = rs.standard_normal((7, 5)).astype(np.float32)
    names = ["v1_frame_0_cls2", "v1_frame_1_cls2", "img9_cls0", "v2_frame_0_cls1", "b", "a", "z"]
    with tempfile.TemporaryDirectory() as d:
        feats.tofile(os.path.join(d, "feature.bin"))
        open(os.path.join(d, "id.txt"), "w", encoding="utf8").write("#".join(names))
        open(os.path.join(d, "shape.txt"), "w").write("7 5")
        bf = ImageBigFile(d)
        req = ["b", "z", "a", "a", "b", "missing", "v1_frame_1_cls2"]
        rn, rv = bf.read(req)
        one = bf.read_one("img9_cls0")
        ri, rvi = bf.read([5, 0, 5], isname=False)
        empty = bf.read(["nope"])
        shape = bf.shape()
    with open(os.path.join(HERE, "bigfile.json"), "w") as f:
        json.dump({"names": names, "feats_seed": 808, "request": req, "read_names": rn, "read_vecs": rv,
                   "read_one": one, "read_idx_names": ri, "read_idx_vecs": rvi,
                   "empty": [list(empty[0]), list(empty[1])], "shape": shape}, f)
    print("wrote bigfile.json")


# ----------------------------------------------------------------------------
# 6. encoder-side Linear layers: MFC and PrjHeadFusionEncoder in eval mode
# ----------------------------------------------------------------------------
def gen_encoder_layers():
    torch.manual_seed(1234)
    out = {}
    mfc = ref_model.MFC([300, 200], 0.2).eval()
    with torch.no_grad():
        mfc.fc1.bias.copy_(torch.randn(200) * 0.1)
    x = torch.from_numpy(synth.gaussian(611, 37, 300))
    with torch.no_grad():
        out["mfc_out"] = mfc(x).numpy()
    for k, v in mfc.state_dict().items():
        out["mfc." + k] = v.numpy()
    opt = types.SimpleNamespace(common_embedding_size=96, visual_mapping_size=[0, 120], text_mapping_size=[0, 72],
                                prj_head_output=False)
    head = ref_model.PrjHeadFusionEncoder(opt).eval()
    bn = head.projection_head[1]
    with torch.no_grad():                      # a BatchNorm that has seen data: non-trivial statistics and affine terms
        bn.running_mean.copy_(torch.randn(512) * 0.05)
        bn.running_var.copy_(torch.rand(512) * 0.5 + 0.5)
        bn.weight.copy_(torch.rand(512) + 0.5)
        bn.bias.copy_(torch.randn(512) * 0.1)
        head.fc2.bias.copy_(torch.randn(96) * 0.1)
    v = torch.from_numpy(synth.gaussian(612, 41, 120))
    t = torch.from_numpy(synth.gaussian(613, 41, 72))
    with torch.no_grad():
        out["head_out"] = head(v, t).numpy()
        opt.prj_head_output = True
        out["head_concat"] = head(v, t).numpy()
    for k, val in head.state_dict().items():
        if not k.startswith(("projection_head.0.", "projection_head.3.")):      # the same tensors as fc1 / fc2
            out["head." + k] = val.numpy()
    save("encoder_layers.npz", **out)


# ----------------------------------------------------------------------------
# 7. tester.py command line: every option of the reference parser with its default / type / choices
# ----------------------------------------------------------------------------
def gen_tester_cli():
    import argparse
    captured = {}
    orig = argparse.ArgumentParser.parse_args

    def spy(self, args=None, namespace=None):
        captured["actions"] = [dict(dest=a.dest, option_strings=list(a.option_strings), default=a.default,
                                    type=getattr(a.type, "__name__", None), choices=list(a.choices) if a.choices else None,
                                    required=a.required)
                               for a in self._actions if a.dest != "help"]
        return orig(self, ["insCartest"], namespace)

    sys.modules.setdefault("tensorboard_logger", types.ModuleType("tensorboard_logger"))
    import tester as ref_tester
    argparse.ArgumentParser.parse_args = spy
    try:
        ns = ref_tester.parse_args()
    finally:
        argparse.ArgumentParser.parse_args = orig
    out = {"actions": captured["actions"], "parsed_defaults": vars(ns)}
    with open(os.path.join(HERE, "tester_cli.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("wrote tester_cli.json")


if __name__ == "__main__":
    gen_ndcg_metric()
    gen_ranking()
    gen_losses()
    gen_finalize()
    gen_bigfile()
    gen_encoder_layers()
    gen_tester_cli()
