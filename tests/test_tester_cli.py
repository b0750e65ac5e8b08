"""`python -m fancyrec_b200.tester` is the reference's tester.py from the outside -- the same
command line (every option, type, default, choice: tests/golden/tester_cli.json is the reference parser, introspected),
exit code 0 for a missing checkpoint (tester.py:59-61) and for an existing result without --overwrite (tester.py:74-75),
the reference's checkpoint layout, and (GPU test below) the same four printed lines."""
import argparse
import json
import os
import subprocess
import sys
import types

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _our_actions():
    from fancyrec_b200 import tester
    captured = {}
    orig = argparse.ArgumentParser.parse_args

    def spy(self, args=None, namespace=None):
        captured["actions"] = [dict(dest=a.dest, option_strings=list(a.option_strings), default=a.default,
                                    type=getattr(a.type, "__name__", None), choices=list(a.choices) if a.choices else None,
                                    required=a.required)
                               for a in self._actions if a.dest != "help"]
        return orig(self, args, namespace)

    argparse.ArgumentParser.parse_args = spy
    try:
        ns = tester.parse_args(["insCartest"])
    finally:
        argparse.ArgumentParser.parse_args = orig
    return captured["actions"], vars(ns)


def test_command_line_surface_equals_the_reference(golden_dir):
    g = json.load(open(os.path.join(golden_dir, "tester_cli.json")))
    actions, defaults = _our_actions()
    assert actions == g["actions"]
    assert defaults == g["parsed_defaults"]


def test_missing_checkpoint_exits_zero_like_the_reference(tmp_path):
    p = subprocess.run([sys.executable, "-m", "fancyrec_b200.tester", "insCartest", "--logger_name", str(tmp_path)],
                       cwd=ROOT, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=300)
    assert p.returncode == 0, p.stderr[-1000:]
    assert json.loads(p.stdout)["testCollection"] == "insCartest"          # json.dumps(vars(opt), indent=2), tester.py:53


class _Enc(torch.nn.Module):
    """Stand-in for a learned encoder (out of scope): one Linear layer."""

    def __init__(self, i, o):
        super().__init__()
        self.fc = torch.nn.Linear(i, o)

    def forward(self, data):
        return self.fc(data[0] if isinstance(data, (tuple, list)) else data)


class _Fuse(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.gain = torch.nn.Parameter(torch.ones(1))

    def forward(self, v, t):
        return torch.cat((v, t), 1) * self.gain


def _options(nb, d):
    return argparse.Namespace(trainCollection="insCartrain", cv_name="cv_x", brand_num=nb, brand_aspect=24,
                              common_embedding_size=d, metric="auc", single_modal_text=False, single_modal_visual=False,
                              text_net="bi-gru", fusion_style="ph")


def _build(n, dv, dt, nb, seed):
    from fancyrec_b200 import model

    def build(opt, options, checkpoint):
        mdl = model.FancyRec(options, vid_encoding=_Enc(dv, 40), text_encoding=_Enc(dt, 24), fusion_encoding=_Fuse())
        g = torch.Generator().manual_seed(seed)
        labels = torch.randint(0, nb, (n,), generator=g)
        vis, txt = torch.randn((n, dv), generator=g), torch.randn((n, dt), generator=g)

        class DS(torch.utils.data.Dataset):
            def __len__(self):
                return n

        def batches():
            for lo in range(0, n, opt.batch_size):
                idx = list(range(lo, min(n, lo + opt.batch_size)))
                v = vis[idx]
                videos = (v, v, [1] * len(idx), torch.ones(len(idx), 1))
                yield labels[idx], videos, (txt[idx],), idx, idx, idx
        loader = types.SimpleNamespace(dataset=DS(), __iter__=None)

        class Loader:
            dataset = DS()

            def __iter__(self):
                return batches()

            def __len__(self):
                return (n + opt.batch_size - 1) // opt.batch_size
        return mdl, Loader()
    return build


@pytest.mark.gpu
def test_cli_flow_on_a_reference_layout_checkpoint(tmp_path, capsys):
    from fancyrec_b200 import evaluator, model, tester
    nb, d, n, dv, dt = 9, 64, 400, 30, 20
    options = _options(nb, d)
    torch.manual_seed(11)
    src = model.FancyRec(options, vid_encoding=_Enc(dv, 40), text_encoding=_Enc(dt, 24), fusion_encoding=_Fuse())
    run = tmp_path / "insCartrain" / "cv_x" / "run0"
    run.mkdir(parents=True)
    # trainer.py:294-301 -- the reference's checkpoint dict; 'model' is the LIST of four state_dicts (model.py:637-649)
    torch.save({"epoch": 3, "model": src.state_dict(), "best_rsum": 1.0, "opt": options, "Eiters": 77},
               str(run / "model_best.pth.tar"))
    argv = ["insCartest", "--rootpath", str(tmp_path), "--overwrite", "1", "--n_caption", "1", "--batch_size", "64",
            "--logger_name", str(run)]
    res = tester.main(argv, build=_build(n, dv, dt, nb, seed=5))
    out = capsys.readouterr().out
    assert "=> loaded!" in out
    lines = [l for l in out.splitlines() if l.startswith(("AUC[0-1]:", "NDCG@10[0-1]:", "NDCG@50[0-1]:", "recall@1:"))]
    assert [l.split(":")[0] for l in lines] == ["AUC[0-1]", "NDCG@10[0-1]", "NDCG@50[0-1]", "recall@1"]
    assert lines[0] == "AUC[0-1]: %s" % res[2] and lines[3] == "recall@1: %s" % res[5]
    # the output directory rule of tester.py:69-72 was applied (trainCollection -> testCollection, cv_name -> results/...)
    assert os.path.isdir(str(tmp_path / "insCartest" / "results" / "insCartrain" / "run0" / "model_best.pth.tar"))
    # the same evaluation done by hand from the checkpointed weights gives the same tuple
    mdl, loader = _build(n, dv, dt, nb, seed=5)(tester.parse_args(argv), options, None)
    mdl = mdl.to(torch.device("cuda:0"))
    mdl.load_state_dict(torch.load(str(run / "model_best.pth.tar"), weights_only=False)["model"])
    brands, posts = evaluator.encode_data(mdl, loader, 10, lambda *_: None)
    want = evaluator.test_post_ranking(nb, "auc", mdl, posts, brands)
    assert tuple(map(float, res)) == tuple(map(float, want))
    # an existing result without --overwrite: skip with exit code 0 (tester.py:74-75)
    marker = tmp_path / "insCartest" / "results" / "insCartrain" / "run0" / "model_best.pth.tar" / "pred_errors_matrix.pth.tar"
    marker.write_text("x")
    with pytest.raises(SystemExit) as ex:
        tester.main(argv[:4] + ["0"] + argv[5:], build=_build(n, dv, dt, nb, seed=5))
    assert ex.value.code == 0
