"""bench.py contract on a box without a GPU: the reference arm (the reference's algorithm through the oracle port on the
host cores) prints ONE JSON line with the agreed keys; our arm refuses to run without a B200 (no CPU fallback)."""
import json
import os
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _run(args, env=None):
    e = dict(os.environ)
    e.update(env or {})
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py")] + args, cwd=ROOT, env=e,
                          stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, timeout=600)


def test_reference_arm_line():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--brands", "40", "--cpu-sample-posts", "3000"])
    assert p.returncode == 0, p.stderr[-2000:]
    lines = [l for l in p.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "brand_post_pairs_scored_and_ranked_per_sec"
    assert d["unit"] == "pairs/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("configs[1]")


def test_reference_arm_other_ranks_exit_quietly():
    p = _run(["--impl", "reference", "--steps", "1", "--warmup", "0", "--gpus", "2"], env={"RANK": "1", "WORLD_SIZE": "2"})
    assert p.returncode == 0 and p.stdout.strip() == ""


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box WITHOUT a GPU")
def test_our_arm_has_no_cpu_fallback():
    p = _run(["--steps", "1", "--warmup", "0", "--no-e2e", "--no-cpu-baseline"])
    assert p.returncode != 0 and "no CPU fallback" in (p.stderr + p.stdout)
