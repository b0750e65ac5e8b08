"""A10 on the device: util/metric.py's scorers over batches of sorted label lists (frx_metric_scores, one warp per
list) against the host scorers -- themselves pinned to the reference's values in tests/golden/ndcg_metric.json -- bit
for bit in float64: random graded lists, ragged lengths, every cut-off regime of MetricScorer.getLength (k = 0,
k < len, k == len, k > len), negative grades, and the reference's ZeroDivisionError cases."""
import numpy as np
import pytest
import torch

from tests.gpu_util import dev

pytestmark = pytest.mark.gpu

NAMES = ["P", "P@1", "P@5", "P@50", "AP", "AP@2", "AP@10", "AP@500", "RR", "RR@3", "NDCG", "NDCG@1", "NDCG@10", "NDCG@50",
         "NDCG@77", "NDCG@500", "DCG@1", "DCG@5", "DCG@25", "DCG@500", "DCG"]


def _lists(seed, n, max_len, ragged):
    rs = np.random.RandomState(seed)
    lab = rs.choice([0, 0, 0, 0, 1, 1, 2, 3, -1], size=(n, max_len)).astype(np.int32)
    lab[0, :] = 0
    lab[0, max_len // 2] = 1                       # a single late positive
    lab[1, :] = 3                                   # all relevant, one grade
    lens = rs.randint(1, max_len + 1, size=n).astype(np.int32) if ragged else np.full(n, max_len, np.int32)
    lens[:2] = max_len
    for i in range(n):                              # every list keeps a positive grade inside its length (NDCG defined)
        if lab[i, :lens[i]].max() <= 0:
            lab[i, 0] = 2
    return lab, lens


@pytest.mark.parametrize("max_len,ragged", [(77, False), (77, True), (1000, True), (5, True)])
def test_device_scorers_equal_host_scorers_bit_for_bit(max_len, ragged):
    from fancyrec_b200.util.metric import getScorer
    lab, lens = _lists(max_len + ragged, 64, max_len, ragged)
    lab_t = torch.from_numpy(lab).to(dev())
    lens_t = torch.from_numpy(lens).to(dev()) if ragged else None
    for name in NAMES:
        scorer = getScorer(name)
        got = scorer.score_device(lab_t, lens_t).cpu().numpy()
        want = np.array([scorer.score([int(v) for v in lab[i, :lens[i]]]) for i in range(len(lab))], dtype=np.float64)
        assert got.dtype == np.float64
        assert np.array_equal(got.view(np.int64), want.view(np.int64)), (name, np.abs(got - want).max())


def test_device_scorers_reproduce_the_reference_goldens():
    from fancyrec_b200.util.metric import getScorer
    for labels in ([1, 1, 0, 0, 0], [3, 2, 3, 0, 1, 2]):
        t = torch.tensor([labels], dtype=torch.int32, device=dev())
        for name in ["P@1", "AP", "AP@2", "NDCG", "NDCG@10", "RR", "DCG@5"]:
            assert float(getScorer(name).score_device(t)[0]) == getScorer(name).score(labels)
    # the values SURVEY.md 4 records from the reference run
    t = torch.tensor([[3, 2, 3, 0, 1, 2]], dtype=torch.int32, device=dev())
    assert float(getScorer("AP").score_device(t)[0]) == 0.9266666666666665
    assert float(getScorer("NDCG@10").score_device(t)[0]) == 0.9315085232327253
    assert float(getScorer("DCG@5").score_device(t)[0]) == 0.22453831113386238
    t = torch.tensor([[1, 1, 0, 0, 0]], dtype=torch.int32, device=dev())
    assert float(getScorer("DCG@5").score_device(t)[0]) == 0.028655435770250502
    assert float(getScorer("AP@2").score_device(t)[0]) == 1.0


def test_device_scorers_raise_like_the_reference():
    from fancyrec_b200.util.metric import MetricScorer, getScorer
    zeros = torch.zeros((3, 9), dtype=torch.int32, device=dev())
    with pytest.raises(ZeroDivisionError):
        getScorer("NDCG@5").score_device(zeros)            # util/metric.py:74-77: d / d2 with d2 == 0
    assert float(getScorer("AP").score_device(zeros)[0]) == 0.0 and float(getScorer("RR").score_device(zeros)[0]) == 0.0
    assert float(MetricScorer(10).score_device(zeros)[0]) == 0.0


def test_scorers_on_topk_relevance_lists():
    """The natural caller: relevance lists of the fused top-k output (label of the post at each rank == brand)."""
    from fancyrec_b200 import ops, ranking
    from fancyrec_b200.util.metric import getScorer
    g = torch.Generator(device=dev()).manual_seed(3)
    nb, n, d, k = 40, 5000, 64, 50
    brand = torch.randn((nb, d), generator=g, device=dev())
    labels = (torch.randperm(n, generator=g, device=dev()) % nb).to(torch.int32)
    posts = torch.randn((n, d), generator=g, device=dev()) + 0.5 * brand[labels.long()]
    res = ops.score_topk(ranking.to_operand(brand), ranking.to_operand(posts), k, d=d, labels=labels)
    rel = (labels[res["index"].long()] == torch.arange(nb, device=dev(), dtype=torch.int32).unsqueeze(1)).to(torch.int32)
    host = rel.cpu().numpy()
    for name in ["NDCG@10", "P@10", "AP@50", "RR"]:
        sc = getScorer(name)
        want = np.array([sc.score([int(v) for v in host[b]]) for b in range(nb)])
        assert np.array_equal(sc.score_device(rel.contiguous()).cpu().numpy(), want)
