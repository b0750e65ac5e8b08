"""Rank-metric scorers with the reference's names and semantics (util/metric.py:6-123).

`score(sorted_labels)` is the reference's host arithmetic over ONE list (nothing in the reference calls it; the API is
kept because north_star names it).  `score_device(labels)` evaluates the same scorer over a BATCH of integer label
lists on the GPU -- one warp per list, frx_metric_scores -- and returns float64 values bit-identical to `score`.
"""
import math
import random


class MetricScorer:
    def __init__(self, k=0):
        self.k = k

    def score(self, sorted_labels):
        return 0.0

    def getLength(self, sorted_labels):
        n = len(sorted_labels)
        return self.k if 0 < self.k <= n else n

    def name(self):
        base = type(self).__name__.replace("Scorer", "")
        return "%s@%d" % (base, self.k) if self.k > 0 else base

    _KIND = None

    def score_device(self, labels, lengths=None):
        """Batch form on the device: labels int32 CUDA tensor [N, L] of sorted label lists (lengths [N] int32 for ragged
        lists) -> float64 CUDA tensor [N], entry i == self.score(list i) bit for bit.  Raises what `score` raises
        (ZeroDivisionError for an NDCG list without a positive grade or an empty P@k list)."""
        if self._KIND is None:
            import torch
            return torch.zeros(labels.shape[0], dtype=torch.float64, device=labels.device)     # MetricScorer.score -> 0.0
        from .. import ops
        out = ops.metric_scores(labels, self._KIND, self.k, lengths)
        if self._KIND in ("NDCG", "P") and bool((out != out).any()):
            raise ZeroDivisionError("float division by zero")
        return out


class APScorer(MetricScorer):
    _KIND = "AP"

    def __init__(self, k):
        MetricScorer.__init__(self, k)

    def score(self, sorted_labels):
        relevant = sum(1 for x in sorted_labels if x > 0)
        if relevant == 0:
            return 0.0
        total, seen = 0.0, 0
        for pos in range(self.getLength(sorted_labels)):
            if sorted_labels[pos] >= 1:
                seen += 1
                total += float(seen) / (pos + 1.0)
        return total / relevant


class RRScorer(MetricScorer):
    _KIND = "RR"

    def score(self, sorted_labels):
        for pos, lab in enumerate(sorted_labels):
            if lab >= 1:
                return 1.0 / (pos + 1)
        return 0.0


class PrecisionScorer(MetricScorer):
    _KIND = "P"

    def score(self, sorted_labels):
        n = self.getLength(sorted_labels)
        return float(sum(1 for lab in sorted_labels[:n] if lab >= 1)) / n


class NDCGScorer(PrecisionScorer):
    _KIND = "NDCG"

    def score(self, sorted_labels):
        # no zero guard, as in the reference: all-zero labels raise ZeroDivisionError
        return self.getDCG(sorted_labels) / self.getIdealDCG(sorted_labels)

    def getDCG(self, sorted_labels):
        gain = max(sorted_labels[0], 0)
        for pos in range(1, self.getLength(sorted_labels)):
            gain += float(max(sorted_labels[pos], 0)) / math.log(pos + 1, 2)
        return gain

    def getIdealDCG(self, sorted_labels):
        return self.getDCG(sorted(sorted_labels, reverse=True))


class DCGScorer(PrecisionScorer):
    _KIND = "DCG"

    def score(self, sorted_labels):
        return self.getDCG(sorted_labels)

    def getIdealDCG(self, sorted_labels):
        return self.getDCG(sorted(sorted_labels, reverse=True))

    def getRandomDCG(self, sorted_labels):
        random.shuffle(sorted_labels)
        return self.getDCG(sorted_labels)

    def getDCG(self, sorted_labels):
        parts = [(math.pow(2, rel) - 1) / math.log(rank + 1, 2)
                 for rank, rel in enumerate(sorted_labels[:self.k], 1)]
        return 0.01757 * sum(parts)


_SCORERS = {"P": PrecisionScorer, "AP": APScorer, "RR": RRScorer, "NDCG": NDCGScorer, "DCG": DCGScorer}


def getScorer(name):
    parts = name.split("@")
    k = int(parts[1]) if len(parts) == 2 else 0
    return _SCORERS[parts[0]](k)
