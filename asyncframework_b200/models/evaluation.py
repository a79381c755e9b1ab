"""Model evaluation metrics for the GLM family (mllib/evaluation/{RegressionMetrics,BinaryClassificationMetrics}.scala).

Input is a ``ShardedDataset`` of ``(prediction-or-score, label)`` records — what ``model.predict`` zipped with the labels gives —
or two tensors.  Regression metrics are one aggregate pass; the binary curves sort the DISTINCT scores (the reference bins them
when ``numBins`` is set; the exact curve is affordable here because scores arrive as tensors).
"""
from __future__ import annotations

import math
from typing import List, Tuple

import torch


def _pairs(data, labels=None) -> Tuple[torch.Tensor, torch.Tensor]:
    if labels is not None:
        return torch.as_tensor(data, dtype=torch.float64).reshape(-1), torch.as_tensor(labels, dtype=torch.float64).reshape(-1)
    recs = data.collect() if hasattr(data, "collect") else list(data)
    if not recs:
        return torch.zeros(0, dtype=torch.float64), torch.zeros(0, dtype=torch.float64)
    t = torch.as_tensor([(float(a), float(b)) for a, b in recs], dtype=torch.float64)
    return t[:, 0], t[:, 1]


class RegressionMetrics:
    def __init__(self, prediction_and_observations, observations=None):
        p, y = _pairs(prediction_and_observations, observations)
        self.n = p.numel()
        err = y - p
        self._ss_err = float((err * err).sum())
        self._abs_err = float(err.abs().sum())
        self._ss_tot = float(((y - y.mean()) ** 2).sum()) if self.n else 0.0
        self._ss_reg = float(((p - y.mean()) ** 2).sum()) if self.n else 0.0

    @property
    def mean_squared_error(self) -> float: return self._ss_err / self.n
    @property
    def root_mean_squared_error(self) -> float: return math.sqrt(self.mean_squared_error)
    @property
    def mean_absolute_error(self) -> float: return self._abs_err / self.n
    @property
    def r2(self) -> float: return 1.0 - self._ss_err / self._ss_tot
    @property
    def explained_variance(self) -> float: return self._ss_reg / self.n


class BinaryClassificationMetrics:
    def __init__(self, score_and_labels, labels=None):
        s, y = _pairs(score_and_labels, labels)
        order = torch.argsort(s, descending=True, stable=True)
        s, y = s[order], y[order]
        # cumulative counts at the LAST occurrence of every distinct score (threshold = that score)
        last = torch.ones_like(s, dtype=torch.bool)
        if s.numel() > 1:
            last[:-1] = s[:-1] != s[1:]
        self._thresholds = s[last]
        self._tp = torch.cumsum(y, 0)[last]
        self._fp = torch.cumsum(1.0 - y, 0)[last]
        self._pos, self._neg = float(y.sum()), float((1.0 - y).sum())

    def thresholds(self) -> List[float]:
        return self._thresholds.tolist()

    def roc(self) -> List[Tuple[float, float]]:
        """(false positive rate, true positive rate) points from (0,0) to (1,1)."""
        fpr = (self._fp / max(self._neg, 1e-300)).tolist()
        tpr = (self._tp / max(self._pos, 1e-300)).tolist()
        return [(0.0, 0.0)] + list(zip(fpr, tpr)) + [(1.0, 1.0)]

    def pr(self) -> List[Tuple[float, float]]:
        """(recall, precision) points, starting at (0, first precision) like the reference."""
        rec = (self._tp / max(self._pos, 1e-300)).tolist()
        prec = (self._tp / (self._tp + self._fp)).tolist()
        return [(0.0, prec[0] if prec else 1.0)] + list(zip(rec, prec))

    @staticmethod
    def _trapezoid(points) -> float:
        return sum((x1 - x0) * (y0 + y1) / 2.0 for (x0, y0), (x1, y1) in zip(points, points[1:]))

    @property
    def area_under_roc(self) -> float: return self._trapezoid(self.roc())
    @property
    def area_under_pr(self) -> float: return self._trapezoid(self.pr())

    def f_measure_by_threshold(self, beta: float = 1.0) -> List[Tuple[float, float]]:
        p = self._tp / (self._tp + self._fp)
        r = self._tp / max(self._pos, 1e-300)
        f = (1 + beta * beta) * p * r / (beta * beta * p + r).clamp_min(1e-300)
        return list(zip(self._thresholds.tolist(), f.tolist()))
