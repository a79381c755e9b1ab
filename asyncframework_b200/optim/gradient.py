"""Gradient family.

Parity: ``Gradient.compute(data, label, weights, cumGradient) -> loss``
(mllib/optimization/Gradient.scala:30-58); ``LogisticGradient`` binary (:188-196) and
multinomial (:197-273), ``LeastSquaresGradient`` (:285-303), ``HingeGradient`` (:313-343),
``log1pExp`` (mllib/util/MLUtils.scala:548-554); and the ASYNC samples' inlined least-squares
``gradfun`` with coefficient ``r = x·w - y`` (no factor 2; SparkASGDThread.scala:420-435).

Each gradient exposes
  * ``compute(features, label, weights, cum_gradient=None)``  — record level, reference signature;
  * ``coefficients(margin, y) -> (coeff, loss)``             — the elementwise K2 op used by the
    batched torch path *and* mirrored by the CUDA kernels (``LOSS_ID``);
  * ``batch(shard, w) -> (grad_sum[d], loss_sum, count)``    — vectorised oracle.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ..data.batch import Shard
from ..linalg import blas_util
from ..linalg.vectors import DenseVector, Vector

LOSS_IDS = {"least_squares": 0, "logistic": 1, "hinge": 2, "least_squares_half": 3}


def log1p_exp(x: torch.Tensor) -> torch.Tensor:
    """Stable log(1 + e^x)."""
    return torch.where(x > 0, x + torch.log1p(torch.exp(-x.abs())), torch.log1p(torch.exp(-x.abs())))


class Gradient:
    name = "gradient"
    loss_id = -1

    def coefficients(self, margin: torch.Tensor, y: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
        raise NotImplementedError

    def compute(self, features: Vector, label: float, weights: Vector,
                cum_gradient: Optional[DenseVector] = None):
        """Reference signature.  With ``cum_gradient`` the gradient is *added* into it and the
        loss is returned; without it ``(gradient, loss)`` is returned."""
        m = torch.as_tensor(blas_util.dot(features, weights), dtype=torch.float64)
        c, l = self.coefficients(m, torch.as_tensor(float(label), dtype=torch.float64))
        if cum_gradient is None:
            g = DenseVector(torch.zeros(weights.size, dtype=torch.float64))
            blas_util.axpy(float(c), features, g)
            return g, float(l)
        blas_util.axpy(float(c), features, cum_gradient)
        return float(l)

    def batch(self, shard: Shard, w: torch.Tensor, history: Optional[torch.Tensor] = None):
        """(Σ coeff_i x_i, Σ loss_i, n, coeff) over all rows of ``shard``; with ``history`` the
        SAGA correction ``coeff_i - history_i`` is used for the gradient (K4)."""
        margin = shard.matvec(w)
        coeff, loss = self.coefficients(margin, shard.y.to(margin.dtype))
        eff = coeff if history is None else coeff - history.to(coeff.dtype)
        return shard.rmatvec(eff), loss.sum(), shard.num_rows, coeff


class LeastSquaresGradient(Gradient):
    """MLlib flavour: diff = x·w - y, grad = diff·x, loss = diff²/2."""
    name = "least_squares_half"
    loss_id = LOSS_IDS["least_squares_half"]

    def coefficients(self, margin, y):
        diff = margin - y
        return diff, diff * diff / 2.0


class AsyncLeastSquaresGradient(Gradient):
    """The ASYNC samples' ``gradfun``: same coefficient, objective reported as diff² (A19)."""
    name = "least_squares"
    loss_id = LOSS_IDS["least_squares"]

    def coefficients(self, margin, y):
        diff = margin - y
        return diff, diff * diff


class LogisticGradient(Gradient):
    """Binary logistic with labels in {0,1}: margin = -x·w; mult = 1/(1+e^margin) - y;
    loss = log1pExp(margin) [- margin if y == 0 ... i.e. + x·w·(1-y)] (Gradient.scala:188-196)."""
    name = "logistic"
    loss_id = LOSS_IDS["logistic"]

    def __init__(self, num_classes: int = 2):
        self.num_classes = int(num_classes)

    def coefficients(self, margin, y):
        neg = -margin
        mult = torch.sigmoid(margin) - y           # 1/(1+exp(-x·w)) - y
        loss = torch.where(y > 0, log1p_exp(neg), log1p_exp(neg) - neg)
        return mult, loss

    def compute(self, features, label, weights, cum_gradient=None):
        if self.num_classes == 2:
            return super().compute(features, label, weights, cum_gradient)
        return self._compute_multinomial(features, label, weights, cum_gradient)

    def _compute_multinomial(self, features, label, weights, cum_gradient):
        """(K-1)·d weight layout with max-margin stabilisation (Gradient.scala:197-273)."""
        x = features.to_dense().values.double()
        d = x.numel()
        K = self.num_classes
        W = weights.to_dense().values.double().reshape(K - 1, d)
        margins = W @ x
        max_margin = float(margins.max().clamp_min(0.0)) if margins.numel() else 0.0
        lab = int(label)
        margin_y = float(margins[lab - 1]) if lab > 0 else 0.0
        if max_margin > 0:
            shifted = margins - max_margin
            s = torch.exp(shifted).sum() + torch.exp(torch.tensor(-max_margin, dtype=torch.float64))
            mult = torch.exp(shifted) / s
            loss = float(torch.log(s)) + max_margin - margin_y
        else:
            s = torch.exp(margins).sum() + 1.0
            mult = torch.exp(margins) / s
            loss = float(torch.log(s)) - margin_y
        if lab > 0:
            mult = mult.clone()
            mult[lab - 1] -= 1.0
        g = (mult[:, None] * x[None, :]).reshape(-1)
        if cum_gradient is None:
            return DenseVector(g), loss
        cum_gradient.values.add_(g.to(cum_gradient.values.dtype))
        return loss


class HingeGradient(Gradient):
    """SVM hinge with labels {0,1} mapped to ±1 (Gradient.scala:313-343)."""
    name = "hinge"
    loss_id = LOSS_IDS["hinge"]

    def coefficients(self, margin, y):
        s = 2.0 * y - 1.0
        active = (1.0 - s * margin) > 0
        coeff = torch.where(active, -s, torch.zeros_like(s))
        loss = torch.where(active, 1.0 - s * margin, torch.zeros_like(s))
        return coeff, loss


def make_gradient(name: str) -> Gradient:
    name = name.lower()
    if name in ("least_squares", "async_least_squares", "ls"):
        return AsyncLeastSquaresGradient()
    if name in ("least_squares_half", "mllib_least_squares"):
        return LeastSquaresGradient()
    if name in ("logistic", "lr"):
        return LogisticGradient()
    if name in ("hinge", "svm"):
        return HingeGradient()
    raise ValueError(f"unknown gradient {name!r}")
