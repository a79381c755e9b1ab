"""Updater family with the fork's step-size schedules.

Parity: ``Updater.compute(weightsOld, gradient, stepSize, iter, regParam) -> (weights, regVal)``
(mllib/optimization/Updater.scala:70-160).  ASYNC edits: ``SimpleUpdater`` keeps
η_t = η/√t (:77); ``L1Updater`` (:114-129) and ``SquaredL2Updater`` (:151-158) use a **constant**
step (upstream uses η/√t).  ``decay="reference"`` reproduces the fork, ``"sqrt"`` upstream Spark,
``"const"`` constant everywhere.  On CUDA tensors the elementwise work runs in the hand-written
``updater`` kernels (ops/updaters.py).
"""
from __future__ import annotations

import math
from typing import Tuple

import torch


def _step(step_size: float, it: int, decay: str, fork_default: str) -> float:
    mode = fork_default if decay == "reference" else decay
    if mode == "sqrt":
        return step_size / math.sqrt(max(it, 1))
    return step_size


class Updater:
    updater_id = -1

    def __init__(self, decay: str = "reference"):
        if decay not in ("reference", "sqrt", "const"):
            raise ValueError("decay must be reference|sqrt|const")
        self.decay = decay

    def compute(self, weights_old: torch.Tensor, gradient: torch.Tensor, step_size: float, it: int,
                reg_param: float) -> Tuple[torch.Tensor, float]:
        raise NotImplementedError


class SimpleUpdater(Updater):
    """w ← w − (η/√t)·g ; regVal = 0."""
    updater_id = 0

    def compute(self, weights_old, gradient, step_size, it, reg_param):
        eta = _step(step_size, it, self.decay, "sqrt")
        if weights_old.is_cuda:
            from ..ops import updaters
            return updaters.simple_update(weights_old, gradient, eta), 0.0
        return weights_old - eta * gradient.to(weights_old.dtype), 0.0


class L1Updater(Updater):
    """Proximal soft-threshold: w ← sign(w')·max(|w'| − ηλ, 0) with w' = w − ηg ; regVal = λ‖w‖₁."""
    updater_id = 1

    def compute(self, weights_old, gradient, step_size, it, reg_param):
        eta = _step(step_size, it, self.decay, "const")
        if weights_old.is_cuda:
            from ..ops import updaters
            w, l1 = updaters.l1_update(weights_old, gradient, eta, reg_param)
            return w, reg_param * l1
        w = weights_old - eta * gradient.to(weights_old.dtype)
        shrink = reg_param * eta
        w = torch.sign(w) * torch.clamp(w.abs() - shrink, min=0.0)
        return w, float(reg_param * w.abs().sum())


class SquaredL2Updater(Updater):
    """w ← (1 − ηλ)w − ηg ; regVal = ½λ‖w‖²."""
    updater_id = 2

    def compute(self, weights_old, gradient, step_size, it, reg_param):
        eta = _step(step_size, it, self.decay, "const")
        if weights_old.is_cuda:
            from ..ops import updaters
            w, sq = updaters.l2_update(weights_old, gradient, eta, reg_param)
            return w, 0.5 * reg_param * sq
        w = weights_old * (1.0 - eta * reg_param) - eta * gradient.to(weights_old.dtype)
        return w, float(0.5 * reg_param * (w.double() ** 2).sum())


def make_updater(name: str, decay: str = "reference") -> Updater:
    name = name.lower()
    if name in ("simple", "none"):
        return SimpleUpdater(decay)
    if name in ("l1",):
        return L1Updater(decay)
    if name in ("squared_l2", "l2"):
        return SquaredL2Updater(decay)
    raise ValueError(f"unknown updater {name!r}")
