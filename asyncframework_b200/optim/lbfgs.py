"""Limited-memory BFGS over a ``ShardedDataset`` (the other MLlib optimizer next to GradientDescent).

Parity: ``LBFGS`` (mllib/optimization/LBFGS.scala): ``setNumCorrections`` (default 10), ``setConvergenceTol`` (1e-6),
``setMaxNumIterations`` (100), ``setRegParam``; the cost function (``CostFun.calculate``) is the FULL-batch
``treeAggregate`` of (gradSum, lossSum) divided by n, plus the regularisation value / gradient derived from the
``Updater`` with the reference's trick — ``regVal = updater.compute(w, 0, 0, 1, λ)._2`` and ``regGrad = w −
updater.compute(w, 0, 1, 1, λ)._1`` — so any L2-type updater works unchanged.  The reference delegates the quasi-Newton
iteration to breeze's ``LBFGS`` (two-loop recursion + strong-Wolfe line search); here it is written out: two-loop
recursion over the last ``m`` curvature pairs, backtracking Armijo line search with a curvature guard (pairs with
``sᵀy ≤ 0`` are skipped, which keeps the inverse-Hessian estimate positive definite).

Every cost evaluation is ONE job over the cached data (the same seqOp as ``run_mini_batch_sgd``: vectorised
``Gradient.batch`` on CPU shards, the fused CUDA gradient kernel on HBM-resident shards).
"""
from __future__ import annotations

import logging
import time
from typing import List, Tuple

import torch

from ..core.broadcast import Broadcast
from .gradient import Gradient
from .gradient_descent import Optimizer, _partition_grad
from .updater import Updater

log = logging.getLogger("asyncframework_b200.lbfgs")


class LBFGS(Optimizer):
    def __init__(self, gradient: Gradient, updater: Updater):
        self.gradient, self.updater = gradient, updater
        self.num_corrections = 10
        self.convergence_tol = 1e-6
        self.max_num_iterations = 100
        self.reg_param = 0.0
        self.loss_history: List[float] = []
        self.warray: List[Tuple[int, torch.Tensor]] = []
        self.cost_evaluations = 0

    def set_num_corrections(self, m: int):
        if m <= 0:
            raise ValueError(f"Number of corrections must be positive but got {m}")
        self.num_corrections = int(m)
        return self

    def set_convergence_tol(self, t: float):
        if t < 0:
            raise ValueError(f"Convergence tolerance must be nonnegative but got {t}")
        self.convergence_tol = float(t)
        return self

    def set_max_num_iterations(self, n: int): self.max_num_iterations = int(n); return self
    set_num_iterations = set_max_num_iterations
    def set_reg_param(self, r: float): self.reg_param = float(r); return self
    def set_gradient(self, g: Gradient): self.gradient = g; return self
    def set_updater(self, u: Updater): self.updater = u; return self

    def get_all_weights(self):
        return self.warray

    # ------------------------------------------------------------------ cost function
    def _cost(self, data, w: torch.Tensor, n: int) -> Tuple[float, torch.Tensor]:
        self.cost_evaluations += 1
        gradient = self.gradient
        bc = Broadcast(w, snapshot=False)

        def seq(acc, shard, bc=bc):
            out = _partition_grad(gradient, shard, bc.value)
            if out is None:
                return acc
            g, l, _ = out
            if acc is None:
                return [g.clone(), float(l)]
            acc[0] += g
            acc[1] += float(l)
            return acc

        def comb(a, b):
            if a is None:
                return b
            if b is None:
                return a
            a[0] += b[0].to(a[0].device)
            a[1] += b[1]
            return a
        agg = data.map_batch(lambda s: s).tree_aggregate(lambda: None, seq, comb)
        bc.destroy()
        grad_sum, loss_sum = agg if agg is not None else (torch.zeros_like(w), 0.0)
        zero = torch.zeros_like(w)
        _, reg_val = self.updater.compute(w, zero, 0.0, 1, self.reg_param)
        w_reg, _ = self.updater.compute(w, zero, 1.0, 1, self.reg_param)
        grad = grad_sum.to(w.device, w.dtype) / n + (w - w_reg)
        return loss_sum / n + float(reg_val), grad

    # ------------------------------------------------------------------ optimize
    def optimize(self, data, initial_weights: torch.Tensor) -> torch.Tensor:
        n = data.count()
        self.loss_history, self.warray, self.cost_evaluations = [], [], 0
        if n == 0:
            log.warning("LBFGS.optimize returning initial weights, no data found")
            return initial_weights
        w = initial_weights.clone()
        f, g = self._cost(data, w, n)
        self.loss_history.append(f)
        self.warray.append((int(time.time() * 1e3), w.clone()))
        S: List[torch.Tensor] = []
        Y: List[torch.Tensor] = []
        for it in range(1, self.max_num_iterations + 1):
            gnorm = float(torch.linalg.vector_norm(g.double()))
            if gnorm <= 1e-12:
                break
            # two-loop recursion: d = -H g
            q = g.clone()
            alphas = []
            for s, y in zip(reversed(S), reversed(Y)):
                rho = 1.0 / float(torch.dot(y.double(), s.double()))
                a = rho * float(torch.dot(s.double(), q.double()))
                alphas.append((a, rho))
                q.add_(y, alpha=-a)
            if S:
                gamma = float(torch.dot(S[-1].double(), Y[-1].double())) / float(torch.dot(Y[-1].double(), Y[-1].double()))
                q.mul_(gamma)
            for (a, rho), s, y in zip(reversed(alphas), S, Y):
                b = rho * float(torch.dot(y.double(), q.double()))
                q.add_(s, alpha=a - b)
            d = -q
            slope = float(torch.dot(g.double(), d.double()))
            if slope >= 0:                                  # not a descent direction (numerical trouble): restart from steepest descent
                S.clear()
                Y.clear()
                d, slope = -g, -gnorm * gnorm
            step = 1.0 if S else min(1.0, 1.0 / gnorm)
            f_new, g_new, w_new = f, g, w
            for _ in range(30):                             # backtracking Armijo
                w_new = w + step * d
                f_new, g_new = self._cost(data, w_new, n)
                if f_new <= f + 1e-4 * step * slope:
                    break
                step *= 0.5
            else:
                log.info("LBFGS line search failed at iteration %d", it)
                break
            s_k, y_k = w_new - w, g_new - g
            if float(torch.dot(s_k.double(), y_k.double())) > 1e-12:
                S.append(s_k)
                Y.append(y_k)
                if len(S) > self.num_corrections:
                    S.pop(0)
                    Y.pop(0)
            improvement = abs(f - f_new) / max(abs(f), abs(f_new), 1e-12)
            w, f, g = w_new, f_new, g_new
            self.loss_history.append(f)
            self.warray.append((int(time.time() * 1e3), w.clone()))
            if improvement < self.convergence_tol:           # breeze: relative improvement of the objective
                break
        log.info("LBFGS.runLBFGS finished. Last 10 losses %s", self.loss_history[-10:])
        return w


def run_lbfgs(data, gradient: Gradient, updater: Updater, num_corrections: int, convergence_tol: float, max_num_iterations: int,
              reg_param: float, initial_weights: torch.Tensor):
    """``LBFGS.runLBFGS`` -> (weights, loss history)."""
    opt = (LBFGS(gradient, updater).set_num_corrections(num_corrections).set_convergence_tol(convergence_tol)
           .set_max_num_iterations(max_num_iterations).set_reg_param(reg_param))
    w = opt.optimize(data, initial_weights)
    return w, opt.loss_history
