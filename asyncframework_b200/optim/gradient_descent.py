"""Synchronous mini-batch SGD (the MLlib baseline) with the ASYNC weight-history hooks.

Parity: ``GradientDescent.runMiniBatchSGD`` (mllib/optimization/GradientDescent.scala:197-310):
per iteration broadcast weights (:263), ``data.sample(false, frac, 42+i).treeAggregate`` of
(gradSum, lossSum, count) (:267-277), ``updater.compute(w, gradSum/miniBatchSize, stepSize, i,
regParam)`` (:287-291), loss history (:286), convergence test ‖Δw‖ < tol·max(‖w‖,1) (:328-340).
ASYNC additions (A11): ``Warray`` snapshots ``(wall-clock ms, weights)`` at i == 1 and every 100th
iteration (:257-260) surfaced through ``Optimizer.get_all_weights`` (Optimizer.scala:39-40).

Data is a batch-capable ``ShardedDataset``; the per-partition seqOp is the vectorised
``Gradient.batch`` on CPU tensors or the fused CUDA gradient kernel on HBM-resident shards.
"""
from __future__ import annotations

import logging
import time
from typing import List, Optional, Tuple

import torch

from ..core.broadcast import Broadcast
from .gradient import Gradient
from .updater import Updater

log = logging.getLogger("asyncframework_b200.gd")


class Optimizer:
    def optimize(self, data, initial_weights: torch.Tensor) -> torch.Tensor:
        raise NotImplementedError

    def get_all_weights(self) -> List[Tuple[int, torch.Tensor]]:
        """[(wall-clock ms, weights)] snapshots (``Optimizer.getAllWeights``)."""
        raise NotImplementedError


def _is_converged(prev: torch.Tensor, cur: torch.Tensor, tol: float) -> bool:
    diff = float(torch.linalg.vector_norm((prev - cur).double()))
    return diff < tol * max(float(torch.linalg.vector_norm(cur.double())), 1.0)


def _partition_grad(gradient: Gradient, shard, w: torch.Tensor):
    """seqOp over one (already sampled) shard -> (gradSum, lossSum, count)."""
    if shard is None or shard.num_rows == 0:
        return None
    if shard.device.type == "cuda":
        from ..ops.grad import grad_full          # one entry point: dispatches on the shard layout (dense / CSR round kernel)
        g, loss, cnt = grad_full(shard, w, gradient.loss_id)
        return g, loss, cnt
    g, loss, cnt, _ = gradient.batch(shard, w)
    return g, float(loss), cnt


class GradientDescent(Optimizer):
    def __init__(self, gradient: Gradient, updater: Updater):
        self.gradient = gradient
        self.updater = updater
        self.step_size = 1.0
        self.num_iterations = 100
        self.reg_param = 0.0
        self.mini_batch_fraction = 1.0
        self.convergence_tol = 0.001
        self.snapshot_every = 100
        self.seed_base = 42
        self.warray: List[Tuple[int, torch.Tensor]] = []
        self.loss_history: List[float] = []

    # builder-style setters, names as in the reference
    def set_step_size(self, s: float): self.step_size = float(s); return self
    def set_num_iterations(self, n: int): self.num_iterations = int(n); return self
    def set_reg_param(self, r: float): self.reg_param = float(r); return self
    def set_mini_batch_fraction(self, f: float): self.mini_batch_fraction = float(f); return self
    def set_convergence_tol(self, t: float): self.convergence_tol = float(t); return self
    def set_gradient(self, g: Gradient): self.gradient = g; return self
    def set_updater(self, u: Updater): self.updater = u; return self

    def optimize(self, data, initial_weights: torch.Tensor) -> torch.Tensor:
        w, losses, hist = run_mini_batch_sgd(
            data, self.gradient, self.updater, self.step_size, self.num_iterations, self.reg_param,
            self.mini_batch_fraction, initial_weights, self.convergence_tol, self.snapshot_every, self.seed_base)
        self.warray = hist
        self.loss_history = losses
        return w

    def get_all_weights(self):
        return self.warray


def run_mini_batch_sgd(data, gradient: Gradient, updater: Updater, step_size: float, num_iterations: int,
                       reg_param: float, mini_batch_fraction: float, initial_weights: torch.Tensor,
                       convergence_tol: float = 0.001, snapshot_every: int = 100, seed_base: int = 42):
    """Returns (weights, stochastic loss history, [(wall-clock ms, weights)] snapshots)."""
    if mini_batch_fraction < 1.0 and convergence_tol > 0.0:
        log.warning("Testing against a convergenceTol when using miniBatchFraction < 1.0 can be unstable "
                    "because of the stochasticity in sampling.")
    if num_iterations * mini_batch_fraction < 1.0:
        log.warning("Not all examples will be used if numIterations * miniBatchFraction < 1.0")
    losses: List[float] = []
    warray: List[Tuple[int, torch.Tensor]] = []
    num_examples = data.count()
    if num_examples == 0:
        log.warning("GradientDescent.runMiniBatchSGD returning initial weights, no data found")
        return initial_weights, losses, warray
    weights = initial_weights.clone()
    prev: Optional[torch.Tensor] = None
    cur: Optional[torch.Tensor] = None
    # regVal of the initial weights (the reference runs updater.compute with a zero gradient, step 0)
    _, reg_val = updater.compute(weights, torch.zeros_like(weights), 0.0, 1, reg_param)
    converged = False
    i = 1
    while not converged and i <= num_iterations:
        if i == 1 or i % snapshot_every == 0:
            warray.append((int(time.time() * 1e3), weights.detach().clone()))
        bc = Broadcast(weights, snapshot=False)
        sampled = data.sample(False, mini_batch_fraction, seed_base + i) if mini_batch_fraction < 1.0 else data

        def seq(acc, shard, bc=bc):
            out = _partition_grad(gradient, shard, bc.value)
            if out is None:
                return acc
            g, l, c = out
            if acc is None:
                return [g.clone(), float(l), int(c)]
            acc[0] += g
            acc[1] += float(l)
            acc[2] += int(c)
            return acc

        def comb(a, b):
            if a is None:
                return b
            if b is None:
                return a
            a[0] += b[0].to(a[0].device)
            a[1] += b[1]
            a[2] += b[2]
            return a

        agg = sampled.map_batch(lambda s: s).tree_aggregate(lambda: None, seq, comb)
        bc.destroy()
        if agg is not None and agg[2] > 0:
            grad_sum, loss_sum, mb = agg
            losses.append(loss_sum / mb + reg_val)
            weights, reg_val = updater.compute(weights, (grad_sum / mb).to(weights.device), step_size, i, reg_param)
            prev, cur = cur, weights
            if prev is not None and cur is not None and convergence_tol > 0:
                converged = _is_converged(prev, cur, convergence_tol)
        else:
            log.warning("Iteration (%d/%d). The size of sampled batch is zero", i, num_iterations)
        i += 1
    log.info("GradientDescent.runMiniBatchSGD finished. Last 10 stochastic losses %s", losses[-10:])
    return weights, losses, warray
