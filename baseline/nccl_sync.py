"""The comparison baseline named by BASELINE.json: a plain PyTorch + NCCL(+cuBLAS) re-implementation
of the reference's synchronous drivers (SparkASGDSync / SparkASAGASync semantics, Appendix A.3/A.4
of SURVEY.md) — the Spark/JVM reference itself cannot run on GPUs.

Per iteration: ``dist.broadcast(w)`` (driver -> workers), Bernoulli(b) row sample, gather the rows
(``X[idx]``), ``X_b @ w`` and ``X_b.T @ r`` through cuBLAS, ``dist.all_reduce`` of the gradient,
update on every rank.  No hand-written kernels, no fusion: this is "what you get from stock
PyTorch", and is the bar the fused backend is measured against.
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.distributed as dist


class NcclSyncSGD:
    def __init__(self, X: torch.Tensor, y: torch.Tensor, d: int, total_rows: int, batch_rate: float, gamma: float,
                 seed: int = 42, loss: str = "logistic", saga: bool = False):
        self.X, self.y, self.d = X, y, d
        self.N, self.b, self.gamma, self.seed = total_rows, batch_rate, gamma, seed
        self.loss = loss
        self.saga = saga
        self.w = torch.zeros(X.shape[1], dtype=torch.float32, device=X.device)
        self.abar = torch.zeros_like(self.w)
        self.hist = torch.zeros(X.shape[0], dtype=torch.float32, device=X.device) if saga else None
        self.k = 0
        self.gen = torch.Generator(device=X.device)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.last_loss = torch.zeros((), device=X.device)

    def step(self) -> None:
        if self.world > 1:
            dist.broadcast(self.w, 0)
        self.gen.manual_seed(self.seed + self.k + 1 + 7919 * (dist.get_rank() if self.world > 1 else 0))
        mask = torch.rand(self.X.shape[0], generator=self.gen, device=self.X.device) < self.b
        idx = mask.nonzero(as_tuple=True)[0]
        Xb = self.X[idx]
        if Xb.dtype != self.w.dtype:             # bf16 row storage: gather in bf16, widen for the fp32 GEMVs (what stock PyTorch offers)
            Xb = Xb.to(self.w.dtype)
        yb = self.y[idx]
        margin = Xb @ self.w
        if self.loss == "logistic":
            r = torch.sigmoid(margin) - yb
            self.last_loss = torch.nn.functional.softplus(-margin).sum() + (margin * (1 - yb)).sum()
        else:
            r = margin - yb
            self.last_loss = (r * r).sum()
        eff = r
        if self.saga:
            eff = r - self.hist[idx]
            self.hist[idx] = r
        g = Xb.t() @ eff
        if self.world > 1:
            dist.all_reduce(g)
        if self.saga:
            self.w -= self.gamma * (g / (self.b * self.N) + self.abar)
            self.abar += g / self.N
        else:
            self.w -= (self.gamma / math.sqrt(self.k + 1)) * g / (self.b * self.N)
        self.k += 1
