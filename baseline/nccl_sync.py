"""The comparison baseline named by BASELINE.json: a plain PyTorch + NCCL(+cuBLAS) re-implementation of the reference's
synchronous drivers (SparkASGDSync / SparkASAGASync semantics, Appendix A.3/A.4 of SURVEY.md) — the Spark/JVM reference
itself cannot run on GPUs.

Per iteration: ``dist.broadcast(w)`` (driver -> workers), Bernoulli(b) row sample, gather the rows (``X[idx]``), ``X_b @ w``
and ``X_b.T @ r`` through cuBLAS, ``dist.all_reduce`` of the gradient, update on every rank.  No hand-written kernels, no
fusion: this is "what you get from stock PyTorch", and is the bar the fused backend is measured against.

Two variants so the bar is not a straw man:

* ``stock``  — ``mask.nonzero()`` builds the index list: the idiomatic spelling, with one device->host sync per step;
* ``nosync`` — ``torch.nonzero_static`` with a fixed capacity (expected batch + 8 sigma, padded slots get a zero
  coefficient): no host synchronisation anywhere in a step, the fastest this design gets without custom kernels.

``step_e2e`` is the end-to-end flavour used by bench.py: a pinned host descriptor (step size) is copied H2D, the step runs,
the scalar loss is copied D2H into pinned memory (read one step behind, like ``FusedEngine.step(lag=1)``).
"""
from __future__ import annotations

import math
from typing import Optional

import torch
import torch.distributed as dist


class NcclSyncSGD:
    def __init__(self, X: torch.Tensor, y: torch.Tensor, d: int, total_rows: int, batch_rate: float, gamma: float,
                 seed: int = 42, loss: str = "logistic", saga: bool = False, variant: str = "stock",
                 w0: Optional[torch.Tensor] = None):
        self.X, self.y, self.d = X, y, d
        self.N, self.b, self.gamma, self.seed = total_rows, batch_rate, gamma, seed
        self.loss = loss
        self.saga = saga
        self.variant = variant
        self.w = torch.zeros(X.shape[1], dtype=torch.float32, device=X.device)
        if w0 is not None:
            self.w[: w0.numel()].copy_(w0.to(torch.float32))
        self.abar = torch.zeros_like(self.w)
        self.hist = torch.zeros(X.shape[0], dtype=torch.float32, device=X.device) if saga else None
        self.k = 0
        self.gen = torch.Generator(device=X.device)
        self.world = dist.get_world_size() if dist.is_initialized() else 1
        self.rank = dist.get_rank() if dist.is_initialized() else 0
        self.last_loss = torch.zeros((), device=X.device)
        n = X.shape[0]
        self.capacity = int(min(n, math.ceil(batch_rate * n + 8.0 * math.sqrt(max(batch_rate * (1 - batch_rate) * n, 1.0)) + 8)))
        # end-to-end staging (pinned): [gamma] in, [loss] out
        if X.is_cuda:
            self._desc_host = torch.zeros(4, dtype=torch.float32).pin_memory()
            self._desc_dev = torch.zeros(4, dtype=torch.float32, device=X.device)
            self._out_host = torch.zeros(2, 4, dtype=torch.float32).pin_memory()
            self._out_events = [torch.cuda.Event(), torch.cuda.Event()]
            self._pending = [False, False]
            self._idx = 0

    def step(self, gamma_dev: Optional[torch.Tensor] = None) -> None:
        if self.world > 1:
            dist.broadcast(self.w, 0)
        self.gen.manual_seed(self.seed + self.k + 1 + 7919 * self.rank)
        mask = torch.rand(self.X.shape[0], generator=self.gen, device=self.X.device) < self.b
        if self.variant == "nosync":
            idx = torch.nonzero_static(mask, size=self.capacity, fill_value=-1).squeeze(1)
            valid = (idx >= 0).to(torch.float32)
            idx = idx.clamp_min(0)
        else:
            idx = mask.nonzero(as_tuple=True)[0]          # device -> host sync (the output size is data dependent)
            valid = None
        Xb = self.X[idx]
        if Xb.dtype != self.w.dtype:             # bf16 row storage: gather in bf16, widen for the fp32 GEMVs (what stock PyTorch offers)
            Xb = Xb.to(self.w.dtype)
        yb = self.y[idx]
        margin = Xb @ self.w
        if self.loss == "logistic":
            r = torch.sigmoid(margin) - yb
            per = torch.nn.functional.softplus(-margin) + margin * (1 - yb)
        else:
            r = margin - yb
            per = r * r
        if valid is not None:
            r = r * valid
            per = per * valid
        self.last_loss = per.sum()
        eff = r
        if self.saga:
            eff = r - self.hist[idx] * (valid if valid is not None else 1.0)
            if valid is None:
                self.hist[idx] = r
            else:
                self.hist.index_put_((idx,), torch.where(valid > 0, r, self.hist[idx]))
        g = Xb.t() @ eff
        if self.world > 1:
            dist.all_reduce(g)
        if self.saga:
            gam = self.gamma if gamma_dev is None else gamma_dev
            self.w -= gam * (g / (self.b * self.N) + self.abar)
            self.abar += g / self.N
        else:
            gam = self.gamma if gamma_dev is None else gamma_dev
            self.w -= (gam / math.sqrt(self.k + 1)) * g / (self.b * self.N)
        self.k += 1

    # ---- end-to-end step: pinned H2D descriptor -> step -> pinned D2H loss (read one step behind) -----------------
    def step_e2e(self) -> Optional[float]:
        slot = self._idx & 1
        if self._pending[slot]:
            self._out_events[slot].synchronize()
        self._desc_host[0] = self.gamma
        self._desc_dev.copy_(self._desc_host, non_blocking=True)
        self.step(gamma_dev=self._desc_dev[0])
        self._out_host[slot, 0:1].copy_(self.last_loss.reshape(1), non_blocking=True)
        self._out_events[slot].record()
        self._pending[slot] = True
        self._idx += 1
        prev = slot ^ 1
        if self._pending[prev]:
            self._out_events[prev].synchronize()
            self._pending[prev] = False
            return float(self._out_host[prev, 0])
        return None

    def flush_e2e(self) -> Optional[float]:
        last = None
        for slot in ((self._idx & 1), (self._idx & 1) ^ 1):
            if self._pending[slot]:
                self._out_events[slot].synchronize()
                self._pending[slot] = False
                last = float(self._out_host[slot, 0])
        return last

    h2d_bytes_per_step = 16
    d2h_bytes_per_step = 4
