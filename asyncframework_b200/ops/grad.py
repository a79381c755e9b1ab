"""Standalone fused gradient (K1-K4, K7, K11) on an HBM-resident shard: the round kernel in
``grad_only`` mode.  Used by the MLlib-style ``GradientDescent`` GPU path, the NCCL backend and
the numerics tests.  Returns ``(g[d], loss_sum, sampled_rows)``.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ..data.batch import Shard
from .round import RoundKernel, ServerLayout, ServerPointers

def _kernel_for(shard: Shard, loss_id: int) -> RoundKernel:
    """One cached grad-only RoundKernel per shard (kept in ``shard.meta`` so it dies with the shard)."""
    import os
    variant = os.environ.get("AFB_DENSE_VARIANT", "simple")
    rk = shard.meta.get("_grad_kernel")
    if rk is not None and shard.is_dense and rk.variant != variant and not (variant in ("pipe", "ring") and rk.variant == "simple" and rk.ld > 896):
        rk = None
    if rk is None:
        ld = shard.ld if shard.is_dense else (shard.num_features + 3) // 4 * 4
        layout = ServerLayout(ld, 1, ring_size=1)
        rk = RoundKernel(shard, ServerPointers(0, layout), algo="grad_only", loss_id=loss_id, num_workers=1,
                         worker_id=0, total_rows=max(shard.num_rows, 1), batch_rate=1.0, gamma=0.0, seed=0)
        shard.meta["_grad_kernel"] = rk
    rk.params.loss_id = int(loss_id)
    return rk


def grad(shard: Shard, w: torch.Tensor, loss_id: int, fraction: float = 1.0, seed: int = 0,
         history: Optional[torch.Tensor] = None, stream=None) -> Tuple[torch.Tensor, float, int]:
    """Σ_{i∈sample} (coeff_i − hist_i)·x_i, Σ loss_i, |sample| with Bernoulli(fraction) sampling
    keyed by ``seed`` (core/sampler.py).  ``history`` (fp32[n_rows]) turns on the SAGA
    correction and is updated in place with the new residuals."""
    from ..core.sampler import threshold24
    rk = _kernel_for(shard, loss_id)
    d = shard.num_features
    rk.w_local.zero_()
    rk.w_local[:d].copy_(w.to(torch.float32))
    rk.g_local.zero_()
    rk.params.thr24 = threshold24(fraction)
    rk.params.seed = int(seed)
    rk.params.hist = history.data_ptr() if history is not None else None
    rk.launch(stream)
    ctl = rk.read_ctl()
    return rk.g_local[:d].clone(), float(ctl["last_loss"]), int(ctl["last_batch"])


def grad_full(shard: Shard, w: torch.Tensor, loss_id: int):
    return grad(shard, w, loss_id, 1.0, 0)
