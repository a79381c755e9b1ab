"""Updater kernels (K8) on CUDA tensors: simple / L1-prox / squared-L2 (csrc/kernels/level1.cu)."""
from __future__ import annotations

from typing import Tuple

import torch

from . import check, current_stream_ptr, kernels
from .level1 import _dscratch


def _run(mode: int, w: torch.Tensor, g: torch.Tensor, eta: float, lam: float, want_norm: bool):
    w32 = w.to(torch.float32).contiguous()
    g32 = g.to(torch.float32).contiguous()
    out = torch.empty_like(w32)
    norm = _dscratch(w.device) if want_norm else None
    check(kernels().afb_updater(mode, w32.data_ptr(), g32.data_ptr(), out.data_ptr(), w32.numel(), float(eta), float(lam),
                                norm.data_ptr() if norm is not None else None, current_stream_ptr()), "updater")
    return out.to(w.dtype), (float(norm.item()) if norm is not None else 0.0)


def simple_update(w: torch.Tensor, g: torch.Tensor, eta: float) -> torch.Tensor:
    return _run(0, w, g, eta, 0.0, False)[0]


def l1_update(w: torch.Tensor, g: torch.Tensor, eta: float, lam: float) -> Tuple[torch.Tensor, float]:
    return _run(1, w, g, eta, lam, True)


def l2_update(w: torch.Tensor, g: torch.Tensor, eta: float, lam: float) -> Tuple[torch.Tensor, float]:
    return _run(2, w, g, eta, lam, True)
