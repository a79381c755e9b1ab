"""Level-1 BLAS on CUDA tensors (csrc/kernels/level1.cu).  Parity: BLAS.scala axpy/dot/scal."""
from __future__ import annotations

import torch

from . import check, current_stream_ptr, kernels

_scratch = {}


def _dscratch(device) -> torch.Tensor:
    key = (device.index, "d")
    if key not in _scratch:
        _scratch[key] = torch.zeros(1, dtype=torch.float64, device=device)
    return _scratch[key]


def _f32(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise TypeError("CUDA level-1 ops need contiguous float32 tensors")
    return t


def axpy_(a: float, x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    _f32(y)
    x = _f32(x.to(torch.float32).contiguous())
    check(kernels().afb_axpy(float(a), x.data_ptr(), y.data_ptr(), y.numel(), current_stream_ptr()), "axpy")
    return y


def scal_(a: float, x: torch.Tensor) -> torch.Tensor:
    _f32(x)
    check(kernels().afb_scal(float(a), x.data_ptr(), x.numel(), current_stream_ptr()), "scal")
    return x


def dot(x: torch.Tensor, y: torch.Tensor) -> float:
    x = _f32(x.to(torch.float32).contiguous())
    y = _f32(y.to(torch.float32).contiguous())
    out = _dscratch(x.device)
    check(kernels().afb_dot(x.data_ptr(), y.data_ptr(), x.numel(), out.data_ptr(), current_stream_ptr()), "dot")
    return float(out.item())
