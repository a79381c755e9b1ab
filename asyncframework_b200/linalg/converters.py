"""Vector <-> tensor / ndarray conversions (the role of ``BreezeConverters``,
ASYNCsamples/.../BreezeConverters.scala:9-28: implicit mllib.Vector <-> breeze conversions so the
samples can do dense algebra on ``w``).  Here the algebra library is torch."""
from __future__ import annotations

import numpy as np
import torch

from .vectors import DenseVector, SparseVector, Vector


def to_torch(v, dtype: torch.dtype = None, device=None) -> torch.Tensor:
    if isinstance(v, Vector):
        t = v.to_dense().values
    elif isinstance(v, torch.Tensor):
        t = v
    else:
        t = torch.as_tensor(np.asarray(v))
    if dtype is not None or device is not None:
        t = t.to(dtype=dtype or t.dtype, device=device or t.device)
    return t


def to_numpy(v) -> np.ndarray:
    return to_torch(v).detach().cpu().numpy()


def from_torch(t: torch.Tensor, sparse: bool = False) -> Vector:
    dv = DenseVector(t.reshape(-1), dtype=t.dtype)
    return dv.to_sparse() if sparse else dv


def to_dense(v) -> DenseVector:
    return v.to_dense() if isinstance(v, Vector) else DenseVector(to_torch(v))


def to_sparse(v) -> SparseVector:
    return v.to_sparse() if isinstance(v, (DenseVector, SparseVector)) else to_dense(v).to_sparse()
