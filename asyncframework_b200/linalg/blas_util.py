"""Public level-1 BLAS helpers that return the mutated vector so calls can be chained.

Parity: ``BLASUtil.axpyOp/dotOp/scalOp`` (mllib/BLASUtil.scala:6-18) wrapping
``BLAS.axpy`` (dense->dense BLAS.scala:66-69, sparse->dense scatter :74-93),
``BLAS.dot`` (dense :126-129, sparse·dense gather :134-147, sparse·sparse merge :152-176),
``BLAS.scal`` (:220-229) and ``BLAS.copy`` (:181-215).

On CUDA tensors the dense variants dispatch to the hand-written sm_100a level-1 kernels in
``asyncframework_b200.ops`` (fused axpy / dot / scal); on CPU they are plain torch ops in the
vector's dtype (fp64 by default, matching the reference).
"""
from __future__ import annotations

import torch

from .vectors import DenseVector, SparseVector, Vector


def _require_same_size(x: Vector, y: Vector) -> None:
    if x.size != y.size:
        # BLAS.scala:45,106 `require(x.size == y.size)`
        raise ValueError(f"vector size mismatch: {x.size} vs {y.size}")


def axpy(a: float, x: Vector, y: DenseVector) -> None:
    """y += a * x  (y must be dense, as in the reference)."""
    _require_same_size(x, y)
    if not isinstance(y, DenseVector):
        raise TypeError("axpy only supports adding to a DenseVector")
    if isinstance(x, SparseVector):
        if x.values.numel():
            y.values.index_add_(0, x.indices.long().to(y.values.device),
                                (a * x.values).to(dtype=y.values.dtype, device=y.values.device))
    else:
        xv = x.to_dense().values
        if y.values.is_cuda:
            from ..ops import level1
            level1.axpy_(a, xv.to(y.values.dtype), y.values)
        else:
            y.values.add_(xv.to(y.values.dtype), alpha=a)


def dot(x: Vector, y: Vector) -> float:
    _require_same_size(x, y)
    if isinstance(x, SparseVector) and isinstance(y, SparseVector):
        # sparse·sparse merge
        xi, yi = x.indices.long(), y.indices.long()
        common, xpos, ypos = _intersect(xi, yi)
        if common == 0:
            return 0.0
        return float((x.values[xpos].double() * y.values[ypos].double()).sum())
    if isinstance(x, SparseVector):
        return float((x.values.double() * y.to_dense().values[x.indices.long()].double()).sum())
    if isinstance(y, SparseVector):
        return dot(y, x)
    xv, yv = x.to_dense().values, y.to_dense().values
    if xv.is_cuda and yv.is_cuda:
        from ..ops import level1
        return float(level1.dot(xv, yv))
    return float(torch.dot(xv.double(), yv.double()))


def _intersect(a: torch.Tensor, b: torch.Tensor):
    mask_a = torch.isin(a, b)
    mask_b = torch.isin(b, a)
    xpos = torch.nonzero(mask_a).reshape(-1)
    ypos = torch.nonzero(mask_b).reshape(-1)
    return int(xpos.numel()), xpos, ypos


def scal(a: float, x: Vector) -> None:
    """x *= a (in place; works for sparse and dense)."""
    if isinstance(x, (DenseVector, SparseVector)):
        if x.values.is_cuda and isinstance(x, DenseVector):
            from ..ops import level1
            level1.scal_(a, x.values)
        else:
            x.values.mul_(a)
    else:
        raise TypeError(type(x))


def copy(x: Vector, y: DenseVector) -> None:
    _require_same_size(x, y)
    y.values.copy_(x.to_dense().values.to(y.values.dtype))


# --- chaining wrappers (the ASYNC additions) -------------------------------------------

def axpy_op(a: float, x: Vector, y: DenseVector) -> DenseVector:
    axpy(a, x, y)
    return y


def dot_op(x: Vector, y: Vector) -> float:
    return dot(x, y)


def scal_op(a: float, x: Vector) -> Vector:
    scal(a, x)
    return x
