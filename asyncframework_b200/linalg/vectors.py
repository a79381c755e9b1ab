"""Vector types (dense / sparse) and ``LabeledPoint``.

Parity targets: ``DenseVector`` / ``SparseVector`` (mllib/linalg/Vectors.scala:618,756),
``Vectors.dense/sparse/zeros`` factory, ``LabeledPoint(label, features)``
(mllib/regression/LabeledPoint.scala:36).  Storage is a torch tensor so that a vector
can live on the CPU (fp64 oracle paths) or in HBM (fp32 GPU paths) with the same API.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Iterable, Sequence, Union

import numpy as np
import torch

ArrayLike = Union[Sequence[float], np.ndarray, torch.Tensor]


def _as_tensor(x: ArrayLike, dtype: torch.dtype, device=None) -> torch.Tensor:
    if isinstance(x, torch.Tensor):
        return x.to(dtype=dtype, device=device) if (x.dtype != dtype or (device is not None and x.device != torch.device(device))) else x
    return torch.as_tensor(np.asarray(x), dtype=dtype, device=device)


class Vector:
    """Common interface: ``size``, ``to_dense()``, ``to_array()``, ``num_nonzeros``."""
    size: int

    def to_dense(self) -> "DenseVector":
        raise NotImplementedError

    def to_array(self) -> np.ndarray:
        return self.to_dense().values.detach().cpu().numpy()

    def __len__(self) -> int:
        return self.size


class DenseVector(Vector):
    __slots__ = ("values",)

    def __init__(self, values: ArrayLike, dtype: torch.dtype = torch.float64, device=None):
        self.values = _as_tensor(values, dtype, device).reshape(-1)

    @property
    def size(self) -> int:  # type: ignore[override]
        return int(self.values.numel())

    @property
    def num_nonzeros(self) -> int:
        return int((self.values != 0).sum().item())

    def to_dense(self) -> "DenseVector":
        return self

    def copy(self) -> "DenseVector":
        return DenseVector(self.values.clone(), dtype=self.values.dtype)

    def to_sparse(self) -> "SparseVector":
        nz = torch.nonzero(self.values, as_tuple=False).reshape(-1)
        return SparseVector(self.size, nz.to(torch.int32), self.values[nz], dtype=self.values.dtype)

    def __getitem__(self, i):
        return self.values[i].item() if isinstance(i, int) else self.values[i]

    def __eq__(self, other) -> bool:
        if isinstance(other, Vector):
            return self.size == other.size and bool(torch.equal(
                self.values.cpu().double(), other.to_dense().values.cpu().double()))
        return NotImplemented

    def __repr__(self) -> str:
        return f"DenseVector({self.values.tolist()})"


class SparseVector(Vector):
    """``size, indices:int32[], values[]`` with strictly increasing indices."""
    __slots__ = ("_size", "indices", "values")

    def __init__(self, size: int, indices: ArrayLike, values: ArrayLike,
                 dtype: torch.dtype = torch.float64, device=None):
        self._size = int(size)
        self.indices = _as_tensor(indices, torch.int32, device).reshape(-1)
        self.values = _as_tensor(values, dtype, device).reshape(-1)
        if self.indices.numel() != self.values.numel():
            raise ValueError("indices and values must have the same length")
        if self.indices.numel():
            if int(self.indices.min()) < 0 or int(self.indices.max()) >= self._size:
                raise ValueError("index out of range")
            if self.indices.numel() > 1 and not bool((self.indices[1:] > self.indices[:-1]).all()):
                raise ValueError("indices must be strictly increasing")

    @property
    def size(self) -> int:  # type: ignore[override]
        return self._size

    @property
    def num_nonzeros(self) -> int:
        return int((self.values != 0).sum().item())

    def to_dense(self) -> DenseVector:
        out = torch.zeros(self._size, dtype=self.values.dtype, device=self.values.device)
        out[self.indices.long()] = self.values
        return DenseVector(out, dtype=self.values.dtype)

    def to_sparse(self) -> "SparseVector":
        return self

    def __eq__(self, other) -> bool:
        if isinstance(other, Vector):
            return self.to_dense() == other.to_dense()
        return NotImplemented

    def __repr__(self) -> str:
        return f"SparseVector({self._size}, {self.indices.tolist()}, {self.values.tolist()})"


class Vectors:
    """Factory methods, named like ``org.apache.spark.mllib.linalg.Vectors``."""

    @staticmethod
    def dense(*values, dtype: torch.dtype = torch.float64, device=None) -> DenseVector:
        if len(values) == 1 and not isinstance(values[0], (int, float)):
            return DenseVector(values[0], dtype=dtype, device=device)
        return DenseVector(list(values), dtype=dtype, device=device)

    @staticmethod
    def sparse(size: int, *args, dtype: torch.dtype = torch.float64, device=None) -> SparseVector:
        """``sparse(size, indices, values)`` or ``sparse(size, [(i, v), ...])`` or ``sparse(size, {i: v})``."""
        if len(args) == 2:
            return SparseVector(size, args[0], args[1], dtype=dtype, device=device)
        if len(args) == 1:
            pairs = args[0]
            if isinstance(pairs, dict):
                pairs = pairs.items()
            pairs = sorted((int(i), float(v)) for i, v in pairs)
            return SparseVector(size, [p[0] for p in pairs], [p[1] for p in pairs], dtype=dtype, device=device)
        raise TypeError("Vectors.sparse(size, indices, values) or Vectors.sparse(size, pairs)")

    @staticmethod
    def zeros(size: int, dtype: torch.dtype = torch.float64, device=None) -> DenseVector:
        return DenseVector(torch.zeros(size, dtype=dtype, device=device), dtype=dtype)

    @staticmethod
    def norm(v: Vector, p: float = 2.0) -> float:
        return float(torch.linalg.vector_norm(v.to_dense().values.double() if isinstance(v, DenseVector) else v.values.double(), ord=p))

    @staticmethod
    def sqdist(a: Vector, b: Vector) -> float:
        d = a.to_dense().values.double() - b.to_dense().values.double()
        return float((d * d).sum())


@dataclass
class LabeledPoint:
    label: float
    features: Vector

    def __iter__(self):
        yield self.label
        yield self.features

    def __repr__(self) -> str:
        return f"LabeledPoint({self.label}, {self.features!r})"
