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

    def __str__(self) -> str:
        """Spark's text form ``[v0,v1,...]`` (what ``RDD[Vector].saveAsTextFile`` writes and ``Vectors.parse`` reads)."""
        return "[" + ",".join(repr(float(v)) for v in self.values.tolist()) + "]"


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

    def __str__(self) -> str:
        """Spark's text form ``(size,[i0,i1,...],[v0,v1,...])``."""
        return (f"({self._size},[" + ",".join(str(int(i)) for i in self.indices.tolist()) + "],["
                + ",".join(repr(float(v)) for v in self.values.tolist()) + "])")


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
    def parse(s: str, dtype: torch.dtype = torch.float64) -> Vector:
        """``Vectors.parse`` (linalg/Vectors.scala:389-410, util/NumericParser.scala): ``[1.0,2.0]`` -> dense,
        ``(4,[0,2],[1.0,3.0])`` -> sparse."""
        return _vector_from_parsed(_parse_numeric(s), dtype)

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

    def __str__(self) -> str:
        """Spark's text form ``(label,features)`` (regression/LabeledPoint.scala:41-43)."""
        return f"({float(self.label)!r},{self.features})"

    @staticmethod
    def parse(s: str, dtype: torch.dtype = torch.float64) -> "LabeledPoint":
        """``LabeledPoint.parse`` (regression/LabeledPoint.scala:60-74): ``(label,[...])`` / ``(label,(size,[..],[..]))`` or the
        pre-1.0 dense form ``label,f0 f1 f2``."""
        s = s.strip()
        if s.startswith("("):
            parsed = _parse_numeric(s)
            if not (isinstance(parsed, tuple) and len(parsed) == 2 and isinstance(parsed[0], float)):
                raise ValueError(f"Cannot parse {s!r} as a labeled point")
            return LabeledPoint(parsed[0], _vector_from_parsed(parsed[1], dtype))
        label, _, rest = s.partition(",")
        return LabeledPoint(float(label), DenseVector([float(t) for t in rest.split()], dtype=dtype))


def _parse_numeric(s: str):
    """The grammar of util/NumericParser.scala: numbers, ``[a,b,...]`` arrays (-> list of floats) and ``(x,y,...)`` tuples
    (-> tuple), nested.  Raises ``ValueError`` on anything else."""
    s = s.strip()
    pos = 0

    def fail(msg):
        raise ValueError(f"Cannot parse {s!r}: {msg} at position {pos}")

    def skip():
        nonlocal pos
        while pos < len(s) and s[pos].isspace():
            pos += 1

    def value():
        nonlocal pos
        skip()
        if pos >= len(s):
            fail("unexpected end")
        c = s[pos]
        if c in "[(":
            close = "]" if c == "[" else ")"
            pos += 1
            items = []
            skip()
            if pos < len(s) and s[pos] == close:
                pos += 1
                return items if c == "[" else tuple(items)
            while True:
                items.append(value())
                skip()
                if pos >= len(s):
                    fail(f"missing {close!r}")
                if s[pos] == ",":
                    pos += 1
                    continue
                if s[pos] == close:
                    pos += 1
                    break
                fail(f"unexpected {s[pos]!r}")
            if c == "[":
                if not all(isinstance(x, float) for x in items):
                    fail("arrays hold numbers only")
                return items
            return tuple(items)
        start = pos
        while pos < len(s) and s[pos] not in ",[]()" and not s[pos].isspace():
            pos += 1
        try:
            return float(s[start:pos])
        except ValueError:
            fail(f"bad number {s[start:pos]!r}")

    out = value()
    skip()
    if pos != len(s):
        fail("trailing characters")
    return out


def _vector_from_parsed(parsed, dtype: torch.dtype = torch.float64) -> Vector:
    if isinstance(parsed, list):
        return DenseVector(parsed, dtype=dtype)
    if isinstance(parsed, tuple) and len(parsed) == 3 and isinstance(parsed[0], float) and isinstance(parsed[1], list) and isinstance(parsed[2], list):
        return SparseVector(int(parsed[0]), [int(i) for i in parsed[1]], parsed[2], dtype=dtype)
    raise ValueError(f"Cannot parse {parsed!r} as a vector")
