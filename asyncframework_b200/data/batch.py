"""``Shard``: the rows of one partition, resident on one device, dense or CSR.

This replaces "a cached RDD partition of LabeledPoint objects" (core/rdd/RDD.scala:331-345) with
columnar tensors: ``X`` (dense ``[n, ld]`` row-major, ``ld`` = d rounded up to 4 floats so every
row starts 16-byte aligned for vector loads / bulk copies) or CSR ``(indptr, indices, values)``,
labels ``y[n]`` and the global row index ``gidx[n]`` produced by ``zip_with_index``.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Iterator, Optional, Tuple

import torch

from ..linalg.vectors import DenseVector, LabeledPoint, SparseVector


def padded_ld(d: int) -> int:
    return (d + 3) // 4 * 4


@dataclass
class Shard:
    num_features: int
    y: torch.Tensor                              # [n]
    X: Optional[torch.Tensor] = None             # dense [n, ld]
    indptr: Optional[torch.Tensor] = None        # CSR int64 [n+1]
    indices: Optional[torch.Tensor] = None       # CSR int32 [nnz]
    values: Optional[torch.Tensor] = None        # CSR [nnz]
    gidx: Optional[torch.Tensor] = None          # int64 [n] global row index (zip_with_index)
    partition: int = 0
    meta: dict = field(default_factory=dict)

    # ------------------------------------------------------------------
    @property
    def is_dense(self) -> bool:
        return self.X is not None

    @property
    def num_rows(self) -> int:
        return int(self.y.numel())

    @property
    def device(self) -> torch.device:
        return self.y.device

    @property
    def dtype(self) -> torch.dtype:
        return self.X.dtype if self.is_dense else self.values.dtype

    @property
    def ld(self) -> int:
        return int(self.X.shape[1]) if self.is_dense else self.num_features

    @property
    def nnz(self) -> int:
        if self.is_dense:
            return self.num_rows * self.num_features
        return int(self.values.numel())

    def nbytes(self) -> int:
        tot = self.y.numel() * self.y.element_size()
        for t in (self.X, self.indptr, self.indices, self.values, self.gidx):
            if t is not None:
                tot += t.numel() * t.element_size()
        return tot

    # ------------------------------------------------------------------
    def to(self, device=None, dtype: Optional[torch.dtype] = None, non_blocking: bool = False) -> "Shard":
        def mv(t, dt=None):
            if t is None:
                return None
            return t.to(device=device, dtype=dt if dt is not None else t.dtype, non_blocking=non_blocking)
        return Shard(self.num_features, mv(self.y, dtype), mv(self.X, dtype), mv(self.indptr), mv(self.indices),
                     mv(self.values, dtype), mv(self.gidx), self.partition, dict(self.meta))

    def pin(self) -> "Shard":
        def pn(t):
            return None if t is None else (t.pin_memory() if not t.is_cuda and torch.cuda.is_available() else t)
        return Shard(self.num_features, pn(self.y), pn(self.X), pn(self.indptr), pn(self.indices),
                     pn(self.values), pn(self.gidx), self.partition, dict(self.meta))

    def select(self, rows: torch.Tensor) -> "Shard":
        """Sub-shard with the given local row positions (ascending)."""
        rows = rows.to(self.device)
        g = self.gidx[rows] if self.gidx is not None else None
        if self.is_dense:
            return Shard(self.num_features, self.y[rows], self.X[rows], gidx=g, partition=self.partition)
        starts = self.indptr[rows]
        ends = self.indptr[rows + 1]
        lens = ends - starts
        new_indptr = torch.zeros(rows.numel() + 1, dtype=torch.int64, device=self.device)
        new_indptr[1:] = torch.cumsum(lens, 0)
        total = int(new_indptr[-1])
        if total:
            # positions of every kept nnz: start[row] + offset within row
            row_of = torch.repeat_interleave(torch.arange(rows.numel(), device=self.device), lens)
            off = torch.arange(total, device=self.device) - new_indptr[:-1][row_of]
            pos = starts[row_of] + off
            idx, val = self.indices[pos], self.values[pos]
        else:
            idx, val = self.indices[:0], self.values[:0]
        return Shard(self.num_features, self.y[rows], None, new_indptr, idx, val, g, self.partition)

    def to_dense(self) -> "Shard":
        if self.is_dense:
            return self
        n, d = self.num_rows, self.num_features
        X = torch.zeros(n, padded_ld(d), dtype=self.values.dtype, device=self.device)
        if self.nnz:
            lens = self.indptr[1:] - self.indptr[:-1]
            row_of = torch.repeat_interleave(torch.arange(n, device=self.device), lens)
            X[row_of, self.indices.long()] = self.values
        return Shard(d, self.y, X, gidx=self.gidx, partition=self.partition, meta=dict(self.meta))

    def to_csr(self) -> "Shard":
        if not self.is_dense:
            return self
        Xd = self.X[:, : self.num_features]
        nzmask = Xd != 0
        lens = nzmask.sum(1)
        indptr = torch.zeros(self.num_rows + 1, dtype=torch.int64, device=self.device)
        indptr[1:] = torch.cumsum(lens, 0)
        rr, cc = torch.nonzero(nzmask, as_tuple=True)
        return Shard(self.num_features, self.y, None, indptr, cc.to(torch.int32), Xd[rr, cc], self.gidx,
                     self.partition, dict(self.meta))

    def matvec(self, w: torch.Tensor) -> torch.Tensor:
        """X @ w in the shard's dtype (plain torch; the oracle path)."""
        if self.is_dense:
            return self.X[:, : self.num_features] @ w.to(self.X.dtype)
        prod = self.values * w.to(self.values.dtype)[self.indices.long()]
        out = torch.zeros(self.num_rows, dtype=self.values.dtype, device=self.device)
        if self.nnz:
            lens = self.indptr[1:] - self.indptr[:-1]
            row_of = torch.repeat_interleave(torch.arange(self.num_rows, device=self.device), lens)
            out.index_add_(0, row_of, prod)
        return out

    def rmatvec(self, r: torch.Tensor) -> torch.Tensor:
        """Xᵀ @ r (length d)."""
        if self.is_dense:
            return (self.X[:, : self.num_features].T @ r.to(self.X.dtype))
        out = torch.zeros(self.num_features, dtype=self.values.dtype, device=self.device)
        if self.nnz:
            lens = self.indptr[1:] - self.indptr[:-1]
            row_of = torch.repeat_interleave(torch.arange(self.num_rows, device=self.device), lens)
            out.index_add_(0, self.indices.long(), self.values * r.to(self.values.dtype)[row_of])
        return out

    # record-level view (slow; API parity / tests)
    def row(self, i: int) -> LabeledPoint:
        if self.is_dense:
            return LabeledPoint(float(self.y[i]), DenseVector(self.X[i, : self.num_features], dtype=self.X.dtype))
        s, e = int(self.indptr[i]), int(self.indptr[i + 1])
        return LabeledPoint(float(self.y[i]), SparseVector(self.num_features, self.indices[s:e], self.values[s:e],
                                                           dtype=self.values.dtype))

    def records(self) -> Iterator[LabeledPoint]:
        for i in range(self.num_rows):
            yield self.row(i)

    def indexed_records(self) -> Iterator[Tuple[LabeledPoint, int]]:
        for i in range(self.num_rows):
            yield self.row(i), (int(self.gidx[i]) if self.gidx is not None else i)


def concat_shards(shards, partition: int = 0) -> Shard:
    shards = [s for s in shards if s is not None]
    if not shards:
        raise ValueError("no shards")
    d = shards[0].num_features
    y = torch.cat([s.y for s in shards])
    g = torch.cat([s.gidx for s in shards]) if all(s.gidx is not None for s in shards) else None
    if all(s.is_dense for s in shards):
        return Shard(d, y, torch.cat([s.X for s in shards]), gidx=g, partition=partition)
    shards = [s.to_csr() for s in shards]
    indptrs = [shards[0].indptr]
    off = int(shards[0].indptr[-1])
    for s in shards[1:]:
        indptrs.append(s.indptr[1:] + off)
        off += int(s.indptr[-1])
    return Shard(d, y, None, torch.cat(indptrs), torch.cat([s.indices for s in shards]),
                 torch.cat([s.values for s in shards]), g, partition)
