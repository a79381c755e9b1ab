"""Local matrices + level-2/3 BLAS (``mllib.linalg.{DenseMatrix,SparseMatrix,Matrices}`` and
``BLAS.{gemv,gemm,syr,spr}``, mllib/linalg/Matrices.scala, BLAS.scala:232-237 native path, ``spr`` :244, ``syr`` :288,
``gemm`` :348, ``gemv`` :515).

The reference sends level-2/3 to *native* netlib BLAS (the only native numeric code it has) and the ASYNC samples never
call it; here the dense cases are one torch op (cuBLAS on a GPU — a plain library GEMM, which is what cuBLAS is for),
the sparse cases are CSC gathers.  Storage follows MLlib: dense = column-major values (``is_transposed`` flips to
row-major), sparse = CSC (``col_ptrs``, ``row_indices``, ``values``).
"""
from __future__ import annotations

from typing import Sequence

import numpy as np
import torch

from .vectors import DenseVector, SparseVector, Vector


class Matrix:
    num_rows: int
    num_cols: int

    def to_tensor(self) -> torch.Tensor:
        """Dense ``[num_rows, num_cols]`` tensor."""
        raise NotImplementedError

    def to_array(self) -> np.ndarray:
        return self.to_tensor().cpu().numpy()

    def multiply(self, other):
        """``A.multiply(B | x)`` (Matrix.multiply): matrix·matrix -> DenseMatrix, matrix·vector -> DenseVector."""
        if isinstance(other, Vector):
            y = DenseVector(torch.zeros(self.num_rows, dtype=torch.float64))
            gemv(1.0, self, other, 0.0, y)
            return y
        C = DenseMatrix.zeros(self.num_rows, other.num_cols)
        gemm(1.0, self, other, 0.0, C)
        return C


class DenseMatrix(Matrix):
    def __init__(self, num_rows: int, num_cols: int, values, is_transposed: bool = False, dtype=torch.float64):
        self.num_rows, self.num_cols, self.is_transposed = int(num_rows), int(num_cols), bool(is_transposed)
        self.values = torch.as_tensor(values, dtype=dtype).reshape(-1)
        if self.values.numel() != self.num_rows * self.num_cols:
            raise ValueError(f"The length of the values array must be equal to numRows * numCols = "
                             f"{self.num_rows * self.num_cols}, but is {self.values.numel()}")

    def to_tensor(self) -> torch.Tensor:
        if self.is_transposed:                                       # row-major storage
            return self.values.view(self.num_rows, self.num_cols)
        return self.values.view(self.num_cols, self.num_rows).t()    # column-major: a view, no copy

    @property
    def transpose(self) -> "DenseMatrix":
        return DenseMatrix(self.num_cols, self.num_rows, self.values, not self.is_transposed, self.values.dtype)

    def __getitem__(self, ij):
        i, j = ij
        return float(self.to_tensor()[i, j])

    def to_sparse(self) -> "SparseMatrix":
        return SparseMatrix.from_dense(self.to_tensor())

    def copy(self) -> "DenseMatrix":
        return DenseMatrix(self.num_rows, self.num_cols, self.values.clone(), self.is_transposed, self.values.dtype)

    @staticmethod
    def zeros(r: int, c: int, dtype=torch.float64) -> "DenseMatrix":
        return DenseMatrix(r, c, torch.zeros(r * c, dtype=dtype), dtype=dtype)

    @staticmethod
    def ones(r: int, c: int, dtype=torch.float64) -> "DenseMatrix":
        return DenseMatrix(r, c, torch.ones(r * c, dtype=dtype), dtype=dtype)

    @staticmethod
    def eye(n: int, dtype=torch.float64) -> "DenseMatrix":
        return DenseMatrix(n, n, torch.eye(n, dtype=dtype).reshape(-1), dtype=dtype)

    @staticmethod
    def diag(v: Vector) -> "DenseMatrix":
        d = v.to_dense().values
        return DenseMatrix(d.numel(), d.numel(), torch.diag(d).t().reshape(-1), dtype=d.dtype)

    @staticmethod
    def from_tensor(t: torch.Tensor) -> "DenseMatrix":
        return DenseMatrix(t.shape[0], t.shape[1], t.t().contiguous().reshape(-1), dtype=t.dtype)

    def __repr__(self) -> str:
        return f"DenseMatrix({self.num_rows}x{self.num_cols}{', transposed' if self.is_transposed else ''})"


class SparseMatrix(Matrix):
    """CSC; ``is_transposed`` makes the same three arrays CSR of the transposed shape (MLlib semantics)."""

    def __init__(self, num_rows: int, num_cols: int, col_ptrs, row_indices, values, is_transposed: bool = False,
                 dtype=torch.float64):
        self.num_rows, self.num_cols, self.is_transposed = int(num_rows), int(num_cols), bool(is_transposed)
        self.col_ptrs = torch.as_tensor(col_ptrs, dtype=torch.int64)
        self.row_indices = torch.as_tensor(row_indices, dtype=torch.int64)
        self.values = torch.as_tensor(values, dtype=dtype)
        major = self.num_rows if is_transposed else self.num_cols
        if self.col_ptrs.numel() != major + 1:
            raise ValueError(f"Expecting {major + 1} colPtrs when numCols = {major} but got {self.col_ptrs.numel()}")
        if self.row_indices.numel() != self.values.numel() or int(self.col_ptrs[-1]) != self.values.numel():
            raise ValueError("rowIndices / values / last colPtr disagree on the number of non-zeros")

    def _major_ids(self) -> torch.Tensor:
        counts = self.col_ptrs[1:] - self.col_ptrs[:-1]
        return torch.repeat_interleave(torch.arange(counts.numel()), counts)

    def coo(self):
        """(rows, cols, values) of the logical matrix."""
        major = self._major_ids()
        return (major, self.row_indices, self.values) if self.is_transposed else (self.row_indices, major, self.values)

    def to_tensor(self) -> torch.Tensor:
        r, c, v = self.coo()
        out = torch.zeros(self.num_rows, self.num_cols, dtype=self.values.dtype)
        out.index_put_((r, c), v, accumulate=True)
        return out

    def to_dense(self) -> DenseMatrix:
        return DenseMatrix.from_tensor(self.to_tensor())

    @property
    def transpose(self) -> "SparseMatrix":
        return SparseMatrix(self.num_cols, self.num_rows, self.col_ptrs, self.row_indices, self.values,
                            not self.is_transposed, self.values.dtype)

    @property
    def num_nonzeros(self) -> int:
        return int((self.values != 0).sum())

    @staticmethod
    def from_dense(t: torch.Tensor) -> "SparseMatrix":
        tt = t.t()                                            # iterate column-major
        nz = tt.nonzero(as_tuple=False)                       # sorted by (col, row)
        cols, rows = nz[:, 0], nz[:, 1]
        ptrs = torch.zeros(t.shape[1] + 1, dtype=torch.int64)
        ptrs[1:] = torch.cumsum(torch.bincount(cols, minlength=t.shape[1]), 0)
        return SparseMatrix(t.shape[0], t.shape[1], ptrs, rows, tt[cols, rows], dtype=t.dtype)

    @staticmethod
    def from_coo(num_rows: int, num_cols: int, entries: Sequence) -> "SparseMatrix":
        """``SparseMatrix.fromCOO``: duplicates are summed."""
        dense = torch.zeros(num_rows, num_cols, dtype=torch.float64)
        for i, j, v in entries:
            dense[i, j] += v
        return SparseMatrix.from_dense(dense)

    def __repr__(self) -> str:
        return f"SparseMatrix({self.num_rows}x{self.num_cols}, nnz={self.values.numel()})"


class Matrices:
    dense = staticmethod(lambda r, c, values: DenseMatrix(r, c, values))
    sparse = staticmethod(lambda r, c, col_ptrs, row_indices, values: SparseMatrix(r, c, col_ptrs, row_indices, values))
    zeros = staticmethod(DenseMatrix.zeros)
    ones = staticmethod(DenseMatrix.ones)
    eye = staticmethod(DenseMatrix.eye)
    diag = staticmethod(DenseMatrix.diag)

    @staticmethod
    def speye(n: int) -> SparseMatrix:
        return SparseMatrix(n, n, torch.arange(n + 1), torch.arange(n), torch.ones(n, dtype=torch.float64))

    @staticmethod
    def horzcat(ms: Sequence[Matrix]) -> DenseMatrix:
        return DenseMatrix.from_tensor(torch.cat([m.to_tensor() for m in ms], dim=1))

    @staticmethod
    def vertcat(ms: Sequence[Matrix]) -> DenseMatrix:
        return DenseMatrix.from_tensor(torch.cat([m.to_tensor() for m in ms], dim=0))


# ------------------------------------------------------------------------------------------------ level 2 / 3
def _dense_of(x: Vector) -> torch.Tensor:
    return x.to_dense().values


def gemv(alpha: float, A: Matrix, x: Vector, beta: float, y: DenseVector) -> None:
    """``y := alpha * A * x + beta * y`` (BLAS.gemv :515; dense/sparse A × dense/sparse x)."""
    if A.num_cols != x.size:
        raise ValueError(f"The columns of A don't match the number of elements of x. A: {A.num_cols}, x: {x.size}")
    if A.num_rows != y.size:
        raise ValueError(f"The rows of A don't match the number of elements of y. A: {A.num_rows}, y: {y.size}")
    dt = y.values.dtype
    if alpha == 0.0 and beta == 1.0:
        return
    if alpha == 0.0:
        y.values.mul_(beta)
        return
    if isinstance(A, SparseMatrix):
        r, c, v = A.coo()
        xv = _dense_of(x).to(dt)
        prod = torch.zeros(A.num_rows, dtype=dt).index_add_(0, r, v.to(dt) * xv[c])
    elif isinstance(x, SparseVector):                       # gather the needed columns only
        prod = A.to_tensor().to(dt)[:, x.indices.long()] @ x.values.to(dt)
    else:
        prod = A.to_tensor().to(dt) @ _dense_of(x).to(dt)
    y.values.mul_(beta).add_(prod.to(y.values.device), alpha=alpha)


def gemm(alpha: float, A: Matrix, B: Matrix, beta: float, C: DenseMatrix) -> None:
    """``C := alpha * A * B + beta * C`` (BLAS.gemm :348; A dense or sparse, B dense; C column-major, not transposed)."""
    if A.num_cols != B.num_rows:
        raise ValueError(f"The columns of A don't match the rows of B. A: {A.num_cols}, B: {B.num_rows}")
    if A.num_rows != C.num_rows or B.num_cols != C.num_cols:
        raise ValueError(f"C is {C.num_rows}x{C.num_cols}, expected {A.num_rows}x{B.num_cols}")
    if C.is_transposed:
        raise ValueError("The matrix C cannot be the product of a transpose() call. C.isTransposed must be false.")
    if alpha == 0.0 and beta == 1.0:
        return
    dt = C.values.dtype
    Ct = C.to_tensor()                                        # a view onto C.values
    if alpha == 0.0:
        Ct.mul_(beta)
        return
    Bt = B.to_tensor().to(dt)
    if isinstance(A, SparseMatrix):
        r, c, v = A.coo()
        prod = torch.zeros(A.num_rows, B.num_cols, dtype=dt).index_add_(0, r, v.to(dt)[:, None] * Bt[c])
    else:
        prod = A.to_tensor().to(dt) @ Bt
    Ct.mul_(beta).add_(prod, alpha=alpha)


def syr(alpha: float, x: Vector, A: DenseMatrix) -> None:
    """``A := alpha * x * x^T + A`` (BLAS.syr :288; the full symmetric matrix is updated, like the reference does after
    the native call fills one triangle)."""
    if A.num_rows != A.num_cols:
        raise ValueError(f"A is not a square matrix (and thus cannot be symmetric). A: {A.num_rows} x {A.num_cols}")
    if A.num_rows != x.size:
        raise ValueError(f"The size of x doesn't match the rank of A. A: {A.num_rows} x {A.num_cols}, x: {x.size}")
    At = A.to_tensor()
    if isinstance(x, SparseVector):
        idx, v = x.indices.long(), x.values.to(At.dtype)
        At[idx[:, None], idx[None, :]] += alpha * torch.outer(v, v)
    else:
        v = _dense_of(x).to(At.dtype)
        At.add_(torch.outer(v, v), alpha=alpha)


def spr(alpha: float, v: Vector, U) -> None:
    """``U += alpha * v * v^T`` with U the upper triangle packed column-major in an array of n(n+1)/2 entries
    (BLAS.spr :244).  ``U`` is a 1-D tensor / DenseVector, updated in place."""
    n = v.size
    Uv = U.values if isinstance(U, DenseVector) else U
    if Uv.numel() != n * (n + 1) // 2:
        raise ValueError(f"packed upper triangle of a {n}x{n} matrix has {n * (n + 1) // 2} entries, got {Uv.numel()}")
    iu = torch.triu_indices(n, n)                            # row-major order of the upper triangle ...
    order = torch.argsort(iu[1] * n + iu[0], stable=True)      # ... re-ordered column-major = packed "U" layout
    rows, cols = iu[0][order], iu[1][order]
    if isinstance(v, SparseVector):
        dv = torch.zeros(n, dtype=Uv.dtype)
        dv[v.indices.long()] = v.values.to(Uv.dtype)
    else:
        dv = _dense_of(v).to(Uv.dtype)
    Uv.add_(dv[rows] * dv[cols], alpha=alpha)


def unpack_upper(U, n: int) -> torch.Tensor:
    """Packed upper triangle -> full symmetric ``[n, n]`` tensor (``RowMatrix.triuToFull``)."""
    Uv = U.values if isinstance(U, DenseVector) else U
    iu = torch.triu_indices(n, n)
    order = torch.argsort(iu[1] * n + iu[0], stable=True)
    rows, cols = iu[0][order], iu[1][order]
    M = torch.zeros(n, n, dtype=Uv.dtype)
    M[rows, cols] = Uv
    M[cols, rows] = Uv
    return M
