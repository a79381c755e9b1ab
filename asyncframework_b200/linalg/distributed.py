"""Distributed row matrix over a batch-capable ``ShardedDataset`` (mllib/linalg/distributed/RowMatrix.scala — imported, though
unused, by the reference's ``SparkMean`` toy; the natural home of the Gramian / covariance / PCA / SVD routines of MLlib).

A ``RowMatrix`` is the feature block of the dataset: ``n`` rows sharded over the partitions, ``d`` columns.  Every routine is ONE
``tree_aggregate`` of per-shard partials that are genuine GEMMs on the shard (``X_pᵀ X_p``: d x n_p x d — cuBLAS / tensor cores when
the shard lives in HBM, since torch.matmul is a plain library GEMM there), followed by small dense algebra on the driver
(``d x d`` eigendecomposition), exactly how the reference does it (``computeGramianMatrix`` -> ``treeAggregate`` of ``spr`` updates,
``computePrincipalComponents``, ``computeSVD`` in its "local eigs of the Gramian" mode for small d).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from ..data.batch import Shard
from ..data.dataset import ShardedDataset
from .matrices import DenseMatrix, Matrix
from .vectors import DenseVector


class RowMatrix:
    def __init__(self, rows: ShardedDataset, num_rows: Optional[int] = None, num_cols: Optional[int] = None):
        if not rows.batch_capable:
            raise TypeError("RowMatrix needs a batch-capable ShardedDataset (load_libsvm_file / from_arrays / from_shards)")
        self.rows = rows
        self._n, self._d = num_rows, num_cols

    def num_cols(self) -> int:
        if self._d is None:
            self._d = next(s for s in self.rows.shards() if s is not None).num_features
        return self._d

    def num_rows(self) -> int:
        if self._n is None:
            self._n = self.rows.count()
        return self._n

    def _agg(self, fn):
        def seq(acc, s: Shard):
            if s is None or s.num_rows == 0:
                return acc
            part = fn(s)
            return part if acc is None else acc + part

        def comb(a, b):
            return b if a is None else (a if b is None else a + b)
        out = self.rows.map_batch(lambda s: s).tree_aggregate(lambda: None, seq, comb)
        if out is None:
            raise ValueError("empty matrix")
        return out

    def compute_gramian_matrix(self) -> DenseMatrix:
        """``AᵀA`` (d x d), accumulated in fp64 on the host from per-shard GEMMs in the shard's dtype / device."""
        d = self.num_cols()

        def gram(s: Shard):
            X = s.to_dense().X[:, :d]
            return (X.t() @ X).double().cpu()
        return DenseMatrix.from_tensor(self._agg(gram))

    def compute_column_summary_statistics(self):
        from ..models.feature import col_stats
        return col_stats(self.rows)

    def compute_covariance(self) -> DenseMatrix:
        """Sample covariance (n - 1) from the Gramian and the column means (RowMatrix.computeCovariance)."""
        n, d = self.num_rows(), self.num_cols()
        if n <= 1:
            raise ValueError(f"RowMatrix.computeCovariance called on matrix with only {n} rows.  Cannot compute the covariance of a RowMatrix with <= 1 row.")
        G = self.compute_gramian_matrix().to_tensor()
        mean = self._agg(lambda s: s.to_dense().X[:, :d].double().sum(0).cpu()) / n
        return DenseMatrix.from_tensor((G - n * torch.outer(mean, mean)) / (n - 1))

    def compute_principal_components_and_explained_variance(self, k: int) -> Tuple[DenseMatrix, DenseVector]:
        d = self.num_cols()
        if not (0 < k <= d):
            raise ValueError(f"k = {k} out of range (0, n = {d}]")
        evals, evecs = torch.linalg.eigh(self.compute_covariance().to_tensor())
        order = torch.argsort(evals, descending=True)[:k]
        pcs = evecs[:, order]
        pcs = pcs * torch.where(pcs[pcs.abs().argmax(0), torch.arange(k)] < 0, -1.0, 1.0)     # deterministic sign
        return DenseMatrix.from_tensor(pcs), DenseVector(evals[order] / evals.clamp_min(0).sum())

    def compute_principal_components(self, k: int) -> DenseMatrix:
        return self.compute_principal_components_and_explained_variance(k)[0]

    def compute_svd(self, k: int, compute_u: bool = False, r_cond: float = 1e-9):
        """Top-k singular values / right singular vectors from the eigendecomposition of the Gramian (the reference's
        ``LocalARPACK``/``LocalLAPACK`` modes for small d); ``U = A V S⁻¹`` as a new RowMatrix when asked for."""
        d = self.num_cols()
        if not (0 < k <= d):
            raise ValueError(f"Requested k singular values but got k={k} and numCols={d}.")
        evals, evecs = torch.linalg.eigh(self.compute_gramian_matrix().to_tensor())
        order = torch.argsort(evals, descending=True)
        sig = torch.sqrt(evals[order].clamp_min(0.0))
        keep = int(min(k, int((sig > r_cond * sig[0]).sum())))
        s, V = sig[:keep], evecs[:, order[:keep]]
        U = None
        if compute_u:
            U = self.multiply(DenseMatrix.from_tensor(V / s))
        return U, DenseVector(s), DenseMatrix.from_tensor(V)

    def multiply(self, B: Matrix) -> "RowMatrix":
        """``A · B`` with a local ``d x m`` matrix: a new RowMatrix with the same row partitioning (labels carried along)."""
        d = self.num_cols()
        if B.num_rows != d:
            raise ValueError(f"The number of rows of B ({B.num_rows}) must equal the number of columns of A ({d})")
        Bt = B.to_tensor()
        m = B.num_cols
        from ..data.batch import padded_ld

        def fn(s: Shard):
            X = s.to_dense().X[:, :d]
            Y = torch.zeros(s.num_rows, padded_ld(m), dtype=X.dtype, device=X.device)
            Y[:, :m] = X @ Bt.to(X.dtype).to(X.device)
            return Shard(m, s.y, Y, gidx=s.gidx, partition=s.partition)
        return RowMatrix(self.rows.map_shards(fn), self._n, m)

    def to_tensor(self) -> torch.Tensor:
        """Collect to one local tensor (tests / small matrices), rows in partition order."""
        d = self.num_cols()
        return torch.cat([s.to_dense().X[:, :d].cpu() for s in self.rows.shards() if s is not None])
