"""Feature preprocessing used around the GLM baseline: column statistics, ``StandardScaler``, bias append.

Reference: ``Statistics.colStats`` -> ``MultivariateOnlineSummarizer`` (mllib/stat/MultivariateOnlineSummarizer.scala: count,
mean, variance (unbiased), numNonzeros, max, min, normL1, normL2 in one ``treeAggregate`` pass), ``StandardScaler`` /
``StandardScalerModel`` (mllib/feature/StandardScaler.scala: ``withMean`` false and ``withStd`` true by default — the
configuration ``GeneralizedLinearAlgorithm`` uses internally, GeneralizedLinearAlgorithm.scala:283-304), ``MLUtils.appendBias``
(mllib/util/MLUtils.scala:255).  All of it runs as ONE job over a batch-capable ``ShardedDataset`` with vectorised per-shard
partials (dense or CSR, CPU or HBM-resident).
"""
from __future__ import annotations

from dataclasses import dataclass

import torch

from ..data.batch import Shard
from ..data.dataset import ShardedDataset


@dataclass
class ColumnSummary:
    count: int
    mean: torch.Tensor
    variance: torch.Tensor          # unbiased (n - 1)
    num_nonzeros: torch.Tensor
    max: torch.Tensor
    min: torch.Tensor
    norm_l1: torch.Tensor
    norm_l2: torch.Tensor


def col_stats(data: ShardedDataset) -> ColumnSummary:
    """``Statistics.colStats(rdd)``."""
    first = next(s for s in data.shards() if s is not None)
    d = first.num_features

    def seq(acc, s: Shard):
        if s is None or s.num_rows == 0:
            return acc
        X = s.to_dense().X[:, :d].double().cpu()
        part = [torch.tensor(float(s.num_rows), dtype=torch.float64), X.sum(0), (X * X).sum(0), (X != 0).sum(0).double(),
                X.max(0).values, X.min(0).values, X.abs().sum(0)]
        if acc is None:
            return part
        return [acc[0] + part[0], acc[1] + part[1], acc[2] + part[2], acc[3] + part[3], torch.maximum(acc[4], part[4]),
                torch.minimum(acc[5], part[5]), acc[6] + part[6]]

    def comb(a, b):
        if a is None:
            return b
        if b is None:
            return a
        return [a[0] + b[0], a[1] + b[1], a[2] + b[2], a[3] + b[3], torch.maximum(a[4], b[4]), torch.minimum(a[5], b[5]), a[6] + b[6]]
    acc = data.map_batch(lambda s: s).tree_aggregate(lambda: None, seq, comb)
    if acc is None:
        raise ValueError("empty dataset")
    n = float(acc[0])
    mean = acc[1] / n
    var = ((acc[2] - n * mean * mean) / max(n - 1.0, 1.0)).clamp_min(0.0)
    return ColumnSummary(int(n), mean, var, acc[3], acc[4], acc[5], acc[6], torch.sqrt(acc[2]))


def append_bias(data: ShardedDataset) -> ShardedDataset:
    """``MLUtils.appendBias`` on every row: a trailing constant-1 feature (how GLMs fit an intercept)."""
    from .glm import _append_bias
    return data.map_shards(_append_bias)


class StandardScalerModel:
    def __init__(self, std: torch.Tensor, mean: torch.Tensor, with_std: bool = True, with_mean: bool = False):
        self.std, self.mean, self.with_std, self.with_mean = std, mean, with_std, with_mean

    def set_with_mean(self, b: bool): self.with_mean = bool(b); return self
    def set_with_std(self, b: bool): self.with_std = bool(b); return self

    def _scale(self, dtype, device) -> torch.Tensor:
        return torch.where(self.std > 0, 1.0 / self.std, torch.zeros_like(self.std)).to(dtype=dtype, device=device)

    def transform(self, data):
        """A ``ShardedDataset`` (lazily, shard by shard), a 2-D tensor of rows or a single 1-D feature tensor.  With
        ``with_mean`` sparse shards are densified (centering destroys sparsity, like the reference warns)."""
        d = self.std.numel()
        if isinstance(data, torch.Tensor):
            x = data.to(torch.float64)
            if self.with_mean:
                x = x - self.mean
            return x * self._scale(torch.float64, x.device) if self.with_std else x
        std_scale, mean, with_std, with_mean = self._scale, self.mean, self.with_std, self.with_mean

        def fn(s: Shard):
            if with_mean:
                s = s.to_dense()
            if s.is_dense:
                X = s.X.clone()
                if with_mean:
                    X[:, :d] -= mean.to(X.dtype).to(X.device)
                if with_std:
                    X[:, :d] *= std_scale(X.dtype, X.device)
                return Shard(d, s.y, X, gidx=s.gidx, partition=s.partition, meta=dict(s.meta))
            vals = s.values * std_scale(s.values.dtype, s.device)[s.indices.long()] if with_std else s.values
            return Shard(d, s.y, None, s.indptr, s.indices, vals, s.gidx, s.partition, dict(s.meta))
        return data.map_shards(fn)


class StandardScaler:
    """``new StandardScaler(withMean = false, withStd = true).fit(rdd)``."""

    def __init__(self, with_mean: bool = False, with_std: bool = True):
        self.with_mean, self.with_std = with_mean, with_std

    def fit(self, data: ShardedDataset) -> StandardScalerModel:
        st = col_stats(data)
        return StandardScalerModel(torch.sqrt(st.variance), st.mean, self.with_std, self.with_mean)
