"""Synthetic data generators.

* MLlib-compatible ``LinearDataGenerator`` / ``LogisticRegressionDataGenerator``
  (mllib/util/LinearDataGenerator.scala:74-190, LogisticRegressionDataGenerator.scala:48-66) and
  the test generator ``generate_gd_input`` (mllib test GradientDescentSuite.scala:41-62).
* Shape generators for the benchmark configs (no network => no real datasets): ``mnist8m``
  (dense, integer pixels 0..255, ~19 % non-zero), ``epsilon`` (dense, unit-norm rows, ±1 labels),
  ``rcv1`` (CSR, ~74 nnz/row, d = 47 236).  They generate directly on the target device, one
  shard at a time, so an 8.1 M × 784 matrix never has to exist on the host.
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple

import numpy as np
import torch

from ..linalg.vectors import DenseVector, LabeledPoint
from .batch import Shard, padded_ld
from .dataset import ShardedDataset

SHAPES = {
    # name: (rows, features, layout, nnz/row)
    "mnist8m": (8_100_000, 784, "dense", 150),
    "epsilon": (400_000, 2000, "dense", 2000),
    "rcv1": (677_399, 47_236, "csr", 74),
    "mnist": (60_000, 780, "dense", 150),
}


class LinearDataGenerator:
    @staticmethod
    def generate_linear_input(intercept: float, weights, n_points: int, seed: int, eps: float = 0.1,
                              x_mean=None, x_variance=None, sparsity: float = 0.0) -> List[LabeledPoint]:
        rng = np.random.RandomState(seed)
        w = np.asarray(weights, dtype=np.float64)
        d = w.shape[0]
        x_mean = np.zeros(d) if x_mean is None else np.asarray(x_mean, dtype=np.float64)
        x_var = np.full(d, 1.0 / 3.0) if x_variance is None else np.asarray(x_variance, dtype=np.float64)
        X = (rng.rand(n_points, d) - 0.5) * np.sqrt(12.0 * x_var) + x_mean
        if sparsity > 0:
            X = X * (rng.rand(n_points, d) >= sparsity)
        y = X @ w + intercept + eps * rng.randn(n_points)
        return [LabeledPoint(float(y[i]), DenseVector(X[i])) for i in range(n_points)]

    @staticmethod
    def generate_linear_arrays(n: int, d: int, eps: float = 0.1, seed: int = 42, intercept: float = 0.0):
        rng = np.random.RandomState(seed)
        w = rng.rand(d) - 0.5                     # true weights U[-0.5, 0.5)
        X = (rng.rand(n, d) - 0.5) * math.sqrt(12.0 * (1.0 / 3.0))
        y = X @ w + intercept + eps * rng.randn(n)
        return X, y, w

    @staticmethod
    def generate_linear_dataset(scheduler, n: int, d: int, eps: float, num_partitions: int = 2,
                                intercept: float = 0.0, seed: int = 42, dtype=torch.float64, device=None):
        """``generateLinearRDD(sc, n, d, eps, nparts, intercept)``."""
        X, y, w = LinearDataGenerator.generate_linear_arrays(n, d, eps, seed, intercept)
        ds = ShardedDataset.from_arrays(scheduler, X, y, num_partitions, dtype=dtype, device=device, name="linear")
        ds.true_weights = w  # type: ignore[attr-defined]
        return ds


class LogisticRegressionDataGenerator:
    @staticmethod
    def generate_logistic_arrays(n: int, d: int, eps: float = 1.0, prob_one: float = 0.5, seed: int = 42):
        """Label alternates idx % 2 (as in the reference), features N(0,1) + y*eps."""
        rng = np.random.RandomState(seed)
        y = (np.arange(n) % 2).astype(np.float64) if prob_one == 0.5 else (rng.rand(n) < prob_one).astype(np.float64)
        X = rng.randn(n, d) + (y * eps)[:, None]
        return X, y

    @staticmethod
    def generate_logistic_dataset(scheduler, n: int, d: int, eps: float, num_partitions: int = 2,
                                  prob_one: float = 0.5, seed: int = 42, dtype=torch.float64, device=None):
        X, y = LogisticRegressionDataGenerator.generate_logistic_arrays(n, d, eps, prob_one, seed)
        return ShardedDataset.from_arrays(scheduler, X, y, num_partitions, dtype=dtype, device=device, name="logistic")


class SVMDataGenerator:
    """``mllib/util/SVMDataGenerator.scala``: features U[-1, 1)^d, label 1 when ``x . w_true + 0.1 * N(0, 1) >= 0`` else 0,
    ``w_true ~ N(0, 1)^d`` drawn from a fixed global seed."""

    @staticmethod
    def generate_svm_arrays(n: int, d: int = 2, seed: int = 42, weights_seed: int = 94720):
        w = np.random.RandomState(weights_seed).randn(d)
        rng = np.random.RandomState(seed)
        X = rng.rand(n, d) * 2.0 - 1.0
        y = ((X @ w + 0.1 * rng.randn(n)) >= 0).astype(np.float64)
        return X, y, w

    @staticmethod
    def generate_svm_dataset(scheduler, n: int = 1000, d: int = 2, num_partitions: int = 2, seed: int = 42,
                             dtype=torch.float64, device=None):
        X, y, w = SVMDataGenerator.generate_svm_arrays(n, d, seed)
        ds = ShardedDataset.from_arrays(scheduler, X, y, num_partitions, dtype=dtype, device=device, name="svm")
        ds.true_weights = w  # type: ignore[attr-defined]
        return ds

    @staticmethod
    def main(argv=None) -> int:
        """``SVMGenerator <master> <output_dir> [num_examples] [num_features] [num_partitions]``: writes labeled points as text."""
        import sys
        from ..core.session import AsyncSession
        a = list(sys.argv[1:] if argv is None else argv)
        if len(a) < 2:
            print("Usage: SVMGenerator <master> <output_dir> [num_examples] [num_features] [num_partitions]")
            return 1
        n, d, parts = (int(a[2]) if len(a) > 2 else 1000), (int(a[3]) if len(a) > 3 else 2), (int(a[4]) if len(a) > 4 else 2)
        with AsyncSession(a[0], app_name="SVMGenerator") as sc:
            X, y, _ = SVMDataGenerator.generate_svm_arrays(n, d)
            pts = [LabeledPoint(float(y[i]), DenseVector(X[i])) for i in range(n)]
            sc.parallelize(pts, parts).save_as_text_file(a[1])
        return 0


def generate_gd_input(offset: float, scale: float, n_points: int, seed: int) -> List[LabeledPoint]:
    """GradientDescentSuite.generateGDInput: x1 ~ N(0,1); y = 1{sigmoid(offset + scale*x1) > u}."""
    rng = np.random.RandomState(seed)
    x1 = rng.randn(n_points)
    u = rng.rand(n_points)
    p = 1.0 / (1.0 + np.exp(-(offset + scale * x1)))
    y = (u < p).astype(np.float64)
    return [LabeledPoint(float(y[i]), DenseVector([x1[i]])) for i in range(n_points)]


# ------------------------------------------------------------------------------------------
# benchmark-shape synthetic shards, generated on the target device
# ------------------------------------------------------------------------------------------

def _true_weights(d: int, seed: int, device, dtype) -> torch.Tensor:
    g = torch.Generator(device="cpu").manual_seed(seed)
    return (torch.rand(d, generator=g, dtype=torch.float64) - 0.5).to(device=device, dtype=dtype)


def synthetic_shard(shape: str, rows: int, partition: int, num_partitions: int, seed: int = 42,
                    device=None, dtype: torch.dtype = torch.float32, num_features: Optional[int] = None,
                    labels: str = "binary01", row_offset: int = 0, chunk_rows: int = 262_144,
                    row_dtype: Optional[torch.dtype] = None) -> Shard:
    """One shard of a synthetic dataset of the named shape.

    labels: ``binary01`` (logistic, {0,1}), ``pm1`` (±1), ``regression`` (x·w* + noise).
    Generated in chunks so peak temporary memory stays small next to a multi-GB shard.
    ``row_dtype=torch.bfloat16`` (dense shapes) stores the feature rows in bf16 (labels stay ``dtype``); the mnist shapes hold
    integer pixel values 0..255, which bf16 represents exactly.
    """
    if shape not in SHAPES:
        raise KeyError(f"unknown shape {shape!r}; known: {sorted(SHAPES)}")
    _, d0, layout, nnz_row = SHAPES[shape]
    d = int(num_features or d0)
    dev = torch.device(device) if device is not None else torch.device("cpu")
    g = torch.Generator(device=dev).manual_seed(seed * 1_000_003 + partition)
    wstar = _true_weights(d, seed, dev, torch.float32)
    y = torch.empty(rows, dtype=dtype, device=dev)

    def make_labels(margin: torch.Tensor) -> torch.Tensor:
        if labels == "regression":
            return margin + 0.1 * torch.randn(margin.shape, generator=g, device=dev)
        p = torch.sigmoid(margin)
        b = (torch.rand(margin.shape, generator=g, device=dev) < p).to(torch.float32)
        return b if labels == "binary01" else 2.0 * b - 1.0

    if layout == "dense":
        xdt = row_dtype or dtype
        ld = padded_ld(d) if xdt != torch.bfloat16 else (d + 7) // 8 * 8      # rows are whole 16-byte chunks
        X = torch.zeros(rows, ld, dtype=xdt, device=dev)
        for a in range(0, rows, chunk_rows):
            b = min(rows, a + chunk_rows)
            n = b - a
            if shape in ("mnist8m", "mnist"):
                # integer "pixels": ~19 % non-zero, values 1..255 (exact in fp32/bf16)
                nz = torch.rand(n, d, generator=g, device=dev) < 0.19
                pix = torch.randint(1, 256, (n, d), generator=g, device=dev, dtype=torch.int32).to(torch.float32)
                blk = pix * nz
                scale = 12.0 / (255.0 * math.sqrt(0.19 * d))    # margin std ~2: a learnable but noisy problem
            else:  # epsilon: unit-norm gaussian rows
                blk = torch.randn(n, d, generator=g, device=dev)
                blk = blk / blk.norm(dim=1, keepdim=True).clamp_min(1e-12)
                scale = 4.0
            X[a:b, :d] = blk.to(xdt)
            y[a:b] = make_labels((blk @ wstar) * scale).to(dtype)
            del blk
        sh = Shard(d, y, X, partition=partition)
    else:
        k = nnz_row
        indptr = torch.arange(0, (rows + 1) * k, k, dtype=torch.int64, device=dev)
        indices = torch.empty(rows * k, dtype=torch.int32, device=dev)
        values = torch.empty(rows * k, dtype=dtype, device=dev)
        for a in range(0, rows, chunk_rows):
            b = min(rows, a + chunk_rows)
            n = b - a
            # k distinct ascending features per row, Zipf-ish popularity: strata + jitter keeps them distinct
            strata = torch.arange(k, device=dev, dtype=torch.float64)[None, :]
            u = torch.rand(n, k, generator=g, device=dev, dtype=torch.float64)
            pos = ((strata + u) / k) ** 2.0 * d          # quadratic skew towards low feature ids
            idx = pos.floor().clamp_(0, d - 1).to(torch.int64)
            idx, _ = torch.sort(idx, dim=1)
            # enforce strictly increasing (bump duplicates)
            bump = torch.arange(k, device=dev)[None, :]
            idx = torch.maximum(idx, torch.cummax(idx - bump, dim=1).values + bump).clamp_(max=d - 1)
            dup = torch.zeros_like(idx, dtype=torch.bool)
            dup[:, 1:] = idx[:, 1:] <= idx[:, :-1]
            val = torch.rand(n, k, generator=g, device=dev) + 0.1
            val = val * (~dup)
            val = val / val.norm(dim=1, keepdim=True).clamp_min(1e-12)
            # rows that hit the clamp keep duplicates with zero value; make indices strictly ascending anyway
            if bool(dup.any()):
                fix = torch.arange(d - k, d, device=dev)[None, :].expand(n, k)
                bad = dup.any(dim=1)
                idx[bad] = fix[bad]
            margin = (val * wstar[idx]).sum(1) * 8.0
            indices[a * k:b * k] = idx.reshape(-1).to(torch.int32)
            values[a * k:b * k] = val.reshape(-1).to(dtype)
            y[a:b] = make_labels(margin).to(dtype)
        sh = Shard(d, y, None, indptr, indices, values, partition=partition)
    sh.gidx = torch.arange(rows, dtype=torch.int64, device=dev) + row_offset
    sh.meta["true_weights_seed"] = seed
    return sh


def shard_rows(total_rows: int, num_partitions: int) -> List[int]:
    """Row counts of a round-robin repartition of ``total_rows`` rows."""
    base, rem = divmod(total_rows, num_partitions)
    return [base + (1 if p < rem else 0) for p in range(num_partitions)]


def synthetic_dataset(scheduler, shape: str, num_partitions: int, rows: Optional[int] = None, seed: int = 42,
                      device=None, dtype=torch.float32, labels: str = "binary01",
                      num_features: Optional[int] = None) -> ShardedDataset:
    total = int(rows if rows is not None else SHAPES[shape][0])
    sizes = shard_rows(total, num_partitions)
    shards, off = [], 0
    for p, n in enumerate(sizes):
        shards.append(synthetic_shard(shape, n, p, num_partitions, seed, device, dtype, num_features, labels, off))
        off += n
    ds = ShardedDataset.from_shards(scheduler, shards, f"synthetic:{shape}")
    ds.num_features = shards[0].num_features  # type: ignore[attr-defined]
    return ds
