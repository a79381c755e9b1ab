"""Counter-based Bernoulli row sampler shared bit-for-bit by the CPU backend and the CUDA kernels.

The reference samples with ``BernoulliSampler`` seeded per partition (core/util/random/
RandomSampler.scala:144-176, rdd/PartitionwiseSampledRDD.scala:56-69) for ASGD and with
``scala.util.Random(cTime).nextDouble() < b`` for ASAGA so that the driver can replay the
stream (SparkASAGAThread.scala:280-294,365-369).  A sequential JVM RNG is the wrong shape for a
GPU, so membership here is a pure function ``select(seed, global_row_index)``:

    key = mix64(seed)
    x   = mix64(key + idx * GOLDEN)
    selected  <=>  (x >> 40) < floor(fraction * 2**24)

(mix64 = splitmix64 finaliser).  Any party (worker kernel, server, test oracle) can evaluate or
replay it for any row in O(1) — which is exactly what ASAGA's "replay the sample to ship partial
history" trick needs — and the CUDA implementation in csrc/include/afb_hash.cuh is identical.
"""
from __future__ import annotations

import numpy as np
import torch

GOLDEN = np.uint64(0x9E3779B97F4A7C15)
M1 = np.uint64(0xBF58476D1CE4E5B9)
M2 = np.uint64(0x94D049BB133111EB)
MASK64 = (1 << 64) - 1


def mix64_int(x: int) -> int:
    x &= MASK64
    x ^= x >> 30
    x = (x * 0xBF58476D1CE4E5B9) & MASK64
    x ^= x >> 27
    x = (x * 0x94D049BB133111EB) & MASK64
    x ^= x >> 31
    return x


def _mix64_np(x: np.ndarray) -> np.ndarray:
    x = x.astype(np.uint64, copy=True)
    x ^= x >> np.uint64(30)
    x *= M1
    x ^= x >> np.uint64(27)
    x *= M2
    x ^= x >> np.uint64(31)
    return x


def threshold24(fraction: float) -> int:
    """Integer threshold used by every implementation (so CPU and GPU agree exactly)."""
    t = int(np.floor(float(fraction) * float(1 << 24)))
    return max(0, min(t, 1 << 24))


def hash_u24(seed: int, idx) -> np.ndarray:
    """24-bit uniform integer for every global row index in ``idx`` (numpy, uint64 math)."""
    idx = np.asarray(idx.detach().cpu().numpy() if isinstance(idx, torch.Tensor) else idx).astype(np.uint64)
    key = np.uint64(mix64_int(int(seed) & MASK64))
    with np.errstate(over="ignore"):
        x = _mix64_np(key + idx * GOLDEN)
    return (x >> np.uint64(40)).astype(np.int64)


def bernoulli_mask(seed: int, idx, fraction: float) -> np.ndarray:
    """Boolean selection mask for global row indices ``idx``."""
    return hash_u24(seed, idx) < threshold24(fraction)


def bernoulli_mask_torch(seed: int, idx: torch.Tensor, fraction: float) -> torch.Tensor:
    return torch.from_numpy(bernoulli_mask(seed, idx, fraction)).to(idx.device)


def uniform01(seed: int, idx) -> np.ndarray:
    """U[0,1) doubles from the same hash (used by the delay injector's long-tail draws)."""
    return hash_u24(seed, idx).astype(np.float64) / float(1 << 24)


def poisson_counts(seed: int, idx, mean: float) -> np.ndarray:
    """How many times each row appears in a sample WITH replacement: Poisson(mean) per row (``PoissonSampler``,
    util/random/RandomSampler.scala:180-230), drawn by inverse CDF from the row's own 24-bit uniform — so, like the Bernoulli
    mask, it is a pure function of (seed, global row index) that any party can replay."""
    u = uniform01(seed, idx)
    mean = float(mean)
    if mean <= 0:
        return np.zeros(u.shape, dtype=np.int64)
    counts = np.zeros(u.shape, dtype=np.int64)
    p = np.exp(-mean)
    cdf = np.full(u.shape, p)
    k = 0
    todo = u >= cdf
    while todo.any() and k < 64 + int(10 * mean):
        k += 1
        p = p * mean / k
        cdf = cdf + p
        counts[todo] = k
        todo = u >= cdf
    return counts
