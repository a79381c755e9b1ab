#!/usr/bin/env python
"""Asynchronous SGD written against the Python API (the loop of docs/API.md §1, runnable): bucket barrier, staleness
filter, broadcast per submission round, ASYNCreduce over the sampled partitions, server loop on the driver.

    python examples/async_sgd_python_api.py [--master local[4]] [--workers 4] [--updates 200]

Reference shape: ASYNCsamples/src/main/scala/org/apache/async/SparkASGDThread.scala:150-300.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from a source checkout

import torch  # noqa: E402

from asyncframework_b200 import AsyncSession, barriers  # noqa: E402
from asyncframework_b200.data import LinearDataGenerator  # noqa: E402
from asyncframework_b200.optim import make_gradient  # noqa: E402


def main(master="local[4]", P=4, updates=200, n=4000, d=20, b=0.2, gamma=0.5, taw=8, beta=0.5, seed=42):
    sc = AsyncSession(master)
    points = LinearDataGenerator.generate_linear_dataset(sc.scheduler, n, d, 0.05, P, seed=1).zip_with_index().cache()
    N = points.count()
    AC = sc.async_context(P)
    AC.blocking = False                                    # sc.set_mode(1): actions return immediately
    w = torch.zeros(d, dtype=torch.float64)
    grad = make_gradient("least_squares")
    k = submitted = dropped = 0
    while k < updates:
        if submitted == 0 or AC.stat.available_workers() >= max(1, int(P * beta)):
            a = sc.broadcast(w.clone())
            sel = points.async_barrier(barriers.bucket(P, beta), AC.stat)
            sel.sample(False, b, seed + submitted + 1) \
               .map_batch(lambda s, a=a: grad.batch(s, a.value)[0]) \
               .async_reduce(lambda x, y: x.add_(y), AC)
            submitted += 1
        info = AC.collect_all(timeout=0.2)
        while info is not None:
            if info.staleness <= taw:
                w.add_(info.data, alpha=-gamma / (k // P + 1) ** 0.5 / (b * N / P))
                k += 1
            else:
                dropped += 1
            info = AC.collect_all(timeout=0.0) if AC.has_next() else None
    AC.blocking = True                                     # back to blocking jobs for the evaluation pass
    obj = points.map_batch(lambda s: float(((s.matvec(w) - s.y) ** 2).sum())).reduce(lambda p, q: p + q) / N
    print(f"updates {k}  submission rounds {submitted}  dropped (stale) {dropped}  objective {obj:.5f}")
    sc.stop()
    return obj


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--master", default="local[4]")
    ap.add_argument("--workers", type=int, default=4)
    ap.add_argument("--updates", type=int, default=200)
    a = ap.parse_args()
    main(a.master, a.workers, a.updates)
