"""Built-in barrier predicates for ``async_barrier``.

A predicate is ``WorkerState -> bool`` exactly as in the reference
(``f: workerState[U] => Boolean``, core/rdd/RDD.scala:1050); ``async_barrier`` additionally
requires the worker to be available.  The sample drivers use the *bucket* rule
``state.getAvailableWorkers() >= floor(P * bucketRatio)`` (SparkASGDThread.scala:282-283).
"""
from __future__ import annotations

import math
from typing import Callable

from .worker_state import WorkerState

Predicate = Callable[[WorkerState], bool]


def asp() -> Predicate:
    """Fully asynchronous: every available worker participates."""
    return lambda state: True


def bucket(num_workers: int, beta: float) -> Predicate:
    """Partial barrier: proceed only when >= floor(P*beta) workers are free."""
    need = math.floor(num_workers * beta)
    return lambda state: state.available_workers() >= need


def bsp(num_workers: int) -> Predicate:
    """Bulk-synchronous: all workers must be free."""
    return bucket(num_workers, 1.0)


def ssp(bound: int) -> Predicate:
    """Stale-synchronous: a worker may proceed while the slowest worker's staleness <= bound."""
    return lambda state: state.max_staleness() <= bound


def threshold(count: int) -> int:
    return count


def bucket_threshold(num_workers: int, beta: float) -> int:
    return int(math.floor(num_workers * beta))
