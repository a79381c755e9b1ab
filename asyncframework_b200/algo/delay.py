"""Delay / straggler injector.

Reference: inside a ``mapPartitionsWithIndex`` placed before sampling
(SparkASGDThread.scala:126-138, 244-249, 284-309; same blocks in the other drivers).  After the
warm-up (``k > 100*P`` async, ``k > 100`` sync) ``avgDelay = culTime/culCount`` is the mean
observed task time.  *Controlled* mode (``coeff > 0``): partition 0 sleeps ``coeff * avgDelay``
every task.  *Production-cluster* mode (``coeff == -1``): 25 % of the partitions (indices 4c) are
stragglers, the first 20 % of those sleep ``U[2.5,10) * avgDelay`` (long tail), the rest
``U[1.5,2.5) * avgDelay``.  The fused backend runs the same policy on the device
(csrc/include/afb_round_impl.cuh, ``delay_mode``).
"""
from __future__ import annotations

import random
import threading
import time
from typing import Callable, Iterator, List, Optional


def straggler_sets(num_partitions: int):
    length = int(round(0.25 * num_partitions))
    n_normal = int(round(0.8 * length))
    n_long = length - n_normal
    long_tail = [c * 4 for c in range(length) if c < n_long]
    normal = [c * 4 for c in range(length) if c >= n_long]
    return normal, long_tail


class DelayInjector:
    def __init__(self, num_partitions: int, coeff: float, warmup_updates: int, sleep: Callable[[float], None] = time.sleep,
                 rng: Optional[random.Random] = None):
        self.P = num_partitions
        self.coeff = float(coeff)
        self.production = (coeff == -1)
        self.warmup_updates = int(warmup_updates)
        self.normal, self.long_tail = straggler_sets(num_partitions)
        self.cul_time_ms = 0.0
        self.cul_count = 0
        self.avg_delay_ms = 0.0
        self.calibrated = False
        self._lock = threading.Lock()
        self._sleep = sleep
        self._rng = rng or random.Random()
        self.injected_ms = 0.0

    def __getstate__(self):          # shipped inside task closures to executor processes
        st = dict(self.__dict__)
        st["_lock"] = None
        return st

    def __setstate__(self, st):
        self.__dict__.update(st)
        self._lock = threading.Lock()

    @property
    def mode(self) -> str:
        return "production" if self.production else ("controlled" if self.coeff > 0 else "none")

    # -- calibration (updater side) ---------------------------------------
    def observe(self, k: int, task_ms: float) -> None:
        """Accumulate task time during the first ``warmup_updates`` accepted updates."""
        if k < self.warmup_updates:
            with self._lock:
                self.cul_time_ms += task_ms
                self.cul_count += 1

    def maybe_calibrate(self, k: int) -> bool:
        if not self.calibrated and k > self.warmup_updates and self.cul_count > 0:
            self.avg_delay_ms = self.cul_time_ms / self.cul_count
            self.calibrated = True
        return self.calibrated

    # -- injection (worker side) --------------------------------------------
    def delay_ms_for(self, index: int) -> float:
        if not self.calibrated:
            return 0.0
        if not self.production:
            return self.coeff * self.avg_delay_ms if (index == 0 and self.coeff > 0) else 0.0
        if index in self.long_tail:
            return (self._rng.random() * 7.5 + 2.5) * self.avg_delay_ms
        if index in self.normal:
            return (self._rng.random() + 1.5) * self.avg_delay_ms
        return 0.0

    def hook(self) -> Callable[[int, Iterator], Iterator]:
        """``f(index, iterator)`` for ``map_partitions_with_index(..., preserves_batch=True)``."""
        def f(index: int, it: Iterator) -> Iterator:
            ms = self.delay_ms_for(index)
            if ms > 0:
                self.injected_ms += ms
                self._sleep(ms / 1e3)
            return it
        return f
