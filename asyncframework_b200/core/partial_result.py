"""Result envelope delivered by ``AsyncContext.collect_all``.

Parity: ``RDDPartialRes`` (core/rdd/RDDPartialRes.scala:13-36).  The reference stores
*staleness* in the ``ts`` slot for ASYNCreduce but the *submit timestamp* for ASYNCaggregate
(RDD.scala:1151-1152 vs :1333, quirk B.6); we carry both, plus the true batch size
(the reference always reports Int.MinValue, quirk B.7).
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Generic, TypeVar

T = TypeVar("T")


@dataclass
class PartialResult(Generic[T]):
    data: T
    staleness: int
    timestamp: int
    batch_size: int
    worker_id: int
    submit_ms: float = 0.0
    finish_ms: float = 0.0

    # reference-style accessors
    def get_task_result(self) -> T:
        return self.data

    def get_staleness(self) -> int:
        return self.staleness

    def get_timestamp(self) -> int:
        return self.timestamp

    def get_batch_size(self) -> int:
        return self.batch_size

    def get_worker_id(self) -> int:
        return self.worker_id

    @property
    def task_ms(self) -> float:
        return self.finish_ms - self.submit_ms
