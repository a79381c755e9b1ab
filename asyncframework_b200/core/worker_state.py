"""Per-worker status record + the STAT table.

Parity: ``workerState`` (core/rdd/workerState.scala:14-87): staleness, averageTaskTime,
availability, numTasks and the two *global* queries ``getAvailableWorkers`` (:63-72) and
``getMaxStaleness`` (:74-82) exposed on every record so that a barrier predicate
``WorkerState -> bool`` can express BSP / SSP / bucket rules.  Storage is the C++ ``WorkerTable``
(atomics) so submitter, updater and completion threads can touch it without a Python lock —
the reference uses an unsynchronised mutable.HashMap.
"""
from __future__ import annotations

from typing import Dict, Iterator, Optional

from .. import _native


class WorkerState:
    """A live view of one row of the status table."""
    __slots__ = ("_table", "worker_id")

    def __init__(self, table: "StatTable", worker_id: int):
        self._table = table
        self.worker_id = worker_id

    # -- per-worker fields ------------------------------------------------
    @property
    def staleness(self) -> int:
        return int(self._table._rt.afb_table_staleness(self._table._h, self.worker_id))

    @property
    def avg_task_time_ms(self) -> float:
        return float(self._table._rt.afb_table_avg_ms(self._table._h, self.worker_id))

    @property
    def available(self) -> bool:
        return bool(self._table._rt.afb_table_available(self._table._h, self.worker_id))

    @available.setter
    def available(self, v: bool) -> None:
        self._table._rt.afb_table_set_available(self._table._h, self.worker_id, int(bool(v)))

    @property
    def alive(self) -> bool:
        return bool(self._table._rt.afb_table_alive(self._table._h, self.worker_id))

    @property
    def num_tasks(self) -> int:
        return int(self._table._rt.afb_table_num_tasks(self._table._h, self.worker_id))

    @property
    def heartbeat_age_ms(self) -> float:
        return float(self._table._rt.afb_table_heartbeat_age_ms(self._table._h, self.worker_id))

    # -- global queries (reference exposes them on every record) ------------
    def available_workers(self) -> int:
        return self._table.available_workers()

    def max_staleness(self) -> int:
        return self._table.max_staleness()

    # reference-style accessors
    get_staleness = lambda self: self.staleness                 # noqa: E731
    get_availability = lambda self: self.available              # noqa: E731
    get_num_tasks = lambda self: self.num_tasks                 # noqa: E731
    get_average_task_time = lambda self: self.avg_task_time_ms  # noqa: E731
    get_available_workers = available_workers
    get_max_staleness = max_staleness

    def __repr__(self) -> str:
        return (f"WorkerState(worker={self.worker_id}, staleness={self.staleness}, "
                f"avg_task_time_ms={self.avg_task_time_ms:.3f}, available={self.available}, "
                f"num_tasks={self.num_tasks})")


class StatTable:
    """``AC.STAT``: dict-like ``worker -> WorkerState``; entries appear once a worker has
    been submitted to or has reported (the reference's HashMap has no entry before that,
    and ASYNCbarrier treats a missing entry as "eligible", RDD.scala:1053-1064)."""

    def __init__(self, num_workers: int):
        self._rt = _native.runtime()
        self.num_workers = int(num_workers)
        self._h = self._rt.afb_table_create(self.num_workers)

    def __del__(self):
        try:
            self._rt.afb_table_destroy(self._h)
        except Exception:
            pass

    def get(self, worker: int) -> Optional[WorkerState]:
        if 0 <= worker < self.num_workers and self._rt.afb_table_has_entry(self._h, worker):
            return WorkerState(self, worker)
        return None

    def __getitem__(self, worker: int) -> WorkerState:
        ws = self.get(worker)
        if ws is None:
            raise KeyError(worker)
        return ws

    def __contains__(self, worker: int) -> bool:
        return self.get(worker) is not None

    def __len__(self) -> int:
        return sum(1 for _ in self)

    def __iter__(self) -> Iterator[int]:
        for p in range(self.num_workers):
            if self._rt.afb_table_has_entry(self._h, p):
                yield p

    def items(self):
        for p in self:
            yield p, WorkerState(self, p)

    def any_state(self) -> Optional[WorkerState]:
        """Some entry to run global queries on (the reference hard-codes STAT(0), quirk B.14)."""
        for p in self:
            return WorkerState(self, p)
        return None

    # mutation helpers used by the async actions
    def set_available(self, worker: int, available: bool) -> None:
        self._rt.afb_table_set_available(self._h, worker, int(bool(available)))

    def task_done(self, worker: int, staleness: int, task_ms: float) -> int:
        return int(self._rt.afb_table_task_done(self._h, worker, int(staleness), float(task_ms)))

    def set_state(self, worker: int, staleness: int, avg_ms: float, available: bool, num_tasks: int) -> None:
        self._rt.afb_table_set_state(self._h, worker, int(staleness), float(avg_ms), int(bool(available)), int(num_tasks))

    def set_alive(self, worker: int, alive: bool) -> None:
        self._rt.afb_table_set_alive(self._h, worker, int(bool(alive)))

    def heartbeat(self, worker: int) -> None:
        self._rt.afb_table_heartbeat(self._h, worker)

    def available_workers(self) -> int:
        return int(self._rt.afb_table_available_workers(self._h))

    def max_staleness(self) -> int:
        return int(self._rt.afb_table_max_staleness(self._h))

    def snapshot(self) -> Dict[int, dict]:
        return {p: dict(staleness=ws.staleness, avg_task_time_ms=ws.avg_task_time_ms, available=ws.available,
                        num_tasks=ws.num_tasks, alive=ws.alive) for p, ws in self.items()}
