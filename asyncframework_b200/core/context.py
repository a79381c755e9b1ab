"""``AsyncContext`` — the driver-side mailbox + bookkeeping shared by all async actions.

Parity: ``ASYNCcontext`` (core/rdd/ASYNCcontext.scala:14-74): blocking result queue
(``ResultList``), worker-state table (``STAT``), logical clock counting merged non-empty results
(``CurrentTime``/``LastTime`` :24-54), ``ASYNCcollect`` (:56), ``ASYNCcollectAll`` (:60),
``getSize`` (:64), ``hasNext`` (:69), ``setRecordStat`` (:40).

The queue/clock/table are the C++ primitives in csrc/runtime/host_runtime.cpp; payloads (tensors,
Python objects) are parked in a slot table on the Python side and referenced by id.
"""
from __future__ import annotations

import itertools
import threading
from typing import Any, Dict, Generic, Optional, TypeVar

from .. import _native
from .partial_result import PartialResult
from .worker_state import StatTable

T = TypeVar("T")


class QueueClosed(RuntimeError):
    pass


class AsyncContext(Generic[T]):
    def __init__(self, num_workers: int, scheduler=None):
        self._rt = _native.runtime()
        self.num_workers = int(num_workers)
        self._q = self._rt.afb_queue_create()
        self._clock = self._rt.afb_clock_create()
        self.stat = StatTable(self.num_workers)
        self.STAT = self.stat  # reference spelling
        self._payloads: Dict[int, Any] = {}
        self._payload_lock = threading.Lock()
        self._ids = itertools.count(1)
        self._last_time = -(2**31)
        self.record_stat = False
        self.scheduler = scheduler
        self._blocking = True
        self.dropped_empty = 0
        # per-worker in-flight counter: fixes quirk B.3 (an empty completion of a later round
        # must not free a worker that still has a real task in flight)
        self._inflight = [0] * self.num_workers
        self._inflight_lock = threading.Lock()
        self.merge_lock = threading.Lock()   # JobWaiter.taskSucceeded is `synchronized`
        # tasks of async jobs that exhausted their retries: nobody awaits an async JobWaiter (DAGScheduler.scala:641-660,
        # quirk B.15), so the failure is recorded here for the drivers to surface (`raise_if_failed`)
        self.lost: list = []

    def __del__(self):
        try:
            self._rt.afb_queue_destroy(self._q)
            self._rt.afb_clock_destroy(self._clock)
        except Exception:
            pass

    # ---- logical clock ---------------------------------------------------
    @property
    def current_time(self) -> int:
        return int(self._rt.afb_clock_get(self._clock))

    @current_time.setter
    def current_time(self, t: int) -> None:
        self._rt.afb_clock_set(self._clock, int(t))

    def add_to_current_time(self, d: int = 1) -> int:
        """Atomically tick; returns the value *before* the tick."""
        return int(self._rt.afb_clock_add(self._clock, int(d)))

    @property
    def last_time(self) -> int:
        return self._last_time

    @last_time.setter
    def last_time(self, t: int) -> None:
        self._last_time = int(t)

    def is_old(self) -> bool:
        return self.current_time == self._last_time

    # ---- blocking / non-blocking job mode (sc.set_mode, A7) ---------------------
    @property
    def blocking(self) -> bool:
        return self._blocking

    @blocking.setter
    def blocking(self, v: bool) -> None:
        self._blocking = bool(v)
        if self.scheduler is not None:
            self.scheduler.set_mode(0 if v else 1)

    def set_mode(self, mode: int) -> None:
        self.blocking = (mode == 0)

    # ---- producer side (called by the async actions' merge handlers) -------------
    def put(self, result: PartialResult) -> None:
        pid = next(self._ids)
        with self._payload_lock:
            self._payloads[pid] = result.data
        rec = _native.Record(pid, int(result.timestamp), int(result.staleness), int(result.batch_size),
                             int(result.worker_id), 0, float(result.submit_ms), float(result.finish_ms))
        self._rt.afb_queue_put(self._q, rec)

    def mark_submitted(self, worker: int) -> None:
        with self._inflight_lock:
            self._inflight[worker] += 1
        self.stat.heartbeat(worker)          # "busy since": the watchdog measures silence from the submission, not from the last result
        self.stat.set_available(worker, False)

    def mark_finished(self, worker: int) -> bool:
        """Returns True when the worker has no other task in flight (=> may become available)."""
        with self._inflight_lock:
            self._inflight[worker] = max(0, self._inflight[worker] - 1)
            return self._inflight[worker] == 0

    def inflight(self, worker: int) -> int:
        return self._inflight[worker]

    # ---- consumer side -------------------------------------------------------------
    def _take(self, timeout: Optional[float]) -> Optional[PartialResult]:
        rec = _native.Record()
        rc = self._rt.afb_queue_take(self._q, rec, -1.0 if timeout is None else float(timeout) * 1e3)
        if rc == 0:
            return None
        if rc < 0:
            raise QueueClosed("AsyncContext result queue closed")
        with self._payload_lock:
            data = self._payloads.pop(rec.payload_id)
        return PartialResult(data, int(rec.staleness), int(rec.ts), int(rec.batch_size), int(rec.worker),
                             float(rec.t_submit_ms), float(rec.t_finish_ms))

    def collect(self, timeout: Optional[float] = None) -> Optional[T]:
        """Blocking take returning just the data (``ASYNCcollect``)."""
        r = self._take(timeout)
        return None if r is None else r.data

    def collect_all(self, timeout: Optional[float] = None) -> Optional[PartialResult]:
        """Blocking take returning the envelope (``ASYNCcollectAll``)."""
        return self._take(timeout)

    def drain(self, max_results: int = 256, timeout: Optional[float] = None) -> list:
        """Blocking batch take: waits like ``collect_all`` for the first result, then returns every result that is already
        queued (up to ``max_results``) from ONE native call — the server loop's way to merge a burst of completions
        without a Python<->C++ round trip per record."""
        n_max = max(1, int(max_results))
        recs = (_native.Record * n_max)()
        rc = self._rt.afb_queue_take_batch(self._q, recs, n_max, -1.0 if timeout is None else float(timeout) * 1e3)
        if rc == 0:
            return []
        if rc < 0:
            raise QueueClosed("AsyncContext result queue closed")
        out = []
        with self._payload_lock:
            for i in range(rc):
                r = recs[i]
                out.append(PartialResult(self._payloads.pop(r.payload_id), int(r.staleness), int(r.ts), int(r.batch_size),
                                         int(r.worker), float(r.t_submit_ms), float(r.t_finish_ms)))
        return out

    def size(self) -> int:
        return int(self._rt.afb_queue_size(self._q))

    def has_next(self) -> bool:
        return self.size() > 0

    def close(self) -> None:
        self._rt.afb_queue_close(self._q)

    def reopen(self) -> None:
        self._rt.afb_queue_reopen(self._q)

    # ---- broadcast factories ---------------------------------------------------------
    def broadcast(self, value):
        from .broadcast import Broadcast
        return Broadcast(value)

    def async_broadcast(self, value, slots: int = 8):
        from .broadcast import AsyncBroadcast
        return AsyncBroadcast(self, value, slots=slots)

    # reference-style aliases
    ASYNCcollect = collect
    ASYNCcollectAll = collect_all
    get_size = size
    get_current_time = lambda self: self.current_time  # noqa: E731

    def record_lost(self, worker: int, exc) -> None:
        """A task of an async job was lost for good (retries exhausted / result handler raised)."""
        self.lost.append((int(worker), exc))

    def raise_if_failed(self) -> None:
        """BSP drivers call this while waiting for P results: a blocking Spark job would raise on the driver."""
        if self.lost:
            from .scheduler import JobFailed
            worker, exc = self.lost[0]
            raise JobFailed(f"async task of worker {worker} failed permanently: {exc!r}") from (exc if isinstance(exc, BaseException) else None)

    def set_record_stat(self, b: bool) -> None:
        self.record_stat = bool(b)
