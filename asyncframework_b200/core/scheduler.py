"""In-process job scheduler: the ``local[N]`` substrate the reference's samples actually run on.

Replaces SparkContext.runJob -> DAGScheduler -> TaskSchedulerImpl -> LocalSchedulerBackend ->
Executor (core/SparkContext.scala:2038-2064, scheduler/DAGScheduler.scala:587-660,
scheduler/local/LocalSchedulerBackend.scala:83-88, executor/Executor.scala:290-560) with the part
ASYNC needs: single-stage jobs of one task per partition, **hard** partition->executor affinity
(partition p always runs on executor p % N; the reference only has best-effort locality, S4),
per-task retry (spark.task.maxFailures = 4), a result handler invoked under a lock as tasks
finish (JobWaiter.taskSucceeded, scheduler/JobWaiter.scala:56-64), and the ASYNC delta:
**non-blocking mode** — ``set_mode(1)`` makes ``run_job`` return right after submission
(DAGScheduler.scala:641-660), with the reference's "first async job still blocks" quirk
available as an option (``block_first_async_job``; Appendix B.4).

Executors are Python threads; heavy work inside tasks is torch / CUDA which releases the GIL.
When a CUDA device is given each executor owns a CUDA stream so logical workers overlap on one GPU.
"""
from __future__ import annotations

import itertools
import queue
import threading
import time
import traceback
from dataclasses import dataclass, field
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence

import torch

from . import accumulator as _acc


@dataclass
class TaskContext:
    job_id: int
    partition_id: int
    attempt: int
    executor_id: int
    device: Optional[torch.device] = None
    stream: Any = None
    submit_ms: float = 0.0
    metrics: Dict[str, float] = field(default_factory=dict)
    job: Any = None
    killed: bool = False                     # this ATTEMPT was killed (the other copy of a speculated task won)

    def is_cancelled(self) -> bool:
        """``TaskContext.isInterrupted``: true once the job this task belongs to has been cancelled, or this attempt killed."""
        return self.killed or bool(getattr(self.job, "cancelled", False))


_job_group = threading.local()          # sc.setJobGroup: a property of the submitting thread (SparkContext.setLocalProperty)


def set_job_group(group_id: Optional[str]) -> None:
    _job_group.id = group_id


class JobFailed(RuntimeError):
    pass


class JobCancelled(JobFailed):
    """Raised by ``JobWaiter.wait`` after ``cancel()`` (SparkException "Job N cancelled", DAGScheduler.scala:816-823)."""


class JobWaiter:
    """Tracks completion of one job (scheduler/JobWaiter.scala)."""

    def __init__(self, job_id: int, total: int, result_handler: Optional[Callable[[int, Any], None]]):
        self.job_id = job_id
        self.total = total
        self.finished = 0
        self.result_handler = result_handler
        self._lock = threading.Lock()
        self._done = threading.Event()
        self.failure: Optional[BaseException] = None
        self.cancelled = False
        if total == 0:
            self._done.set()

    def cancel(self, reason: str = "cancelled by user") -> None:
        """``JobWaiter.cancel`` / ``SparkContext.cancelJob`` (scheduler/JobWaiter.scala:49-51): tasks of this job that have
        not started are skipped, results of tasks that are already running are discarded, ``wait`` raises
        :class:`JobCancelled`.  Running closures can poll ``TaskContext.is_cancelled()`` to stop early."""
        with self._lock:
            if self._done.is_set():
                return
            self.cancelled = True
            self.failure = JobCancelled(f"job {self.job_id} {reason}")
            self._done.set()

    def task_succeeded(self, index: int, result: Any) -> None:
        with self._lock:  # `synchronized` in the reference
            if self.cancelled:
                return
            if self.result_handler is not None:
                self.result_handler(index, result)
            self.finished += 1
            if self.finished >= self.total:
                self._done.set()

    def job_failed(self, exc: BaseException) -> None:
        with self._lock:
            self.failure = exc
            self._done.set()

    def wait(self, timeout: Optional[float] = None) -> bool:
        ok = self._done.wait(timeout)
        if isinstance(self.failure, JobCancelled):
            raise self.failure
        if self.failure is not None:
            raise JobFailed(f"job {self.job_id} failed: {self.failure}") from self.failure
        return ok

    @property
    def done(self) -> bool:
        return self._done.is_set()


class ListenerBus:
    """Synchronous event fan-out (LiveListenerBus analogue, S14).  Events are plain dicts:
    job_start / task_start / task_end / task_failed / job_end."""

    def __init__(self):
        self._listeners: List[Callable[[dict], None]] = []

    def add(self, fn: Callable[[dict], None]) -> None:
        self._listeners.append(fn)

    def post(self, event: dict) -> None:
        for fn in self._listeners:
            try:
                fn(event)
            except Exception:  # listeners must never take the scheduler down
                traceback.print_exc()


class _Task:
    __slots__ = ("job", "partition", "func", "dataset", "attempt", "submit_ms")

    def __init__(self, job, partition, func, dataset, submit_ms):
        self.job, self.partition, self.func, self.dataset = job, partition, func, dataset
        self.attempt = 0
        self.submit_ms = submit_ms


class Scheduler:
    _job_ids = itertools.count(0)

    def __init__(self, num_executors: int = 8, max_failures: int = 4, device=None,
                 block_first_async_job: bool = True, name: str = "afb"):
        self.num_executors = int(num_executors)
        self.max_failures = int(max_failures)
        self.device = torch.device(device) if device is not None else None
        self.block_first_async_job = block_first_async_job
        self.mode = 0            # 0 = blocking, 1 = non-blocking (sc.set_mode)
        self.first_iter = 1
        self.listeners = ListenerBus()
        self.fault_injector: Optional[Callable[[TaskContext], None]] = None
        self.failed_jobs: List[int] = []
        self._queues: List["queue.Queue[Optional[_Task]]"] = [queue.Queue() for _ in range(self.num_executors)]
        self._threads: List[threading.Thread] = []
        self._stopped = False
        self._active: Dict[int, JobWaiter] = {}
        self._jobs_lock = threading.Lock()
        self._streams = [None] * self.num_executors
        for e in range(self.num_executors):
            t = threading.Thread(target=self._executor_loop, args=(e,), name=f"{name}-executor-{e}", daemon=True)
            t.start()
            self._threads.append(t)

    # ------------------------------------------------------------------ mode
    def set_mode(self, mode: int) -> None:
        self.mode = int(mode)

    # ------------------------------------------------------------------ executors
    def _executor_loop(self, eid: int) -> None:
        # ``local[N]`` = N task slots sharing the machine: every executor thread gets its share of the cores for intra-op
        # parallelism (at::set_num_threads is per calling thread under OpenMP).  Without this, N tasks x all-cores intra-op
        # threads oversubscribe the box — measured here: a 378 KB ``clone()`` costs 8.4 ms with 8 intra-op threads vs 11 us with 1.
        import os as _os
        torch.set_num_threads(max(1, (_os.cpu_count() or 1) // max(1, self.num_executors)))
        stream = None
        if self.device is not None and self.device.type == "cuda":
            torch.cuda.set_device(self.device)
            stream = torch.cuda.Stream(device=self.device)
            self._streams[eid] = stream
        q = self._queues[eid]
        while True:
            task = q.get()
            if task is None:
                return
            try:
                self._run_task(task, eid, stream)
            except BaseException as exc:  # noqa: BLE001 — an executor thread must never die: later tasks would never be dequeued
                self._fail_job(task.job, task.partition, exc)

    def _run_task(self, task: _Task, eid: int, stream) -> None:
        job: JobWaiter = task.job
        # NB: a failed sibling task does not cancel this one — in async mode every worker's result is
        # independent (the reference silently ignores failed async jobs).
        tctx = TaskContext(job.job_id, task.partition, task.attempt, eid, self.device, stream, task.submit_ms, job=job)
        if job.cancelled:                       # cancelled before this task started: skip it
            self.listeners.post({"event": "task_skipped", "job": job.job_id, "partition": task.partition, "executor": eid})
            self._task_lost(job, task.partition)
            return
        self.listeners.post({"event": "task_start", "job": job.job_id, "partition": task.partition,
                             "attempt": task.attempt, "executor": eid, "t_ms": time.perf_counter() * 1e3})
        t0 = time.perf_counter()
        _acc.begin_task()
        try:
            if self.fault_injector is not None:
                self.fault_injector(tctx)
            if stream is not None:
                with torch.cuda.stream(stream):
                    result = task.func(tctx, task.dataset.iterator(task.partition, tctx))
                    stream.synchronize()
            else:
                result = task.func(tctx, task.dataset.iterator(task.partition, tctx))
        except Exception as exc:  # noqa: BLE001
            _acc.end_task(False)                # a failed attempt's accumulator updates are dropped
            if job.cancelled:                   # an interrupted task of a cancelled job is not a failure to retry
                self._task_lost(job, task.partition)
                return
            task.attempt += 1
            self.listeners.post({"event": "task_failed", "job": job.job_id, "partition": task.partition,
                                 "attempt": task.attempt, "error": repr(exc)})
            if task.attempt >= self.max_failures:
                self.failed_jobs.append(job.job_id)
                first_failure = job.failure is None
                job.job_failed(exc)
                if first_failure:
                    self.listeners.post({"event": "job_end", "job": job.job_id, "failed": True, "error": repr(exc)[:300],
                                         "t_ms": time.perf_counter() * 1e3})
                on_fail = getattr(job, "on_task_lost", None)
                if on_fail is not None:
                    on_fail(task.partition, exc)
            else:
                self._queues[eid].put(task)   # retry on the same executor (hard affinity)
            return
        if job.cancelled:                       # finished after the cancellation: the result is discarded
            _acc.end_task(False)
            self._task_lost(job, task.partition)
            return
        _acc.merge_updates(_acc.end_task(True))   # exactly once, for the successful attempt
        dt = (time.perf_counter() - t0) * 1e3
        self.listeners.post({"event": "task_end", "job": job.job_id, "partition": task.partition,
                             "executor": eid, "run_ms": dt, "metrics": tctx.metrics})
        try:
            job.task_succeeded(task.partition, result)
        except Exception as exc:  # noqa: BLE001 — the user's result handler raised (e.g. merge of an empty aggregate with
            # strict_empty=True): the job fails on the driver like a blocking Spark job would (JobWaiter.scala:56-64 runs the
            # handler under the listener's lock and a throw fails the job), the worker is freed, the executor thread lives on
            self._fail_job(job, task.partition, exc)
            return
        if job.done:
            self.listeners.post({"event": "job_end", "job": job.job_id, "t_ms": time.perf_counter() * 1e3})

    def _fail_job(self, job: JobWaiter, partition: int, exc: BaseException) -> None:
        self.failed_jobs.append(job.job_id)
        first_failure = job.failure is None
        try:
            job.job_failed(exc)
        except Exception:  # noqa: BLE001
            pass
        if first_failure:
            self.listeners.post({"event": "job_end", "job": job.job_id, "failed": True, "error": repr(exc)[:300],
                                 "t_ms": time.perf_counter() * 1e3})
        on_fail = getattr(job, "on_task_lost", None)
        if on_fail is not None:
            try:
                on_fail(partition, exc)
            except Exception:  # noqa: BLE001
                pass

    @staticmethod
    def _task_lost(job: JobWaiter, partition: int) -> None:
        """Tell the submitter (async actions free the worker here) that this partition will not deliver a result."""
        on_fail = getattr(job, "on_task_lost", None)
        if on_fail is not None:
            on_fail(partition, job.failure)

    # ------------------------------------------------------------------ jobs
    def submit_job(self, dataset, func: Callable[[TaskContext, Iterable], Any],
                   partitions: Optional[Sequence[int]] = None,
                   result_handler: Optional[Callable[[int, Any], None]] = None,
                   on_task_lost: Optional[Callable[[int, BaseException], None]] = None) -> JobWaiter:
        if self._stopped:
            raise RuntimeError("scheduler stopped")
        if getattr(self, "profiler", None) is not None:
            func = self.profiler.wrap(func)               # spark.python.profile: every task under cProfile
        parts = list(range(dataset.num_partitions)) if partitions is None else list(partitions)
        job = JobWaiter(next(Scheduler._job_ids), len(parts), result_handler)
        job.on_task_lost = on_task_lost  # type: ignore[attr-defined]
        job.group = getattr(_job_group, "id", None)       # type: ignore[attr-defined]  (sc.setJobGroup is per submitting thread)
        with self._jobs_lock:
            self._active = {j: w for j, w in self._active.items() if not w.done}
            self._active[job.job_id] = job
        now = time.perf_counter() * 1e3
        self.listeners.post({"event": "job_start", "job": job.job_id, "partitions": parts, "t_ms": now})
        for p in parts:
            self._queues[p % self.num_executors].put(_Task(job, p, func, dataset, now))
        return job

    def run_job(self, dataset, func, partitions=None, result_handler=None, on_task_lost=None, blocking: Optional[bool] = None):
        """Blocking in mode 0 (returns the per-partition results in order); in mode 1 returns
        immediately after submission (except, optionally, for the very first async job)."""
        results: Dict[int, Any] = {}
        if result_handler is None:
            def result_handler(i, r):  # noqa: E306
                results[i] = r
        waiter = self.submit_job(dataset, func, partitions, result_handler, on_task_lost)
        # `blocking` overrides the scheduler-wide mode for THIS job: blocking actions (collect/reduce/...) pass True and
        # async actions pass False, so concurrent driver threads never observe each other's temporary mode flips
        nonblocking = (self.mode == 1) if blocking is None else (not blocking)
        if not nonblocking or (self.first_iter == 1 and self.block_first_async_job):
            waiter.wait()
            if nonblocking:
                self.first_iter = 0
            parts = list(range(dataset.num_partitions)) if partitions is None else list(partitions)
            return [results.get(p) for p in parts]
        self.first_iter = 0
        return waiter

    def cancel_job(self, job_id: int, reason: str = "cancelled by user") -> bool:
        """``SparkContext.cancelJob``; returns False when the job is unknown or already finished."""
        with self._jobs_lock:
            job = self._active.get(job_id)
        if job is None or job.done:
            return False
        job.cancel(reason)
        self.listeners.post({"event": "job_end", "job": job_id, "cancelled": True, "t_ms": time.perf_counter() * 1e3})
        return True

    def cancel_job_group(self, group_id: str) -> int:
        """``SparkContext.cancelJobGroup``: cancel every active job submitted under ``set_job_group(group_id)``."""
        with self._jobs_lock:
            ids = [j for j, w in self._active.items() if not w.done and getattr(w, "group", None) == group_id]
        return sum(1 for j in ids if self.cancel_job(j, f"cancelled as part of job group {group_id}"))

    def active_job_ids(self, group_id: Optional[str] = None) -> List[int]:
        with self._jobs_lock:
            return sorted(j for j, w in self._active.items() if not w.done and (group_id is None or getattr(w, "group", None) == group_id))

    def cancel_all_jobs(self) -> int:
        """``SparkContext.cancelAllJobs``: returns how many jobs were cancelled."""
        with self._jobs_lock:
            ids = [j for j, w in self._active.items() if not w.done]
        return sum(1 for j in ids if self.cancel_job(j))

    def stream_for(self, partition: int):
        return self._streams[partition % self.num_executors]

    def stop(self) -> None:
        if self._stopped:
            return
        self._stopped = True
        for q in self._queues:
            q.put(None)
        for t in self._threads:
            t.join(timeout=5)

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.stop()
