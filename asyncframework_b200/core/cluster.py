"""``local-cluster[n]``: a scheduler whose executors are separate OS processes.

This is the analogue of Spark's ``local-cluster[n, cores, memMB]`` master (core/deploy/
LocalSparkCluster.scala:49-70) that the reference's test estate uses for "multi-process on one box"
(SURVEY §4), and it exercises the pieces the in-process ``Scheduler`` short-circuits:

  * **task binary** — ``(dataset lineage, func)`` is serialised ONCE per job with cloudpickle (closures,
    captured broadcast snapshots, delay-injector state…) and cached per executor, like
    ``DAGScheduler.submitMissingTasks`` broadcasting the task binary (scheduler/DAGScheduler.scala:1051-1063);
  * **task launch / status RPC** — small control messages over per-executor pipes
    (CoarseGrainedSchedulerBackend ``LaunchTask`` / ``StatusUpdate``);
  * **task-result path** — pickled results come back to a driver-side receiver thread which invokes the
    job's result handler (JobWaiter.taskSucceeded) — so ``async_reduce`` / ``async_aggregate`` deliver
    results one at a time exactly as on the in-process backend;
  * **executor-side block store** — ``cache()``d partitions live in the executor process across jobs,
    keyed by dataset id (BlockManager MEMORY_ONLY);
  * **failure handling** — a task exception is retried on the same executor up to ``max_failures``; an
    executor process that dies is detected (its pipe closes), its in-flight tasks fail the job and the
    async bookkeeping frees the worker.

Interface-compatible with ``core.scheduler.Scheduler`` (``submit_job`` / ``run_job`` / ``set_mode`` /
``listeners`` / ``stop``), so every driver and the whole ShardedDataset API run unchanged.
"""
from __future__ import annotations

import itertools
import multiprocessing as mp
import os
import pickle
import threading
import time
import traceback
from typing import Any, Callable, Dict, Iterable, List, Optional, Sequence

import cloudpickle

from . import accumulator as _acc
from .scheduler import JobFailed, JobWaiter, ListenerBus, TaskContext  # noqa: F401


def _executor_main(eid: int, conn) -> None:
    """Executor process.  A receiver thread owns the pipe: task binaries and control messages ("drop" = the job finished
    or was cancelled, "stop") take effect immediately, task launches are queued for the single task thread — so a
    cancellation overtakes the tasks that are still queued behind a running one."""
    import queue as _queue
    import torch
    torch.set_num_threads(1)
    binaries: Dict[int, Any] = {}
    tasks: "_queue.Queue" = _queue.Queue()
    send_lock = threading.Lock()

    def send(msg) -> None:
        with send_lock:
            conn.send(msg)

    def receiver() -> None:
        while True:
            try:
                msg = conn.recv()
            except (EOFError, OSError):
                tasks.put(None)
                return
            kind = msg[0]
            if kind == "stop":
                tasks.put(None)
                return
            if kind == "binary":
                try:
                    binaries[msg[1]] = cloudpickle.loads(msg[2])
                except BaseException as exc:  # noqa: BLE001  (reported when the job's tasks run)
                    binaries[msg[1]] = exc
            elif kind == "drop":
                binaries.pop(msg[1], None)
            elif kind == "task":
                tasks.put(msg)

    threading.Thread(target=receiver, daemon=True, name=f"executor-{eid}-recv").start()
    while True:
        msg = tasks.get()
        if msg is None:
            return
        _, job_id, partition, attempt, submit_ms = msg
        t0 = time.perf_counter()
        binary = binaries.get(job_id)
        if binary is None:                        # job finished or cancelled while this task was queued
            send(("gone", job_id, partition, attempt, "", 0.0))
            continue
        try:
            if isinstance(binary, BaseException):
                raise binary
            dataset, func, fault = binary
            tctx = TaskContext(job_id, partition, attempt, eid, None, None, submit_ms)
            _acc.begin_task()
            if fault is not None:
                fault(tctx)
            result = func(tctx, dataset.iterator(partition, tctx))
            data = cloudpickle.dumps((result, _acc.end_task(True)))      # accumulator updates ride with the result
            send(("ok", job_id, partition, attempt, data, (time.perf_counter() - t0) * 1e3))
        except BaseException as exc:  # noqa: BLE001
            _acc.end_task(False)
            send(("err", job_id, partition, attempt, repr(exc) + "\n" + traceback.format_exc(limit=5), 0.0))


class ClusterScheduler:
    _job_ids = itertools.count(1_000_000)

    def __init__(self, num_executors: int = 2, max_failures: int = 4, block_first_async_job: bool = True,
                 name: str = "afb-cluster"):
        self.num_executors = int(num_executors)
        self.max_failures = int(max_failures)
        self.block_first_async_job = block_first_async_job
        self.mode = 0
        self.first_iter = 1
        self.device = None
        self.listeners = ListenerBus()
        self.fault_injector: Optional[Callable[[TaskContext], None]] = None
        self.failed_jobs: List[int] = []
        self._ctx = mp.get_context("spawn")
        self._conns = []
        self._procs = []
        self._send_locks = []
        self._jobs: Dict[int, dict] = {}
        self._lock = threading.Lock()
        self._stopped = False
        self._blocks_held: Dict[int, set] = {}       # source block id -> executors that already hold its data
        self.last_task_binary_bytes = 0
        for e in range(self.num_executors):
            parent, child = self._ctx.Pipe()
            p = self._ctx.Process(target=_executor_main, args=(e, child), name=f"{name}-executor-{e}", daemon=True)
            p.start()
            child.close()
            self._conns.append(parent)
            self._procs.append(p)
            self._send_locks.append(threading.Lock())
        self._receivers = [threading.Thread(target=self._recv_loop, args=(e,), daemon=True, name=f"{name}-recv-{e}")
                           for e in range(self.num_executors)]
        for t in self._receivers:
            t.start()

    # ------------------------------------------------------------------ mode
    def set_mode(self, mode: int) -> None:
        self.mode = int(mode)

    def stream_for(self, partition: int):
        return None

    # ------------------------------------------------------------------ plumbing
    def _send(self, eid: int, msg) -> None:
        with self._send_locks[eid]:
            self._conns[eid].send(msg)

    def _recv_loop(self, eid: int) -> None:
        conn = self._conns[eid]
        while not self._stopped:
            try:
                msg = conn.recv()
            except (EOFError, OSError):
                self._executor_lost(eid)
                return
            kind, job_id, partition, attempt, data, run_ms = msg
            with self._lock:
                job = self._jobs.get(job_id)
            if job is None:
                continue
            waiter: JobWaiter = job["waiter"]
            if kind == "gone":
                continue
            if kind == "ok":
                self.listeners.post({"event": "task_end", "job": job_id, "partition": partition, "executor": eid,
                                     "run_ms": run_ms, "metrics": {"result_bytes": len(data)}})
                try:
                    result, acc_updates = pickle.loads(data)
                    if not waiter.cancelled:
                        _acc.merge_updates(acc_updates)
                    waiter.task_succeeded(partition, result)
                except Exception as exc:  # a raising result handler fails the job, not the receiver thread
                    waiter.job_failed(exc)
                job["pending"].discard(partition)
                if waiter.done:
                    self._finish_job(job_id)
            else:
                attempt += 1
                self.listeners.post({"event": "task_failed", "job": job_id, "partition": partition, "attempt": attempt, "error": data})
                if attempt >= self.max_failures:
                    self.failed_jobs.append(job_id)
                    exc = RuntimeError(f"task {partition} of job {job_id} failed {attempt} times: {data.splitlines()[0]}")
                    if job["on_task_lost"] is not None:
                        job["on_task_lost"](partition, exc)
                    waiter.job_failed(exc)
                    job["pending"].discard(partition)
                else:
                    self._send(eid, ("task", job_id, partition, attempt, job["submit_ms"]))

    def cancel_job(self, job_id: int, reason: str = "cancelled by user") -> bool:
        """``SparkContext.cancelJob`` over executor processes: the waiter raises JobCancelled, the task binary is dropped on
        the executors (queued tasks of the job then return immediately), results still in flight are ignored and the
        submitter is told about every partition that will not deliver (async actions free their workers)."""
        with self._lock:
            job = self._jobs.pop(job_id, None)
        if job is None or job["waiter"].done:
            return False
        job["waiter"].cancel(reason)
        for e in job["executors"]:
            try:
                self._send(e, ("drop", job_id))
            except (OSError, BrokenPipeError):
                pass
        if job["on_task_lost"] is not None:
            for p in sorted(job["pending"]):
                job["on_task_lost"](p, job["waiter"].failure)
        self.listeners.post({"event": "job_end", "job": job_id, "cancelled": True, "t_ms": time.perf_counter() * 1e3})
        return True

    def cancel_all_jobs(self) -> int:
        with self._lock:
            ids = list(self._jobs)
        return sum(1 for j in ids if self.cancel_job(j))

    def _executor_lost(self, eid: int) -> None:
        if self._stopped:
            return
        with self._lock:
            jobs = list(self._jobs.items())
        for job_id, job in jobs:
            lost = [p for p in list(job["pending"]) if p % self.num_executors == eid]
            for p in lost:
                exc = RuntimeError(f"executor {eid} lost while running partition {p} of job {job_id}")
                if job["on_task_lost"] is not None:
                    job["on_task_lost"](p, exc)
                job["waiter"].job_failed(exc)
                job["pending"].discard(p)

    def _finish_job(self, job_id: int) -> None:
        with self._lock:
            job = self._jobs.pop(job_id, None)
        if job is None:
            return
        for e in job["executors"]:
            try:
                self._send(e, ("drop", job_id))
            except (OSError, BrokenPipeError):
                pass
        self.listeners.post({"event": "job_end", "job": job_id, "t_ms": time.perf_counter() * 1e3})

    # ------------------------------------------------------------------ jobs
    def submit_job(self, dataset, func, partitions: Optional[Sequence[int]] = None,
                   result_handler: Optional[Callable[[int, Any], None]] = None,
                   on_task_lost: Optional[Callable[[int, BaseException], None]] = None) -> JobWaiter:
        if self._stopped:
            raise RuntimeError("scheduler stopped")
        parts = list(range(dataset.num_partitions)) if partitions is None else list(partitions)
        job_id = next(ClusterScheduler._job_ids)
        waiter = JobWaiter(job_id, len(parts), result_handler)
        now = time.perf_counter() * 1e3
        executors = sorted({p % self.num_executors for p in parts})
        # source data travels once per executor: blocks every target executor already holds are pickled as references
        from ..data.dataset import ShippingContext
        with self._lock:
            held = {b: set(es) for b, es in self._blocks_held.items()}
        with ShippingContext(lambda b: all(e in held.get(b, ()) for e in executors)) as ship:
            payload = cloudpickle.dumps((dataset, func, self.fault_injector))
        self.last_task_binary_bytes = len(payload)
        with self._lock:
            self._jobs[job_id] = {"waiter": waiter, "pending": set(parts), "on_task_lost": on_task_lost,
                                  "executors": executors, "submit_ms": now}
        self.listeners.post({"event": "job_start", "job": job_id, "partitions": parts, "t_ms": now,
                             "task_binary_bytes": len(payload)})
        for e in executors:
            self._send(e, ("binary", job_id, payload))
        with self._lock:                              # only now may a later job refer to these blocks by id (pipe order)
            for b in ship.shipped:
                self._blocks_held.setdefault(b, set()).update(executors)
        for p in parts:
            self.listeners.post({"event": "task_start", "job": job_id, "partition": p, "attempt": 0,
                                 "executor": p % self.num_executors, "t_ms": now})
            self._send(p % self.num_executors, ("task", job_id, p, 0, now))
        if not parts:
            self._finish_job(job_id)
        return waiter

    def run_job(self, dataset, func, partitions=None, result_handler=None, on_task_lost=None, blocking: Optional[bool] = None):
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

    def kill_executor(self, eid: int) -> None:
        """Fault injection: kill one executor process (FailureSuite / DistributedSuite style)."""
        self._procs[eid].kill()

    def stop(self) -> None:
        if self._stopped:
            return
        self._stopped = True
        for e, c in enumerate(self._conns):
            try:
                with self._send_locks[e]:
                    c.send(("stop",))
            except (OSError, BrokenPipeError):
                pass
        for p in self._procs:
            p.join(timeout=5)
            if p.is_alive():
                p.terminate()
        for c in self._conns:
            try:
                c.close()
            except OSError:
                pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.stop()
