"""Cluster-mode scheduler: executors are separate OS processes (``local-cluster[n]``) or processes launched by
standalone Worker daemons on any host (``spark://host:port``, see :mod:`asyncframework_b200.deploy`).

This is the analogue of Spark's ``local-cluster[n, cores, memMB]`` master (core/deploy/LocalSparkCluster.scala:49-70)
and of ``CoarseGrainedSchedulerBackend`` / ``CoarseGrainedExecutorBackend`` (core/scheduler/cluster/
CoarseGrainedSchedulerBackend.scala:121,239,288-311; core/executor/CoarseGrainedExecutorBackend.scala:92,143), and
it exercises the pieces the in-process ``Scheduler`` short-circuits:

  * **task binary** — ``(dataset lineage, func)`` is serialised ONCE per job with cloudpickle (closures, captured
    broadcast handles, delay-injector state…) and cached per executor, like ``DAGScheduler.submitMissingTasks``
    broadcasting the task binary (scheduler/DAGScheduler.scala:1051-1063);
  * **task launch / status RPC** — small control messages (``LaunchTask`` / ``StatusUpdate``) over a pluggable
    transport: per-executor pipes (``local-cluster``) or the TCP :class:`~asyncframework_b200.rpc.RpcEnv`
    (standalone);
  * **task-result path** — a serialised result up to ``max_direct_result_size`` (1 MiB) rides inside the status
    message; a larger one is stored in the executor's BlockManager and the driver **pulls** it over the bulk
    chunk-fetch path on a 4-thread result-getter pool, then deletes the block (``IndirectTaskResult``,
    executor/Executor.scala:453-480; scheduler/TaskResultGetter.scala:57-125); results that would take the
    job's total above ``max_result_size`` abort the job (TaskSetManager.scala:706-712);
  * **per-process runtime** — the driver and every executor own an :class:`~asyncframework_b200.core.env.AsyncEnv`
    (RpcEnv + BlockManager registered with the driver's BlockManagerMaster): ``cache()``d partitions are
    ``rdd_<id>_<p>`` blocks of the executor's store, broadcasts are piece-wise TorrentBroadcasts;
  * **failure handling** — a task exception is retried on the same executor up to ``max_failures``; an executor
    process that dies is detected (pipe EOF / RPC disconnect).  With ``recover_lost_executors`` the slot is
    re-filled (a new process for ``local-cluster``, a Master-scheduled replacement for standalone), the lost
    tasks are re-launched there and the blocks it held are re-shipped (executor-lost resubmission,
    DAGScheduler.scala:1520); otherwise its in-flight tasks fail the job and the async bookkeeping frees the worker.

Interface-compatible with ``core.scheduler.Scheduler`` (``submit_job`` / ``run_job`` / ``set_mode`` /
``listeners`` / ``stop``), so every driver and the whole ShardedDataset API run unchanged.
"""
from __future__ import annotations

import itertools
import multiprocessing as mp
import os
import pickle
import queue as _queue
import threading
import time
import traceback
from concurrent.futures import ThreadPoolExecutor
from typing import Any, Callable, Dict, List, Optional, Sequence

import cloudpickle

from . import accumulator as _acc
from .scheduler import JobFailed, JobWaiter, ListenerBus, TaskContext  # noqa: F401


def thread_dump() -> List[dict]:
    """Stack of every thread of this process (``Utils.getThreadDump``; served by the UI's thread-dump page)."""
    import sys
    names = {t.ident: t.name for t in threading.enumerate()}
    return [{"thread_id": tid, "name": names.get(tid, "?"), "stack": "".join(traceback.format_stack(frame))}
            for tid, frame in sorted(sys._current_frames().items())]


# =================================================================================================== executor side
class TaskKilled(BaseException):
    """Raised INSIDE a running task when the driver kills that attempt (``TaskKilledException``; the reference interrupts
    the task thread, executor/Executor.scala:223-243 ``kill(interruptThread)``).  A BaseException so that a task's own
    ``except Exception`` blocks do not swallow it."""


def _interrupt_thread(thread_id: int, exc_type) -> bool:
    """``Thread.interrupt`` for a python thread: the exception is raised at that thread's next bytecode boundary (a task
    blocked inside a C call sees it when the call returns)."""
    import ctypes
    n = ctypes.pythonapi.PyThreadState_SetAsyncExc(ctypes.c_ulong(thread_id), ctypes.py_object(exc_type))
    if n > 1:                                          # "should never happen": undo
        ctypes.pythonapi.PyThreadState_SetAsyncExc(ctypes.c_ulong(thread_id), None)
    return n == 1


class ExecutorBackend:
    """Transport-agnostic executor.  Feed it driver messages with :meth:`handle` (from a receiver thread / an RPC
    inbox) and run :meth:`run_tasks` on the task thread; status messages leave through ``send``.

    Task binaries and control messages ("drop" = the job finished or was cancelled, "stop") take effect
    immediately, task launches are queued for the task thread — so a cancellation overtakes the tasks that are
    still queued behind a running one."""

    def __init__(self, eid: int, send: Callable[[tuple], None], cores: int = 1):
        self.eid = eid
        self.cores = max(1, int(cores))              # spark.executor.cores: task slots = concurrent task threads
        self._send = send
        self.binaries: Dict[int, Any] = {}
        self.tasks: "_queue.Queue" = _queue.Queue()
        self._send_lock = threading.Lock()
        self.tasks_run = 0
        self.tasks_killed = 0
        self._run_lock = threading.Lock()
        self._running: Dict[tuple, tuple] = {}       # (job, partition, attempt) -> (thread id, TaskContext) while the body runs
        self._kill_requests: set = set()             # attempts killed before they left the queue

    def send(self, msg: tuple) -> None:
        with self._send_lock:
            self._send(msg)

    def handle(self, msg: tuple) -> bool:
        kind = msg[0]
        if kind == "stop":
            for _ in range(self.cores):
                self.tasks.put(None)
            return False
        if kind == "binary":
            try:
                self.binaries[msg[1]] = cloudpickle.loads(msg[2])
            except BaseException as exc:  # noqa: BLE001  (reported when the job's tasks run)
                self.binaries[msg[1]] = exc
        elif kind == "drop":
            self.binaries.pop(msg[1], None)
        elif kind == "task":
            self.tasks.put(msg)
        elif kind == "kill_task":                        # TaskSetManager.scala:736: the other copy of this task won
            self.kill_task(msg[1], msg[2], msg[3])
        elif kind == "drop_source":                      # ContextCleaner: the driver's copy of this source block is gone
            from ..data.dataset import drop_source_block
            drop_source_block(msg[1])
        elif kind == "thread_dump":                      # ExecutorThreadDumpPage: answered by the receiver thread, even mid-task
            self.send(("threads", -1, -1, msg[1], thread_dump(), 0.0))
        return True

    def kill_task(self, job_id: int, partition: int, attempt: int) -> None:
        """Interrupt one attempt: running -> ``TaskKilled`` is raised in its thread (and ``TaskContext.is_cancelled()`` turns
        true for cooperative bodies); still queued -> it is skipped when it reaches the head of the queue."""
        key = (job_id, partition, attempt)
        with self._run_lock:
            running = self._running.get(key)
            if running is None:
                self._kill_requests.add(key)
                return
            tid, tctx = running
            tctx.killed = True
            _interrupt_thread(tid, TaskKilled)

    def close(self) -> None:
        self._closed = True
        for _ in range(self.cores):
            self.tasks.put(None)

    def start_heartbeats(self, interval: float) -> None:
        """Executor -> driver liveness + progress (``Executor.reportHeartBeat``, executor/Executor.scala:770-815): a
        separate thread, so a long task still heartbeats but a wedged or stopped PROCESS does not."""
        def loop():
            import os
            while not getattr(self, "_closed", False):
                time.sleep(interval)
                try:
                    from .env import AsyncEnv
                    env = AsyncEnv.get()
                    mem = env.block_manager.memory_status() if env is not None else (0, 0)
                    self.send(("heartbeat", -1, -1, 0, {"tasks_run": self.tasks_run, "queued": self.tasks.qsize(),
                                                       "active": self.active_task, "pid": os.getpid(),
                                                       "storage_max": mem[0], "storage_free": mem[1]}, 0.0))
                except Exception:  # noqa: BLE001  (the driver is gone: the receiver will shut us down)
                    return
        threading.Thread(target=loop, daemon=True, name=f"executor-{self.eid}-heartbeat").start()

    active_task = None

    def run_tasks(self) -> None:
        """Runs the task loop on the calling thread plus ``cores - 1`` helper threads; returns when all have stopped."""
        helpers = [threading.Thread(target=self._task_loop, daemon=True, name=f"executor-{self.eid}-task-{i}")
                   for i in range(1, self.cores)]
        for t in helpers:
            t.start()
        self._task_loop()
        for t in helpers:
            t.join(timeout=5)

    def _task_loop(self) -> None:
        from .env import AsyncEnv
        from ..storage.block_manager import StorageLevel
        while True:
            msg = self.tasks.get()
            if msg is None:
                return
            _, job_id, partition, attempt, submit_ms = msg
            t0 = time.perf_counter()
            binary = self.binaries.get(job_id)
            if binary is None:                        # job finished or cancelled while this task was queued
                self.send(("gone", job_id, partition, attempt, "", 0.0))
                continue
            key = (job_id, partition, attempt)
            with self._run_lock:
                if key in self._kill_requests:            # killed while it was still queued
                    self._kill_requests.discard(key)
                    self.tasks_killed += 1
                    self.send(("killed", job_id, partition, attempt, "before start", 0.0))
                    continue
            self.send(("started", job_id, partition, attempt, "", 0.0))     # queued != running: speculation needs to know
            self.active_task = (job_id, partition, attempt)
            registered = False
            try:
                if isinstance(binary, BaseException):
                    raise binary
                dataset, func, fault = binary
                tctx = TaskContext(job_id, partition, attempt, self.eid, None, None, submit_ms)
                with self._run_lock:
                    self._running[key] = (threading.get_ident(), tctx)
                    registered = True
                _acc.begin_task()
                if fault is not None:
                    fault(tctx)
                result = func(tctx, dataset.iterator(partition, tctx))
                with self._run_lock:                      # past this point a kill can no longer interrupt the thread
                    self._running.pop(key, None)
                    registered = False
                t1 = time.perf_counter()
                data = cloudpickle.dumps((result, _acc.end_task(True)))      # accumulator updates ride with the result
                t2 = time.perf_counter()
                # TaskMetrics (executor/TaskMetrics.scala): executorRunTime, resultSerializationTime
                run_ms = {"run_ms": (t1 - t0) * 1e3, "serialize_ms": (t2 - t1) * 1e3}
                # what the task recorded itself: memoryBytesSpilled / diskBytesSpilled / peakExecutionMemory of its sorters
                run_ms.update({k: v for k, v in tctx.metrics.items() if isinstance(v, (int, float)) and k not in run_ms})
                self.tasks_run += 1
                env = AsyncEnv.get()
                if env is not None and len(data) > env.limits.max_direct_result_size:
                    block_id = f"taskresult_{job_id}_{partition}_{attempt}"
                    env.block_manager.put_bytes(block_id, data, StorageLevel.MEMORY_AND_DISK_SER)  # type: ignore[attr-defined]
                    self.send(("indirect", job_id, partition, attempt, (block_id, len(data), tuple(env.block_manager.id)), run_ms))
                else:
                    self.send(("ok", job_id, partition, attempt, data, run_ms))
            except TaskKilled:
                _acc.end_task(False)
                self.tasks_killed += 1
                self.send(("killed", job_id, partition, attempt, "interrupted", (time.perf_counter() - t0) * 1e3))
            except BaseException as exc:  # noqa: BLE001
                _acc.end_task(False)
                self.send(("err", job_id, partition, attempt, repr(exc) + "\n" + traceback.format_exc(limit=5), 0.0))
            finally:
                if registered:
                    with self._run_lock:
                        self._running.pop(key, None)
                self.active_task = None


def _executor_main(eid: int, conn, driver_address=None, limits=None, cores: int = 1) -> None:
    """``local-cluster`` executor process: control messages over a pipe, blocks over its own RpcEnv."""
    import torch
    torch.set_num_threads(1)
    env = None
    if driver_address is not None:
        from .env import AsyncEnv
        env = AsyncEnv.create_executor(str(eid), driver_address, limits)
    backend = ExecutorBackend(eid, conn.send, cores)

    def receiver() -> None:
        while True:
            try:
                msg = conn.recv()
            except (EOFError, OSError):
                backend.close()
                return
            if not backend.handle(msg):
                return

    threading.Thread(target=receiver, daemon=True, name=f"executor-{eid}-recv").start()
    if limits is not None:
        backend.start_heartbeats(limits.heartbeat_interval)
    try:
        backend.run_tasks()
    finally:
        if env is not None:
            env.stop()


# =================================================================================================== transports
class PipeTransport:
    """``local-cluster[n]``: the scheduler forks ``n`` executor processes itself and talks to them over pipes."""

    def __init__(self, num_executors: int, name: str, cores_per_executor: int = 1):
        self.num_executors = num_executors
        self.cores_per_executor = max(1, int(cores_per_executor))
        self.name = name
        # executor processes come from a FORK SERVER that has torch and this package imported already: a new executor starts in
        # tens of milliseconds instead of re-importing torch (~1.5 s), and — unlike a plain fork of the driver — never inherits
        # the driver's threads, locks or CUDA context (the server process itself never touches CUDA).  AFB_EXECUTOR_START=spawn
        # restores the slower, fully independent start method.
        method = os.environ.get("AFB_EXECUTOR_START", "forkserver")
        self._ctx = mp.get_context(method)
        if method == "forkserver":
            try:
                self._ctx.set_forkserver_preload(["torch", "cloudpickle", "asyncframework_b200", "asyncframework_b200.core.cluster",
                                                  "asyncframework_b200.core.env", "asyncframework_b200.data.dataset"])
            except Exception:  # noqa: BLE001  (the server is already running with its preload list: fine)
                pass
        self._conns: List[Any] = [None] * num_executors
        self._procs: List[Any] = [None] * num_executors
        self._send_locks = [threading.Lock() for _ in range(num_executors)]
        self._gen = [0] * num_executors
        self.sched: "ClusterScheduler" = None  # type: ignore[assignment]

    def start(self, sched: "ClusterScheduler") -> None:
        self.sched = sched
        for e in range(self.num_executors):
            self._launch(e)

    def _launch(self, e: int) -> None:
        parent, child = self._ctx.Pipe()
        env = self.sched.env
        p = self._ctx.Process(target=_executor_main, name=f"{self.name}-executor-{e}", daemon=True,
                              args=(e, child, tuple(env.rpc_env.address) if env is not None else None,
                                    env.limits if env is not None else None, self.cores_per_executor))
        p.start()
        child.close()
        self._conns[e], self._procs[e] = parent, p
        self._gen[e] += 1
        threading.Thread(target=self._recv_loop, args=(e, parent, self._gen[e]), daemon=True,
                         name=f"{self.name}-recv-{e}").start()

    def _recv_loop(self, eid: int, conn, gen: int) -> None:
        while not self.sched._stopped:
            try:
                msg = conn.recv()
            except (EOFError, OSError):
                if gen == self._gen[eid]:
                    self.sched._executor_lost(eid)
                return
            self.sched._on_message(eid, msg)

    def send(self, eid: int, msg: tuple) -> None:
        with self._send_locks[eid]:
            self._conns[eid].send(msg)

    def relaunch(self, eid: int) -> bool:
        try:
            self._conns[eid].close()
        except OSError:
            pass
        old = self._procs[eid]
        if old is not None and old.is_alive():
            old.kill()
        self._launch(eid)
        return True

    def kill(self, eid: int) -> None:
        self._procs[eid].kill()

    def stop(self) -> None:
        for e, c in enumerate(self._conns):
            try:
                with self._send_locks[e]:
                    c.send(("stop",))
            except (OSError, BrokenPipeError, AttributeError):
                pass
        for p in self._procs:
            if p is None:
                continue
            p.join(timeout=5)
            if p.is_alive():
                p.terminate()
        for c in self._conns:
            try:
                c.close()
            except (OSError, AttributeError):
                pass


# =================================================================================================== driver side
class ClusterScheduler:
    _job_ids = itertools.count(1_000_000)

    def __init__(self, num_executors: int = 2, max_failures: int = 4, block_first_async_job: bool = True,
                 name: str = "afb-cluster", transport=None, limits=None, recover_lost_executors: bool = False,
                 use_env: bool = True, speculation: bool = False, speculation_multiplier: float = 1.5,
                 speculation_quantile: float = 0.75, speculation_interval: float = 0.1, speculation_min_ms: float = 100.0,
                 blacklist: bool = False, blacklist_max_failures: int = 2, blacklist_timeout: float = 3600.0,
                 cores_per_executor: int = 1, max_task_binary_bytes: int = 128 << 20):
        self.num_executors = int(num_executors)
        self.max_failures = int(max_failures)
        self.block_first_async_job = block_first_async_job
        self.recover_lost_executors = recover_lost_executors
        self.mode = 0
        self.first_iter = 1
        self.device = None
        self.listeners = ListenerBus()
        self.fault_injector: Optional[Callable[[TaskContext], None]] = None
        self.failed_jobs: List[int] = []
        self._jobs: Dict[int, dict] = {}
        self._lock = threading.Lock()
        self._stopped = False
        self.max_task_binary_bytes = int(max_task_binary_bytes)      # spark.rpc.message.maxSize
        self._blocks_held: Dict[int, set] = {}       # source block id -> executors that already hold its data
        self._cleaner_registered: set = set()
        self.last_task_binary_bytes = 0
        self.indirect_results = 0
        self.executors_recovered = 0
        # speculation (TaskSetManager.checkSpeculatableTasks, scheduler/TaskSetManager.scala:975) and executor blacklisting
        # (scheduler/BlacklistTracker.scala:50).  Both are off by default, as in the reference: asynchrony is this
        # framework's answer to stragglers; they exist for the synchronous drivers and for flaky hosts.
        self.speculation = speculation
        self.speculation_multiplier, self.speculation_quantile = speculation_multiplier, speculation_quantile
        self.speculation_interval, self.speculation_min_ms = speculation_interval, speculation_min_ms
        self.executor_last_active: Dict[int, float] = {}   # monotonic time of the last task launch / completion per executor slot
        self.allocation_manager = None       # ExecutorAllocationManager once enable_dynamic_allocation() ran
        self._submit_lock = threading.RLock()
        self.speculated_tasks = 0
        self.speculative_wins = 0
        self.speculative_kills_sent = 0      # kill_task messages to the losing copy (TaskSetManager.scala:736)
        self.tasks_killed = 0                # "killed" acknowledgements from executors
        self.blacklist_enabled = blacklist
        self.blacklist_max_failures, self.blacklist_timeout = blacklist_max_failures, blacklist_timeout
        self._dump_waiters: Dict[int, Any] = {}
        self._dump_ids = itertools.count(1)
        self.executor_heartbeats: Dict[int, dict] = {}     # executor -> last heartbeat payload (+ "t" = monotonic arrival time)
        self.executors_expired = 0
        self._exec_failures: Dict[int, int] = {}
        self._blacklisted: Dict[int, float] = {}          # executor -> expiry (monotonic)
        self._getter = ThreadPoolExecutor(4, thread_name_prefix=f"{name}-task-result-getter")
        self.env = None
        if use_env:
            from .env import AsyncEnv, RuntimeLimits
            self.env = AsyncEnv.create_driver(limits or RuntimeLimits(), name=f"{name}-driver")
        self.transport = transport if transport is not None else PipeTransport(self.num_executors, name, cores_per_executor)
        try:
            self.transport.start(self)
        except BaseException:
            self.stop()                    # do not leak the driver env / half-started executors
            raise
        if self.env is not None:
            threading.Thread(target=self._heartbeat_monitor, daemon=True, name=f"{name}-heartbeat-receiver").start()
        if self.speculation:
            threading.Thread(target=self._speculation_loop, daemon=True, name=f"{name}-speculation").start()

    # ------------------------------------------------------------------ mode
    def set_mode(self, mode: int) -> None:
        self.mode = int(mode)

    def stream_for(self, partition: int):
        return None

    def _make_block_dropper(self, block_id: int):
        def drop() -> None:
            if self._stopped:
                return
            with self._lock:
                holders = self._blocks_held.pop(block_id, set())
            self._cleaner_registered.discard(block_id)
            for e in holders:
                try:
                    self._send(e, ("drop_source", block_id))
                except Exception:  # noqa: BLE001
                    pass
        return drop

    # ------------------------------------------------------------------ placement
    def is_blacklisted(self, eid: int) -> bool:
        exp = self._blacklisted.get(eid)
        if exp is None:
            return False
        if time.monotonic() >= exp:                     # BlacklistTracker.applyBlacklistTimeout
            self._blacklisted.pop(eid, None)
            self._exec_failures.pop(eid, None)
            self.listeners.post({"event": "executor_unblacklisted", "executor": eid})
            return False
        return True

    def _route(self, partition: int, avoid: Optional[int] = None) -> int:
        """Home executor of a partition is ``p % n`` (hard affinity: that is where its cached blocks live).  A blacklisted
        (or explicitly avoided) home falls over to the next healthy executor, which recomputes from lineage."""
        n = self.num_executors
        home = partition % n
        for step in range(n):
            e = (home + step) % n
            if e != avoid and not self.is_blacklisted(e):
                return e
        return home                                     # everyone is blacklisted: Spark aborts; we keep the home executor

    def _note_failure(self, eid: int) -> None:
        if not self.blacklist_enabled:
            return
        self._exec_failures[eid] = self._exec_failures.get(eid, 0) + 1
        healthy = sum(1 for e in range(self.num_executors) if e != eid and not self.is_blacklisted(e))
        if self._exec_failures[eid] >= self.blacklist_max_failures and eid not in self._blacklisted and healthy > 0:
            self._blacklisted[eid] = time.monotonic() + self.blacklist_timeout
            self.listeners.post({"event": "executor_blacklisted", "executor": eid, "failures": self._exec_failures[eid]})

    def _launch(self, job_id: int, job: dict, partition: int, attempt: int, eid: int, speculative: bool = False) -> None:
        """Send one task attempt to ``eid``, shipping the job's binary there first if that executor has not seen it."""
        if eid not in job["executors"]:
            self._send(eid, ("binary", job_id, job["make_payload"]([eid])))
            job["executors"].append(eid)
            with self._lock:
                for b in job["make_payload"].shipped:     # type: ignore[attr-defined]
                    self._blocks_held.setdefault(b, set()).add(eid)
        if not speculative:
            job["assigned"][partition] = eid
            job.setdefault("attempts", {})[partition] = attempt
            job["started"].pop(partition, None)       # set again by the executor's "started" message
        self.executor_last_active[eid] = time.monotonic()
        self._send(eid, ("task", job_id, partition, attempt, job["submit_ms"]))

    # ------------------------------------------------------------------ liveness
    def _heartbeat_monitor(self) -> None:
        """``HeartbeatReceiver.expireDeadHosts`` (core/HeartbeatReceiver.scala:193-220): an executor that has heartbeated at
        least once and then stays silent for ``network_timeout`` is wedged (stopped process, dead host behind a live TCP
        connection): kill it, which takes the ordinary lost-executor path (replacement + resubmission, or job failure)."""
        timeout = self.env.limits.network_timeout
        period = max(0.05, min(timeout / 4, 5.0))
        while not self._stopped:
            time.sleep(period)
            now = time.monotonic()
            for eid, hb in list(self.executor_heartbeats.items()):
                if now - hb["t"] > timeout:
                    self.executor_heartbeats.pop(eid, None)
                    self.executors_expired += 1
                    self.listeners.post({"event": "executor_expired", "executor": eid, "silent_s": now - hb["t"]})
                    try:
                        self.transport.kill(eid)
                    except Exception:  # noqa: BLE001
                        threading.Thread(target=self._executor_lost, args=(eid,), daemon=True).start()

    # ------------------------------------------------------------------ speculation
    def _speculation_loop(self) -> None:
        while not self._stopped:
            time.sleep(self.speculation_interval)
            with self._lock:
                jobs = list(self._jobs.items())
            for job_id, job in jobs:
                try:
                    self._check_speculatable(job_id, job)
                except Exception:  # noqa: BLE001
                    traceback.print_exc()

    def _check_speculatable(self, job_id: int, job: dict) -> None:
        done = job["durations"]
        total = len(done) + len(job["pending"])
        if total < 2 or len(done) < max(1, int(self.speculation_quantile * total)) or not job["pending"]:
            return
        med = sorted(done)[len(done) // 2]
        threshold = max(self.speculation_multiplier * med, self.speculation_min_ms)
        now = time.perf_counter()
        for p in sorted(job["pending"]):
            t0 = job["started"].get(p)
            if t0 is None or p in job["speculated"] or (now - t0) * 1e3 <= threshold:
                continue
            target = self._route(p, avoid=job["assigned"].get(p))
            if target == job["assigned"].get(p):
                continue
            job["speculated"][p] = target
            self.speculated_tasks += 1
            self.listeners.post({"event": "task_speculated", "job": job_id, "partition": p, "executor": target,
                                 "running_ms": (now - t0) * 1e3, "median_ms": med})
            self._launch(job_id, job, p, 1000, target, speculative=True)       # attempt >= 1000 marks a speculative copy

    # ------------------------------------------------------------------ plumbing
    def _send(self, eid: int, msg) -> None:
        self.transport.send(eid, msg)

    def _on_message(self, eid: int, msg: tuple) -> None:
        kind, job_id, partition, attempt, data, run_ms = msg
        if kind in ("ok", "indirect", "err", "started", "killed"):
            self.executor_last_active[eid] = time.monotonic()
        with self._lock:
            job = self._jobs.get(job_id)
        if kind == "heartbeat":
            self.executor_heartbeats[eid] = {**data, "t": time.monotonic()}
            return
        if kind == "threads":
            fut = self._dump_waiters.pop(attempt, None)
            if fut is not None:
                fut.set_result(data)
            return
        if kind == "started":
            if job is not None and attempt < 1000:
                job["started"][partition] = time.perf_counter()
            return
        if kind == "killed":                              # the losing copy of a speculated task was interrupted
            self.tasks_killed += 1
            self.listeners.post({"event": "task_killed", "job": job_id, "partition": partition, "attempt": attempt,
                                 "executor": eid, "reason": data, "ran_ms": run_ms})
            return
        if job is None or kind == "gone":
            if kind == "indirect" and self.env is not None:       # nobody will read it: free the executor's block
                self._getter.submit(self._drop_result_block, data[0])
            return
        if kind == "ok":
            self._task_succeeded(eid, job_id, job, partition, data, run_ms, len(data))
        elif kind == "indirect":
            self._getter.submit(self._fetch_indirect, eid, job_id, job, partition, data, run_ms)
        else:
            if attempt >= 1000:                       # a failed speculative copy is simply dropped; the original still runs
                job["speculated"].pop(partition, None)
                return
            if partition not in job["pending"]:
                return
            if "BlockMissing" in data and job.get("reshipped", 0) < 2 * self.num_executors:
                # the executor was sent a block REFERENCE for data it does not hold (a replacement that raced with the
                # bookkeeping of its predecessor's loss): forget what we think it holds and re-ship the binary with the data
                job["reshipped"] = job.get("reshipped", 0) + 1
                with self._lock:
                    for held in self._blocks_held.values():
                        held.discard(eid)
                if eid in job["executors"]:
                    job["executors"].remove(eid)
                self._launch(job_id, job, partition, attempt, eid)
                return
            attempt += 1
            self.listeners.post({"event": "task_failed", "job": job_id, "partition": partition, "attempt": attempt,
                                 "executor": eid, "error": data})
            self._note_failure(eid)
            if attempt >= self.max_failures:
                self.failed_jobs.append(job_id)
                exc = RuntimeError(f"task {partition} of job {job_id} failed {attempt} times: {data.splitlines()[0]}")
                self._fail_task(job, partition, exc)
            else:                                      # same executor unless it just got blacklisted
                self._launch(job_id, job, partition, attempt, self._route(partition))

    def _fail_task(self, job: dict, partition: int, exc: BaseException) -> None:
        if job["on_task_lost"] is not None:
            job["on_task_lost"](partition, exc)
        job["waiter"].job_failed(exc)
        job["pending"].discard(partition)
        if not job.get("failed_posted"):
            job["failed_posted"] = True
            self.listeners.post({"event": "job_end", "job": job["waiter"].job_id, "failed": True, "error": str(exc)[:300],
                                 "t_ms": time.perf_counter() * 1e3})
        if not job["pending"]:                        # nothing of this job is in flight any more: forget it (and its binary)
            self._finish_job(job["waiter"].job_id)

    def _task_succeeded(self, eid: int, job_id: int, job: dict, partition: int, data: bytes, run_ms, nbytes: int,
                        extra: Optional[dict] = None) -> None:
        waiter: JobWaiter = job["waiter"]
        with self._lock:
            if partition not in job["pending"] or partition in job["claimed"]:
                return                                 # the other copy of a speculated / resubmitted task already delivered
            job["claimed"].add(partition)
        twin = job["speculated"].get(partition)
        if twin is not None:                              # two copies ran: stop the one that lost (TaskSetManager.scala:736)
            if twin == eid:
                self.speculative_wins += 1
                loser, loser_attempt = job["assigned"].get(partition), job.get("attempts", {}).get(partition, 0)
            else:
                loser, loser_attempt = twin, 1000
            if loser is not None and loser != eid:
                try:
                    self._send(loser, ("kill_task", job_id, partition, loser_attempt))
                    self.speculative_kills_sent += 1
                except Exception:  # noqa: BLE001  (the executor is gone: nothing left to kill)
                    pass
        metrics = dict(run_ms) if isinstance(run_ms, dict) else {"run_ms": float(run_ms)}
        metrics.update(extra or {})
        metrics["result_bytes"] = nbytes
        run_ms = metrics.pop("run_ms")
        limit = self.env.limits.max_result_size if self.env is not None else 0
        with self._lock:
            job["result_bytes"] += nbytes
            too_big = bool(limit) and job["result_bytes"] > limit
        if too_big:
            exc = RuntimeError(f"Total size of serialized results of job {job_id} ({job['result_bytes']} bytes) is bigger "
                               f"than max_result_size ({limit} bytes)")
            self.failed_jobs.append(job_id)
            self._fail_task(job, partition, exc)
            return
        job["durations"].append(float(metrics.get("run_ms", 0.0)))
        t0 = time.perf_counter()
        try:
            decoded = pickle.loads(data)
        except Exception as exc:  # noqa: BLE001
            self._fail_task(job, partition, RuntimeError(f"cannot deserialize the result of task {partition}: {exc!r}"))
            return
        metrics["deserialize_ms"] = (time.perf_counter() - t0) * 1e3           # TaskMetrics.resultDeserializationTime (driver side)
        self.listeners.post({"event": "task_end", "job": job_id, "partition": partition, "executor": eid,
                             "run_ms": run_ms, "metrics": metrics})
        try:
            result, acc_updates = decoded
            if not waiter.cancelled:
                _acc.merge_updates(acc_updates)
            waiter.task_succeeded(partition, result)
        except Exception as exc:  # a raising result handler fails the job, not the receiver thread
            waiter.job_failed(exc)
        job["pending"].discard(partition)
        if waiter.done:
            self._finish_job(job_id)

    def _fetch_indirect(self, eid: int, job_id: int, job: dict, partition: int, ref, run_ms: float) -> None:
        """``TaskResultGetter.enqueueSuccessfulTask`` for an IndirectTaskResult: pull the block, then delete it."""
        from ..storage.block_manager import BlockManagerId
        block_id, size, holder = ref
        limit = self.env.limits.max_result_size
        if limit and size > limit:
            self._drop_result_block(block_id)
            self.failed_jobs.append(job_id)
            self._fail_task(job, partition, RuntimeError(
                f"result of task {partition} of job {job_id} ({size} bytes) is bigger than max_result_size ({limit} bytes)"))
            return
        t0 = time.perf_counter()
        data = self.env.block_manager.get_remote_bytes(block_id, [BlockManagerId(*holder)])
        fetch_ms = (time.perf_counter() - t0) * 1e3
        self._drop_result_block(block_id)
        if data is None:                               # the executor died between the status message and the fetch
            self._fail_task(job, partition, RuntimeError(f"lost the result block {block_id} of job {job_id}"))
            return
        self.indirect_results += 1
        self._task_succeeded(eid, job_id, job, partition, data, run_ms, size, {"fetch_wait_ms": fetch_ms, "indirect": 1})

    def _drop_result_block(self, block_id: str) -> None:
        try:
            self.env.bm_master.remove_block(block_id)
        except Exception:  # noqa: BLE001
            pass

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
            except Exception:  # noqa: BLE001  (a dead executor has nothing to drop)
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

    def cancel_job_group(self, group_id: str) -> int:
        with self._lock:
            ids = [j for j, job in self._jobs.items() if getattr(job["waiter"], "group", None) == group_id]
        return sum(1 for j in ids if self.cancel_job(j, f"cancelled as part of job group {group_id}"))

    def active_job_ids(self, group_id: Optional[str] = None) -> List[int]:
        with self._lock:
            return sorted(j for j, job in self._jobs.items()
                          if not job["waiter"].done and (group_id is None or getattr(job["waiter"], "group", None) == group_id))

    def _executor_lost(self, eid: int) -> None:
        if self._stopped:
            return
        self.executor_heartbeats.pop(eid, None)        # its replacement starts with a clean liveness record
        self.listeners.post({"event": "executor_lost", "executor": eid, "t_ms": time.perf_counter() * 1e3})
        with self._lock:
            jobs = list(self._jobs.items())
            for held in self._blocks_held.values():       # whatever it cached is gone: re-ship on next use
                held.discard(eid)
        if self.env is not None:
            try:
                self.env.bm_master.remove_executor(str(eid))
            except Exception:  # noqa: BLE001
                pass
        if self.recover_lost_executors and self._recover(eid, jobs):
            return
        for job_id, job in jobs:
            lost = [p for p in list(job["pending"]) if job["assigned"].get(p, p % self.num_executors) == eid]
            for p in lost:
                self._fail_task(job, p, RuntimeError(f"executor {eid} lost while running partition {p} of job {job_id}"))

    def _recover(self, eid: int, jobs) -> bool:
        """Re-fill the slot and re-launch what the dead executor was running (DAGScheduler.handleExecutorLost ->
        resubmit).  The replacement starts with an empty store, so task binaries are re-sent WITH their source
        blocks."""
        try:
            if not self.transport.relaunch(eid):
                return False
        except Exception:  # noqa: BLE001
            traceback.print_exc()
            return False
        self.executors_recovered += 1
        self.listeners.post({"event": "executor_added", "executor": eid, "t_ms": time.perf_counter() * 1e3})
        for job_id, job in jobs:
            lost = sorted(p for p in job["pending"] if job["assigned"].get(p, p % self.num_executors) == eid)
            if eid in job["executors"]:
                job["executors"].remove(eid)           # the replacement has no binary yet: _launch re-ships it
            if not lost:
                continue
            job["lost_count"] = job.get("lost_count", 0) + 1
            if job["lost_count"] >= self.max_failures:
                for p in lost:
                    self._fail_task(job, p, RuntimeError(f"executor {eid} lost {job['lost_count']} times during job {job_id}"))
                continue
            try:
                for p in lost:
                    self.listeners.post({"event": "task_resubmitted", "job": job_id, "partition": p, "executor": eid})
                    self._launch(job_id, job, p, 0, eid)
            except Exception as exc:  # noqa: BLE001
                for p in lost:
                    self._fail_task(job, p, exc)
        return True

    def _finish_job(self, job_id: int) -> None:
        with self._lock:
            job = self._jobs.pop(job_id, None)
        if job is None:
            return
        for e in job["executors"]:
            try:
                self._send(e, ("drop", job_id))
            except Exception:  # noqa: BLE001
                pass
        if not job.get("failed_posted"):              # a failed job already announced its end
            self.listeners.post({"event": "job_end", "job": job_id, "t_ms": time.perf_counter() * 1e3})

    # ------------------------------------------------------------------ jobs
    def submit_job(self, dataset, func, partitions: Optional[Sequence[int]] = None,
                   result_handler: Optional[Callable[[int, Any], None]] = None,
                   on_task_lost: Optional[Callable[[int, BaseException], None]] = None) -> JobWaiter:
        if self._stopped:
            raise RuntimeError("scheduler stopped")
        if getattr(self, "profiler", None) is not None:
            func = self.profiler.wrap(func)               # spark.python.profile: every task under cProfile
        parts = list(range(dataset.num_partitions)) if partitions is None else list(partitions)
        am = self.allocation_manager
        if am is not None:                                # dynamic allocation: a pending executor target is applied while no
            with self._submit_lock:                       # job is routed (partition p lives on slot p % num_executors)
                am.before_job()
                return self._submit_job_routed(dataset, func, parts, result_handler, on_task_lost)
        return self._submit_job_routed(dataset, func, parts, result_handler, on_task_lost)

    def _submit_job_routed(self, dataset, func, parts, result_handler, on_task_lost) -> JobWaiter:
        job_id = next(ClusterScheduler._job_ids)
        waiter = JobWaiter(job_id, len(parts), result_handler)
        from .scheduler import _job_group
        waiter.group = getattr(_job_group, "id", None)   # type: ignore[attr-defined]
        now = time.perf_counter() * 1e3
        assigned = {p: self._route(p) for p in parts}
        executors = sorted(set(assigned.values()))
        # source data travels once per executor: blocks every target executor already holds are pickled as references
        from ..data.dataset import ShippingContext
        fault = self.fault_injector

        def make_payload(targets: List[int]) -> bytes:
            with self._lock:
                held = {b: set(es) for b, es in self._blocks_held.items()}
            with ShippingContext(lambda b: all(e in held.get(b, ()) for e in targets)) as ship:
                try:
                    payload = cloudpickle.dumps((dataset, func, fault))
                except Exception as exc:  # noqa: BLE001  ("Task not serializable": name the culprit, not just its type)
                    from .closure import TaskNotSerializable, explain_pickle_failure
                    why = explain_pickle_failure({"function": func, "dataset lineage": dataset, "fault injector": fault}, "task binary")
                    raise TaskNotSerializable(f"Task not serializable: {why or exc!r}") from exc
            make_payload.shipped = ship.shipped        # type: ignore[attr-defined]
            if not ship.shipped:                       # no source data travels with this binary: what is in it was CAPTURED
                if len(payload) > self.max_task_binary_bytes:
                    raise RuntimeError(
                        f"Serialized task binary of job {job_id} is {len(payload)} bytes, which exceeds max allowed: "
                        f"spark.rpc.message.maxSize ({self.max_task_binary_bytes} bytes).  Consider using broadcast variables for "
                        f"large values captured by the closure.")
                if len(payload) > 1000 * 1024 and not getattr(self, "_warned_big_task", False):
                    self._warned_big_task = True
                    import logging
                    logging.getLogger("asyncframework_b200.scheduler").warning(
                        "Job %d contains a task binary of very large size (%d KB). The maximum recommended task size is 1000 KB: "
                        "a closure is probably capturing data that should be a broadcast.", job_id, len(payload) // 1024)
            for bid, blk in ship.blocks.items():       # ContextCleaner: free the executors' copies when the driver's block dies
                if bid not in self._cleaner_registered:
                    self._cleaner_registered.add(bid)
                    from .cleaner import ContextCleaner
                    ContextCleaner.get().register(blk, "source_block", self._make_block_dropper(bid))
            return payload

        payload = make_payload(executors)
        self.last_task_binary_bytes = len(payload)
        with self._lock:
            self._jobs[job_id] = {"waiter": waiter, "pending": set(parts), "on_task_lost": on_task_lost,
                                  "executors": executors, "submit_ms": now, "result_bytes": 0,
                                  "make_payload": make_payload, "assigned": assigned, "started": {}, "durations": [],
                                  "speculated": {}, "claimed": set()}
        self.listeners.post({"event": "job_start", "job": job_id, "partitions": parts, "t_ms": now,
                             "task_binary_bytes": len(payload)})
        job = self._jobs[job_id]
        unreachable: Dict[int, BaseException] = {}
        for e in executors:
            try:
                self._send(e, ("binary", job_id, payload))
            except Exception as exc:  # noqa: BLE001  (dead pipe / slot down): its tasks fail, the job is not left dangling
                unreachable[e] = exc
        reached = [e for e in executors if e not in unreachable]
        with self._lock:                              # only now may a later job refer to these blocks by id (message order)
            for b in make_payload.shipped:            # type: ignore[attr-defined]
                self._blocks_held.setdefault(b, set()).update(reached)
        for p in parts:
            e = assigned[p]
            if e not in unreachable:
                self.listeners.post({"event": "task_start", "job": job_id, "partition": p, "attempt": 0, "executor": e, "t_ms": now})
                self.executor_last_active[e] = time.monotonic()
                try:
                    self._send(e, ("task", job_id, p, 0, now))
                    continue
                except Exception as exc:  # noqa: BLE001
                    unreachable[e] = exc
            self._fail_task(job, p, RuntimeError(f"executor {e} is unreachable, cannot launch partition {p} of job {job_id}: {unreachable[e]}"))
        if not parts or (unreachable and not job["pending"]):
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

    def request_total_executors(self, total: int) -> bool:
        """``sc.requestTotalExecutors``: supported by transports with a cluster manager behind them (standalone)."""
        if not hasattr(self.transport, "request_total_executors"):
            raise NotImplementedError(f"{type(self.transport).__name__} has a fixed set of executors")
        return self.transport.request_total_executors(int(total))

    def enable_dynamic_allocation(self, policy=None, start: bool = True):
        """``spark.dynamicAllocation.enabled``: start the backlog-driven policy loop (core/allocation.py)."""
        from .allocation import ExecutorAllocationManager
        if not hasattr(self.transport, "request_total_executors"):
            raise NotImplementedError(f"{type(self.transport).__name__} has a fixed set of executors")
        if self.allocation_manager is None:
            self.allocation_manager = ExecutorAllocationManager(self, policy)
            if start:
                self.allocation_manager.start()
        return self.allocation_manager

    def kill_executor(self, eid: int) -> None:
        """Fault injection: kill one executor process (FailureSuite / DistributedSuite style)."""
        self.transport.kill(eid)

    def thread_dump(self, eid, timeout: float = 10.0) -> List[dict]:
        """Thread stacks of the driver (``eid="driver"``) or of an executor process."""
        if eid == "driver":
            return thread_dump()
        from concurrent.futures import Future
        token, fut = next(self._dump_ids), Future()
        self._dump_waiters[token] = fut
        self._send(int(eid), ("thread_dump", token))
        try:
            return fut.result(timeout)
        finally:
            self._dump_waiters.pop(token, None)

    def storage_status(self) -> dict:
        """Per-process block-store usage as the driver's BlockManagerMaster sees it (``sc.getExecutorStorageStatus``)."""
        return {} if self.env is None else self.env.bm_master.storage_status()

    def stop(self) -> None:
        if self._stopped:
            return
        self._stopped = True
        if self.allocation_manager is not None:
            self.allocation_manager.stop()
        self.transport.stop()
        self._getter.shutdown(wait=False)
        if self.env is not None:
            self.env.stop()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.stop()
