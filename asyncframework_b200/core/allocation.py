"""Backlog-driven dynamic allocation.

Reference: ``ExecutorAllocationManager.scala`` — ``schedule()`` runs every 100 ms (:228), ``updateAndSyncNumExecutorsTarget``
(:292) lowers the target to what the load needs, or — once tasks have been backlogged for ``schedulerBacklogTimeout`` —
raises it by 1, 2, 4, ... executors every ``sustainedSchedulerBacklogTimeout`` until the backlog is covered; executors idle
for ``executorIdleTimeout`` are released, never below ``minExecutors``.

How it maps onto this runtime: a partition lives on slot ``p % num_executors`` (shard affinity is hard here: the executor
owns that shard's rows), so the executor set can only change BETWEEN jobs.  The manager therefore keeps a *target* that the
policy loop moves with the load, and applies it at the first moment no job is in flight: from its own thread when the
cluster is idle, or from ``ClusterScheduler.submit_job`` right before the next job is routed (iterative drivers submit one
job per round, so a target raised during round r is in force for round r+1).  The load signal is read straight from the
scheduler's job table (pending = not finished, running = reported "started"), not reconstructed from listener events.
"""
from __future__ import annotations

import math
import threading
import time
from dataclasses import dataclass
from typing import Callable, Dict, List, Optional


@dataclass
class AllocationPolicy:
    min_executors: int = 1
    max_executors: int = 1 << 30
    initial_executors: int = 0                 # 0 = whatever the application started with
    backlog_timeout: float = 1.0               # spark.dynamicAllocation.schedulerBacklogTimeout
    sustained_backlog_timeout: float = 1.0     # spark.dynamicAllocation.sustainedSchedulerBacklogTimeout
    idle_timeout: float = 60.0                 # spark.dynamicAllocation.executorIdleTimeout
    interval: float = 0.1                      # ExecutorAllocationManager.intervalMillis

    @staticmethod
    def from_conf(conf) -> "AllocationPolicy":
        g = conf.get
        return AllocationPolicy(g("spark.dynamicAllocation.minExecutors"), g("spark.dynamicAllocation.maxExecutors"),
                                g("spark.dynamicAllocation.initialExecutors"), g("spark.dynamicAllocation.schedulerBacklogTimeout"),
                                g("spark.dynamicAllocation.sustainedSchedulerBacklogTimeout"), g("spark.dynamicAllocation.executorIdleTimeout"))

    def validate(self) -> "AllocationPolicy":
        if self.min_executors < 1 or self.max_executors < self.min_executors:
            raise ValueError(f"dynamic allocation needs 1 <= minExecutors ({self.min_executors}) <= maxExecutors ({self.max_executors})")
        if self.initial_executors and not (self.min_executors <= self.initial_executors <= self.max_executors):
            raise ValueError("spark.dynamicAllocation.initialExecutors must lie between minExecutors and maxExecutors")
        return self


class ExecutorAllocationManager:
    """Moves ``scheduler.num_executors`` with the task backlog.  ``tick()`` is one pass of the policy (the thread started by
    :meth:`start` calls it every ``policy.interval``; tests call it by hand with an injected clock)."""

    def __init__(self, scheduler, policy: Optional[AllocationPolicy] = None, tasks_per_executor: Optional[int] = None,
                 clock: Callable[[], float] = time.monotonic):
        self.sched = scheduler
        self.policy = (policy or AllocationPolicy()).validate()
        self.tasks_per_executor = max(1, int(tasks_per_executor or getattr(scheduler, "cores_per_executor", 1) or 1))
        self.clock = clock
        self.target = max(self.policy.min_executors, min(self.policy.max_executors,
                                                          self.policy.initial_executors or scheduler.num_executors))
        self._to_add = 1                       # numExecutorsToAdd: doubles while the backlog is sustained
        self._add_time: Optional[float] = None
        self._idle_since: Dict[int, float] = {}
        self._lock = threading.RLock()
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None
        self.history: List[dict] = []          # every target change / application, for the UI and the tests
        self.applied = 0
        self.failed_requests = 0

    # ---- load signal ---------------------------------------------------------------------------------------------------
    def load(self) -> dict:
        """``{"pending", "running", "backlog", "busy": set of executors with unfinished tasks, "jobs"}`` right now."""
        s = self.sched
        pending = running = 0
        busy = set()
        with s._lock:
            jobs = list(s._jobs.values())
        for job in jobs:
            for _ in range(4):                       # receiver threads mutate these containers: retry a torn snapshot
                try:
                    pend = set(job["pending"])
                    started = set(job["started"])
                    assigned = dict(job["assigned"])
                    spec = dict(job["speculated"])
                    break
                except RuntimeError:
                    continue
            else:
                continue
            pending += len(pend)
            running += sum(1 for p in pend if p in started)
            for p in pend:
                e = assigned.get(p)
                if e is not None:
                    busy.add(e)
            for p, e in spec.items():
                if p in pend:
                    busy.add(e)
        return {"pending": pending, "running": running, "backlog": pending - running, "busy": busy, "jobs": len(jobs)}

    def max_needed(self, pending: int) -> int:
        return int(math.ceil(pending / self.tasks_per_executor))

    # ---- one policy pass -------------------------------------------------------------------------------------------------
    def tick(self) -> int:
        """Update the target from the current load and apply it if no job is in flight.  Returns the target."""
        with self._lock:
            now = self.clock()
            ld = self.load()
            pol = self.policy
            needed = max(pol.min_executors, min(pol.max_executors, self.max_needed(ld["pending"])))
            old = self.target
            if ld["backlog"] > 0:
                if self._add_time is None:
                    self._add_time = now + pol.backlog_timeout
                elif now >= self._add_time:
                    if needed > self.target:
                        self.target = min(needed, pol.max_executors, self.target + self._to_add)
                        # keep doubling only while the request was fully usable (ExecutorAllocationManager.scala:385-391)
                        self._to_add = self._to_add * 2 if self.target < needed else 1
                    else:
                        self._to_add = 1
                    self._add_time = now + pol.sustained_backlog_timeout
            else:
                self._add_time = None
                self._to_add = 1
            # idle executors: a slot without unfinished tasks since idle_timeout is released (highest slots first: slots
            # are dense, partition p lives on slot p % n)
            n = self.sched.num_executors
            # an executor is idle since the later of: the first tick that saw it without unfinished tasks, and its last task
            # launch / status message (short tasks start and finish between two ticks: polling alone would call a busy
            # executor idle; ExecutorAllocationManager resets the idle timer in onTaskStart for the same reason)
            last_active = getattr(self.sched, "executor_last_active", {})
            mono_to_clock = now - time.monotonic() if self.clock is time.monotonic else None
            for e in range(n):
                if e in ld["busy"]:
                    self._idle_since.pop(e, None)
                else:
                    since = self._idle_since.setdefault(e, now)
                    la = last_active.get(e)
                    if la is not None and mono_to_clock is not None and la + mono_to_clock > since:
                        self._idle_since[e] = la + mono_to_clock
            for e in list(self._idle_since):
                if e >= n:
                    self._idle_since.pop(e, None)
            if ld["backlog"] <= 0:
                expired = sum(1 for e, t0 in self._idle_since.items() if now - t0 >= pol.idle_timeout)
                if expired:
                    floor = max(pol.min_executors, self.max_needed(ld["pending"]))
                    self.target = max(floor, min(self.target, n - expired))
            if self.target != old:
                self.history.append({"t": now, "event": "target", "from": old, "to": self.target, **{k: ld[k] for k in ("pending", "running", "backlog")}})
            if ld["jobs"] == 0:
                self._apply(now)
            return self.target

    def _apply(self, now: float) -> bool:
        n = self.sched.num_executors
        if self.target == n:
            return True
        try:
            ok = bool(self.sched.request_total_executors(self.target))
        except (RuntimeError, NotImplementedError) as exc:      # a job slipped in / fixed executor set
            self.failed_requests += 1
            self.history.append({"t": now, "event": "request_failed", "target": self.target, "error": repr(exc)})
            return False
        if ok:
            self.applied += 1
            self.history.append({"t": now, "event": "applied", "from": n, "to": self.target})
            for e in list(self._idle_since):                    # fresh executors start their idle clock now
                self._idle_since[e] = now
            self.sched.listeners.post({"event": "executors_resized", "from": n, "to": self.target, "reason": "dynamic_allocation"})
        else:
            self.failed_requests += 1
            self.history.append({"t": now, "event": "request_failed", "target": self.target, "error": "cluster manager refused"})
        return ok

    def before_job(self) -> None:
        """Called by ``submit_job`` before routing: apply a pending target while the job table is still empty."""
        with self._lock:
            with self.sched._lock:
                idle = not self.sched._jobs
            if idle and self.target != self.sched.num_executors:
                self._apply(self.clock())

    # ---- thread ----------------------------------------------------------------------------------------------------------
    def start(self) -> "ExecutorAllocationManager":
        if self._thread is None:
            def loop():
                while not self._stop.wait(self.policy.interval):
                    if getattr(self.sched, "_stopped", False):
                        return
                    try:
                        self.tick()
                    except Exception:  # noqa: BLE001  (the policy must never take the driver down)
                        import traceback
                        traceback.print_exc()
            self._thread = threading.Thread(target=loop, name="afb-dynamic-allocation", daemon=True)
            self._thread.start()
        return self

    def stop(self) -> None:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=5)
            self._thread = None
