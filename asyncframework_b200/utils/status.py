"""Live application status store.

Reference: ``AppStatusListener`` (core/status/AppStatusListener.scala:43) folds the ``SparkListener*`` events
(core/scheduler/SparkListener.scala:42-168) into an ``AppStatusStore`` (kvstore-backed) that the web UI
(core/ui/SparkUI.scala:39) and the REST API (``/api/v1/applications/...``, core/status/api/v1) read;
``TaskMetrics`` (core/executor/TaskMetrics.scala) are aggregated per stage / executor.

Here the events are the scheduler's listener-bus dicts (``job_start`` / ``task_start`` / ``task_end`` /
``task_failed`` / ``task_resubmitted`` / ``job_end`` / ``executor_lost`` / ``executor_added``, emitted identically by
the in-process, ``local-cluster`` and standalone schedulers) plus the ASYNC-specific view: per-worker staleness,
average task time and availability straight from the :class:`AsyncContext`'s worker table — the live version of
the "Average waiting time per worker" / staleness telemetry the reference only prints at the end of a run
(ASYNCsamples/.../SparkASGDThread.scala:344-359).
"""
from __future__ import annotations

import threading
import time
from collections import OrderedDict
from typing import Any, Dict, List, Optional


def _quantiles(vals: List[float]) -> Dict[str, float]:
    if not vals:
        return {}
    s = sorted(vals)
    pick = lambda q: s[min(len(s) - 1, int(q * len(s)))]  # noqa: E731
    return {"min": s[0], "p25": pick(0.25), "p50": pick(0.5), "p75": pick(0.75), "p99": pick(0.99), "max": s[-1],
            "mean": sum(s) / len(s)}


class AppStatusStore:
    def __init__(self, app_name: str = "app", app_id: Optional[str] = None, max_jobs: int = 1000):
        self.app_name = app_name
        self.app_id = app_id or f"local-{int(time.time() * 1000)}"
        self.start_time = time.time()
        self.end_time: Optional[float] = None
        self.max_jobs = max_jobs
        self.jobs: "OrderedDict[int, dict]" = OrderedDict()
        self.executors: Dict[Any, dict] = {}
        self.environment: Dict[str, Any] = {}
        self._lock = threading.Lock()
        self._contexts: List[Any] = []
        self._scheduler = None
        self.killed_tasks = 0
        self.allocation_events: List[dict] = []

    # ------------------------------------------------------------------ wiring
    def attach(self, scheduler, environment: Optional[dict] = None) -> "AppStatusStore":
        self._scheduler = scheduler
        scheduler.listeners.add(self)
        with self._lock:
            for e in range(getattr(scheduler, "num_executors", 0)):
                self._executor(e)
            if environment:
                self.environment.update(environment)
        return self

    def watch_context(self, ctx) -> None:
        """Expose an AsyncContext's worker table (staleness / availability / task time) under ``/async``."""
        self._contexts.append(ctx)

    def _executor(self, eid) -> dict:
        ex = self.executors.get(eid)
        if ex is None:
            ex = self.executors[eid] = {"id": eid, "active": True, "active_tasks": 0, "completed_tasks": 0, "failed_tasks": 0,
                                        "total_run_ms": 0.0, "result_bytes": 0, "added": time.time(), "removed": None,
                                        "restarts": 0}
        return ex

    # ------------------------------------------------------------------ listener
    def __call__(self, ev: dict) -> None:
        kind = ev.get("event")
        now = time.time()
        with self._lock:
            if kind == "job_start":
                self.jobs[ev["job"]] = {"job_id": ev["job"], "status": "RUNNING", "submission_time": now, "completion_time": None,
                                        "num_tasks": len(ev.get("partitions", [])), "num_active_tasks": 0,
                                        "num_completed_tasks": 0, "num_failed_tasks": 0, "num_skipped_tasks": 0,
                                        "num_resubmitted_tasks": 0, "task_binary_bytes": ev.get("task_binary_bytes"),
                                        "run_ms": [], "result_bytes": 0, "tasks": {}}
                while len(self.jobs) > self.max_jobs:
                    self.jobs.popitem(last=False)
                return
            if kind == "executors_resized":                  # dynamic allocation applied a new executor target
                self.allocation_events.append({"time": now, "from": ev.get("from"), "to": ev.get("to"), "reason": ev.get("reason")})
                del self.allocation_events[:-200]
                return
            if kind == "task_killed":                        # the losing copy of a speculated task was interrupted
                self.killed_tasks += 1
                j = self.jobs.get(ev.get("job"))
                if j is not None:
                    j["num_killed_tasks"] = j.get("num_killed_tasks", 0) + 1
                return
            if kind == "executor_lost":
                ex = self._executor(ev["executor"])
                ex["active"], ex["removed"], ex["active_tasks"] = False, now, 0
                return
            if kind == "executor_added":
                ex = self._executor(ev["executor"])
                ex["active"], ex["removed"] = True, None
                ex["restarts"] += 1
                return
            job = self.jobs.get(ev.get("job"))
            if job is None:
                return
            p = ev.get("partition")
            if kind == "task_start":
                job["num_active_tasks"] += 1
                job["tasks"][p] = {"partition": p, "attempt": ev.get("attempt", 0), "executor": ev.get("executor"),
                                   "status": "RUNNING", "launch_time": now}
                self._executor(ev.get("executor"))["active_tasks"] += 1
            elif kind == "task_end":
                job["num_active_tasks"] = max(0, job["num_active_tasks"] - 1)
                job["num_completed_tasks"] += 1
                run_ms = float(ev.get("run_ms", 0.0))
                job["run_ms"].append(run_ms)
                m = ev.get("metrics") or {}
                job["result_bytes"] += int(m.get("result_bytes", 0) or 0)
                # shuffle sorters that spilled (TaskMetrics.memoryBytesSpilled / diskBytesSpilled, stage page "Spill" columns)
                job["memory_bytes_spilled"] = job.get("memory_bytes_spilled", 0) + int(m.get("memory_bytes_spilled", 0) or 0)
                job["disk_bytes_spilled"] = job.get("disk_bytes_spilled", 0) + int(m.get("disk_bytes_spilled", 0) or 0)
                t = job["tasks"].setdefault(p, {"partition": p})
                t.update(status="SUCCESS", run_ms=run_ms, executor=ev.get("executor"), metrics=dict(m), finish_time=now)
                ex = self._executor(ev.get("executor"))
                ex["active_tasks"] = max(0, ex["active_tasks"] - 1)
                ex["completed_tasks"] += 1
                ex["total_run_ms"] += run_ms
                ex["result_bytes"] += int(m.get("result_bytes", 0) or 0)
            elif kind == "task_failed":
                job["num_failed_tasks"] += 1
                t = job["tasks"].setdefault(p, {"partition": p})
                t.update(status="FAILED", attempt=ev.get("attempt"), error=str(ev.get("error", ""))[:500])
                ex = self.executors.get(t.get("executor"))
                if ex is not None:
                    ex["failed_tasks"] += 1
            elif kind == "task_skipped":
                job["num_skipped_tasks"] += 1
                job["num_active_tasks"] = max(0, job["num_active_tasks"] - 1)
            elif kind == "task_resubmitted":
                job["num_resubmitted_tasks"] += 1
            elif kind == "job_end":
                job["completion_time"] = now
                if ev.get("cancelled"):
                    job["status"] = "CANCELLED"
                elif ev.get("failed"):
                    job["status"] = "FAILED"
                    job["error"] = ev.get("error")
                elif job["num_completed_tasks"] >= job["num_tasks"]:
                    job["status"] = "SUCCEEDED"
                else:
                    job["status"] = "FAILED"

    # ------------------------------------------------------------------ views (REST payloads)
    def application_info(self) -> dict:
        return {"id": self.app_id, "name": self.app_name,
                "attempts": [{"start_time": self.start_time, "end_time": self.end_time,
                              "duration_ms": ((self.end_time or time.time()) - self.start_time) * 1e3,
                              "completed": self.end_time is not None}]}

    def job_list(self, status: Optional[str] = None) -> List[dict]:
        with self._lock:
            out = [self._job_view(j) for j in self.jobs.values() if status is None or j["status"] == status.upper()]
        return out[::-1]

    def job(self, job_id: int) -> Optional[dict]:
        with self._lock:
            j = self.jobs.get(job_id)
            return None if j is None else self._job_view(j, tasks=True)

    @staticmethod
    def _job_view(j: dict, tasks: bool = False) -> dict:
        v = {k: j[k] for k in ("job_id", "status", "submission_time", "completion_time", "num_tasks", "num_active_tasks",
                               "num_completed_tasks", "num_failed_tasks", "num_skipped_tasks", "num_resubmitted_tasks",
                               "task_binary_bytes", "result_bytes")}
        if j.get("error"):
            v["error"] = j["error"]
        for k in ("memory_bytes_spilled", "disk_bytes_spilled", "num_killed_tasks"):
            v[k] = j.get(k, 0)
        v["task_run_ms"] = _quantiles(j["run_ms"])
        if j["completion_time"] is not None:
            v["duration_ms"] = (j["completion_time"] - j["submission_time"]) * 1e3
        if tasks:
            v["tasks"] = [dict(t) for _, t in sorted(j["tasks"].items())]
        return v

    def executor_list(self) -> List[dict]:
        storage = {}
        sched = self._scheduler
        if sched is not None and hasattr(sched, "storage_status"):
            try:
                for bm_id, st in sched.storage_status().items():
                    storage[str(bm_id.executor_id)] = {"max_memory": st["max_memory"], "memory_used": st["mem_used"],
                                                       "disk_used": st["disk_used"], "rdd_blocks": sum(1 for b in st["blocks"] if b.startswith("rdd_")),
                                                       "host_port": f"{bm_id.host}:{bm_id.port}"}
            except Exception:  # noqa: BLE001
                pass
        with self._lock:
            out = []
            for eid, ex in sorted(self.executors.items(), key=lambda kv: str(kv[0])):
                v = dict(ex)
                v.update(storage.get(str(eid), {}))
                out.append(v)
        if "driver" in storage:
            out.insert(0, {"id": "driver", "active": True, **storage["driver"]})
        return out

    def rdd_storage(self) -> List[dict]:
        sched = self._scheduler
        if sched is None or not hasattr(sched, "storage_status"):
            return []
        rdds: Dict[str, dict] = {}
        for bm_id, st in sched.storage_status().items():
            for b, (mem, disk) in st["blocks"].items():
                if not b.startswith("rdd_"):
                    continue
                rid = b.split("_")[1]
                r = rdds.setdefault(rid, {"id": int(rid), "num_cached_partitions": 0, "memory_used": 0, "disk_used": 0, "executors": set()})
                r["num_cached_partitions"] += 1
                r["memory_used"] += mem
                r["disk_used"] += disk
                r["executors"].add(str(bm_id.executor_id))
        return [{**r, "executors": sorted(r["executors"])} for r in rdds.values()]

    def async_workers(self) -> List[dict]:
        """One row per (context, worker): staleness, average task time, availability, tasks — A2's table, live."""
        out = []
        for ci, ctx in enumerate(self._contexts):
            try:
                for w, row in sorted(ctx.stat.snapshot().items()):
                    out.append({"context": ci, "worker": w, "clock": ctx.current_time, "queued_results": ctx.size(), **row})
            except Exception:  # noqa: BLE001
                continue
        return out

    def summary(self) -> dict:
        with self._lock:
            by = {}
            for j in self.jobs.values():
                by[j["status"]] = by.get(j["status"], 0) + 1
            tasks = sum(j["num_completed_tasks"] for j in self.jobs.values())
            fails = sum(j["num_failed_tasks"] for j in self.jobs.values())
            spilled = sum(j.get("disk_bytes_spilled", 0) for j in self.jobs.values())
        return {"application": self.application_info(), "jobs_by_status": by, "completed_tasks": tasks, "failed_task_attempts": fails,
                "killed_task_attempts": self.killed_tasks, "disk_bytes_spilled": spilled,
                "executor_allocation_events": list(self.allocation_events[-20:]),
                "executors": len(self.executors), "active_executors": sum(1 for e in self.executors.values() if e["active"])}

    def stop(self) -> None:
        self.end_time = time.time()
