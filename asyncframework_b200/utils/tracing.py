"""Tracing / profiling hooks.

Reference: hand-rolled wall-clock telemetry + Spark listener events / TaskMetrics.  Here:
  * ``TraceRecorder`` — subscribes to the scheduler's listener bus and writes a Chrome-trace JSON
    (one lane per executor, one slice per task) plus aggregate TaskMetrics-style counters;
  * ``nvtx_range`` — NVTX ranges around host-side phases when CUDA is present;
  * ``cuda_timer`` — CUDA-event timing on the current stream (the device-timed metric of BASELINE.md);
  * device-side timestamps: every fused round stamps ``globaltimer`` at pull and at apply into the
    result ring (ResultRec.t_start_ns / t_end_ns), so per-round device time needs no host sync.
"""
from __future__ import annotations

import contextlib
import json
import threading
import time
from typing import Dict, List, Optional

import torch


class TraceRecorder:
    def __init__(self):
        self.events: List[dict] = []
        self._open: Dict[tuple, float] = {}
        self._lock = threading.Lock()
        self.counters = {"tasks": 0, "failed": 0, "run_ms": 0.0, "jobs": 0}

    def __call__(self, ev: dict) -> None:
        kind = ev.get("event")
        with self._lock:
            if kind == "task_start":
                self._open[(ev["job"], ev["partition"])] = ev.get("t_ms", time.perf_counter() * 1e3)
            elif kind == "task_end":
                t0 = self._open.pop((ev["job"], ev["partition"]), None)
                self.counters["tasks"] += 1
                self.counters["run_ms"] += ev.get("run_ms", 0.0)
                if t0 is not None:
                    self.events.append({"name": f"job{ev['job']}/p{ev['partition']}", "ph": "X", "pid": 0,
                                        "tid": ev.get("executor", 0), "ts": t0 * 1e3, "dur": ev.get("run_ms", 0.0) * 1e3})
            elif kind == "task_failed":
                self.counters["failed"] += 1
            elif kind == "job_start":
                self.counters["jobs"] += 1

    def attach(self, scheduler) -> "TraceRecorder":
        scheduler.listeners.add(self)
        return self

    def dump(self, path: str) -> None:
        with open(path, "w") as f:
            json.dump({"traceEvents": self.events, "counters": self.counters}, f)


@contextlib.contextmanager
def nvtx_range(name: str):
    use = torch.cuda.is_available()
    if use:
        torch.cuda.nvtx.range_push(name)
    try:
        yield
    finally:
        if use:
            torch.cuda.nvtx.range_pop()


class cuda_timer:
    """``with cuda_timer() as t: ...; t.ms`` — CUDA events on the current stream, synchronised on exit."""

    def __enter__(self):
        self.e0 = torch.cuda.Event(enable_timing=True)
        self.e1 = torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        self.e0.record()
        return self

    def __exit__(self, *a):
        self.e1.record()
        torch.cuda.synchronize()
        self.ms = self.e0.elapsed_time(self.e1)
        return False
