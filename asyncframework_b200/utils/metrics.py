"""Metrics registry with pluggable sinks (Dropwizard MetricsSystem with Console / CSV /
Slf4j / JMX / Graphite sinks, core/metrics/MetricsSystem.scala:97-190; sources DAGSchedulerSource,
ExecutorSource...).  Gauges, counters and timers are registered by name; ``report()`` pushes a
snapshot to every sink.  ``SchedulerSource`` wires the scheduler's listener bus in."""
from __future__ import annotations

import csv
import json
import os
import threading
import time
from typing import Callable, Dict, List, Optional


class MetricsRegistry:
    def __init__(self):
        self._lock = threading.Lock()
        self.counters: Dict[str, float] = {}
        self.gauges: Dict[str, Callable[[], float]] = {}
        self.timers: Dict[str, List[float]] = {}
        self.sinks: List["Sink"] = []

    def inc(self, name: str, by: float = 1.0) -> None:
        with self._lock:
            self.counters[name] = self.counters.get(name, 0.0) + by

    def gauge(self, name: str, fn: Callable[[], float]) -> None:
        self.gauges[name] = fn

    def observe(self, name: str, value: float) -> None:
        with self._lock:
            self.timers.setdefault(name, []).append(float(value))

    def snapshot(self) -> Dict[str, float]:
        with self._lock:
            out = dict(self.counters)
            for k, vals in self.timers.items():
                if vals:
                    s = sorted(vals)
                    out[k + ".count"] = len(s)
                    out[k + ".mean"] = sum(s) / len(s)
                    out[k + ".p50"] = s[len(s) // 2]
                    out[k + ".p99"] = s[min(len(s) - 1, int(len(s) * 0.99))]
                    out[k + ".max"] = s[-1]
        for k, fn in self.gauges.items():
            try:
                out[k] = float(fn())
            except Exception:  # noqa: BLE001
                pass
        return out

    def add_sink(self, sink: "Sink") -> "MetricsRegistry":
        self.sinks.append(sink)
        return self

    def report(self) -> Dict[str, float]:
        snap = self.snapshot()
        for s in self.sinks:
            s.write(time.time(), snap)
        return snap


class Sink:
    def write(self, ts: float, snap: Dict[str, float]) -> None:
        raise NotImplementedError


class ConsoleSink(Sink):
    def __init__(self, out=print):
        self.out = out

    def write(self, ts, snap):
        self.out("-- metrics %s --" % time.strftime("%H:%M:%S", time.localtime(ts)))
        for k in sorted(snap):
            self.out(f"  {k} = {snap[k]:.6g}")


class CsvSink(Sink):
    """One ``<name>.csv`` per metric with ``t,value`` rows (like Spark's CsvSink)."""

    def __init__(self, directory: str):
        self.dir = directory
        os.makedirs(directory, exist_ok=True)

    def write(self, ts, snap):
        for k, v in snap.items():
            path = os.path.join(self.dir, k.replace("/", "_") + ".csv")
            new = not os.path.exists(path)
            with open(path, "a", newline="") as f:
                w = csv.writer(f)
                if new:
                    w.writerow(["t", "value"])
                w.writerow([f"{ts:.3f}", v])


class JsonlSink(Sink):
    def __init__(self, path: str):
        self.path = path

    def write(self, ts, snap):
        with open(self.path, "a") as f:
            f.write(json.dumps({"t": ts, **snap}) + "\n")


class SchedulerSource:
    """DAGSchedulerSource/ExecutorSource analogue: counts jobs/tasks/failures and task run time."""

    def __init__(self, registry: MetricsRegistry, scheduler, prefix: str = "scheduler"):
        self.r, self.prefix = registry, prefix
        scheduler.listeners.add(self)

    def __call__(self, ev: dict) -> None:
        kind = ev.get("event")
        if kind == "job_start":
            self.r.inc(self.prefix + ".jobs")
        elif kind == "task_end":
            self.r.inc(self.prefix + ".tasks")
            self.r.observe(self.prefix + ".task_run_ms", ev.get("run_ms", 0.0))
        elif kind == "task_failed":
            self.r.inc(self.prefix + ".task_failures")
