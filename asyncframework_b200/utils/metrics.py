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


def _prom_name(name: str) -> str:
    import re
    n = re.sub(r"[^a-zA-Z0-9_:]", "_", name)
    return n if not n[0].isdigit() else "_" + n


def prometheus_text(snap: Dict[str, float], prefix: str = "afb_") -> str:
    """Prometheus text exposition format 0.0.4 of a snapshot (PrometheusServlet's role; every value is a gauge)."""
    lines = []
    for k in sorted(snap):
        n = prefix + _prom_name(k)
        lines.append(f"# TYPE {n} gauge")
        lines.append(f"{n} {float(snap[k]):.10g}")
    return "\n".join(lines) + "\n"


class PrometheusFileSink(Sink):
    """Writes the exposition text atomically to a file (node-exporter textfile collector style)."""

    def __init__(self, path: str, prefix: str = "afb_"):
        self.path, self.prefix = path, prefix

    def write(self, ts, snap):
        tmp = self.path + ".tmp"
        with open(tmp, "w") as f:
            f.write(prometheus_text(snap, self.prefix))
        os.replace(tmp, self.path)


class StatsdSink(Sink):
    """``StatsdSink`` (core/metrics/sink/StatsdSink.scala): one UDP datagram per batch of ``name:value|g`` lines."""

    def __init__(self, host: str = "127.0.0.1", port: int = 8125, prefix: str = "afb", max_datagram: int = 1400):
        import socket
        self.addr, self.prefix, self.max_datagram = (host, port), prefix, max_datagram
        self.sock = socket.socket(socket.AF_INET, socket.SOCK_DGRAM)

    def write(self, ts, snap):
        batch, size = [], 0
        for k in sorted(snap):
            line = f"{self.prefix}.{k}:{float(snap[k]):.10g}|g"
            if size + len(line) + 1 > self.max_datagram and batch:
                self.sock.sendto("\n".join(batch).encode(), self.addr)
                batch, size = [], 0
            batch.append(line)
            size += len(line) + 1
        if batch:
            self.sock.sendto("\n".join(batch).encode(), self.addr)


class GraphiteSink(Sink):
    """``GraphiteSink`` (core/metrics/sink/GraphiteSink.scala): plaintext protocol ``<path> <value> <unix ts>\\n`` over TCP;
    reconnects on every report so a restarted carbon daemon does not wedge the reporter."""

    def __init__(self, host: str, port: int = 2003, prefix: str = "afb", timeout: float = 2.0):
        self.addr, self.prefix, self.timeout = (host, port), prefix, timeout
        self.errors = 0

    def write(self, ts, snap):
        import socket
        payload = "".join(f"{self.prefix}.{k} {float(snap[k]):.10g} {int(ts)}\n" for k in sorted(snap)).encode()
        try:
            with socket.create_connection(self.addr, timeout=self.timeout) as s:
                s.sendall(payload)
        except OSError:
            self.errors += 1                # a metrics sink must never take the job down


class Slf4jSink(Sink):
    """``Slf4jSink``: the snapshot goes through the framework logger at INFO."""

    def __init__(self, logger_name: str = "asyncframework_b200.metrics"):
        import logging
        self.log = logging.getLogger(logger_name)

    def write(self, ts, snap):
        for k in sorted(snap):
            self.log.info("type=GAUGE, name=%s, value=%.10g", k, snap[k])


_SINK_CLASSES = {"console": ConsoleSink, "csv": CsvSink, "jsonl": JsonlSink, "prometheus": PrometheusFileSink,
                 "statsd": StatsdSink, "graphite": GraphiteSink, "slf4j": Slf4jSink}


class MetricsSystem:
    """Registry + sinks + a periodic reporter thread, configurable from ``metrics.properties``-style keys
    (core/metrics/MetricsSystem.scala:97-190, MetricsConfig): ``*.sink.<name>.class`` picks the sink,
    ``*.sink.<name>.<option>`` are its constructor arguments, ``*.sink.<name>.period`` (seconds) the report period —
    one reporter thread at the smallest period drives all sinks."""

    def __init__(self, registry: Optional[MetricsRegistry] = None):
        self.registry = registry or MetricsRegistry()
        self.period = 10.0
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None

    @classmethod
    def from_properties(cls, props: Dict[str, str], registry: Optional[MetricsRegistry] = None) -> "MetricsSystem":
        ms = cls(registry)
        sinks: Dict[str, Dict[str, str]] = {}
        for key, val in props.items():
            parts = key.split(".")
            if len(parts) >= 4 and parts[1] == "sink":          # <instance>.sink.<name>.<option>
                sinks.setdefault(parts[2], {})[".".join(parts[3:])] = val
        periods = []
        for name, opts in sinks.items():
            kind = opts.pop("class", name).rsplit(".", 1)[-1].lower().replace("sink", "")
            if kind not in _SINK_CLASSES:
                raise ValueError(f"unknown metrics sink {kind!r} (have: {sorted(_SINK_CLASSES)})")
            if "period" in opts:
                periods.append(float(opts.pop("period")))
            opts.pop("unit", None)
            kwargs = {k: (int(v) if k in ("port", "max_datagram") else v) for k, v in opts.items()}
            ms.registry.add_sink(_SINK_CLASSES[kind](**kwargs))
        if periods:
            ms.period = min(periods)
        return ms

    @staticmethod
    def load_properties(path: str) -> Dict[str, str]:
        out = {}
        with open(path) as f:
            for line in f:
                line = line.strip()
                if line and not line.startswith("#") and "=" in line:
                    k, v = line.split("=", 1)
                    out[k.strip()] = v.strip()
        return out

    def start(self, period: Optional[float] = None) -> "MetricsSystem":
        if period is not None:
            self.period = period
        if self._thread is None:
            self._stop.clear()
            self._thread = threading.Thread(target=self._loop, daemon=True, name="metrics-reporter")
            self._thread.start()
        return self

    def _loop(self) -> None:
        while not self._stop.wait(self.period):
            try:
                self.registry.report()
            except Exception:  # noqa: BLE001
                pass

    def stop(self, final_report: bool = True) -> None:
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)
            self._thread = None
        if final_report:
            self.registry.report()


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
            m = ev.get("metrics") or {}
            for k in ("result_bytes",):
                if k in m:
                    self.r.inc(self.prefix + "." + k, float(m[k]))
            for k in ("deserialize_ms", "serialize_ms", "fetch_wait_ms"):
                if k in m:
                    self.r.observe(self.prefix + ".task_" + k, float(m[k]))
        elif kind == "task_failed":
            self.r.inc(self.prefix + ".task_failures")
        elif kind == "executor_lost":
            self.r.inc(self.prefix + ".executors_lost")
        elif kind == "task_resubmitted":
            self.r.inc(self.prefix + ".tasks_resubmitted")
