"""Event log + replay (EventLoggingListener JSON event log replayable by the
History Server, core/scheduler/EventLoggingListener.scala:48-52, deploy/history/FsHistoryProvider.scala:80).

``EventLog(path).attach(scheduler)`` appends every scheduler event (job_start / task_start / task_end /
task_failed / job_end) as one JSON object per line; ``replay(path)`` rebuilds the per-job summary the
History Server would show: task counts, failures, run-time statistics, wall time, executors used."""
from __future__ import annotations

import json
import threading
import time
from collections import defaultdict
from typing import Dict, List


class EventLog:
    def __init__(self, path: str):
        self.path = path
        self._f = open(path, "a")
        self._lock = threading.Lock()

    def __call__(self, ev: dict) -> None:
        rec = dict(ev)
        rec.setdefault("wall", time.time())
        with self._lock:
            self._f.write(json.dumps(rec, default=repr) + "\n")
            self._f.flush()

    def attach(self, scheduler) -> "EventLog":
        scheduler.listeners.add(self)
        return self

    def close(self) -> None:
        with self._lock:
            self._f.close()


def replay(path: str) -> Dict[int, dict]:
    """job id -> {partitions, tasks, failures, run_ms_sum, run_ms_max, executors, wall_ms, finished}."""
    jobs: Dict[int, dict] = defaultdict(lambda: {"partitions": 0, "tasks": 0, "failures": 0, "run_ms_sum": 0.0,
                                                 "run_ms_max": 0.0, "executors": set(), "t_start": None, "t_end": None,
                                                 "finished": False})
    with open(path) as f:
        for line in f:
            ev = json.loads(line)
            j = jobs[ev.get("job", -1)]
            kind = ev.get("event")
            if kind == "job_start":
                j["partitions"] = len(ev.get("partitions", []))
                j["t_start"] = ev.get("t_ms")
            elif kind == "task_end":
                j["tasks"] += 1
                j["run_ms_sum"] += ev.get("run_ms", 0.0)
                j["run_ms_max"] = max(j["run_ms_max"], ev.get("run_ms", 0.0))
                j["executors"].add(ev.get("executor"))
            elif kind == "task_failed":
                j["failures"] += 1
            elif kind == "job_end":
                j["t_end"] = ev.get("t_ms")
                j["finished"] = True
                if ev.get("failed"):
                    j["failed"] = True
                if ev.get("cancelled"):
                    j["cancelled"] = True
    out = {}
    for jid, j in jobs.items():
        j["executors"] = sorted(e for e in j["executors"] if e is not None)
        j["wall_ms"] = (j["t_end"] - j["t_start"]) if (j["t_end"] is not None and j["t_start"] is not None) else None
        out[jid] = j
    return out


def summarize(path: str) -> List[str]:
    lines = []
    for jid, j in sorted(replay(path).items()):
        lines.append(f"job {jid}: {j['tasks']}/{j['partitions']} tasks, {j['failures']} failed attempts, "
                     f"run {j['run_ms_sum']:.1f} ms total / {j['run_ms_max']:.1f} ms max, executors {j['executors']}, "
                     f"wall {j['wall_ms'] if j['wall_ms'] is None else round(j['wall_ms'], 1)} ms, finished={j['finished']}")
    return lines
