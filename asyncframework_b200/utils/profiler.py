"""Task profiler (``spark.python.profile``; python/pyspark/profiler.py:26-146 ``BasicProfiler`` + ``ProfilerCollector``).

Every task of a profiled scheduler runs under ``cProfile``; the task's stats travel back to the driver through an
accumulator (exactly-once for successful attempts, like PySpark's ``PStatsParam`` accumulator) keyed by job, and
``show_profiles()`` / ``dump_profiles(dir)`` print or write them (``pstats`` files, readable with ``python -m pstats``).
Works identically on the in-process, ``local-cluster`` and standalone masters.
"""
from __future__ import annotations

import cProfile
import io
import marshal
import os
import pstats
from typing import Dict

import threading

from ..core.accumulator import Accumulator

_PROFILE_LOCK = threading.Lock()


def _merge_stats(a: dict, b: dict) -> dict:
    """Sum two ``pstats``-style tables ``{func: (cc, nc, tt, ct, callers)}`` (``pstats.Stats.add``)."""
    if not a:
        return dict(b)
    out = dict(a)
    for func, (cc, nc, tt, ct, callers) in b.items():
        if func in out:
            occ, onc, ott, oct_, ocallers = out[func]
            merged = dict(ocallers)
            for c, v in callers.items():
                merged[c] = tuple(x + y for x, y in zip(merged[c], v)) if c in merged and isinstance(v, tuple) else (merged.get(c, 0) + v if not isinstance(v, tuple) else v)
            out[func] = (occ + cc, onc + nc, ott + tt, oct_ + ct, merged)
        else:
            out[func] = (cc, nc, tt, ct, dict(callers))
    return out


def _merge_by_job(a: Dict[int, dict], b: Dict[int, dict]) -> Dict[int, dict]:
    out = dict(a)
    for job, st in b.items():
        out[job] = _merge_stats(out.get(job, {}), st)
    return out


def _run_profiled(func, tctx, it, acc):
    # CPython allows ONE active cProfile per interpreter (sys.monitoring tool id): executor THREADS of an in-process master
    # take turns while profiling is on; executor processes have one task thread each and never wait here
    with _PROFILE_LOCK:
        pr = cProfile.Profile()
        try:
            return pr.runcall(func, tctx, it)
        finally:
            pr.create_stats()
            acc.add({tctx.job_id: pr.stats})


class _Table:
    """Adapter so ``pstats.Stats`` can load a raw table."""

    def __init__(self, stats: dict):
        self.stats = stats

    def create_stats(self) -> None:
        pass


class TaskProfiler:
    def __init__(self):
        self.acc = Accumulator({}, "task-profiles", _merge_by_job)

    def wrap(self, func):
        acc = self.acc

        def profiled(tctx, it):
            return _run_profiled(func, tctx, it, acc)      # module-level helper: pickled by reference (it touches a lock)
        return profiled

    def attach(self, scheduler) -> "TaskProfiler":
        scheduler.profiler = self
        return self

    # ---- driver side ---------------------------------------------------------------------------------------------------
    def stats(self, job_id: int) -> pstats.Stats:
        return pstats.Stats(_Table(self.acc.value[job_id]))

    def jobs(self):
        return sorted(self.acc.value)

    def show_profiles(self, top: int = 10, out=None) -> str:
        buf = io.StringIO()
        for job in self.jobs():
            buf.write("=" * 60 + f"\nProfile of job {job}\n" + "=" * 60 + "\n")
            st = pstats.Stats(_Table(self.acc.value[job]), stream=buf)
            st.sort_stats("cumulative").print_stats(top)
        text = buf.getvalue()
        (out or print)(text)
        return text

    def dump_profiles(self, directory: str) -> None:
        os.makedirs(directory, exist_ok=True)
        for job in self.jobs():
            with open(os.path.join(directory, f"job_{job}.pstats"), "wb") as f:
                marshal.dump(self.acc.value[job], f)
