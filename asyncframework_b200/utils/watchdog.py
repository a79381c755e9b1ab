"""Failure detection: heartbeat watchdog over the worker-status table.

Spark: executors heartbeat every 10 s and are expired after ``spark.network.timeout``
(core/HeartbeatReceiver.scala:75-87,123-134,193-220); ASYNC itself has no recovery — a lost worker
stays ``available=false`` forever.  Here every worker stamps a heartbeat (host table:
``StatTable.heartbeat``; fused backend: ``WorkerSlot.heartbeat_ns`` written by the round kernel) and
this watchdog marks silent workers dead.  Async modes continue with P−1 workers (dead workers are
excluded from ``available_workers`` and from barrier selection); sync modes abort.
"""
from __future__ import annotations

import threading
import time
from typing import Callable, List, Optional


class HeartbeatWatchdog:
    def __init__(self, stat, timeout_s: float = 30.0, period_s: float = 0.5,
                 on_dead: Optional[Callable[[int], None]] = None, only_busy: bool = True):
        self.stat = stat
        self.timeout_ms = timeout_s * 1e3
        self.period_s = period_s
        self.on_dead = on_dead
        self.only_busy = only_busy
        self.dead: List[int] = []
        self._stop = threading.Event()
        self._thread: Optional[threading.Thread] = None

    def check_once(self) -> List[int]:
        newly = []
        for p in list(self.stat):
            ws = self.stat[p]
            if not ws.alive:
                continue
            busy = not ws.available
            if (busy or not self.only_busy) and ws.heartbeat_age_ms > self.timeout_ms:
                self.stat.set_alive(p, False)
                self.stat.set_available(p, False)
                self.dead.append(p)
                newly.append(p)
                if self.on_dead:
                    self.on_dead(p)
        return newly

    def _loop(self):
        while not self._stop.wait(self.period_s):
            self.check_once()

    def start(self) -> "HeartbeatWatchdog":
        self._thread = threading.Thread(target=self._loop, name="afb-watchdog", daemon=True)
        self._thread.start()
        return self

    def stop(self) -> None:
        self._stop.set()
        if self._thread:
            self._thread.join(timeout=2)


class DeviceWatchdog:
    """Same policy for the fused backend: polls the device-side status table of the server block
    (heartbeat_ns is a globaltimer stamp; compared against the newest stamp seen) — and ACTS on what it finds:

    * asynchronous modes: ``engine.declare_dead(p)`` marks the worker dead on the server (``alive = 0``), removes it from the
      idle count / enable mask and lowers the bucket threshold to ``floor(P_alive * beta)``: the run continues with P-1
      workers and still reaches its target k (HeartbeatReceiver.scala:193-220 is the Spark side);
    * synchronous modes: ``engine.request_stop_local()`` — the in-kernel BSP flag waits observe the stop word and the round
      kernels exit with the abort flag set, so the driver raises instead of hanging on the dead peer's flag forever.

    ``expect_running=True`` (a run is in progress) also treats an IDLE worker with a stale heartbeat as dead: with persistent
    round kernels a live worker is busy almost all the time, so "idle and silent while the others advance" means its process
    is gone."""

    def __init__(self, engine, timeout_s: float = 30.0, act: bool = True):
        self.engine = engine
        self.timeout_ns = int(timeout_s * 1e9)
        self.dead: List[int] = []
        self.act = act

    def start(self, period_s: float = 0.2, expect_running: bool = True) -> "DeviceWatchdog":
        """Poll from a daemon thread (on the engine's side stream: the launching stream may be occupied by a persistent round
        kernel that is itself waiting for the dead worker)."""
        self._stop = threading.Event()

        def loop():
            from ..parallel import _devapi as _dev
            _dev.set_device(self.engine.device)
            while not self._stop.wait(period_s):
                try:
                    self.check_once(expect_running=expect_running)
                except Exception as exc:  # noqa: BLE001 — the watchdog must never take the run down
                    self.error = exc
        self._thread = threading.Thread(target=loop, name="afb-device-watchdog", daemon=True)
        self._thread.start()
        return self

    def stop(self) -> None:
        if getattr(self, "_stop", None) is not None:
            self._stop.set()
            self._thread.join(timeout=2)

    def check_once(self, expect_running: bool = False) -> List[int]:
        ctx = self.engine.side_stream_ctx() if hasattr(self.engine, "side_stream_ctx") else None
        if ctx is None:
            return self._check(expect_running)
        with ctx:
            out = self._check(expect_running)
            if hasattr(self.engine, "_side_stream") and getattr(self.engine, "_side_stream", None) is not None:
                self.engine._side_stream.synchronize()
            return out

    def _check(self, expect_running: bool = False) -> List[int]:
        st = self.engine.worker_status()
        newest = int(st["heartbeat_ns"].max())
        if newest and not getattr(self, "_base_ns", 0):
            self._base_ns = newest                  # device time of the first sign of life seen by this watchdog
        newly = []
        for p in range(len(st)):
            busy = int(st[p]["available"]) == 0
            hb = int(st[p]["heartbeat_ns"])
            if not hb and expect_running:
                hb = getattr(self, "_base_ns", 0)   # a worker that NEVER reported (its process died before its first round)
            if (busy or expect_running) and hb and newest - hb > self.timeout_ns and p not in self.dead:
                self.dead.append(p)
                newly.append(p)
        if self.act and newly:
            if getattr(self.engine, "is_async", True):
                for p in newly:
                    self.engine.declare_dead(p)
            else:
                self.engine.request_stop_local()
        return newly
