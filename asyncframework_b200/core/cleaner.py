"""``ContextCleaner`` (core/ContextCleaner.scala:60): asynchronous cleanup of state that outlives its driver-side handle.

The reference registers RDDs, shuffles, broadcasts and accumulators with weak references; when the driver object is garbage
collected a cleaner thread removes the blocks on the executors.  Iterative optimisers need this more than most applications:
ASGD creates ONE BROADCAST PER ROUND, so without cleanup every block store of the cluster would accumulate thousands of dead
``broadcast_<id>_piece*`` blocks (spilling to disk once the memory store is full).

Here: ``weakref.finalize`` on the handle enqueues a cleanup task; one daemon thread per process executes them (finalizers may
fire on any thread, including inside the allocator — they must not do RPC themselves).  Registered today: torrent broadcasts
(pieces + cached values everywhere), shipped source blocks of datasets (executor-side copies), shuffle outputs.
"""
from __future__ import annotations

import queue
import threading
import weakref
from typing import Callable, Optional


class ContextCleaner:
    _instance: Optional["ContextCleaner"] = None
    _lock = threading.Lock()

    def __init__(self):
        self._q: "queue.Queue" = queue.Queue()
        self.cleaned = {"broadcast": 0, "source_block": 0, "shuffle": 0}
        self._thread = threading.Thread(target=self._loop, daemon=True, name="context-cleaner")
        self._thread.start()

    @classmethod
    def get(cls) -> "ContextCleaner":
        with cls._lock:
            if cls._instance is None:
                cls._instance = cls()
            return cls._instance

    def register(self, obj, kind: str, fn: Callable[[], None]) -> None:
        """Run ``fn`` (on the cleaner thread) after ``obj`` has been garbage collected.  ``fn`` must not reference ``obj``."""
        weakref.finalize(obj, self._q.put, (kind, fn))

    def _loop(self) -> None:
        while True:
            kind, fn = self._q.get()
            try:
                fn()
                self.cleaned[kind] = self.cleaned.get(kind, 0) + 1
            except Exception:  # noqa: BLE001  (the scheduler may already be gone: cleanup is best effort)
                pass
            finally:
                self._q.task_done()

    def flush(self, timeout: float = 5.0) -> None:
        """Wait until everything enqueued so far has been processed (tests)."""
        done = threading.Event()
        self._q.put(("flush", done.set))
        done.wait(timeout)
