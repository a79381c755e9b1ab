"""Accumulators: task-side counters merged into a driver-side value (``AccumulatorV2``, core/util/AccumulatorV2.scala:43;
``sc.longAccumulator`` / ``doubleAccumulator``, SparkContext.scala:1338-1375).

Semantics kept from Spark: tasks only ``add``; the driver reads ``value``; updates of a task are applied exactly once, when the
task SUCCEEDS (a failed attempt's partial updates are dropped, a retried task does not double-count).  Works on both schedulers:
in-process executors merge under the accumulator's lock; executor processes return the task's pending updates with its result
(``core/cluster.py``) and the driver merges them.
"""
from __future__ import annotations

import itertools
import threading
from typing import Any, Callable, Dict, Optional

_ids = itertools.count(1)
_registry: Dict[int, "Accumulator"] = {}        # driver-side: id -> accumulator
_task_local = threading.local()                  # executor-side: updates of the task running on this thread


def begin_task() -> None:
    _task_local.pending = {}


def end_task(success: bool) -> Dict[int, Any]:
    """Returns the task's updates (and forgets them).  The caller merges them only for a successful task."""
    pending = getattr(_task_local, "pending", None) or {}
    _task_local.pending = None
    return pending if success else {}


def merge_updates(updates: Dict[int, Any]) -> None:
    for acc_id, delta in updates.items():
        acc = _registry.get(acc_id)
        if acc is not None:
            acc._merge(delta)


class Accumulator:
    def __init__(self, zero: Any = 0, name: Optional[str] = None, op: Optional[Callable[[Any, Any], Any]] = None):
        self.acc_id = next(_ids)
        self.name = name
        self.zero = zero
        self._op = op or (lambda a, b: a + b)
        self._value = zero
        self._lock = threading.Lock()
        self._on_driver = True
        _registry[self.acc_id] = self

    # ---- task side ---------------------------------------------------------------------------------------------
    def add(self, v: Any) -> None:
        pending = getattr(_task_local, "pending", None)
        if pending is None:                       # outside a task (driver code): apply directly
            self._merge(v)
            return
        pending[self.acc_id] = self._op(pending[self.acc_id], v) if self.acc_id in pending else self._op(self.zero, v)

    __iadd__ = lambda self, v: (self.add(v), self)[1]   # noqa: E731  (acc += 1 inside closures)

    # ---- driver side -------------------------------------------------------------------------------------------
    def _merge(self, delta: Any) -> None:
        with self._lock:
            self._value = self._op(self._value, delta)

    @property
    def value(self) -> Any:
        if not self._on_driver:
            raise RuntimeError("Accumulator.value can only be read on the driver (tasks may only add)")
        with self._lock:
            return self._value

    def reset(self) -> None:
        with self._lock:
            self._value = self.zero

    # ---- shipping to executor processes: a write-only proxy with the same id ----------------------------------------
    def __reduce__(self):
        return (_proxy, (type(self), self.acc_id, self.name, self.zero, self._op))

    def __repr__(self) -> str:
        return f"{type(self).__name__}(id={self.acc_id}, name={self.name!r}, value={self.value if self._on_driver else '<task side>'!r})"


def _pair_add(a, b):
    return (a[0] + b[0], a[1] + b[1])


class NumericAccumulator(Accumulator):
    """``LongAccumulator`` / ``DoubleAccumulator`` (AccumulatorV2.scala:290-420): besides the running ``sum`` (= ``value``) the
    driver can read how many values were added (``count``) and their mean (``avg``)."""

    def __init__(self, zero: Any = 0, name: Optional[str] = None):
        super().__init__((zero, 0), name, _pair_add)

    def add(self, v: Any) -> None:
        super().add((v, 1))

    __iadd__ = lambda self, v: (self.add(v), self)[1]   # noqa: E731

    @property
    def value(self) -> Any:
        return Accumulator.value.fget(self)[0]

    @property
    def sum(self) -> Any:
        return self.value

    @property
    def count(self) -> int:
        return Accumulator.value.fget(self)[1]

    @property
    def avg(self) -> float:
        s, n = Accumulator.value.fget(self)
        return float("nan") if n == 0 else s / n


def _list_add(a, b):
    return a + (b if isinstance(b, list) else [b])


class CollectionAccumulator(Accumulator):
    """``sc.collectionAccumulator`` (AccumulatorV2.scala:446-480): tasks append elements, the driver reads the list (order across
    tasks is the order in which their results were merged)."""

    def __init__(self, name: Optional[str] = None):
        super().__init__([], name, _list_add)

    def add(self, v: Any) -> None:
        super().add([v])

    @property
    def value(self) -> list:
        return list(Accumulator.value.fget(self))


def _proxy(cls, acc_id: int, name, zero, op) -> Accumulator:
    existing = _registry.get(acc_id)
    if existing is not None:                      # unpickled in the driver process itself (in-process scheduler, tests)
        return existing
    a = cls.__new__(cls)
    a.acc_id, a.name, a.zero, a._op = acc_id, name, zero, op
    a._value, a._lock, a._on_driver = zero, threading.Lock(), False
    return a
