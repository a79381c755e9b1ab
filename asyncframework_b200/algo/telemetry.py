"""Wait-time & convergence telemetry shared by the drivers.

Reference behaviour: snapshots ``optVars += (elapsed ms, copy of w)`` every ``printerFreq`` applied
updates (SparkASGDThread.scala:192-195); per-worker ``WaitingTime`` = Σ(submit − last finish)
(:112-116, 328-335); final report "Elapsed time(ms)", per-worker waiting times, "Average waiting
time(ms) per worker and iteration" (:344-359) and one ``time,objective`` line per snapshot where
objective = Σ_i (x_i·w_t − y_i)² / N over ALL rows in one pass (:386-401).
Here the same numbers are also emitted as structured JSONL (``MetricsLog``).
"""
from __future__ import annotations

import json
import os
import threading
import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple

import torch


class MetricsLog:
    """Append-only JSONL sink (one object per line); no-op when path is None."""

    def __init__(self, path: Optional[str]):
        self.path = path
        self._lock = threading.Lock()
        self._f = open(path, "a") if path else None

    def emit(self, kind: str, **fields) -> None:
        if self._f is None:
            return
        rec = {"t": time.time(), "kind": kind}
        rec.update(fields)
        with self._lock:
            self._f.write(json.dumps(rec, default=float) + "\n")
            self._f.flush()

    def close(self) -> None:
        if self._f is not None:
            self._f.close()
            self._f = None


@dataclass
class RunReport:
    algorithm: str
    num_partitions: int
    iterations: int = 0
    elapsed_ms: float = 0.0
    snapshots: List[Tuple[float, torch.Tensor]] = field(default_factory=list)   # (elapsed ms, w copy)
    objective: List[Tuple[float, float]] = field(default_factory=list)          # (elapsed ms, objective)
    waiting_ms: Dict[int, float] = field(default_factory=dict)
    avg_waiting_ms: float = 0.0
    accepted: int = 0
    dropped: int = 0
    staleness_hist: Dict[int, int] = field(default_factory=dict)
    final_weights: Optional[torch.Tensor] = None
    extra: Dict[str, float] = field(default_factory=dict)

    def print_reference_style(self, out=print) -> None:
        """The reference's end-of-run report, line for line (SparkASGDThread.scala:346-401; the MLlib driver prints only the
        objective table and ``finished``, SparkSGDMLLIB.scala:75-111)."""
        if self.algorithm != "sgd_mllib":
            out(f"Elapsed time(ms): {self.elapsed_ms:.0f}")
            out("*********************************")
            out("Individual waiting times:")
            for p in sorted(self.waiting_ms):
                out(f"{p},{self.waiting_ms[p]:.0f}")
            out(f"Average waiting time(ms) per worker and iteration:{self.avg_waiting_ms}")
        out("*********************************")
        for t, obj in self.objective:
            out(f"{t:.0f},{obj}")
        if self.algorithm == "sgd_mllib":
            out("finished")

    def to_dict(self) -> dict:
        return {"algorithm": self.algorithm, "num_partitions": self.num_partitions, "iterations": self.iterations,
                "elapsed_ms": self.elapsed_ms, "objective": self.objective, "waiting_ms": self.waiting_ms,
                "avg_waiting_ms": self.avg_waiting_ms, "accepted": self.accepted, "dropped": self.dropped,
                "staleness_hist": self.staleness_hist, "extra": self.extra}


class WaitTimer:
    """Per-worker accumulated waiting time: submit time − last finish time."""

    def __init__(self):
        self.waiting: Dict[int, float] = {}
        self.finish: Dict[int, float] = {}
        self.submit: Dict[int, float] = {}
        self._lock = threading.Lock()

    def on_finish(self, worker: int, now_ms: float) -> None:
        with self._lock:
            self.finish[worker] = now_ms

    def on_submit(self, workers, now_ms: float) -> None:
        with self._lock:
            for p in workers:
                self.submit[p] = now_ms
                self.waiting[p] = self.waiting.get(p, 0.0) + (now_ms - self.finish.get(p, now_ms))

    def task_ms(self, worker: int, now_ms: float) -> float:
        return now_ms - self.submit.get(worker, now_ms)


def _eval_on_device(s, W, loss_id):
    """Module-level on purpose: the closure below is pickled BY VALUE for executor processes, and a ``from ..ops import``
    inside it makes cloudpickle chase ``torch.ops`` (registered in ``sys.modules``, not picklable)."""
    from ..ops.eval import eval_loss
    return eval_loss(s, W.to(s.device), loss_id).double().cpu()


def evaluate_objective(points, snapshots: List[Tuple[float, torch.Tensor]], loss: str = "least_squares") -> List[Tuple[float, float]]:
    """One pass over the data computing the objective of EVERY snapshot (K12: X[N×d]·W[d×T]).

    least_squares: Σ (x·w_t − y)² / N (the reference's report); logistic / hinge: mean loss.
    On HBM-resident shards this is the evaluation GEMM kernel (ops/eval.py)."""
    from ..optim.gradient import make_gradient
    if not snapshots:
        return []
    grad = make_gradient(loss)
    first = next(s for s in points.shards() if s is not None)
    W = torch.stack([w.to(first.dtype) for _, w in snapshots], dim=1)   # [d, T]

    def seq(acc, s):
        if s is None or s.num_rows == 0:
            return acc
        if s.device.type == "cuda":
            part = _eval_on_device(s, W, grad.loss_id)
        else:
            Wd = W.to(s.device)
            if s.is_dense:
                M = s.X[:, : s.num_features] @ Wd
            else:
                M = torch.stack([s.matvec(Wd[:, t]) for t in range(Wd.shape[1])], dim=1)
            _, L = grad.coefficients(M, s.y.to(M.dtype)[:, None].expand_as(M))
            part = L.sum(0).double().cpu()
        n = torch.tensor(float(s.num_rows), dtype=torch.float64)
        if acc is None:
            return [part, n]
        acc[0] += part
        acc[1] += n
        return acc

    def comb(a, b):
        if a is None:
            return b
        if b is None:
            return a
        return [a[0] + b[0], a[1] + b[1]]

    tot = points.map_batch(lambda s: s).aggregate(lambda: None, seq, comb)
    obj = (tot[0] / tot[1]).tolist()
    return [(t, float(o)) for (t, _), o in zip(snapshots, obj)]
