"""Event-ordered NumPy fp64 oracle of the reference algorithms.

The reference cannot run here (no JVM) and is racy anyway, so correctness is anchored on this
single-threaded transcription of Appendix A: an explicit event list
``("submit", [workers])`` / ``("complete", worker)`` is replayed deterministically and produces
``w``, ``k``, the staleness of every result, accepted/dropped flags and the availability table.
The CPU backend is tested against it under scripted schedules; CUDA backends are tested against
the CPU backend.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

from ..core.sampler import bernoulli_mask


def _coeff(loss: str, margin: np.ndarray, y: np.ndarray) -> np.ndarray:
    if loss in ("least_squares", "least_squares_half"):
        return margin - y
    if loss == "logistic":
        return 1.0 / (1.0 + np.exp(-margin)) - y
    if loss == "hinge":
        s = 2.0 * y - 1.0
        return np.where(1.0 - s * margin > 0, -s, 0.0)
    raise ValueError(loss)


@dataclass
class OracleResult:
    w: np.ndarray
    k: int
    clock: int
    staleness: List[int] = field(default_factory=list)
    accepted: List[bool] = field(default_factory=list)
    workers: List[int] = field(default_factory=list)
    available: Dict[int, bool] = field(default_factory=dict)
    alpha_bar: Optional[np.ndarray] = None
    hist: Optional[Dict[int, np.ndarray]] = None
    w_trace: List[np.ndarray] = field(default_factory=list)


class AsyncOracle:
    """Shared state + sampling/gradient helpers.  ``parts[p] = (X_p, y_p, gidx_p)``."""

    def __init__(self, parts: Sequence[Tuple[np.ndarray, np.ndarray, np.ndarray]], loss: str, batch_rate: float,
                 gamma: float, taw: int, seed: int):
        self.parts = [(np.asarray(X, np.float64), np.asarray(y, np.float64), np.asarray(g, np.int64)) for X, y, g in parts]
        self.P = len(parts)
        self.N = sum(p[0].shape[0] for p in self.parts)
        self.d = self.parts[0][0].shape[1]
        self.loss, self.b, self.gamma, self.taw, self.seed = loss, batch_rate, gamma, taw, seed

    def sample_rows(self, p: int, seed_eff: int) -> np.ndarray:
        return np.nonzero(bernoulli_mask(seed_eff, self.parts[p][2], self.b))[0]

    def grad(self, p: int, w: np.ndarray, rows: np.ndarray, hist: Optional[np.ndarray] = None):
        X, y, _ = self.parts[p]
        Xb, yb = X[rows], y[rows]
        r = _coeff(self.loss, Xb @ w, yb)
        eff = r if hist is None else r - hist[rows]
        return Xb.T @ eff, r


def run_asgd_oracle(orc: AsyncOracle, events: Sequence[Tuple[str, object]], iters: int) -> OracleResult:
    """A.1.  ``("submit", workers)`` snapshots w, stamps ts=clock, marks workers busy, fixes each
    worker's sample (seed+k+1 with k at submit).  ``("complete", p)`` merges p's oldest in-flight
    task: staleness = clock − ts; clock += 1; accept iff staleness <= taw."""
    w = np.zeros(orc.d)
    k = clock = 0
    inflight: Dict[int, List[Tuple[np.ndarray, int, int]]] = {p: [] for p in range(orc.P)}
    res = OracleResult(w, 0, 0, available={p: True for p in range(orc.P)})
    par_recs = orc.b * orc.N / orc.P
    for ev, arg in events:
        if k >= iters:
            break
        if ev == "submit":
            w_snap = w.copy()
            for p in arg:
                inflight[p].append((w_snap, clock, orc.seed + k + 1))
                res.available[p] = False
        elif ev == "complete":
            p = int(arg)
            w_snap, ts, seed_eff = inflight[p].pop(0)
            rows = orc.sample_rows(p, seed_eff)
            res.available[p] = len(inflight[p]) == 0
            if rows.size == 0:
                continue                                   # empty sample: availability only (RDD.scala:1160-1164)
            g, _ = orc.grad(p, w_snap, rows)
            staleness = clock - ts
            clock += 1
            ok = staleness <= orc.taw
            res.staleness.append(staleness); res.accepted.append(ok); res.workers.append(p)
            if ok:
                eta = -orc.gamma / math.sqrt(k // orc.P + 1)
                w = w + eta * g / par_recs
                k += 1
                res.w_trace.append(w.copy())
        else:
            raise ValueError(ev)
    res.w, res.k, res.clock = w, k, clock
    return res


def run_asaga_oracle(orc: AsyncOracle, events: Sequence[Tuple[str, object]], iters: int) -> OracleResult:
    """A.2 with the history sharded per worker; accept iff k − ts <= taw; discarded results leave
    the history untouched."""
    w = np.zeros(orc.d)
    abar = np.zeros(orc.d)
    hist = {p: np.zeros(orc.parts[p][0].shape[0]) for p in range(orc.P)}
    k = clock = 0
    inflight: Dict[int, List[Tuple[np.ndarray, int, int]]] = {p: [] for p in range(orc.P)}
    res = OracleResult(w, 0, 0, available={p: True for p in range(orc.P)})
    par_recs = orc.b * orc.N / orc.P
    for ev, arg in events:
        if k >= iters:
            break
        if ev == "submit":
            w_snap = w.copy()
            for p in arg:
                inflight[p].append((w_snap, clock, orc.seed + k + 1))
                res.available[p] = False
        else:
            p = int(arg)
            w_snap, ts, seed_eff = inflight[p].pop(0)
            rows = orc.sample_rows(p, seed_eff)
            res.available[p] = len(inflight[p]) == 0
            if rows.size == 0:
                continue
            # history is read when the task RUNS; in the scripted oracle a task runs at completion
            g, r = orc.grad(p, w_snap, rows, hist[p])
            lag = k - ts
            clock += 1
            ok = lag <= orc.taw
            res.staleness.append(lag); res.accepted.append(ok); res.workers.append(p)
            if ok:
                hist[p][rows] = r
                w = w - orc.gamma * (g / par_recs + abar)
                abar = abar + g / orc.N
                k += 1
                res.w_trace.append(w.copy())
    res.w, res.k, res.clock, res.alpha_bar, res.hist = w, k, clock, abar, hist
    return res


def run_sgd_sync_oracle(orc: AsyncOracle, iters: int) -> OracleResult:
    """A.3: w <- w − (γ/√(k+1)) Σ_p g_p / (bN), sample seed+k+1."""
    w = np.zeros(orc.d)
    res = OracleResult(w, 0, 0)
    for k in range(iters):
        G = np.zeros(orc.d)
        for p in range(orc.P):
            rows = orc.sample_rows(p, orc.seed + k + 1)
            if rows.size:
                G += orc.grad(p, w, rows)[0]
        w = w - (orc.gamma / math.sqrt(k + 1)) * G / (orc.b * orc.N)
        res.w_trace.append(w.copy())
    res.w, res.k, res.clock = w, iters, iters * orc.P
    return res


def run_saga_sync_oracle(orc: AsyncOracle, iters: int) -> OracleResult:
    """A.4: w <- w − γ G/(bN) − γ ᾱ ; ᾱ <- ᾱ + G/N."""
    w = np.zeros(orc.d)
    abar = np.zeros(orc.d)
    hist = {p: np.zeros(orc.parts[p][0].shape[0]) for p in range(orc.P)}
    res = OracleResult(w, 0, 0)
    for k in range(iters):
        G = np.zeros(orc.d)
        for p in range(orc.P):
            rows = orc.sample_rows(p, orc.seed + k + 1)
            if rows.size:
                g, r = orc.grad(p, w, rows, hist[p])
                hist[p][rows] = r
                G += g
        w = w - orc.gamma * (G / (orc.b * orc.N) + abar)
        abar = abar + G / orc.N
        res.w_trace.append(w.copy())
    res.w, res.k, res.alpha_bar, res.hist = w, iters, abar, hist
    return res
