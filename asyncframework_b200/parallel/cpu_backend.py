"""Multi-process CPU backend over gloo: the ``local-cluster[n,c,m]`` analogue and the
home of BASELINE config 1 (SparkASGDSync logistic regression on 1k x 20 synthetic LIBSVM, 2 CPU
workers).  SPMD: every rank runs the same driver; rank 0 is parameter server + worker.

  * ``run_sync_gloo``  - BSP mini-batch SGD / SAGA: broadcast(w) -> local sampled gradient ->
                         reduce to rank 0 -> update (SparkASGDSync / SparkASAGASync semantics).
  * ``run_async_gloo`` - parameter-server ASGD / ASAGA: workers ``send`` their result to rank 0 and
                         ``recv`` a fresh ``w``; the server thread applies results in arrival order
                         with the staleness rule (SparkASGDThread / SparkASAGAThread semantics).
Launch with torchrun (``--master-addr 127.0.0.1``) or ``spawn_local``.
"""
from __future__ import annotations

import math
import os
import queue
import threading
import time
from typing import List, Optional

import numpy as np
import torch
import torch.distributed as dist

from ..config import AsyncConfig
from ..core.sampler import bernoulli_mask
from ..data.batch import Shard
from ..optim.gradient import make_gradient


def _init(backend: str = "gloo"):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend, rank=rank, world_size=world)
    return rank, world


def _load_local(cfg: AsyncConfig, rank: int, world: int):
    from ..core.scheduler import Scheduler
    from ..algo.common import load_points
    sched = Scheduler(1)
    try:
        c = AsyncConfig.from_dict(cfg.to_dict())
        c.num_partitions = world
        ds = load_points(c, sched, torch.device("cpu"), torch.float64).zip_with_index()
        sizes = ds.partition_sizes()
        return ds.batch(rank), sum(sizes)
    finally:
        sched.stop()


def _sampled(shard: Shard, seed_eff: int, b: float) -> torch.Tensor:
    return torch.from_numpy(np.nonzero(bernoulli_mask(seed_eff, shard.gidx, b))[0])


def run_sync_gloo(cfg: AsyncConfig):
    from ..algo.telemetry import RunReport
    rank, world = _init()
    shard, N = _load_local(cfg, rank, world)
    grad = make_gradient(cfg.loss)
    d = shard.num_features
    saga = cfg.algorithm == "saga_sync"
    w = torch.zeros(d, dtype=torch.float64)
    abar = torch.zeros(d, dtype=torch.float64)
    hist = torch.zeros(shard.num_rows, dtype=torch.float64) if saga else None
    snaps = [(0.0, w.clone())]
    losses = []
    t0 = time.perf_counter()
    for k in range(cfg.num_iterations):
        if world > 1:
            dist.broadcast(w, 0)                                        # driver -> workers
        rows = _sampled(shard, cfg.seed + k + 1, cfg.batch_rate)
        sub = shard.select(rows)
        if rows.numel():
            g, l, n, coeff = grad.batch(sub, w, hist[rows] if saga else None)
            if saga:
                hist[rows] = coeff
        else:
            g, l, n = torch.zeros(d, dtype=torch.float64), torch.zeros(()), 0
        pack = torch.cat([g.double(), torch.tensor([float(l), float(n)], dtype=torch.float64)])
        if world > 1:
            dist.reduce(pack, 0)                                        # workers -> driver
        if rank == 0:
            G, loss_sum, cnt = pack[:d], float(pack[d]), float(pack[d + 1])
            if saga:
                w = w - cfg.step_size * (G / (cfg.batch_rate * N) + abar)
                abar = abar + G / N
            else:
                w = w - (cfg.step_size / math.sqrt(k + 1)) * G / (cfg.batch_rate * N)
            losses.append(loss_sum / max(cnt, 1.0))
            if k % cfg.printer_freq == 0:
                snaps.append(((time.perf_counter() - t0) * 1e3, w.clone()))
    if world > 1:
        dist.broadcast(w, 0)
    rep = RunReport(cfg.algorithm, world, iterations=cfg.num_iterations, elapsed_ms=(time.perf_counter() - t0) * 1e3,
                    snapshots=snaps, accepted=cfg.num_iterations, final_weights=w.clone())
    rep.extra["loss_history"] = losses
    # objective of the final model over all rows
    m = shard.matvec(w)
    _, lv = grad.coefficients(m, shard.y)
    tot = torch.tensor([float(lv.sum())], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot)
    rep.objective = [(rep.elapsed_ms, float(tot) / N)]
    return rep


TAG_RESULT, TAG_MODEL = 11, 12


def run_async_gloo(cfg: AsyncConfig):
    """Parameter-server ASGD/ASAGA over gloo point-to-point.  Message to the server:
    [worker, ts, n, loss, g(d)]; reply: [clock, k, stop, w(d)]."""
    from ..algo.telemetry import RunReport
    rank, world = _init()
    shard, N = _load_local(cfg, rank, world)
    grad = make_gradient(cfg.loss)
    d = shard.num_features
    saga = cfg.algorithm == "asaga"
    P = world
    par_recs = cfg.batch_rate * N / P
    iters = cfg.num_iterations
    state = {"w": torch.zeros(d, dtype=torch.float64), "abar": torch.zeros(d, dtype=torch.float64), "k": 0, "clock": 0,
             "dropped": 0}
    stal: List[int] = []
    lock = threading.Lock()
    t0 = time.perf_counter()

    def apply(worker: int, ts: int, g: torch.Tensor) -> None:
        with lock:
            k, clock = state["k"], state["clock"]
            staleness = (k - ts) if saga else (clock - ts)
            state["clock"] = clock + 1
            stal.append(staleness)
            if k >= iters:
                return
            if staleness <= cfg.taw:
                if saga:
                    state["w"] = state["w"] - cfg.step_size * (g / par_recs + state["abar"])
                    state["abar"] = state["abar"] + g / N
                else:
                    state["w"] = state["w"] - (cfg.step_size / math.sqrt(k // P + 1)) * g / par_recs
                state["k"] = k + 1
            else:
                state["dropped"] += 1

    def server_loop():
        live = set(range(1, world))
        buf = torch.zeros(4 + d, dtype=torch.float64)
        while live:
            src = dist.recv(buf, tag=TAG_RESULT)
            worker, ts = int(buf[0]), int(buf[1])
            if buf[2] >= 0:
                apply(worker, ts, buf[4:].clone())
            with lock:
                stop = 1.0 if state["k"] >= iters else 0.0
                reply = torch.cat([torch.tensor([float(state["clock"]), float(state["k"]), stop], dtype=torch.float64), state["w"]])
            dist.send(reply, src, tag=TAG_MODEL)
            if stop:
                live.discard(src)

    hist = torch.zeros(shard.num_rows, dtype=torch.float64) if saga else None
    server = None
    if rank == 0 and world > 1:
        server = threading.Thread(target=server_loop, daemon=True)
        server.start()
    # worker loop (rank 0 is also a worker and talks to the server state directly)
    w_local = torch.zeros(d, dtype=torch.float64)
    ts, k_seen = 0, 0
    last_k, resubmits = -1, 0
    while True:
        resubmits = resubmits + 1 if k_seen == last_k else 0       # same empty-sample guard as algo/asgd.py
        last_k = k_seen
        rows = _sampled(shard, cfg.seed + k_seen + 1 + 7919 * resubmits, cfg.batch_rate)
        if rows.numel():
            sub = shard.select(rows)
            g, l, n, coeff = grad.batch(sub, w_local, hist[rows] if saga else None)
        else:
            g, l, n, coeff = torch.zeros(d, dtype=torch.float64), 0.0, -1, None
        if rank == 0:
            if n >= 0:
                before = state["k"]
                apply(0, ts, g.double())
                if saga and state["k"] > before:
                    hist[rows] = coeff
            with lock:
                w_local, ts, k_seen = state["w"].clone(), (state["k"] if saga else state["clock"]), state["k"]
                stop = state["k"] >= iters
        else:
            msg = torch.cat([torch.tensor([float(rank), float(ts), float(n), float(l)], dtype=torch.float64), g.double()])
            dist.send(msg, 0, tag=TAG_RESULT)
            reply = torch.zeros(3 + d, dtype=torch.float64)
            dist.recv(reply, 0, tag=TAG_MODEL)
            k_new = int(reply[1])
            if saga and n >= 0 and k_new > k_seen:
                hist[rows] = coeff           # optimistic commit: server accepted something since we pulled
            ts = k_new if saga else int(reply[0])
            k_seen = k_new
            w_local = reply[3:].clone()
            stop = bool(reply[2])
        if stop:
            break
    if server is not None:
        server.join(timeout=30)
    w = state["w"].clone() if rank == 0 else w_local
    if world > 1:
        dist.broadcast(w, 0)
    rep = RunReport(cfg.algorithm, world, iterations=state["k"] if rank == 0 else k_seen,
                    elapsed_ms=(time.perf_counter() - t0) * 1e3, accepted=state["k"], dropped=state["dropped"],
                    final_weights=w)
    m = shard.matvec(w)
    _, lv = grad.coefficients(m, shard.y)
    tot = torch.tensor([float(lv.sum())], dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tot)
    rep.objective = [(rep.elapsed_ms, float(tot) / N)]
    return rep


def _worker_entry(rank: int, world: int, port: int, cfg_dict: dict, mode: str, out_q):
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "MASTER_ADDR": "127.0.0.1", "MASTER_PORT": str(port)})
    cfg = AsyncConfig.from_dict(cfg_dict)
    rep = run_sync_gloo(cfg) if mode == "sync" else run_async_gloo(cfg)
    if rank == 0:
        out_q.put({"objective": rep.objective, "w": rep.final_weights.tolist(), "iterations": rep.iterations,
                   "loss_history": rep.extra.get("loss_history", []), "dropped": rep.dropped})
    dist.barrier()
    dist.destroy_process_group()


def spawn_local(cfg: AsyncConfig, world: int = 2, mode: str = "sync", port: Optional[int] = None, timeout: float = 120.0) -> dict:
    """Run ``world`` CPU worker processes on this box (gloo over 127.0.0.1) and return rank 0's summary."""
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = port or (29600 + os.getpid() % 300)
    procs = [ctx.Process(target=_worker_entry, args=(r, world, port, cfg.to_dict(), mode, q)) for r in range(world)]
    for p in procs:
        p.start()
    try:
        out = q.get(timeout=timeout)
    finally:
        for p in procs:
            p.join(timeout=30)
            if p.is_alive():
                p.terminate()
    return out
