"""The four ASYNC drivers on the fused backend (one process per GPU, in-kernel NVLink push/pull).

Same knobs and report as the in-process drivers; the difference is *where* the protocol runs:
submitter gating, logical clock, staleness filter, apply and availability bookkeeping all execute
on the device inside the round kernel (csrc/include/afb_round_impl.cuh).  The host only enqueues
rounds (``rounds_per_graph`` at a time), polls the accepted-update counter, takes the periodic
``(elapsed ms, w)`` snapshots and finally evaluates the objective of every snapshot in one pass.
"""
from __future__ import annotations

import time
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.distributed as dist

from ..config import AsyncConfig
from ..data.batch import Shard
from ..data.generators import SHAPES, shard_rows, synthetic_shard
from ..optim.gradient import LOSS_IDS
from ..parallel.comm import init_distributed
from ..parallel.fused_backend import FusedEngine
from .common import say
from .telemetry import MetricsLog, RunReport


def _local_shards(cfg: AsyncConfig, points, rank: int, world: int, device) -> Tuple[List[Shard], List[int], int]:
    P = cfg.num_partitions
    ids = [p for p in range(P) if p % world == rank]
    if points is not None:
        if points.num_partitions != P:
            points = points.repartition(P)
        sizes = points.partition_sizes()
        total = sum(sizes)
        idx = points.zip_with_index()
        shards = [idx.batch(p).to(device, torch.float32) for p in ids]
        return shards, ids, total
    if cfg.synthetic in SHAPES:
        total = int(cfg.num_rows or SHAPES[cfg.synthetic][0])
        sizes = shard_rows(total, P)
        labels = "regression" if cfg.loss.startswith("least_squares") else "binary01"
        shards = [synthetic_shard(cfg.synthetic, sizes[p], p, P, cfg.seed, device, torch.float32,
                                  cfg.num_features or None, labels, sum(sizes[:p])) for p in ids]
        return shards, ids, total
    # LIBSVM / small generators: every rank parses the file and keeps its partitions (load-time only)
    from ..core.scheduler import Scheduler
    from .common import load_points
    sched = Scheduler(1)
    try:
        ds = load_points(cfg, sched, torch.device("cpu"), torch.float32).zip_with_index()
        total = ds.count()
        shards = []
        for p in ids:
            s = ds.batch(p)
            if cfg.layout == "dense" or (cfg.layout == "auto" and (s.is_dense or s.nnz >= cfg.density_threshold * s.num_rows * s.num_features)):
                s = s.to_dense()
            else:
                s = s.to_csr()
            shards.append(s.to(device, torch.float32))
        return shards, ids, total
    finally:
        sched.stop()


def evaluate_snapshots(engine: FusedEngine, shards: List[Shard], snaps: List[Tuple[float, torch.Tensor]], loss: str,
                       total_rows: int) -> List[Tuple[float, float]]:
    """Objective of every snapshot in one pass over all shards (K12) + cross-rank sum (NCCL: not a hot path)."""
    from ..ops.eval import eval_loss
    if not snaps:
        return []
    dev = engine.device
    T = len(snaps)
    W = torch.stack([w.to(dev, torch.float32) for _, w in snaps], dim=1) if engine.rank == 0 else \
        torch.empty(engine.d, T, dtype=torch.float32, device=dev)
    if engine.world > 1:
        dist.broadcast(W, 0)
    tot = torch.zeros(T, dtype=torch.float64, device=dev)
    for s in shards:
        if s.device.type == "cuda":
            tot += eval_loss(s, W, LOSS_IDS[loss])
        else:
            from ..optim.gradient import make_gradient
            M = torch.stack([s.matvec(W[:, t]) for t in range(T)], dim=1).double()
            _, L = make_gradient(loss).coefficients(M, s.y.double()[:, None].expand_as(M))
            tot += L.sum(0)
    if engine.world > 1:
        dist.all_reduce(tot)
    vals = (tot / total_rows).tolist()
    return [(t, float(v)) for (t, _), v in zip(snaps, vals)]


def run_fused(cfg: AsyncConfig, points=None) -> RunReport:
    rank, local_rank, world = init_distributed()
    if torch.cuda.is_available():
        device = torch.device("cuda", local_rank)
        torch.cuda.set_device(device)
    else:
        # no GPU: the same host logic runs against the pure-Python round simulator (ops/round_sim.py)
        device = torch.device("cpu")
    if cfg.num_partitions % world != 0:
        raise ValueError(f"num_partitions ({cfg.num_partitions}) must be a multiple of the number of ranks ({world})")
    metrics = MetricsLog(cfg.metrics_path if rank == 0 else None)
    shards, ids, total = _local_shards(cfg, points, rank, world, device)
    w0 = None
    k0 = 0
    if cfg.resume:
        from ..utils.checkpoint import load_checkpoint
        ck = load_checkpoint(cfg.resume, map_location=device)
        w0 = ck["w"]
        k0 = int(ck.get("k", 0))
    engine = FusedEngine(cfg, shards, total, ids, device)
    if w0 is not None:
        engine.block.w[: engine.d].copy_(w0.to(device, torch.float32))
        if ck.get("alpha_bar") is not None:
            engine.block.abar[: engine.d].copy_(ck["alpha_bar"].to(device, torch.float32))
        for rk, wid in zip(engine.kernels, ids):
            h = ck.get("history", {}).get(wid)
            if h is not None and rk.hist is not None:
                rk.hist.copy_(h.to(device))
        # resume is EXACT: the accepted-update counter k, the logical clock and the per-worker round counters continue, so the
        # step-size schedule (gamma/sqrt(floor(k/P)+1), rsqrt(round+1)) and the seed+round+1 sampling sequence pick up where
        # the checkpoint left off, and `num_iterations` is a TOTAL (as in the host drivers)
        rounds = (ck.get("extra") or {}).get("rounds")
        if rounds is not None and len(rounds) != len(engine.kernels):
            rounds = None
        engine.restore_progress(k0, ck.get("clock"), rounds)
    engine.barrier()
    is_async = engine.is_async
    iters = cfg.num_iterations
    P = cfg.num_partitions
    snaps: List[Tuple[float, torch.Tensor]] = [(0.0, engine.weights().clone())] if rank == 0 else []
    next_snap = cfg.printer_freq
    R = max(1, int(cfg.rounds_per_graph))
    t0 = time.perf_counter()
    k = k0
    while next_snap <= k:
        next_snap += cfg.printer_freq
    ck_next = (k0 // cfg.checkpoint_every + 1) * cfg.checkpoint_every if cfg.checkpoint_every else None
    ck_writer = None
    if ck_next is not None and cfg.checkpoint_dir:
        from ..utils.checkpoint import AsyncCheckpointWriter
        ck_writer = AsyncCheckpointWriter()          # snapshots are written in the background (copy-on-snapshot)
    # failure detection: rank 0 polls the device-side status table at every snapshot; a worker that has been
    # busy without a heartbeat for cfg.heartbeat_timeout_s is reported (asynchronous runs simply continue without it)
    from ..utils.watchdog import DeviceWatchdog
    watchdog = DeviceWatchdog(engine, timeout_s=cfg.heartbeat_timeout_s) if (rank == 0 and is_async) else None
    while True:
        if is_async:
            engine.run_rounds(R)
            st = engine.server_state()
            k = min(st["k"], iters)
        else:
            n = min(R, iters - k)
            engine.run_rounds(n)
            engine.synchronize()
            k += n
        if rank == 0 and k >= next_snap:
            snaps.append(((time.perf_counter() - t0) * 1e3, engine.weights().clone()))
            say(cfg, f"Iteration {k} is finished")
            metrics.emit("snapshot", k=k, elapsed_ms=snaps[-1][0])
            if watchdog is not None:
                for wdead in watchdog.check_once(expect_running=True):
                    say(cfg, f"worker {wdead} declared dead (no device heartbeat for {cfg.heartbeat_timeout_s} s): continuing with "
                             f"{P - len(watchdog.dead)} workers")
                    metrics.emit("worker_dead", worker=wdead)
            while next_snap <= k:
                next_snap += cfg.printer_freq
        if ck_next is not None and k >= ck_next and cfg.checkpoint_dir:
            from ..utils.checkpoint import save_engine_checkpoint
            save_engine_checkpoint(engine, cfg, k, writer=ck_writer)
            ck_next += cfg.checkpoint_every
        if k >= iters:
            break
    engine.synchronize()
    elapsed = (time.perf_counter() - t0) * 1e3
    if ck_writer is not None:
        ck_writer.close()
    engine.barrier()
    rep = RunReport(cfg.algorithm, P, iterations=k, elapsed_ms=elapsed, snapshots=snaps, accepted=k)
    if rank == 0:
        status = engine.worker_status()
        rep.waiting_ms = {p: float(status[p]["wait_ns"]) * 1e-6 for p in range(P)}
        if k:
            rep.avg_waiting_ms = sum(rep.waiting_ms.values()) / (P * k) if is_async else sum(rep.waiting_ms.values()) / (P * k * P)
        res = engine.drain_results()
        rep.dropped = sum(1 for r in res if not r.data["accepted"])
        for r in res:
            rep.staleness_hist[r.staleness] = rep.staleness_hist.get(r.staleness, 0) + 1
        rep.extra["avg_task_ms"] = {p: float(status[p]["avg_task_ms"]) for p in range(P)}
        rep.extra["avg_delay_ms"] = engine.server_state()["avg_delay_ns"] * 1e-6
        if watchdog is not None:
            watchdog.check_once()
            rep.extra["dead_workers"] = list(watchdog.dead)
        snaps.append((elapsed, engine.weights().clone()))
    rep.final_weights = engine.weights().clone()
    if cfg.eval_objective:
        obj = evaluate_snapshots(engine, shards, snaps, cfg.loss, total)
        if rank == 0:
            rep.objective = obj
    rep.extra["launches"] = engine.launches
    metrics.emit("report", **{k2: v for k2, v in rep.to_dict().items()})
    metrics.close()
    return rep
