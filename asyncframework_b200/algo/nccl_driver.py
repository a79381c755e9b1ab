"""``--backend=nccl``: the synchronous drivers on stock PyTorch collectives (SURVEY §7.1 "cuda-nccl (baseline)", §7.2 step 5).

One process per GPU under ``torchrun`` (or ``asyncframework_b200.launch``); every iteration is ``dist.broadcast(w)`` +
Bernoulli sample + cuBLAS GEMVs + ``dist.all_reduce(g)`` + the SGD / SAGA update (``baseline/nccl_sync.py::NcclSyncSGD``,
the class ``bench.py --impl nccl`` times).  This file wraps it in the same driver contract as every other backend — the
reference's positional CLI, ``(elapsed ms, w)`` snapshots every ``printer_freq`` iterations, the post-run
``time,objective`` table, JSONL metrics — so the bar the fused backend is measured against is runnable exactly like the
product: ``torchrun --nproc-per-node 8 -m asyncframework_b200.algo.sgd_sync <11 args> --backend=nccl``.
Without CUDA the same code runs over gloo (used by the CPU test-suite).
"""
from __future__ import annotations

import time
from typing import List, Tuple

import torch
import torch.distributed as dist

from ..config import AsyncConfig
from ..parallel.comm import init_distributed
from .common import say
from .fused_driver import _local_shards
from .telemetry import MetricsLog, RunReport


def run_nccl(cfg: AsyncConfig, points=None) -> RunReport:
    if cfg.algorithm not in ("sgd_sync", "saga_sync"):
        raise ValueError("--backend=nccl implements the synchronous drivers (sgd_sync, saga_sync); "
                         "the asynchronous ones need one-sided communication: use --backend=fused")
    import sys
    import os
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
    from baseline.nccl_sync import NcclSyncSGD
    owns_group = not dist.is_initialized()
    rank, local_rank, world = init_distributed()
    device = torch.device("cuda", local_rank) if torch.cuda.is_available() else torch.device("cpu")
    if cfg.num_partitions % world != 0:
        raise ValueError(f"num_partitions ({cfg.num_partitions}) must be a multiple of the number of ranks ({world})")
    metrics = MetricsLog(cfg.metrics_path if rank == 0 else None)
    shards, ids, total = _local_shards(cfg, points, rank, world, device)
    dense = [s.to_dense() if not s.is_dense else s for s in shards]      # stock PyTorch path: dense rows + cuBLAS
    X = torch.cat([s.X for s in dense], 0) if dense else torch.zeros(0, 4, device=device)
    y = torch.cat([s.y for s in dense], 0) if dense else torch.zeros(0, device=device)
    d = dense[0].num_features
    loss = "logistic" if cfg.loss == "logistic" else "least_squares"
    if cfg.loss == "hinge":
        raise ValueError("the NCCL baseline implements least_squares and logistic")
    opt = NcclSyncSGD(X.float(), y.float(), d, total, cfg.batch_rate, cfg.step_size, cfg.seed, loss, saga=cfg.algorithm == "saga_sync")
    snaps: List[Tuple[float, torch.Tensor]] = [(0.0, opt.w[:d].clone())] if rank == 0 else []
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for k in range(cfg.num_iterations):
        opt.step()
        if rank == 0 and (k + 1) % cfg.printer_freq == 0:
            if device.type == "cuda":
                torch.cuda.synchronize()
            snaps.append(((time.perf_counter() - t0) * 1e3, opt.w[:d].clone()))
            say(cfg, f"Iteration {k + 1} is finished")
            metrics.emit("snapshot", k=k + 1, elapsed_ms=snaps[-1][0])
    if device.type == "cuda":
        torch.cuda.synchronize()
    elapsed = (time.perf_counter() - t0) * 1e3
    rep = RunReport(cfg.algorithm, cfg.num_partitions, iterations=cfg.num_iterations, elapsed_ms=elapsed, snapshots=snaps,
                    accepted=cfg.num_iterations, final_weights=opt.w[:d].clone())
    rep.extra["backend"] = f"nccl-baseline ({dist.get_backend() if dist.is_initialized() else 'single process'})"
    if cfg.eval_objective:
        if rank == 0:
            snaps.append((elapsed, opt.w[:d].clone()))
        T = torch.tensor([len(snaps)], device=device)
        if world > 1:
            dist.broadcast(T, 0)
        W = torch.stack([w.float() for _, w in snaps], 1) if rank == 0 else torch.empty(d, int(T), device=device)
        if world > 1:
            dist.broadcast(W, 0)
        M = X[:, :d].float() @ W
        yy = y.float()[:, None]
        L = (torch.nn.functional.softplus(-M) + M * (1 - yy)) if loss == "logistic" else (M - yy) ** 2
        tot = L.double().sum(0)
        if world > 1:
            dist.all_reduce(tot)
        if rank == 0:
            rep.objective = [(t, float(v)) for (t, _), v in zip(snaps, (tot / total).tolist())]
    metrics.emit("report", **rep.to_dict())
    metrics.close()
    if owns_group and dist.is_initialized():
        # leave the process group cleanly: a rank that exits with live gloo / NCCL threads can abort during interpreter teardown
        try:
            dist.barrier()
            dist.destroy_process_group()
        except Exception:  # noqa: BLE001
            pass
    return rep
