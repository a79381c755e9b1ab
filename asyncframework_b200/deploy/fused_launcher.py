"""Launch the fused multi-GPU backend THROUGH a cluster master instead of ``torchrun`` (the ``spark-submit`` deployment shape for
the product path; multi-host capable because the standalone Master places one executor per GPU worker on any host).

``run_fused_on_cluster(sc, cfg)`` submits ONE job with ``world = sc.scheduler.num_executors`` partitions; task ``r`` runs on
executor slot ``r`` (hard affinity), exports the ``torchrun`` contract (``RANK``, ``WORLD_SIZE``, ``LOCAL_RANK``,
``MASTER_ADDR``, ``MASTER_PORT``) and calls :func:`asyncframework_b200.algo.fused_driver.run_fused` — so every executor process
becomes one rank of the fused engine (NCCL rendezvous through a TCPStore on rank 0's host, symmetric HBM, in-kernel NVLink
push/pull within a node).  With workers started as ``deploy.worker --gpu I`` each executor sees exactly its GPU as ``cuda:0``
(``CUDA_VISIBLE_DEVICES``), hence ``LOCAL_RANK=0``.  Rank 0's :class:`RunReport` is the job's result.

Requirements: one task slot per executor (``spark.executor.cores=1``: all ranks must run concurrently — the function checks),
and ``cfg.num_partitions`` a multiple of the executor count.  Without CUDA the single-rank case runs on the round simulator
(that is what the CPU test-suite exercises); multi-rank needs GPUs.
"""
from __future__ import annotations

import os
import socket
from typing import Optional

from ..config import AsyncConfig


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("", 0))
        return s.getsockname()[1]


def _rank_main(tctx, it, cfg_dict: dict, world: int, master_addr: str, master_port: int):
    rank = tctx.partition_id
    os.environ.update({"RANK": str(rank), "WORLD_SIZE": str(world), "LOCAL_RANK": os.environ.get("AFB_LOCAL_RANK", "0"),
                       "MASTER_ADDR": master_addr, "MASTER_PORT": str(master_port)})
    import torch
    torch.set_num_threads(1)
    from ..algo.fused_driver import run_fused
    rep = run_fused(AsyncConfig.from_dict(cfg_dict))
    if rank != 0:
        return None
    d = rep.to_dict()
    d["final_weights"] = rep.final_weights.cpu() if rep.final_weights is not None else None
    d["objective"] = rep.objective
    d["snapshots"] = [(t, w.cpu()) for t, w in rep.snapshots]
    return d


def run_fused_on_cluster(sc, cfg: AsyncConfig, master_addr: Optional[str] = None, master_port: Optional[int] = None) -> dict:
    sched = sc.scheduler
    world = sched.num_executors
    if getattr(sched, "transport", None) is None:
        raise TypeError("run_fused_on_cluster needs a cluster master (local-cluster[n] or spark://...): ranks are executor PROCESSES")
    if cfg.num_partitions % world != 0:
        raise ValueError(f"num_partitions ({cfg.num_partitions}) must be a multiple of the number of executors ({world})")
    cores = getattr(sched.transport, "cores_per_executor", 1)
    if cores != 1:
        raise ValueError("run the fused backend with one task slot per executor (spark.executor.cores=1): every rank needs its own process")
    if master_addr is None:
        # rank 0 = executor slot 0; its host is where the TCPStore listens.  Standalone: the address the executor registered
        # from; local-cluster: this host.
        addr = getattr(sched.transport, "addresses", {}).get(0)
        master_addr = addr.host if addr is not None else "127.0.0.1"
    master_port = master_port or _free_port()
    cfg_dict = {**cfg.to_dict(), "backend": "fused"}
    ranks = sc.parallelize(range(world), world)
    out = sched.run_job(ranks, lambda t, it: _rank_main(t, it, cfg_dict, world, master_addr, master_port), blocking=True)
    rep = next(r for r in out if r is not None)
    return rep
