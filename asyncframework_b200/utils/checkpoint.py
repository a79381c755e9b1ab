"""Checkpoint / resume.

The reference has none for ASYNC (model snapshots live only in the driver's ``optVars``; the CSV
writers are commented out, SparkASGDThread.scala:370-385); Spark's own mechanisms are RDD
checkpointing and MLlib model save/load.  Here a checkpoint captures everything needed to resume an
optimisation run exactly: ``w``, ``alpha_bar``, the per-worker SAGA history shards, the accepted-
update counter ``k``, the logical clock, the RNG seed and the config.  On the fused backend the
snapshot is taken between round batches from the server block without stalling in-flight kernels
on other ranks (each rank writes its own history shard: ``<dir>/rank<r>.pt``).
"""
from __future__ import annotations

import json
import os
import time
from typing import Any, Dict, Optional

import torch


def save_checkpoint(path: str, *, w: torch.Tensor, k: int, clock: int = 0, alpha_bar: Optional[torch.Tensor] = None,
                    history: Optional[Dict[int, torch.Tensor]] = None, config: Optional[dict] = None,
                    extra: Optional[dict] = None) -> str:
    os.makedirs(os.path.dirname(os.path.abspath(path)) or ".", exist_ok=True)
    payload = {
        "format": "asyncframework_b200.checkpoint.v1", "time": time.time(), "k": int(k), "clock": int(clock),
        "w": w.detach().cpu(), "alpha_bar": None if alpha_bar is None else alpha_bar.detach().cpu(),
        "history": {int(p): h.detach().cpu() for p, h in (history or {}).items()},
        "config": config or {}, "extra": extra or {},
    }
    tmp = path + ".tmp"
    torch.save(payload, tmp)
    os.replace(tmp, path)          # atomic publish
    return path


def load_checkpoint(path: str, map_location="cpu") -> Dict[str, Any]:
    if os.path.isdir(path):
        rank = int(os.environ.get("RANK", "0"))
        cand = os.path.join(path, f"rank{rank}.pt")
        path = cand if os.path.exists(cand) else os.path.join(path, "rank0.pt")
    ck = torch.load(path, map_location=map_location, weights_only=False)
    if ck.get("format") != "asyncframework_b200.checkpoint.v1":
        raise ValueError(f"{path} is not an asyncframework_b200 checkpoint")
    return ck


def save_engine_checkpoint(engine, cfg, k: int) -> str:
    """Snapshot a FusedEngine: rank-local file with this rank's history shards + (rank 0) the model."""
    d = cfg.checkpoint_dir
    os.makedirs(d, exist_ok=True)
    st = engine.server_state()
    hist = {wid: rk.hist for rk, wid in zip(engine.kernels, engine.local_worker_ids) if rk.hist is not None}
    path = os.path.join(d, f"rank{engine.rank}.pt")
    save_checkpoint(path, w=engine.weights(), k=k, clock=st["clock"], alpha_bar=engine.alpha_bar(), history=hist,
                    config=cfg.to_dict(), extra={"world": engine.world})
    if engine.rank == 0:
        with open(os.path.join(d, "latest.json"), "w") as f:
            json.dump({"k": int(k), "time": time.time(), "world": engine.world}, f)
    return path


def save_model_text(path: str, weights: torch.Tensor, intercept: float = 0.0) -> None:
    """MLlib-compatible *content* (weights + intercept) as plain text."""
    with open(path, "w") as f:
        f.write(f"intercept {float(intercept)!r}\n")
        f.write(" ".join(repr(float(v)) for v in weights.detach().cpu().double().tolist()) + "\n")


def load_model_text(path: str):
    with open(path) as f:
        first = f.readline().split()
        w = torch.tensor([float(t) for t in f.readline().split()], dtype=torch.float64)
    return w, float(first[1])
