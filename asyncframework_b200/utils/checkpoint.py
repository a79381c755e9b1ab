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
import queue
import threading
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


class AsyncCheckpointWriter:
    """Copy-on-snapshot checkpointing that does not stall the update stream: the caller takes device-side
    copies of the state (a few KB of model + the history shards, D2D on the launching stream) and hands them over with a
    CUDA event; a background thread waits for the event, moves the copies to the host and writes the file.  Rounds
    enqueued after the snapshot never wait for the D2H transfer or the disk.  At most ``max_pending`` snapshots are
    buffered (a slower disk then back-pressures ``submit``)."""

    def __init__(self, max_pending: int = 2):
        self._q: "queue.Queue" = queue.Queue(maxsize=max_pending)
        self.written = 0
        self.error: Optional[BaseException] = None
        self._thread = threading.Thread(target=self._loop, name="afb-checkpoint-writer", daemon=True)
        self._thread.start()

    def _loop(self) -> None:
        while True:
            job = self._q.get()
            try:
                if job is None:
                    return
                event, kwargs, after = job
                if event is not None:
                    event.synchronize()               # the device-side copies are complete
                save_checkpoint(**kwargs)
                if after is not None:
                    after()
                self.written += 1
            except BaseException as exc:  # noqa: BLE001  (surfaced by flush/close)
                self.error = exc
            finally:
                self._q.task_done()

    def submit(self, event, after=None, **save_kwargs) -> None:
        if self.error is not None:
            raise RuntimeError("checkpoint writer failed") from self.error
        self._q.put((event, save_kwargs, after))

    def flush(self) -> None:
        self._q.join()
        if self.error is not None:
            raise RuntimeError("checkpoint writer failed") from self.error

    def close(self) -> None:
        self._q.join()
        self._q.put(None)
        self._thread.join(timeout=30)
        if self.error is not None:
            raise RuntimeError("checkpoint writer failed") from self.error


def save_engine_checkpoint(engine, cfg, k: int, writer: Optional[AsyncCheckpointWriter] = None) -> str:
    """Snapshot a FusedEngine: rank-local file with this rank's history shards + (rank 0) the model.  With a ``writer`` the
    state is copied on the device and written in the background; without one the call returns after the file is on disk."""
    d = cfg.checkpoint_dir
    os.makedirs(d, exist_ok=True)
    st = engine.server_state()
    path = os.path.join(d, f"rank{engine.rank}.pt")
    rank, world = engine.rank, engine.world

    def publish_latest():
        if rank == 0:
            tmp = os.path.join(d, "latest.json.tmp")
            with open(tmp, "w") as f:
                json.dump({"k": int(k), "time": time.time(), "world": world}, f)
            os.replace(tmp, os.path.join(d, "latest.json"))

    if writer is None:
        hist = {wid: rk.hist for rk, wid in zip(engine.kernels, engine.local_worker_ids) if rk.hist is not None}
        save_checkpoint(path, w=engine.weights(), k=k, clock=st["clock"], alpha_bar=engine.alpha_bar(), history=hist,
                        config=cfg.to_dict(), extra={"world": world, "rounds": [int(rk.rounds_launched) for rk in engine.kernels]})
        publish_latest()
        return path
    from ..parallel import _devapi as _dev
    # copy-on-snapshot: device-side clones ordered after the rounds already enqueued on the worker streams
    cur = _dev.current_stream(engine.device)
    for stt in engine.streams:
        if stt is not cur:
            ev = _dev.new_event(engine.device)
            ev.record(stt)
            cur.wait_event(ev)
    hist = {wid: rk.hist.clone() for rk, wid in zip(engine.kernels, engine.local_worker_ids) if rk.hist is not None}
    w = engine.weights().clone()
    ab = engine.alpha_bar()
    ab = None if ab is None else ab.clone()
    done = _dev.new_event(engine.device)
    done.record(cur)
    writer.submit(done, after=publish_latest, path=path, w=w, k=k, clock=st["clock"], alpha_bar=ab, history=hist,
                  config=cfg.to_dict(), extra={"world": world, "rounds": [int(rk.rounds_launched) for rk in engine.kernels]})
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
