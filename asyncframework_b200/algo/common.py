"""Shared driver plumbing: data loading, gradient closure, CLI entry, result printing."""
from __future__ import annotations

import logging
import math
import sys
import time
from typing import Callable, List, Optional, Tuple

import torch

from ..config import AsyncConfig, parse_cli
from ..core.context import AsyncContext
from ..core.scheduler import Scheduler
from ..data.batch import Shard
from ..data.dataset import ShardedDataset
from ..data.generators import (LinearDataGenerator, LogisticRegressionDataGenerator, SHAPES, synthetic_dataset)
from ..data.libsvm import load_libsvm_file
from ..optim.gradient import Gradient, make_gradient

log = logging.getLogger("asyncframework_b200.algo")


def now_ms() -> float:
    return time.perf_counter() * 1e3


def resolve_device(cfg: AsyncConfig) -> torch.device:
    if cfg.backend in ("cpu",):
        return torch.device("cpu")
    if torch.cuda.is_available():
        return torch.device("cuda", torch.cuda.current_device())
    return torch.device("cpu")


def compute_dtype(cfg: AsyncConfig, device: torch.device) -> torch.dtype:
    if device.type == "cuda":
        return torch.float32
    return torch.float64 if cfg.dtype in ("float64", "auto", "float32") and cfg.extra.get("cpu_dtype", "float64") == "float64" else torch.float32


def load_points(cfg: AsyncConfig, scheduler: Scheduler, device: torch.device, dtype: torch.dtype) -> ShardedDataset:
    """``MLUtils.loadLibSVMFile(sc, path+file).repartition(numPart)`` (+ the optional unseeded 10 %
    subsample that 4 of the 5 reference drivers apply at load, quirk B.9 -> ``load_subsample``)."""
    P = cfg.num_partitions
    if cfg.synthetic:
        name = cfg.synthetic
        rows = cfg.num_rows or None
        if name == "linear":
            ds = LinearDataGenerator.generate_linear_dataset(scheduler, rows or 1000, cfg.num_features or 20, 0.1, P,
                                                             seed=cfg.seed, dtype=dtype, device=device)
        elif name == "logistic":
            ds = LogisticRegressionDataGenerator.generate_logistic_dataset(scheduler, rows or 1000, cfg.num_features or 20, 1.0, P,
                                                                           seed=cfg.seed, dtype=dtype, device=device)
        elif name in SHAPES:
            labels = "regression" if cfg.loss.startswith("least_squares") else "binary01"
            ds = synthetic_dataset(scheduler, name, P, rows, cfg.seed, device, dtype, labels, cfg.num_features or None)
        else:
            raise ValueError(f"unknown synthetic dataset {name!r}")
    else:
        ds = load_libsvm_file(scheduler, cfg.data_path, cfg.num_features if cfg.num_features > 0 else -1, P,
                              dtype=dtype, layout=cfg.layout, density_threshold=cfg.density_threshold)
        if ds.num_features != cfg.num_features and cfg.num_features > 0:
            raise ValueError(f"num_features mismatch: file has {ds.num_features}, CLI says {cfg.num_features}")
        ds = ds.repartition(P)
        if device.type != "cpu":
            ds = ShardedDataset.from_shards(scheduler, [s.to(device) for s in ds.shards()], ds.name)
    if cfg.load_subsample:
        ds = ds.sample(False, cfg.load_subsample, None)
        ds = ShardedDataset.from_shards(scheduler, [s for s in ds.shards()], "subsampled")
    return ds


def make_scheduler(cfg: AsyncConfig, device: torch.device) -> Scheduler:
    master = getattr(cfg, "master", "") or ""
    if master.startswith(("local-cluster", "spark://", "async://")):
        # ``--master`` / ``spark.master``: the driver's closures run in executor processes (forked locally or by standalone
        # Workers); the data then lives on the CPU of those processes — GPU training at scale is the fused backend's job
        from ..core.session import AsyncSession
        from ..conf import AsyncConf
        rc = AsyncConf()                                    # submit exported --name / --conf as AFB_CONF_* (system properties)
        sess = AsyncSession(master, app_name=None if rc.contains("spark.app.name") else cfg.algorithm, conf=cfg, runtime_conf=rc)
        sess.scheduler.block_first_async_job = bool(cfg.extra.get("block_first_async_job", True))
        sess.scheduler._owning_session = sess
        return sess.scheduler
    return Scheduler(num_executors=cfg.num_partitions, device=device if device.type == "cuda" else None,
                     block_first_async_job=bool(cfg.extra.get("block_first_async_job", True)))


def partition_gradient(gradient: Gradient, shard: Shard, w: torch.Tensor, history: Optional[torch.Tensor] = None):
    """(g[d], loss_sum, n, coeff or None) for one (already sampled) shard — torch oracle path on CPU
    tensors, fused CUDA kernel on HBM-resident shards."""
    if shard.device.type == "cuda" and history is None:
        from ..ops.grad import grad
        g, l, n = grad(shard, w, gradient.loss_id)
        return g, l, n, None
    g, l, n, coeff = gradient.batch(shard, w, history)
    return g, float(l), n, coeff


def start_watchdog(cfg: AsyncConfig, AC, metrics=None):
    """Failure detection for the host drivers: a worker that stays busy without a heartbeat for
    ``cfg.heartbeat_timeout_s`` is marked dead in ``AC.stat`` — async drivers then continue with the remaining workers
    (dead workers are never selected by ``async_barrier`` nor counted as available), sync drivers abort."""
    from ..utils.watchdog import HeartbeatWatchdog

    def on_dead(worker: int) -> None:
        say(cfg, f"worker {worker} declared dead (no heartbeat for {cfg.heartbeat_timeout_s} s)")
        if metrics is not None:
            metrics.emit("worker_dead", worker=worker)
    timeout = float(cfg.heartbeat_timeout_s)
    return HeartbeatWatchdog(AC.stat, timeout_s=timeout, period_s=max(0.01, min(0.5, timeout / 4)), on_dead=on_dead).start()


class HostCheckpointer:
    """``--checkpoint_dir/--checkpoint_every/--resume`` for the host (cpu) drivers — the same file format as the fused
    backend (utils/checkpoint.py); files are written by a background thread from copies of the state, so the updater never
    waits for the disk."""

    def __init__(self, cfg: AsyncConfig):
        self.cfg = cfg
        self.every = int(cfg.checkpoint_every) if (cfg.checkpoint_every and cfg.checkpoint_dir) else 0
        self.next = self.every
        self._writer = None

    def resume(self, device, dtype):
        """Returns the checkpoint dict (tensors on ``device``/``dtype``) or None when ``cfg.resume`` is unset."""
        if not self.cfg.resume:
            return None
        from ..utils.checkpoint import load_checkpoint
        ck = load_checkpoint(self.cfg.resume, map_location="cpu")
        ck["w"] = ck["w"].to(device=device, dtype=dtype)
        if ck.get("alpha_bar") is not None:
            ck["alpha_bar"] = ck["alpha_bar"].to(device=device, dtype=dtype)
        ck["history"] = {int(p): h.to(device=device, dtype=dtype) for p, h in (ck.get("history") or {}).items()}
        while self.every and self.next <= int(ck["k"]):
            self.next += self.every
        return ck

    def maybe_save(self, k: int, make_state) -> bool:
        """``make_state()`` -> dict(w=..., alpha_bar=..., history={p: tensor}) of COPIES, called only when a snapshot is due."""
        if not self.every or k < self.next:
            return False
        import json
        import os
        from ..utils.checkpoint import AsyncCheckpointWriter
        if self._writer is None:
            self._writer = AsyncCheckpointWriter()
        d = self.cfg.checkpoint_dir
        os.makedirs(d, exist_ok=True)

        def publish(k=k):
            tmp = os.path.join(d, "latest.json.tmp")
            with open(tmp, "w") as f:
                json.dump({"k": int(k), "world": 1}, f)
            os.replace(tmp, os.path.join(d, "latest.json"))
        self._writer.submit(None, after=publish, path=os.path.join(d, "rank0.pt"), k=int(k), config=self.cfg.to_dict(), **make_state())
        while self.next <= k:
            self.next += self.every
        return True

    def close(self) -> None:
        if self._writer is not None:
            self._writer.close()
            self._writer = None


class host_math_threads:
    """Driver-side vector math (``w`` updates, snapshots, broadcasts of d <= ~50 k elements) is far below the size where
    intra-op threading pays, and on a shared box the OpenMP fork/join costs milliseconds per op (measured: 8.4 ms vs 11 us for a
    378 KB clone).  ``with host_math_threads():`` pins the CALLING thread to one intra-op thread and restores the old value —
    the setting is per thread, so executor threads and other sessions are unaffected.  Also usable as a plain call at the top of
    a helper thread (``host_math_threads.pin()``)."""

    def __init__(self, n: int = 1):
        self.n = n
        self.prev = None

    def __enter__(self):
        self.prev = torch.get_num_threads()
        torch.set_num_threads(self.n)
        return self

    def __exit__(self, *exc):
        if self.prev is not None:
            torch.set_num_threads(self.prev)
        return False

    @staticmethod
    def pin(n: int = 1) -> None:
        torch.set_num_threads(n)


class WorkerLost(RuntimeError):
    """A synchronous (BSP) run cannot make progress without every worker."""


def say(cfg: AsyncConfig, *args) -> None:
    """The reference drivers narrate on stdout (partition sizes, "Iteration k is finished"); ``cfg.quiet`` mutes that."""
    if not getattr(cfg, "quiet", False):
        print(*args)


def run_cli(algorithm: str, main_fn: Callable[[AsyncConfig], object], argv: Optional[List[str]] = None):
    logging.basicConfig(level=logging.WARNING, format="%(levelname)s %(name)s: %(message)s")
    cfg = parse_cli(algorithm, sys.argv[1:] if argv is None else argv)
    # the reference's banner and parameter echo, label for label (SparkASGDThread.scala:26-62, SparkASGDSync.scala:25-57,
    # SparkSGDMLLIB.scala:16-42) — log scrapers written for the reference keep working
    from ..config import LAYOUTS
    banner = {"asgd": "Spark ASGD application started", "asaga": "Spark ASAGA application started", "sgd_sync": "Spark SGD application started",
              "saga_sync": "Spark SAGA application started", "sgd_mllib": "Spark MLlib SGD application started"}[algorithm]
    labels = {"path": "path name", "file": "file name", "num_features": "num columns", "num_rows": "num rows",
              "num_partitions": "num partitions", "num_iterations": "num iterations", "step_size": "step size", "taw": "taw",
              "batch_rate": "batch rate", "bucket_ratio": "bucket ratio", "printer_freq": "printer freq", "coeff": "coeff", "seed": "seed"}
    print(banner)
    print("Input arguments:")
    for name, _ in LAYOUTS[algorithm]:
        print(f"{labels[name]}: {getattr(cfg, name)}")
    with host_math_threads():
        report = main_fn(cfg)
    if report is not None and hasattr(report, "print_reference_style"):
        report.print_reference_style()
    return report
