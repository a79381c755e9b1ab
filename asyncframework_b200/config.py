"""Configuration dataclass + positional-CLI parsing.

The reference parses hand-rolled positional arguments per driver
(ASYNCsamples/.../SparkASGDThread.scala:28-62, SparkASGDSync.scala:28-46,
SparkSGDMLLIB.scala:20-34).  We keep the *same positional order* so README command
lines map 1:1, and add ``--key=value`` overrides on top.
"""
from __future__ import annotations

import dataclasses
import json
import os
from dataclasses import dataclass, field
from typing import Any, Dict, List, Optional, Sequence, Tuple

# Positional layouts (name, type).  Order is the reference's order.
ASYNC_LAYOUT: Tuple[Tuple[str, type], ...] = (
    ("path", str), ("file", str), ("num_features", int), ("num_rows", int),
    ("num_partitions", int), ("num_iterations", int), ("step_size", float),
    ("taw", int), ("batch_rate", float), ("bucket_ratio", float),
    ("printer_freq", int), ("coeff", float), ("seed", int),
)
SYNC_LAYOUT: Tuple[Tuple[str, type], ...] = (
    ("path", str), ("file", str), ("num_features", int), ("num_rows", int),
    ("num_partitions", int), ("num_iterations", int), ("step_size", float),
    ("batch_rate", float), ("printer_freq", int), ("coeff", float), ("seed", int),
)
MLLIB_LAYOUT: Tuple[Tuple[str, type], ...] = (
    ("path", str), ("file", str), ("num_features", int), ("num_rows", int),
    ("num_partitions", int), ("num_iterations", int), ("step_size", float),
    ("batch_rate", float),
)

LAYOUTS = {
    "asgd": ASYNC_LAYOUT, "asaga": ASYNC_LAYOUT,
    "sgd_sync": SYNC_LAYOUT, "saga_sync": SYNC_LAYOUT,
    "sgd_mllib": MLLIB_LAYOUT,
}


@dataclass
class AsyncConfig:
    """All knobs of the five reference drivers plus the B200 runtime knobs.

    Field names follow the reference's CLI (Appendix A of SURVEY.md):
    ``taw`` = max staleness, ``batch_rate`` = b, ``bucket_ratio`` = beta,
    ``coeff`` = delay intensity (-1 = production-cluster straggler model).
    """
    # --- reference knobs -------------------------------------------------
    path: str = ""
    file: str = ""
    num_features: int = 0            # d
    num_rows: int = 0                # N (re-measured after load, like the reference)
    num_partitions: int = 1          # P  (logical workers)
    num_iterations: int = 100
    step_size: float = 1.0           # gamma
    taw: int = 2**31 - 1             # tau (max staleness)
    batch_rate: float = 0.1          # b
    bucket_ratio: float = 1.0        # beta
    printer_freq: int = 100
    coeff: float = 0.0               # delay intensity; -1 => production cluster model
    seed: int = 42
    # --- framework knobs ---------------------------------------------------
    algorithm: str = "asgd"          # asgd | asaga | sgd_sync | saga_sync | sgd_mllib
    loss: str = "least_squares"      # least_squares | logistic | hinge  (Appendix B.17)
    backend: str = "auto"            # cpu | nccl | fused | auto
    dtype: str = "float32"           # storage/compute dtype on GPU backends (cpu: float64)
    layout: str = "auto"             # dense | csr | auto (density threshold at load)
    density_threshold: float = 0.25
    load_subsample: Optional[float] = None  # the reference's unseeded .sample(false,0.1) at load (B.9)
    reg_param: float = 0.0
    updater: str = "simple"          # simple | l1 | squared_l2 (mllib baseline)
    step_decay: str = "reference"    # reference (fork schedules, A12) | sqrt | const
    barrier: str = "bucket"          # bucket | bsp | ssp | asp  (device-side gates)
    ssp_bound: int = 0
    history_slots: int = 8           # AsyncBroadcast ring depth
    warmup_updates_per_worker: int = 100   # delay-injector calibration (A18)
    eval_objective: bool = True
    checkpoint_dir: Optional[str] = None
    checkpoint_every: int = 0
    resume: Optional[str] = None
    metrics_path: Optional[str] = None
    heartbeat_timeout_s: float = 30.0
    rounds_per_graph: int = 16       # CUDA-graph batch of rounds on the fused backend
    synthetic: Optional[str] = None  # mnist8m | epsilon | rcv1 | linear | logistic
    master: str = ""                 # "" = local[P] in-process; local-cluster[n] / spark://host:port run the closures in executor processes
    quiet: bool = False              # suppress the reference-style progress lines on stdout ("Iteration k is finished", ...)
    extra: Dict[str, Any] = field(default_factory=dict)

    # ------------------------------------------------------------------
    @property
    def data_path(self) -> str:
        """Reference concatenates path+file for the ASYNC drivers
        (SparkASGDThread.scala:73) and path+"/"+file for the MLlib driver
        (SparkSGDMLLIB.scala:56)."""
        if self.algorithm == "sgd_mllib":
            return os.path.join(self.path, self.file)
        return self.path + self.file

    @property
    def is_async(self) -> bool:
        return self.algorithm in ("asgd", "asaga")

    @property
    def production_cluster(self) -> bool:
        return self.coeff == -1

    def validate(self) -> "AsyncConfig":
        if self.algorithm not in LAYOUTS:
            raise ValueError(f"unknown algorithm {self.algorithm!r}")
        if self.num_partitions < 1:
            raise ValueError("num_partitions must be >= 1")
        if not (0.0 < self.batch_rate <= 1.0):
            raise ValueError("batch_rate must be in (0, 1]")
        if not (0.0 <= self.bucket_ratio <= 1.0):
            raise ValueError("bucket_ratio must be in [0, 1]")
        if self.loss not in ("least_squares", "logistic", "hinge"):
            raise ValueError(f"unknown loss {self.loss!r}")
        if self.coeff < 0 and self.coeff != -1:
            raise ValueError("coeff must be >= 0 or exactly -1 (production-cluster mode)")
        return self

    def to_dict(self) -> Dict[str, Any]:
        return dataclasses.asdict(self)

    def to_json(self) -> str:
        return json.dumps(self.to_dict(), sort_keys=True)

    @classmethod
    def from_dict(cls, d: Dict[str, Any]) -> "AsyncConfig":
        names = {f.name for f in dataclasses.fields(cls)}
        known = {k: v for k, v in d.items() if k in names}
        extra = {k: v for k, v in d.items() if k not in names}
        cfg = cls(**known)
        cfg.extra.update(extra)
        return cfg


_REF_LABELS = {"path": "path name", "file": "file name", "num_features": "num columns", "num_rows": "num rows",
               "num_partitions": "num partitions", "num_iterations": "num iterations", "step_size": "step size", "taw": "taw",
               "batch_rate": "batch rate", "bucket_ratio": "bucket ratio", "printer_freq": "printer freq", "coeff": "coeff", "seed": "seed"}


def usage(algorithm: str) -> str:
    """The reference's ``Input format: [path name] [file name] …`` line (SparkASGDThread.scala:28-30) + how to call this driver."""
    layout = LAYOUTS[algorithm]
    names = " ".join(f"<{n}>" for n, _ in layout)
    return ("Input format: " + " ".join(f"[{_REF_LABELS[n]}]" for n, _ in layout)
            + f"\nusage: python -m asyncframework_b200.algo.{algorithm} {names} "
            f"[--key=value ...]\n  ({len(layout)} positional arguments, same order as the reference driver)")


def _coerce(tp: type, text: str) -> Any:
    if tp is int:
        # the reference accepts e.g. "20000000" for taw; tolerate "2e7"
        try:
            return int(text)
        except ValueError:
            return int(float(text))
    if tp is float:
        return float(text)
    return text


def _coerce_override(cfg: AsyncConfig, key: str, text: str) -> Any:
    flds = {f.name: f for f in dataclasses.fields(cfg)}
    if key not in flds:
        return text
    cur = getattr(cfg, key)
    if isinstance(cur, bool):
        return text.lower() in ("1", "true", "yes", "on")
    if isinstance(cur, int):
        return _coerce(int, text)
    if isinstance(cur, float):
        return float(text)
    if cur is None:
        low = text.lower()
        if low in ("none", "null"):
            return None
        for tp in (int, float):
            try:
                return tp(text)
            except ValueError:
                pass
        return text
    return text


def parse_cli(algorithm: str, argv: Sequence[str]) -> AsyncConfig:
    """Parse ``argv`` (without the program name) for ``algorithm``.

    Positional tokens are matched against the reference layout; ``--key=value`` or
    ``--key value`` tokens override any ``AsyncConfig`` field.  Like the reference
    (``if (args.length != 13) { usage; exit }``), a wrong positional count is an error.
    """
    layout = LAYOUTS[algorithm]
    positional: List[str] = []
    overrides: List[Tuple[str, str]] = []
    i = 0
    argv = list(argv)
    while i < len(argv):
        tok = argv[i]
        if tok.startswith("--"):
            body = tok[2:]
            if "=" in body:
                k, v = body.split("=", 1)
            else:
                if i + 1 >= len(argv):
                    raise SystemExit(f"missing value for {tok}\n{usage(algorithm)}")
                k, v = body, argv[i + 1]
                i += 1
            overrides.append((k.replace("-", "_"), v))
        else:
            positional.append(tok)
        i += 1
    if len(positional) != len(layout):
        raise SystemExit(f"The input arguments are wrong. (expected {len(layout)} positional arguments, got {len(positional)})\n"
                         + usage(algorithm))
    cfg = AsyncConfig(algorithm=algorithm)
    for (name, tp), text in zip(layout, positional):
        setattr(cfg, name, _coerce(tp, text))
    for k, v in overrides:
        val = _coerce_override(cfg, k, v)
        if k in {f.name for f in dataclasses.fields(cfg)}:
            setattr(cfg, k, val)
        else:
            cfg.extra[k] = val
    return cfg.validate()


def dist_env() -> Tuple[int, int, int]:
    """(rank, local_rank, world_size) from the torchrun environment."""
    return (int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0")),
            int(os.environ.get("WORLD_SIZE", "1")))
