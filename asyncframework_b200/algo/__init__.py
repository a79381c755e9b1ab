"""The five reference drivers (SparkASGDThread, SparkASGDSync, SparkASAGAThread, SparkASAGASync,
SparkSGDMLLIB) as Python entry points with the reference's positional CLIs.

Driver modules are imported lazily so that ``python -m asyncframework_b200.algo.<driver>`` does not find its own module
pre-imported by the package."""
import importlib

_DRIVER_NAMES = ("asgd", "asaga", "sgd_sync", "saga_sync", "sgd_mllib")
_LAZY = {"DelayInjector": "delay", "MetricsLog": "telemetry", "RunReport": "telemetry", "evaluate_objective": "telemetry"}


def __getattr__(name):
    if name in _DRIVER_NAMES:
        return importlib.import_module(f"{__name__}.{name}")
    if name in _LAZY:
        return getattr(importlib.import_module(f"{__name__}.{_LAZY[name]}"), name)
    if name == "DRIVERS":
        return {n: importlib.import_module(f"{__name__}.{n}").run for n in _DRIVER_NAMES}
    raise AttributeError(name)


def run(cfg, points=None, scheduler=None):
    from .common import host_math_threads
    if cfg.algorithm not in _DRIVER_NAMES:
        raise ValueError(f"unknown algorithm {cfg.algorithm!r}")
    driver = importlib.import_module(f"{__name__}.{cfg.algorithm}").run
    with host_math_threads():          # the driver thread does small-vector math only; see host_math_threads
        return driver(cfg, points, scheduler)
