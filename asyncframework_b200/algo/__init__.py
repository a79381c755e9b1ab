"""The five reference drivers (SparkASGDThread, SparkASGDSync, SparkASAGAThread, SparkASAGASync,
SparkSGDMLLIB) as Python entry points with the reference's positional CLIs."""
from . import asaga, asgd, saga_sync, sgd_mllib, sgd_sync  # noqa: F401
from .delay import DelayInjector
from .telemetry import MetricsLog, RunReport, evaluate_objective

DRIVERS = {"asgd": asgd.run, "asaga": asaga.run, "sgd_sync": sgd_sync.run, "saga_sync": saga_sync.run,
           "sgd_mllib": sgd_mllib.run}


def run(cfg, points=None, scheduler=None):
    from .common import host_math_threads
    with host_math_threads():          # the driver thread does small-vector math only; see host_math_threads
        return DRIVERS[cfg.algorithm](cfg, points, scheduler) if cfg.algorithm != "sgd_mllib" else sgd_mllib.run(cfg, points, scheduler)
