"""``spark-submit`` analogue: resolve configuration, then run a driver or a user script.

Reference: ``SparkSubmit`` + ``SparkSubmitArguments`` (core/deploy/SparkSubmit.scala:125,170,245,817;
SparkSubmitArguments.scala): options ``--master``, ``--name``, ``--class``, ``--conf k=v``, ``--properties-file``,
``--executor-memory``, ``--driver-memory``, ``--executor-cores``, ``--total-executor-cores``, ``--num-executors``,
then the application and its arguments.  Precedence (highest first): command-line flags, ``--conf`` pairs, the
properties file, ``conf/async-defaults.conf``, built-in defaults.  README.md:46 launches the paper's runs as::

    spark-submit --class org.apache.async.SparkASGDThread --executor-memory 15g --driver-memory 20g <jar> <13 args>

which here is::

    python -m asyncframework_b200.submit --class org.apache.async.SparkASGDThread --executor-memory 15g \\
        --driver-memory 20g --master spark://host:7077  <13 args>

``--class`` takes a reference class name (mapped to the matching driver), a driver short name (``asgd`` …) or is
omitted when the application is a ``.py`` script — the script then finds the resolved configuration through
``AsyncConf()`` (exported as ``AFB_CONF_*`` environment variables, the role of ``-Dspark.*`` system properties).
``--gpus N`` launches the driver on N GPUs of this node through ``asyncframework_b200.launch`` (one process per GPU,
fused backend).
"""
from __future__ import annotations

import argparse
import os
import runpy
import sys
from typing import List, Optional

from .conf import AsyncConf

CLASS_TO_DRIVER = {
    "org.apache.async.SparkASGDThread": "asgd", "SparkASGDThread": "asgd",
    "org.apache.async.SparkASGDSync": "sgd_sync", "SparkASGDSync": "sgd_sync",
    "org.apache.async.SparkASAGAThread": "asaga", "SparkASAGAThread": "asaga",
    "org.apache.async.SparkASAGASync": "saga_sync", "SparkASAGASync": "saga_sync",
    "org.apache.async.SparkSGDMLLIB": "sgd_mllib", "SparkSGDMLLIB": "sgd_mllib",
}
DRIVERS = ("asgd", "asaga", "sgd_sync", "saga_sync", "sgd_mllib")


def build_parser() -> argparse.ArgumentParser:
    ap = argparse.ArgumentParser(prog="asyncframework_b200.submit", description=__doc__.split("\n\n")[0])
    ap.add_argument("--master")
    ap.add_argument("--name")
    ap.add_argument("--class", dest="klass")
    ap.add_argument("--conf", action="append", default=[], metavar="KEY=VALUE")
    ap.add_argument("--properties-file")
    ap.add_argument("--executor-memory")
    ap.add_argument("--driver-memory")
    ap.add_argument("--executor-cores", type=int)
    ap.add_argument("--total-executor-cores", type=int)
    ap.add_argument("--num-executors", type=int)
    ap.add_argument("--gpus", type=int, default=0, help="run the driver on N GPUs of this node (fused backend)")
    ap.add_argument("--verbose", "-v", action="store_true")
    ap.add_argument("app", nargs=argparse.REMAINDER, help="[script.py | jar placeholder] application arguments")
    return ap


def resolve_conf(ns: argparse.Namespace) -> AsyncConf:
    conf = AsyncConf(load_defaults=False)
    for kv in ns.conf:                                   # --conf wins over files
        if "=" not in kv:
            raise SystemExit(f"--conf expects KEY=VALUE, got {kv!r}")
        k, v = kv.split("=", 1)
        conf.set(k.strip(), v.strip())
    if ns.properties_file:
        conf.load_properties_file(ns.properties_file)     # set_if_missing semantics
    else:
        d = AsyncConf.defaults_file()
        if d:
            conf.load_properties_file(d)
    flags = {"spark.master": ns.master, "spark.app.name": ns.name, "spark.executor.memory": ns.executor_memory,
             "spark.driver.memory": ns.driver_memory, "spark.executor.cores": ns.executor_cores,
             "spark.cores.max": ns.total_executor_cores, "spark.executor.instances": ns.num_executors}
    for k, v in flags.items():
        if v is not None:
            conf.set(k, v)
    if ns.total_executor_cores and not ns.num_executors and not conf.contains("spark.executor.instances"):
        conf.set("spark.executor.instances", max(1, ns.total_executor_cores // conf.get("spark.executor.cores")))
    return conf.validate()


def export_conf(conf: AsyncConf, env=None) -> dict:
    env = os.environ if env is None else env
    for k, v in conf.get_all():
        env["AFB_CONF_" + k.replace("-", "__").replace(".", "_")] = v
    return env


def main(argv: Optional[List[str]] = None) -> int:
    ns = build_parser().parse_args(argv)
    conf = resolve_conf(ns)
    app = list(ns.app)
    if app and app[0] == "--":
        app = app[1:]
    if ns.verbose:
        print("Resolved properties:\n" + conf.to_debug_string(), file=sys.stderr)
        for w in conf.warnings:
            print("Warning: " + w, file=sys.stderr)
    driver = None
    if ns.klass:
        driver = CLASS_TO_DRIVER.get(ns.klass, ns.klass if ns.klass in DRIVERS else None)
        if driver is None:
            raise SystemExit(f"--class {ns.klass!r}: not a known driver (have {sorted(set(CLASS_TO_DRIVER) | set(DRIVERS))})")
        if app and app[0].endswith(".jar"):               # README command lines carry the application jar: ignore it
            app = app[1:]
    elif app and app[0] in DRIVERS:
        driver, app = app[0], app[1:]
    export_conf(conf)
    if driver is not None:
        master = conf.get("spark.master")
        if conf.contains("spark.master") and not any(a.startswith("--master") for a in app):
            app = app + [f"--master={master}"]
        if ns.gpus and ns.gpus > 1:
            from .launch import main as launch_main
            return launch_main(["--gpus", str(ns.gpus), driver] + app)
        import importlib
        mod = importlib.import_module(f"asyncframework_b200.algo.{driver}")
        mod.main(app)
        return 0
    if not app:
        raise SystemExit("nothing to run: give --class <driver> or a .py script")
    script, sys.argv = app[0], app
    runpy.run_path(script, run_name="__main__")
    return 0


if __name__ == "__main__":
    sys.exit(main())
