"""``spark-submit`` analogue: launch a driver on N GPUs of this node.

    python -m asyncframework_b200.launch --gpus 8 asgd <13 positional args> [--key=value ...]
    python -m asyncframework_b200.launch --cpu-workers 2 sgd_sync <11 positional args> --backend=gloo

One process per GPU via ``torch.distributed.run`` (rendezvous on 127.0.0.1), rank 0 = parameter server
+ worker.  With ``--gpus 1`` (or no GPU) the driver runs in this process.
"""
from __future__ import annotations

import argparse
import os
import subprocess
import sys

DRIVERS = ("asgd", "asaga", "sgd_sync", "saga_sync", "sgd_mllib")


def main(argv=None) -> int:
    ap = argparse.ArgumentParser(prog="asyncframework_b200.launch", add_help=True)
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--cpu-workers", type=int, default=0, help="multi-process CPU run over gloo")
    ap.add_argument("--master-port", type=int, default=29555)
    ap.add_argument("driver", choices=DRIVERS)
    ap.add_argument("args", nargs=argparse.REMAINDER)
    ns = ap.parse_args(argv)
    nproc = ns.cpu_workers or ns.gpus
    mod = f"asyncframework_b200.algo.{ns.driver}"
    extra = list(ns.args)
    if ns.gpus > 1 and not any(a.startswith("--backend") for a in extra):
        extra.append("--backend=fused")
    if ns.cpu_workers and not any(a.startswith("--backend") for a in extra):
        extra.append("--backend=gloo")
    if nproc <= 1:
        cmd = [sys.executable, "-m", mod] + extra
    else:
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
               "--master-addr", "127.0.0.1", "--master-port", str(ns.master_port), "-m", mod] + extra
    return subprocess.call(cmd, env=dict(os.environ))


if __name__ == "__main__":
    sys.exit(main())
