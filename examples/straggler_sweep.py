#!/usr/bin/env python
"""The reference's headline experiment in miniature: ASAGA vs synchronous SAGA while worker 0 is slowed down by a
multiple (`coeff`, the "delay intensity") of the mean task time — and the production-cluster mode (`coeff = -1`: 25 % of
the workers are stragglers).  Uses the five-driver API (`algo.run(AsyncConfig)`), CPU threads backend.

    python examples/straggler_sweep.py [--rows 20000] [--workers 4] [--master local-cluster[4]] [--json out.json]

``--master`` runs the same sweep with the workers as executor PROCESSES (``local-cluster[n]``) or on a standalone cluster
(``spark://host:port``): the delay injector then sleeps inside the executor that owns the slow partition.

Reference: README.md:52-60 (delay intensity), SparkASAGAThread.scala:300-345 (delay injector), paper figures 5-7.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))   # run from a source checkout

from asyncframework_b200 import algo  # noqa: E402
from asyncframework_b200.config import AsyncConfig  # noqa: E402


def main(rows=20000, d=50, P=4, updates=240, master="", json_path=None):
    print(f"{'algorithm':<10} {'delay':>6} {'elapsed ms':>11} {'avg wait ms':>12} {'accepted':>9} {'objective':>10}")
    out = {}
    for coeff in (0.0, 1.0, 3.0, -1.0):
        for name, iters in (("asaga", updates), ("saga_sync", updates // P)):
            cfg = AsyncConfig(algorithm=name, num_features=d, num_rows=rows, num_partitions=P, num_iterations=iters, step_size=0.05,
                              batch_rate=0.1, bucket_ratio=0.5, coeff=coeff, loss="least_squares", synthetic="linear", seed=42,
                              printer_freq=max(1, iters // 4), warmup_updates_per_worker=5, backend="cpu", quiet=True, master=master)
            rep = algo.run(cfg)
            obj = rep.objective[-1][1] if rep.objective else float("nan")
            out[(name, coeff)] = rep
            print(f"{name:<10} {coeff:>6} {rep.elapsed_ms:>11.0f} {rep.avg_waiting_ms:>12.2f} {rep.accepted:>9} {obj:>10.5f}")
    if json_path:
        import json
        with open(json_path, "w") as f:
            json.dump({"master": master or f"local[{P}]", "rows": rows, "features": d, "workers": P,
                       "runs": [{"algorithm": n, "delay": c, "elapsed_ms": round(r.elapsed_ms, 1), "accepted": r.accepted,
                                 "avg_wait_ms": round(r.avg_waiting_ms, 3), "objective": r.objective[-1][1] if r.objective else None}
                                for (n, c), r in out.items()]}, f, indent=1)
    return out


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=20000)
    ap.add_argument("--workers", type=int, default=4)
    ap.add_argument("--master", default="")
    ap.add_argument("--json", default=None)
    a = ap.parse_args()
    main(rows=a.rows, P=a.workers, master=a.master, json_path=a.json)
