"""A standalone cluster on one box, end to end: Master + 2 Workers (real processes for the executors, real TCP), an
application that trains ASGD through the closure API on it, the live UI, a killed executor that the Master replaces.

    python examples/standalone_cluster.py            # CPU only; ~20 s

Reference: the paper's experiments ran on a 1-master + 8-worker standalone cluster (README.md:56); this is that
deployment shape in miniature (``LocalSparkCluster``, core/deploy/LocalSparkCluster.scala:49-70).
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from asyncframework_b200 import AsyncSession  # noqa: E402
from asyncframework_b200.conf import AsyncConf  # noqa: E402
from asyncframework_b200 import barriers  # noqa: E402
from asyncframework_b200.optim import make_gradient  # noqa: E402
from asyncframework_b200.data import LinearDataGenerator  # noqa: E402
from asyncframework_b200.deploy import LocalStandaloneCluster  # noqa: E402


def main(quiet: bool = False) -> float:
    say = (lambda *a: None) if quiet else print
    killed = False
    with LocalStandaloneCluster(num_workers=2, cores_per_worker=2, memory_per_worker=2048) as cluster:
        conf = (AsyncConf(load_defaults=False).set_master(cluster.master_url).set_app_name("asgd-on-standalone")
                .set("spark.io.compression.codec", "lz4").set("spark.ui.enabled", "true").set("spark.ui.port", 0))
        with AsyncSession(conf=conf) as sc:
            say("master", cluster.master_url, "| UI", sc.ui.url, "| executors", sc.scheduler.num_executors)
            P, d, n, b, gamma, taw, beta = 4, 10, 4000, 0.2, 0.5, 8, 0.5
            points = LinearDataGenerator.generate_linear_dataset(sc.scheduler, n, d, 0.05, P, seed=7).zip_with_index().cache()
            N = points.count()
            AC = sc.async_context(P)
            AC.blocking = False
            w = torch.zeros(d, dtype=torch.float64)
            grad = make_gradient("least_squares")

            def objective():
                AC.blocking = True
                try:
                    return points.map_batch(lambda s, w=w.clone(): float(((s.matvec(w) - s.y) ** 2).sum())).reduce(lambda p, q: p + q) / N
                finally:
                    AC.blocking = False
            first = objective()
            k = submitted = 0
            while k < 120:
                if submitted == 0 or AC.stat.available_workers() >= max(1, int(P * beta)):
                    a = sc.broadcast(w.clone())                            # torrent broadcast: only the handle is in the task binary
                    sel = points.async_barrier(barriers.bucket(P, beta), AC.stat)
                    sel.sample(False, b, 42 + submitted + 1).map_batch(lambda s, a=a: grad.batch(s, a.value)[0]) \
                       .async_reduce(lambda x, y: x.add_(y), AC)
                    submitted += 1
                info = AC.collect_all(timeout=0.5)
                while info is not None:
                    if info.staleness <= taw:
                        w.add_(info.data, alpha=-gamma / (k // P + 1) ** 0.5 / (b * N / P))
                        k += 1
                    info = AC.collect_all(timeout=0.0) if AC.has_next() else None
                if k >= 40 and not killed:
                    killed = True
                    say("killing executor 1 ...")
                    sc.scheduler.kill_executor(1)                          # the Master re-fills the slot, lost tasks are re-run
            last = objective()
            say(f"objective {first:.4f} -> {last:.4f} after {k} accepted updates; executors recovered: "
                f"{sc.scheduler.executors_recovered}; jobs: {sc.status_store.summary()['jobs_by_status']}")
            return last / first


if __name__ == "__main__":
    main()
