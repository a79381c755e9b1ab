"""Data preparation on an elastic standalone cluster: a shuffle that spills, a slow stage that makes the application grow,
an idle phase that makes it shrink again.

    python examples/data_prep_on_a_cluster.py            # CPU only; ~15 s

What it shows (all of it is the Spark substrate ASYNC's drivers rely on before the optimisation starts — `repartition`,
`reduceByKey`, `sortBy` in the reference's data-prep lines, SparkASGDThread.scala:60-90):

* `spark.shuffle.spill.*`: map and reduce sides of the shuffle push their records through a spillable sorter
  (`storage/external_sorter.py`); the spill counters come back in the task metrics;
* `spark.dynamicAllocation.*`: the application starts with one executor; backlogged tasks raise the executor target, which
  is applied at the next job boundary (a partition lives on slot p % n, so the set never changes under a running job);
  after `executorIdleTimeout` without task activity the application gives the executors back (`core/allocation.py`);
* stratified sampling, k-fold splits and the text round trip of labeled points (`MLUtils`).
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from asyncframework_b200 import AsyncSession  # noqa: E402
from asyncframework_b200.conf import AsyncConf  # noqa: E402
from asyncframework_b200.data.libsvm import MLUtils  # noqa: E402
from asyncframework_b200.deploy import LocalStandaloneCluster  # noqa: E402


def main() -> None:
    with LocalStandaloneCluster(num_workers=4) as cluster:
        conf = (AsyncConf(load_defaults=False).set_master(cluster.master_url).set_app_name("data-prep")
                .set("spark.executor.instances", "1")
                .set("spark.dynamicAllocation.enabled", "true").set("spark.dynamicAllocation.maxExecutors", "4")
                .set("spark.dynamicAllocation.schedulerBacklogTimeout", "100ms")
                .set("spark.dynamicAllocation.sustainedSchedulerBacklogTimeout", "100ms")
                .set("spark.dynamicAllocation.executorIdleTimeout", "2s")
                .set("spark.shuffle.spill.numElementsForceSpillThreshold", "500"))
        with AsyncSession(conf=conf) as sc:
            spills = []
            sc.scheduler.listeners.add(lambda ev: spills.append(ev["metrics"].get("spills", 0)) if ev["event"] == "task_end" else None)
            am = sc.scheduler.allocation_manager
            print(f"executors at start: {sc.scheduler.num_executors}")

            # 1. a shuffle whose sorters spill: token counts of a synthetic corpus, then the top tokens by count
            words = sc.parallelize(range(40_000), 8).map(lambda i: (f"tok{(i * 7919) % 997}", 1))
            counts = words.reduce_by_key(lambda a, b: a + b, 4)
            top = counts.sort_by(lambda kv: (-kv[1], kv[0]), num_partitions=2).take(3)
            print(f"distinct tokens: {counts.count()}, top: {top}, sorter spills reported by tasks: {sum(spills)}")

            # 2. a slow stage: by now the backlog of the jobs above has grown the application; the stage spreads over the executors
            def featurise(i):
                time.sleep(0.01)
                return (i % 2, float(i % 13))
            slow = sc.parallelize(range(400), 8).map(featurise)
            first = len(set(slow.map_partitions(lambda it: [os.getpid()]).collect()))
            second = len(set(slow.map_partitions(lambda it: [os.getpid()]).collect()))
            print(f"executor processes used by the slow job: {first} -> {second} (target {am.target})")

            # 3. stratified sample, 3-fold split, text round trip
            strat = slow.sample_by_key(False, {0: 0.5, 1: 0.1}, seed=3)
            folds = MLUtils.k_fold(slow, 3, seed=11)
            sizes = [(tr.count(), va.count()) for tr, va in folds]
            print(f"stratified sample: {strat.count_by_key()}, 3-fold sizes (train, validation): {sizes}")

            # 4. idle: the executors go back
            t0 = time.time()
            while sc.scheduler.num_executors > 1 and time.time() - t0 < 30:
                time.sleep(0.2)
            events = [(h["from"], h["to"]) for h in am.history if h["event"] == "applied"]
            print(f"executors after the idle timeout: {sc.scheduler.num_executors}; allocation changes applied: {events}")


if __name__ == "__main__":
    main()
