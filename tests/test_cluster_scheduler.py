"""`local-cluster[n]`: executors in separate processes, task binaries shipped with cloudpickle
(the DistributedSuite / LocalClusterSparkContext analogue)."""
import os
import time

import numpy as np
import pytest
import torch

from asyncframework_b200 import AsyncSession, algo
from asyncframework_b200.config import AsyncConfig
from asyncframework_b200.core import AsyncContext, ClusterScheduler, JobFailed
from asyncframework_b200.data import LinearDataGenerator, ShardedDataset


@pytest.fixture(scope="module")
def cluster():
    s = ClusterScheduler(2, max_failures=3, block_first_async_job=False)
    yield s
    s.stop()


def test_closures_run_in_other_processes(cluster):
    ds = ShardedDataset.from_records(cluster, list(range(20)), 4)
    pids = set(ds.map(lambda x: os.getpid()).collect())
    assert os.getpid() not in pids and len(pids) == 2                     # two executor processes did the work
    k = 7
    assert ds.map(lambda x: x * k).reduce(lambda a, b: a + b) == 7 * 190  # captured variable travels with the closure
    assert ds.aggregate(lambda: 0, lambda a, x: a + x, lambda a, b: a + b) == 190
    assert ds.count() == 20 and sorted(ds.collect()) == list(range(20))


def test_executor_side_cache_survives_across_jobs(cluster):
    ds = ShardedDataset.from_records(cluster, list(range(8)), 2).map(lambda x: (x, time.perf_counter_ns())).cache()
    first = dict(ds.collect())
    second = dict(ds.collect())
    assert first == second                                                 # second job hit the executors' block store


def test_batch_level_ops_and_async_reduce_over_processes(cluster):
    X, y, w = LinearDataGenerator.generate_linear_arrays(400, 5, seed=3)
    ds = ShardedDataset.from_arrays(cluster, X, y, 4).zip_with_index().cache()
    assert ds.partition_sizes() == [100, 100, 100, 100]
    ctx = AsyncContext(4, cluster)
    ctx.set_record_stat(True)
    wt = torch.zeros(5, dtype=torch.float64)
    from asyncframework_b200.optim import make_gradient
    grad = make_gradient("least_squares")
    ds.sample(False, 0.5, 11).map_batch(lambda s: grad.batch(s, wt)[0]).async_reduce(lambda a, b: a + b, ctx)
    got = [ctx.collect_all(timeout=30) for _ in range(4)]
    assert all(g is not None for g in got) and sorted(g.worker_id for g in got) == [0, 1, 2, 3]
    assert ctx.current_time == 4 and ctx.stat.available_workers() == 4
    total = sum(g.data for g in got)
    from asyncframework_b200.core.sampler import bernoulli_mask
    Xall = torch.cat([s.X[:, :5] for s in ds.shards()])
    yall = torch.cat([s.y for s in ds.shards()])
    gall = torch.cat([s.gidx for s in ds.shards()])
    m = torch.from_numpy(bernoulli_mask(11, gall, 0.5))
    ref = Xall[m].T @ (Xall[m] @ wt - yall[m])
    assert torch.allclose(total, ref, atol=1e-9)


def test_task_retry_and_job_failure_over_processes(cluster):
    ds = ShardedDataset.from_records(cluster, list(range(6)), 2)

    def flaky(tctx):
        if tctx.partition_id == 1 and tctx.attempt < 2:
            raise RuntimeError("injected")
    cluster.fault_injector = flaky
    try:
        assert ds.count() == 6                                             # retried on the same executor, then succeeded
        cluster.fault_injector = lambda t: (_ for _ in ()).throw(RuntimeError("always")) if t.partition_id == 0 else None
        with pytest.raises(JobFailed):
            ds.count()
    finally:
        cluster.fault_injector = None
    assert ds.count() == 6                                                 # the cluster is still healthy


def test_driver_runs_on_local_cluster_and_matches_inprocess():
    cfg = AsyncConfig(algorithm="sgd_sync", synthetic="linear", num_rows=600, num_features=8, num_partitions=2,
                      num_iterations=15, step_size=0.8, batch_rate=0.3, printer_freq=5, backend="cpu", eval_objective=False)
    with AsyncSession("local-cluster[2,1,1024]") as sc:
        assert isinstance(sc.scheduler, ClusterScheduler)
        ds = LinearDataGenerator.generate_linear_dataset(sc.scheduler, 600, 8, 0.1, 2, seed=42)
        rep_cluster = algo.run(cfg, points=ds, scheduler=sc.scheduler)
    with AsyncSession("local[2]") as sc:
        ds = LinearDataGenerator.generate_linear_dataset(sc.scheduler, 600, 8, 0.1, 2, seed=42)
        rep_local = algo.run(cfg, points=ds, scheduler=sc.scheduler)
    assert torch.allclose(rep_cluster.final_weights, rep_local.final_weights, rtol=1e-12, atol=1e-14)


def test_lost_executor_fails_the_job_and_frees_the_worker():
    s = ClusterScheduler(2, block_first_async_job=False)
    try:
        ds = ShardedDataset.from_records(s, list(range(4)), 2)
        ctx = AsyncContext(2, s)

        def slow(p, it):
            if p == 0:
                time.sleep(30)
            return it
        ds.map_partitions_with_index(slow).async_reduce(lambda a, b: a + b, ctx)
        r = ctx.collect_all(timeout=30)
        assert r.worker_id == 1
        assert not ctx.stat[0].available
        s.kill_executor(0)                                                 # executor 0 dies mid-task
        t0 = time.time()
        while not ctx.stat[0].available and time.time() - t0 < 10:
            time.sleep(0.05)
        assert ctx.stat[0].available                                       # its worker is not stranded (quirk B.15 fixed)
    finally:
        s.stop()


def test_task_binary_is_small_after_the_first_job(cluster):
    """Source data travels to an executor once; later jobs over the same dataset ship only the lineage + closure
    (the LinearRegressionClusterSuite "task size should be small" check)."""
    rng = np.random.RandomState(0)
    X = rng.randn(4000, 64)
    y = rng.randn(4000)
    sizes = []
    cluster.listeners.add(lambda ev: sizes.append(ev["task_binary_bytes"]) if ev.get("event") == "job_start" else None)
    ds = ShardedDataset.from_arrays(cluster, X, y, 4).zip_with_index().cache()      # zip_with_index runs the first job
    n1 = ds.count()
    first = max(sizes)
    assert sum(1 for b in sizes if b > 100_000) == 1                               # ... and only that one carried the data
    w = torch.zeros(64, dtype=torch.float64)
    tot = ds.map_batch(lambda s, w=w: float((s.matvec(w) - s.y).pow(2).sum())).reduce(lambda a, b: a + b)
    second = cluster.last_task_binary_bytes
    assert n1 == 4000 and abs(tot - float((y ** 2).sum())) < 1e-6
    assert first > 4000 * 64 * 8 and second < 20_000, (first, second)
    # a different dataset over new data ships its own block once
    ds2 = ShardedDataset.from_records(cluster, list(range(50_000)), 4)
    assert ds2.count() == 50_000
    big = cluster.last_task_binary_bytes
    assert ds2.map(lambda x: x % 7).reduce(lambda a, b: a + b) == sum(i % 7 for i in range(50_000))
    assert cluster.last_task_binary_bytes < big // 10


def test_cancel_job_over_executor_processes(cluster):
    from asyncframework_b200.core import JobCancelled
    ds = ShardedDataset.from_records(cluster, list(range(16)), 8)

    def slow(tctx, it):
        import time as _t
        _t.sleep(0.4)
        return sum(it)

    w = cluster.submit_job(ds, slow)                 # 8 tasks x 0.4 s on 2 executors = 1.6 s if it ran to completion
    time.sleep(0.1)
    t0 = time.perf_counter()
    assert cluster.cancel_job(w.job_id) and not cluster.cancel_job(w.job_id)
    with pytest.raises(JobCancelled):
        w.wait(5)
    assert ds.map(lambda x: x + 1).reduce(lambda a, b: a + b) == 136      # queued tasks of the dropped job returned at once
    assert time.perf_counter() - t0 < 1.4


def test_concurrent_job_submission_from_many_threads(cluster):
    """ThreadingSuite analogue: several driver threads submit jobs over shared and private datasets at once; every job gets its
    own results, source blocks are shipped exactly once per executor even when first uses race."""
    import threading
    shared = ShardedDataset.from_records(cluster, list(range(2000)), 4)
    errors, sums = [], {}

    def driver(i):
        try:
            mine = ShardedDataset.from_records(cluster, list(range(i, i + 500)), 3)
            for r in range(4):
                a = shared.map(lambda x, i=i, r=r: x * (i + 1) + r).reduce(lambda p, q: p + q)
                b = mine.map(lambda x: x * x).reduce(lambda p, q: p + q)
                assert a == (i + 1) * sum(range(2000)) + r * 2000, (i, r, a)
                assert b == sum(x * x for x in range(i, i + 500)), (i, r, b)
            sums[i] = True
        except BaseException as exc:  # noqa: BLE001
            errors.append(repr(exc))

    threads = [threading.Thread(target=driver, args=(i,)) for i in range(6)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(60)
    assert not errors, errors[:3]
    assert len(sums) == 6 and not any(t.is_alive() for t in threads)


def test_accumulators_over_executor_processes(cluster):
    from asyncframework_b200.core import Accumulator
    seen, total = Accumulator(0, "seen"), Accumulator(0.0, "total")
    ds = ShardedDataset.from_records(cluster, list(range(60)), 4)

    def f(x):
        seen.add(1)
        total.add(x * 0.5)
        return x
    assert sorted(ds.map(f).collect()) == list(range(60))
    assert seen.value == 60 and total.value == 0.5 * sum(range(60))
    assert ds.map(f).count() == 60 and seen.value == 120          # a second job keeps accumulating
    with pytest.raises(Exception):
        ds.map(lambda x: seen.value).collect()                     # tasks may only add: reading on an executor is an error


def test_key_value_shuffles_agree_across_executor_processes(cluster):
    """Each executor process salts hash() differently; the shuffle partitioner must not depend on it."""
    words = ShardedDataset.from_records(cluster, ("a b a c b a d e f a " * 20).split(), 4)
    counts = dict(words.map(lambda w: (w, 1)).reduce_by_key(lambda x, y: x + y, 3).collect())
    assert counts == {"a": 80, "b": 40, "c": 20, "d": 20, "e": 20, "f": 20}
    assert sorted(words.distinct().collect()) == list("abcdef")
    a = ShardedDataset.from_records(cluster, list(range(12)), 3)
    assert sorted(a.union(a.map(lambda x: x + 100)).collect()) == list(range(12)) + list(range(100, 112))
    assert a.zip(a.map(lambda x: -x)).collect() == [(i, -i) for i in range(12)]


# ------------------------------------------------------------------------------------------------------------------
# runtime around the executors: block managers, torrent broadcast, indirect results, executor recovery
def test_cached_partitions_are_blocks_of_the_executor_block_managers(cluster):
    from asyncframework_b200.data.dataset import ShardedDataset
    ds = ShardedDataset.from_records(cluster, list(range(40)), 4).map(lambda x: x * 2).cache()
    assert ds.count() == 40
    st = cluster.storage_status()
    rdd_blocks = {b for m, s in st.items() for b in s["blocks"] if b.startswith(f"rdd_{ds.ds_id}_")}
    assert rdd_blocks == {f"rdd_{ds.ds_id}_{p}r" for p in range(4)}
    assert all(s["mem_used"] > 0 for m, s in st.items() if m.executor_id != "driver")
    assert cluster.env.bm_master.remove_rdd(ds.ds_id) == 4                    # unpersist
    assert ds.map(lambda x: x + 1).reduce(lambda a, b: a + b) == sum(2 * x + 1 for x in range(40))   # recomputed


def test_torrent_broadcast_and_async_broadcast_history_in_executor_processes():
    import torch
    from asyncframework_b200.core.session import AsyncSession
    from asyncframework_b200.core.torrent_broadcast import TorrentBroadcast
    with AsyncSession("local-cluster[2]") as sc:
        w = torch.arange(200_000, dtype=torch.float64)
        bc = sc.broadcast(w)
        assert isinstance(bc, TorrentBroadcast)
        ds = sc.parallelize(range(8), 4)
        out = ds.map(lambda i: float(bc.value[i])).collect()
        assert out == [float(i) for i in range(8)]
        assert sc.scheduler.last_task_binary_bytes < 50_000                 # the 1.6 MB tensor did not ride in the task binary
        ctx = sc.async_context(4)
        v0 = sc.async_broadcast(ctx, torch.zeros(4))
        v1 = sc.async_broadcast(ctx, torch.ones(4))
        i0 = v0.index
        got = ds.map(lambda i: (float(v1.value[0]), float(v1.value_at(i0)[0]), v1.index)).collect()
        assert got == [(1.0, 0.0, v1.index)] * 8                            # current version AND an older one, by id
        bc.destroy()


def test_large_results_take_the_indirect_path_and_max_result_size_aborts():
    import numpy as np
    from asyncframework_b200.core.cluster import ClusterScheduler
    from asyncframework_b200.core.env import RuntimeLimits
    from asyncframework_b200.core.scheduler import JobFailed
    from asyncframework_b200.data.dataset import ShardedDataset
    lim = RuntimeLimits(max_direct_result_size=64 * 1024, max_result_size=3 << 20)
    with ClusterScheduler(2, limits=lim) as s:
        ds = ShardedDataset.from_records(s, list(range(4)), 4)
        small = ds.map(lambda i: i).collect()
        assert small == [0, 1, 2, 3] and s.indirect_results == 0
        big = ds.map_partitions(lambda it: [np.full(40_000, float(next(iter(it))))]).collect()    # 320 KB per task
        assert [float(a[0]) for a in big] == [0.0, 1.0, 2.0, 3.0] and s.indirect_results == 4
        time.sleep(0.3)
        st = s.storage_status()
        assert not any(b.startswith("taskresult_") for m in st.values() for b in m["blocks"])      # fetched, then deleted
        with pytest.raises(JobFailed, match="max_result_size"):
            ds.map_partitions(lambda it: [np.zeros(200_000)]).collect()                              # 4 x 1.6 MB > 3 MiB


def test_lost_executor_is_replaced_and_its_tasks_are_resubmitted():
    from asyncframework_b200.core.cluster import ClusterScheduler
    from asyncframework_b200.data.dataset import ShardedDataset
    events = []
    with ClusterScheduler(2, recover_lost_executors=True) as s:
        s.listeners.add(lambda e: events.append(e["event"]))
        ds = ShardedDataset.from_records(s, list(range(20)), 4)
        assert ds.count() == 20                                           # executors now hold the source blocks

        def slow(tctx, it):
            time.sleep(1.0)
            return sum(it)
        w = s.submit_job(ds, slow, result_handler=(res := {}).__setitem__)
        time.sleep(0.3)
        s.kill_executor(1)                                                # dies mid-task; partitions 1 and 3 are lost
        assert w.wait(30)
        assert res == {p: sum(range(5 * p, 5 * p + 5)) for p in range(4)}
        assert s.executors_recovered == 1
        assert "executor_lost" in events and "executor_added" in events and "task_resubmitted" in events
        assert ds.map(lambda x: x + 1).reduce(lambda a, b: a + b) == 210   # the replacement got the data re-shipped


def test_speculative_copy_of_a_straggling_task_wins():
    """TaskSetManagerSuite "speculative tasks": one executor is slow for one partition; a copy on the other executor
    delivers first and the straggler's late result is ignored."""
    from asyncframework_b200.core.cluster import ClusterScheduler
    from asyncframework_b200.data.dataset import ShardedDataset
    events = []
    with ClusterScheduler(2, speculation=True, speculation_multiplier=2.0, speculation_quantile=0.5,
                          speculation_interval=0.05, speculation_min_ms=50.0) as s:
        s.listeners.add(events.append)
        ds = ShardedDataset.from_records(s, list(range(40)), 4)
        assert ds.count() == 40

        def work(tctx, it):
            if tctx.partition_id == 3 and tctx.attempt < 1000:
                time.sleep(5.0)                    # the straggler: only the ORIGINAL attempt is slow
            else:
                time.sleep(0.1)
            return sum(it)
        t0 = time.time()
        res = {}
        w = s.submit_job(ds, work, result_handler=res.__setitem__)
        assert w.wait(20)
        assert time.time() - t0 < 4.0                                         # did not wait for the 5 s straggler
        assert res == {p: sum(range(10 * p, 10 * p + 10)) for p in range(4)}
        assert s.speculated_tasks >= 1 and s.speculative_wins >= 1             # (a loaded CI box may speculate another task too)
        spec = {e["partition"]: e for e in events if e["event"] == "task_speculated"}
        assert spec[3]["executor"] == 0                                        # home of partition 3 is executor 1
        time.sleep(0.2)
        assert len(res) == 4                                                   # nothing delivered twice
        assert s.speculative_kills_sent >= 1


def test_losing_copy_of_a_speculated_task_is_interrupted():
    """TaskSetManager.scala:736 / Executor.killTask: once the speculative copy has delivered, the straggling original is
    interrupted — its python loop stops within milliseconds instead of burning the slot to the end."""
    from asyncframework_b200.core.cluster import ClusterScheduler
    from asyncframework_b200.data.dataset import ShardedDataset
    events = []
    with ClusterScheduler(2, speculation=True, speculation_multiplier=2.0, speculation_quantile=0.5,
                          speculation_interval=0.05, speculation_min_ms=50.0) as s:
        s.listeners.add(events.append)
        ds = ShardedDataset.from_records(s, list(range(40)), 4)
        assert ds.count() == 40

        def work(tctx, it):
            if tctx.partition_id == 3 and tctx.attempt < 1000:
                t_end = time.time() + 20.0         # a busy python loop for 20 s unless somebody interrupts it
                while time.time() < t_end:
                    time.sleep(0.005)
            else:
                time.sleep(0.1)
            return sum(it)
        res = {}
        w = s.submit_job(ds, work, result_handler=res.__setitem__)
        assert w.wait(20)
        t0 = time.time()
        while not any(e["event"] == "task_killed" for e in events) and time.time() - t0 < 5.0:
            time.sleep(0.02)
        killed = [e for e in events if e["event"] == "task_killed"]
        assert killed and killed[0]["partition"] == 3 and killed[0]["executor"] == 1 and killed[0]["reason"] == "interrupted"
        assert killed[0]["ran_ms"] < 10_000                                    # nowhere near the 20 s it wanted
        # the freed slot is usable at once: a follow-up job on the same executors returns promptly
        t1 = time.time()
        assert ds.map(lambda x: 2 * x).reduce(lambda a, b: a + b) == 2 * sum(range(40))
        assert time.time() - t1 < 5.0


def test_failing_executor_is_blacklisted_and_its_partitions_move():
    """BlacklistTrackerSuite: after two failed tasks executor 1 is excluded; retries and later jobs run on executor 0."""
    from asyncframework_b200.core.cluster import ClusterScheduler
    from asyncframework_b200.data.dataset import ShardedDataset
    events = []
    with ClusterScheduler(2, max_failures=4, blacklist=True, blacklist_max_failures=2, blacklist_timeout=3.0) as s:
        s.listeners.add(events.append)
        ds = ShardedDataset.from_records(s, list(range(40)), 4)

        def flaky(tctx, it):
            if tctx.executor_id == 1:
                raise RuntimeError("bad disk on executor 1")
            return (tctx.executor_id, sum(it))
        out = s.run_job(ds, flaky)
        assert [o[1] for o in out] == [sum(range(10 * p, 10 * p + 10)) for p in range(4)]
        assert {o[0] for o in out} == {0}                                      # everything ended up on executor 0
        assert s.is_blacklisted(1) and not s.is_blacklisted(0)
        assert any(e["event"] == "executor_blacklisted" and e["executor"] == 1 for e in events)
        out2 = s.run_job(ds, lambda tctx, it: tctx.executor_id)
        assert out2 == [0, 0, 0, 0]                                            # new jobs avoid it too
        time.sleep(3.2)                                                        # blacklist timeout expires
        assert not s.is_blacklisted(1)
        assert s.run_job(ds, lambda tctx, it: tctx.executor_id) == [0, 1, 0, 1]
        assert any(e["event"] == "executor_unblacklisted" for e in events)


@pytest.mark.parametrize("algorithm", ["asgd", "asaga", "sgd_sync", "saga_sync", "sgd_mllib"])
def test_every_driver_runs_end_to_end_on_executor_processes(algorithm):
    """``--master=local-cluster[2]``: data prep, the training loop AND the final objective evaluation all ship closures to
    executor processes (regression: the evaluation closure used to drag ``torch.ops`` into the task binary)."""
    cfg = AsyncConfig(algorithm=algorithm, synthetic="linear", num_rows=400, num_features=6, num_partitions=2,
                      num_iterations=12 if "sync" in algorithm or algorithm == "sgd_mllib" else 40, step_size=0.3,
                      batch_rate=0.4, printer_freq=4, backend="cpu", master="local-cluster[2]", quiet=True, bucket_ratio=0.5)
    rep = algo.run(cfg)
    assert rep.objective, "objective-vs-time table missing"
    first, last = rep.objective[0][1], rep.objective[-1][1]
    assert np.isfinite(last) and last < first, (first, last)


def test_wedged_executor_is_expired_by_missing_heartbeats_and_replaced():
    """HeartbeatReceiverSuite: a SIGSTOPped executor keeps its pipe open but stops heartbeating; the driver expires it
    after ``network_timeout``, kills it and the slot is re-filled."""
    import signal
    from asyncframework_b200.core.cluster import ClusterScheduler
    from asyncframework_b200.core.env import RuntimeLimits
    from asyncframework_b200.data.dataset import ShardedDataset
    events = []
    lim = RuntimeLimits(heartbeat_interval=0.1, network_timeout=2.5)      # generous: only the STOPPED process may expire
    with ClusterScheduler(2, limits=lim, recover_lost_executors=True) as s:
        s.listeners.add(events.append)
        ds = ShardedDataset.from_records(s, list(range(20)), 4)
        assert ds.count() == 20
        t0 = time.time()
        while len(s.executor_heartbeats) < 2 and time.time() - t0 < 10:
            time.sleep(0.05)
        hb = s.executor_heartbeats[1]
        assert hb["tasks_run"] == 2 and hb["storage_max"] > 0
        os.kill(hb["pid"], signal.SIGSTOP)                      # wedged, not dead
        res = {}
        w = s.submit_job(ds, lambda tctx, it: sum(it), result_handler=res.__setitem__)
        assert w.wait(30)
        assert res == {p: sum(range(5 * p, 5 * p + 5)) for p in range(4)}
        assert s.executors_expired == 1 and s.executors_recovered == 1
        assert [e["executor"] for e in events if e["event"] == "executor_expired"] == [1]


def test_context_cleaner_frees_broadcast_pieces_and_shipped_source_blocks():
    """ContextCleanerSuite: state on the executors dies with its driver-side handle — one broadcast per round must not pile up."""
    import gc
    from asyncframework_b200.core.cleaner import ContextCleaner
    from asyncframework_b200.core.session import AsyncSession
    with AsyncSession("local-cluster[2]") as sc:
        ds = sc.parallelize(range(8), 4)
        ids, keep = [], []
        for r in range(5):                                        # the ASGD pattern: a new broadcast every round
            bc = sc.broadcast(torch.full((50_000,), float(r)))
            ids.append(bc.id)
            keep.append(bc)                                       # (held on purpose: a dropped handle is cleaned immediately)
            assert ds.map(lambda i, bc=bc: float(bc.value[0])).collect() == [float(r)] * 8
        def pieces():
            st = sc.scheduler.storage_status()
            return sorted({b.split("_")[1] for s in st.values() for b in s["blocks"] if b.startswith("broadcast_")})
        assert len(pieces()) == 5
        del bc, keep

        def eventually(pred, timeout=10.0):                       # cleanup is asynchronous by design (ContextCleanerSuite uses `eventually`)
            t0 = time.time()
            while time.time() - t0 < timeout:
                gc.collect()
                ContextCleaner.get().flush()
                if pred():
                    return True
                time.sleep(0.1)
            return False
        assert eventually(lambda: pieces() == []), pieces()       # every piece, on the driver and on both executors, is gone
        # shipped source data of a dataset that is no longer referenced
        big = sc.parallelize([bytes(1000)] * 400, 4)
        assert big.count() == 400
        held = lambda: sorted(set(sc.scheduler.run_job(ds, lambda t, it: __import__("asyncframework_b200.data.dataset", fromlist=["x"])._SOURCE_STORE.keys().__len__())))  # noqa: E731
        before = held()
        del big
        assert eventually(lambda: max(held()) < max(before))


def test_executor_cores_are_concurrent_task_slots():
    """``local-cluster[n, c, m]`` / ``spark.executor.cores``: an executor with c cores runs c tasks at the same time."""
    from asyncframework_b200.core.session import AsyncSession

    def slow(tctx, it):
        t0 = time.time()
        time.sleep(0.5)
        return (os.getpid(), t0, time.time())
    for cores, expect_overlap in ((1, False), (2, True)):
        with AsyncSession(f"local-cluster[1,{cores},512]") as sc:
            ds = sc.parallelize(range(4), 2)
            assert ds.count() == 4                                           # warm the executor up
            t0 = time.time()
            (p0, a0, b0), (p1, a1, b1) = sc.scheduler.run_job(ds, slow)
            wall = time.time() - t0
            assert p0 == p1                                                  # one executor process
            overlap = min(b0, b1) - max(a0, a1) > 0.2
            assert overlap == expect_overlap and (wall < 0.95 if expect_overlap else wall > 0.95)


def test_fused_backend_launched_through_the_cluster_master_single_rank_on_the_simulator():
    """``deploy.run_fused_on_cluster``: executor processes become ranks of the fused engine (torchrun contract exported by the task).
    Without CUDA the single-rank engine runs on the round simulator, so the launcher mechanics are testable here."""
    from asyncframework_b200.core.session import AsyncSession
    from asyncframework_b200.deploy import run_fused_on_cluster
    cfg = AsyncConfig(algorithm="asgd", synthetic="mnist8m", num_rows=4000, num_features=40, num_partitions=2, num_iterations=30,
                      step_size=1e-6, batch_rate=0.2, printer_freq=10, loss="logistic", backend="fused", quiet=True)
    with AsyncSession("local-cluster[1]") as sc:
        rep = run_fused_on_cluster(sc, cfg)
    assert rep["accepted"] >= 30 and rep["final_weights"].numel() == 40 and float(rep["final_weights"].abs().sum()) > 0
    assert rep["objective"][-1][1] < rep["objective"][0][1]
    with AsyncSession("local[2]") as sc:
        with pytest.raises(TypeError, match="cluster master"):
            run_fused_on_cluster(sc, cfg)
    with AsyncSession("local-cluster[2]") as sc:
        with pytest.raises(ValueError, match="multiple"):
            run_fused_on_cluster(sc, AsyncConfig(algorithm="asgd", synthetic="mnist8m", num_partitions=3, backend="fused"))


def test_task_not_serializable_names_the_offending_reference(cluster):
    """SerializationDebugger behaviour: the error shows the chain from the task to the object that cannot be pickled."""
    import threading
    from asyncframework_b200.core.closure import TaskNotSerializable, explain_pickle_failure

    class Holder:
        def __init__(self):
            self.name = "cfg"
            self.guard = threading.Lock()
    h = Holder()
    ds = ShardedDataset.from_records(cluster, list(range(8)), 4)
    with pytest.raises(TaskNotSerializable) as ei:
        ds.map(lambda x: (x, h.name)).collect()
    msg = str(ei.value)
    assert "Task not serializable" in msg and "<closure h>" in msg and ".guard" in msg and "lock" in msg.lower() and "hint:" in msg
    assert ds.map(lambda x: x + 1).reduce(lambda a, b: a + b) == 36             # the scheduler is still healthy
    assert explain_pickle_failure(lambda x: x + 1) is None

    def uses_torch_ops(x):
        return torch.ops is not None and x
    why = explain_pickle_failure(uses_torch_ops)
    assert why is not None and "_Ops" in why or why is None                     # (by-reference functions of importable modules pickle fine)


def test_rpc_message_max_size_catches_data_captured_by_a_closure():
    """LocalClusterSparkContext sets spark.rpc.message.maxSize = 1 MiB so tests catch closures that drag data along."""
    from asyncframework_b200.conf import AsyncConf
    from asyncframework_b200.core.session import AsyncSession
    with AsyncSession("local-cluster[2]", conf=AsyncConf(load_defaults=False).set("spark.rpc.message.maxSize", "1")) as sc:
        ds = sc.parallelize(range(8), 4)
        assert ds.count() == 8                                                  # the source data itself may be as big as it likes
        big = torch.zeros(400_000, dtype=torch.float64)                          # 3.2 MB captured by the closure
        with pytest.raises(RuntimeError, match="spark.rpc.message.maxSize"):
            ds.map(lambda i: float(big[i])).collect()
        bc = sc.broadcast(big)                                                   # the fix the message recommends
        assert ds.map(lambda i: float(bc.value[i])).collect() == [0.0] * 8


def test_numeric_and_collection_accumulators_over_executor_processes(cluster):
    """LongAccumulator sum / count / avg and CollectionAccumulator (AccumulatorV2.scala:290-480) with updates travelling back
    from executor processes with the task results."""
    from asyncframework_b200.core.accumulator import CollectionAccumulator, NumericAccumulator
    acc, odd = NumericAccumulator(0, "x"), CollectionAccumulator("odd")
    ds = ShardedDataset.from_records(cluster, list(range(1, 41)), 4)

    def f(x):
        acc.add(x)
        if x % 2:
            odd.add(x)
        return x
    assert ds.map(f).count() == 40
    assert acc.value == acc.sum == 820 and acc.count == 40 and acc.avg == 20.5
    assert sorted(odd.value) == list(range(1, 41, 2))
    acc.reset()
    assert acc.count == 0 and acc.value == 0 and acc.avg != acc.avg          # NaN on an empty accumulator
    with pytest.raises(Exception):
        ds.map(lambda x: len(odd.value)).collect()                            # tasks may only add
