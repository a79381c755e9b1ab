"""RDD-slice semantics + the ASYNC runtime (context / barrier / async actions / scheduler modes),
driven deterministically like DAGSchedulerSuite drives completions by hand."""
import threading
import time

import numpy as np
import pytest
import torch

from asyncframework_b200.core import AsyncBroadcast, AsyncContext, Scheduler, barriers
from asyncframework_b200.core.sampler import bernoulli_mask, mix64_int, threshold24
from asyncframework_b200.data import LinearDataGenerator, ShardedDataset


def test_sampler_is_deterministic_and_unbiased():
    idx = np.arange(200000)
    m1 = bernoulli_mask(43, idx, 0.1)
    assert np.array_equal(m1, bernoulli_mask(43, idx, 0.1))
    assert abs(m1.mean() - 0.1) < 0.003
    assert abs((m1 & bernoulli_mask(44, idx, 0.1)).mean() - 0.01) < 0.002      # independent across seeds
    assert bernoulli_mask(1, idx, 1.0).all() and not bernoulli_mask(1, idx, 0.0).any()
    assert threshold24(0.5) == 1 << 23 and mix64_int(0) == 0


def test_parallelize_map_reduce_aggregate(scheduler):
    ds = ShardedDataset.from_records(scheduler, list(range(100)), 7)
    assert ds.count() == 100 and ds.map(lambda x: x * 2).reduce(lambda a, b: a + b) == 9900
    assert ds.filter(lambda x: x % 2 == 0).count() == 50
    assert ds.aggregate(lambda: 0, lambda a, x: a + x, lambda a, b: a + b) == 4950
    assert ds.tree_aggregate(lambda: 0, lambda a, x: a + x, lambda a, b: a + b, depth=2) == 4950
    assert sorted(ds.collect()) == list(range(100)) and ds.first() == 0 and len(ds.take(5)) == 5
    assert sum(len(g) for g in ds.glom().collect()) == 100


def test_repartition_zip_with_index_sample(scheduler):
    X, y, _ = LinearDataGenerator.generate_linear_arrays(1003, 5, seed=1)
    ds = ShardedDataset.from_arrays(scheduler, X, y, 3).repartition(4)
    sizes = ds.partition_sizes()
    assert sum(sizes) == 1003 and max(sizes) - min(sizes) <= 1           # round-robin balance
    zi = ds.zip_with_index()
    g = torch.cat([s.gidx for s in zi.shards()])
    assert sorted(g.tolist()) == list(range(1003))
    assert zi.batch(1).gidx[0].item() == sizes[0]                          # offset = Σ previous sizes
    s1 = zi.sample(False, 0.3, 7)
    s2 = zi.sample(False, 0.3, 7)
    assert s1.partition_sizes() == s2.partition_sizes()
    n = s1.count()
    assert abs(n - 300) < 60
    # record-level and batch-level sampling select the same rows
    rec_idx = sorted(i for _, i in s1.map(lambda r: r).collect())
    bat_idx = sorted(torch.cat([s.gidx for s in s1.shards()]).tolist())
    assert rec_idx == bat_idx
    wr = zi.sample(True, 0.5, 1)                                             # with replacement: Poisson(0.5) copies per row
    n_wr = wr.count()
    assert abs(n_wr - 500) < 90 and wr.count() == n_wr


def test_cache_materialises_once(scheduler):
    calls = {"n": 0}
    base = ShardedDataset.from_records(scheduler, list(range(20)), 2)

    def f(x):
        calls["n"] += 1
        return x + 1
    ds = base.map(f).cache()
    assert ds.count() == 20 and ds.count() == 20
    assert calls["n"] == 20


def test_async_context_queue_clock_stat():
    ctx = AsyncContext(4)
    assert ctx.size() == 0 and not ctx.has_next() and ctx.collect(timeout=0.01) is None
    from asyncframework_b200.core import PartialResult
    ctx.put(PartialResult("a", 1, 5, 10, 2))
    ctx.put(PartialResult("b", 0, 6, 11, 3))
    assert ctx.size() == 2 and ctx.has_next()
    assert ctx.collect() == "a"
    env = ctx.collect_all()
    assert (env.data, env.staleness, env.timestamp, env.batch_size, env.worker_id) == ("b", 0, 6, 11, 3)
    assert ctx.add_to_current_time(1) == 0 and ctx.current_time == 1
    assert ctx.stat.get(0) is None and len(ctx.stat) == 0
    ctx.stat.task_done(1, 3, 10.0)
    ctx.stat.task_done(1, 5, 20.0)
    st = ctx.stat[1]
    assert st.staleness == 5 and st.num_tasks == 2 and st.avg_task_time_ms == pytest.approx(15.0) and st.available
    ctx.stat.set_available(2, False)
    assert st.available_workers() == 1 and st.max_staleness() == 5
    # blocking take wakes up when another thread puts
    t = threading.Timer(0.05, lambda: ctx.put(PartialResult("late", 0, 0, 0, 0)))
    t.start()
    assert ctx.collect(timeout=2.0) == "late"


def test_async_broadcast_versions():
    ctx = AsyncContext(2)
    w = torch.zeros(3)
    a0 = AsyncBroadcast(ctx, w, slots=2)
    w += 1
    a1 = AsyncBroadcast(ctx, w, slots=2)
    assert a0.value.tolist() == [0, 0, 0] and a1.value.tolist() == [1, 1, 1]      # snapshot at broadcast time
    assert a1.value_at(a0.index).tolist() == [0, 0, 0] and a1(a0.index).tolist() == [0, 0, 0]
    w += 1
    a2 = AsyncBroadcast(ctx, w, slots=2)
    with pytest.raises(KeyError):
        a2.value_at(a0.index)                                                    # ring of 2 evicted version 0
    a2.set_id(a1.index)
    assert a2.value.tolist() == [1, 1, 1] and a2.get_index() == a1.index


def test_barrier_predicates_and_selection(scheduler):
    ds = ShardedDataset.from_records(scheduler, list(range(8)), 4)
    ctx = AsyncContext(4, scheduler)
    # no entries yet -> every partition is selected (RDD.scala:1053-1064)
    assert ds.async_barrier(barriers.bsp(4), ctx.stat).selected_workers == [0, 1, 2, 3]
    for p in range(4):
        ctx.stat.task_done(p, p, 1.0)
    ctx.stat.set_available(3, False)
    assert ds.async_barrier(barriers.asp(), ctx.stat).selected_workers == [0, 1, 2]
    assert ds.async_barrier(barriers.bucket(4, 0.75), ctx.stat).selected_workers == [0, 1, 2]
    assert ds.async_barrier(barriers.bsp(4), ctx.stat).selected_workers == []
    assert ds.async_barrier(barriers.ssp(2), ctx.stat).selected_workers == []      # max staleness is 3
    assert ds.async_barrier(lambda st: st.staleness <= 1, ctx.stat).selected_workers == [0, 1]
    sel = ds.async_barrier(barriers.asp(), ctx.stat)
    assert sorted(sel.collect()) == [0, 1, 2, 3, 4, 5]                              # partition 3 is empty


def test_first_async_job_blocks_like_the_reference(scheduler):
    """DAGScheduler.first_iter quirk (Appendix B.4): the first async job waits for all tasks."""
    ds = ShardedDataset.from_records(scheduler, [1, 2, 3, 4], 2)
    ctx = AsyncContext(2, scheduler)
    ds.async_reduce(lambda a, b: a + b, ctx)
    assert ctx.size() == 2 and scheduler.first_iter == 0              # both results already there
    out = ds.async_reduce(lambda a, b: a + b, ctx)                      # later jobs return a waiter immediately
    out.wait(5)
    assert ctx.size() == 4


def test_async_reduce_scripted_completions():
    """Release tasks by hand and check staleness / clock / availability bookkeeping."""
    scheduler = Scheduler(3, block_first_async_job=False)
    P = 3
    ds = ShardedDataset.from_records(scheduler, [1, 2, 3, 4, 5, 6], P)
    ctx = AsyncContext(P, scheduler)
    ctx.set_record_stat(True)
    gates = {p: threading.Event() for p in range(P)}

    def gated(p, it):
        gates[p].wait(5)
        return it
    job = ds.map_partitions_with_index(gated).async_reduce(lambda a, b: a + b, ctx)
    # first async job blocks by default (quirk B.4) => run it in a thread
    assert scheduler.mode == 1
    for p in range(P):
        assert not ctx.stat[p].available
    for p in (2, 0, 1):
        gates[p].set()
        t0 = time.time()
        while ctx.size() == 0 and time.time() - t0 < 5:
            time.sleep(0.001)
        env = ctx.collect_all(timeout=1)
        assert env.worker_id == p and env.batch_size == 2 and env.data == [3, 7, 11][p]
    assert ctx.current_time == 3
    assert all(ctx.stat[p].available for p in range(P))
    scheduler.stop()


def test_async_reduce_nonblocking_and_staleness():
    sched = Scheduler(2, block_first_async_job=False)
    try:
        ds = ShardedDataset.from_records(sched, [10, 20], 2)
        ctx = AsyncContext(2, sched)
        gates = [threading.Event(), threading.Event()]
        t0 = time.time()
        ds.map_partitions_with_index(lambda p, it: (gates[p].wait(5), it)[1]).async_reduce(lambda a, b: a + b, ctx)
        assert time.time() - t0 < 1.0 and ctx.size() == 0              # returned before any task finished
        gates[0].set()
        r0 = ctx.collect_all(timeout=5)
        assert r0.staleness == 0 and r0.timestamp == 0
        gates[1].set()
        r1 = ctx.collect_all(timeout=5)
        assert r1.staleness == 1 and r1.timestamp == 0                  # one result merged since its submit
        assert ctx.stat.max_staleness() == 1
    finally:
        sched.stop()


def test_async_aggregate_and_empty_partitions(scheduler):
    ds = ShardedDataset.from_records(scheduler, [1, 2, 3], 3).filter(lambda x: x != 2)
    ctx = AsyncContext(3, scheduler)
    ds.async_aggregate(lambda: [0, 0], lambda a, x: [a[0] + x, a[1] + 1], lambda a, b: [a[0] + b[0], a[1] + b[1]], ctx)
    got = sorted((ctx.collect_all(timeout=2).data[0]) for _ in range(2))
    assert got == [1, 3] and ctx.collect_all(timeout=0.05) is None
    assert ctx.dropped_empty == 1 and ctx.stat[1].available             # empty partition freed, no queue entry


def test_task_retry_and_failure_isolation():
    sched = Scheduler(2, max_failures=3)
    try:
        ds = ShardedDataset.from_records(sched, list(range(10)), 2)
        attempts = {"n": 0}

        def inject(tctx):
            if tctx.partition_id == 1 and tctx.attempt < 2:
                attempts["n"] += 1
                raise RuntimeError("injected")
        sched.fault_injector = inject
        assert ds.count() == 10 and attempts["n"] == 2                  # FailureSuite: retried then succeeded
        sched.fault_injector = lambda t: (_ for _ in ()).throw(RuntimeError("always")) if t.partition_id == 0 else None
        from asyncframework_b200.core import JobFailed
        with pytest.raises(JobFailed):
            ds.count()
        # async job failure must not strand the worker (quirk B.15 fixed)
        ctx = AsyncContext(2, sched)
        sched.block_first_async_job = False
        ds.async_reduce(lambda a, b: a + b, ctx)
        r = ctx.collect_all(timeout=5)
        assert r.worker_id == 1
        t0 = time.time()
        while not ctx.stat[0].available and time.time() - t0 < 5:
            time.sleep(0.01)
        assert ctx.stat[0].available
    finally:
        sched.stop()


def test_listener_bus_events(scheduler):
    events = []
    scheduler.listeners.add(events.append)
    ShardedDataset.from_records(scheduler, list(range(4)), 2).count()
    kinds = [e["event"] for e in events]
    assert kinds.count("job_start") == 1 and kinds.count("task_end") == 2 and kinds.count("job_end") == 1


def test_job_cancellation_skips_pending_tasks_and_frees_async_workers():
    """JobWaiter.cancel / cancel_job / cancel_all_jobs (JobCancellationSuite analogue): pending tasks are skipped, running
    closures can poll is_cancelled(), wait() raises JobCancelled, and an ASYNC job's workers become available again."""
    import threading
    import time
    from asyncframework_b200.core import AsyncContext, JobCancelled, Scheduler
    from asyncframework_b200.data import ShardedDataset
    sched = Scheduler(2, block_first_async_job=False)
    try:
        started, release, ran = threading.Event(), threading.Event(), []

        def slow(tctx, it):
            ran.append(tctx.partition_id)
            started.set()
            while not release.is_set() and not tctx.is_cancelled():
                time.sleep(0.002)
            return sum(it)

        ds = ShardedDataset.from_records(sched, list(range(40)), 8)          # 8 tasks on 2 executors: 6 stay queued
        events = []
        sched.listeners.add(events.append)
        w = sched.submit_job(ds, slow)
        assert started.wait(5)
        assert sched.cancel_job(w.job_id) and not sched.cancel_job(w.job_id)
        with pytest.raises(JobCancelled):
            w.wait(5)
        release.set()
        assert ds.count() == 40                                              # executors drained the skipped tasks and still work
        assert len(ran) <= 2 + 0 and sum(1 for e in events if e["event"] == "task_skipped") >= 6
        # asynchronous action: cancelling returns every selected worker to the available pool, no result is queued
        AC = AsyncContext(8, sched)
        AC.blocking = False
        release.clear(); started.clear()
        ds.map_partitions_with_index(lambda p, it: (started.set(), [release.wait(5)], list(it))[2]).async_reduce(lambda a, b: a + b, AC)
        assert started.wait(5) and AC.stat.available_workers() < 8
        assert sched.cancel_all_jobs() == 1
        release.set()
        deadline = time.time() + 5
        while AC.stat.available_workers() < 8 and time.time() < deadline:
            time.sleep(0.01)
        assert AC.stat.available_workers() == 8 and AC.size() == 0
    finally:
        release.set()
        sched.stop()


def test_blocking_and_async_actions_from_concurrent_driver_threads():
    """ThreadingSuite analogue for the in-process scheduler: blocking actions keep returning results (never a JobWaiter) while
    other driver threads issue ASYNC actions that leave the scheduler in non-blocking mode."""
    import threading
    from asyncframework_b200.core import AsyncContext, Scheduler
    from asyncframework_b200.data import ShardedDataset
    sched = Scheduler(4, block_first_async_job=False)
    try:
        ds = ShardedDataset.from_records(sched, list(range(400)), 4).cache()
        errors = []

        def blocking_driver(i):
            try:
                for r in range(40):
                    assert ds.map(lambda x, i=i: x + i).reduce(lambda a, b: a + b) == sum(range(400)) + 400 * i
                    assert ds.count() == 400
            except BaseException as exc:  # noqa: BLE001
                errors.append(repr(exc))

        def async_driver():
            try:
                AC = AsyncContext(4, sched)
                got = 0
                for r in range(25):
                    ds.map(lambda x: 1).async_reduce(lambda a, b: a + b, AC)
                    for _ in range(4):
                        info = AC.collect_all(timeout=10)
                        assert info is not None and info.data == 100
                        got += 1
                assert got == 100
            except BaseException as exc:  # noqa: BLE001
                errors.append(repr(exc))

        threads = [threading.Thread(target=blocking_driver, args=(i,)) for i in range(3)] + [threading.Thread(target=async_driver)]
        for t in threads:
            t.start()
        for t in threads:
            t.join(60)
        assert not errors, errors[:3]
        assert not any(t.is_alive() for t in threads)
    finally:
        sched.stop()


def test_accumulators_count_successful_attempts_exactly_once():
    """sc.longAccumulator semantics (AccumulatorSuite): tasks add, driver reads; a failed attempt's updates are dropped, the
    retried attempt counts once."""
    from asyncframework_b200 import AsyncSession
    sc = AsyncSession("local[3]")
    try:
        rows, big = sc.long_accumulator("rows"), sc.double_accumulator("sum")
        ds = sc.parallelize(list(range(100)), 4)

        def f(x):
            rows.add(1)
            big.add(float(x))
            return x
        assert ds.map(f).count() == 100
        assert rows.value == 100 and big.value == float(sum(range(100)))
        assert rows.count == 100 and rows.avg == 1.0 and big.count == 100 and big.avg == 49.5      # LongAccumulator.count / avg
        names = sc.collection_accumulator("names")
        ds.foreach(lambda x: names.add(x) if x % 25 == 0 else None)
        assert sorted(names.value) == [0, 25, 50, 75]
        # a task that fails on its first attempt after adding: only the successful retry is counted
        attempts = {}

        def flaky(tctx):
            n = attempts.get(tctx.partition_id, 0)
            attempts[tctx.partition_id] = n + 1
            if tctx.partition_id == 1 and n == 0:
                rows.add(1000)                           # belongs to the attempt that is about to fail
                raise RuntimeError("injected")
        sc.scheduler.fault_injector = flaky
        rows.reset()
        assert ds.map(f).count() == 100 and rows.value == 100 and attempts[1] == 2
        sc.scheduler.fault_injector = None
        rows.add(5)                                      # driver-side add applies directly
        assert rows.value == 105
        mx = sc.accumulator(float("-inf"), "max", op=max)
        ds.map(lambda x: mx.add(float(x))).count()
        assert mx.value == 99.0
    finally:
        sc.stop()


def test_dataset_checkpoint_truncates_lineage(tmp_path, scheduler):
    """rdd.checkpoint(): partitions are written once, the re-opened dataset has no parent and gives the same answers."""
    from asyncframework_b200.data import LinearDataGenerator, ShardedDataset
    calls = {"n": 0}

    def expensive(x):
        calls["n"] += 1
        return x * 3
    ds = ShardedDataset.from_records(scheduler, list(range(50)), 4).map(expensive)
    ck = ds.checkpoint(str(tmp_path / "rec"))
    n_after_write = calls["n"]
    assert ck.parent is None and sorted(ck.collect()) == [3 * i for i in range(50)] and ck.count() == 50
    assert calls["n"] == n_after_write == 50                      # the lineage was not recomputed by the later jobs
    pts = LinearDataGenerator.generate_linear_dataset(scheduler, 200, 6, 0.1, 3, seed=2).zip_with_index()
    ckb = pts.checkpoint(str(tmp_path / "batch"))
    assert ckb.batch_capable and ckb.parent is None and ckb.partition_sizes() == pts.partition_sizes()
    a = torch.cat([s.X for s in pts.shards()])
    b = torch.cat([s.X for s in ShardedDataset.from_checkpoint(scheduler, str(tmp_path / "batch")).shards()])
    assert torch.equal(a, b)
    assert ckb.sample(False, 0.5, 7).count() == pts.sample(False, 0.5, 7).count()      # gidx survived: same Bernoulli sample


def test_wider_rdd_surface(scheduler):
    """union / zip / key-value ops / shuffles / ordering / numeric actions (RDDSuite + PairRDDFunctionsSuite analogue)."""
    from asyncframework_b200.data import ShardedDataset
    a = ShardedDataset.from_records(scheduler, list(range(10)), 3)
    b = ShardedDataset.from_records(scheduler, list(range(100, 105)), 2)
    u = a.union(b)
    assert u.num_partitions == 5 and sorted(u.collect()) == list(range(10)) + list(range(100, 105))
    assert a.zip(a.map(lambda x: x * x)).collect() == [(i, i * i) for i in range(10)]
    words = ShardedDataset.from_records(scheduler, "a b a c b a d".split(), 3)
    counts = dict(words.map(lambda w: (w, 1)).reduce_by_key(lambda x, y: x + y).collect())
    assert counts == {"a": 3, "b": 2, "c": 1, "d": 1} == words.count_by_value()
    groups = dict(words.key_by(lambda w: w).group_by_key(2).map_values(len).collect())
    assert groups == counts
    assert sorted(words.distinct().collect()) == ["a", "b", "c", "d"]
    kv = a.key_by(lambda x: x % 3)
    assert kv.count_by_key() == {0: 4, 1: 3, 2: 3} and sorted(kv.keys().distinct().collect()) == [0, 1, 2]
    assert sorted(kv.map_values(lambda v: -v).values().collect()) == sorted(-i for i in range(10))
    assert a.sort_by(lambda x: -x).collect() == list(range(9, -1, -1)) and a.top(3) == [9, 8, 7]
    assert a.sum() == 45 and a.max() == 9 and a.min() == 0 and a.mean() == 4.5
    assert len(a.take_sample(False, 4, seed=1)) == 4 and len(a.take_sample(True, 20, seed=1)) == 20
    ids = a.zip_with_unique_id().collect()
    assert len({i for _, i in ids}) == 10 and sorted(r for r, _ in ids) == list(range(10))
    assert a.coalesce(2).num_partitions == 2 and sorted(a.coalesce(2).collect()) == list(range(10))
    seen = []
    a.foreach(seen.append)
    assert sorted(seen) == list(range(10))


def test_async_grad_fast_path_matches_manual_gradients_and_saga_correction():
    """``ds.async_grad(loss, w_bcast, ctx[, history])``: one fused pass per selected worker, results in ctx."""
    import numpy as np
    from asyncframework_b200 import AsyncSession, barriers
    with AsyncSession("local[3]") as sc:
        rng = np.random.default_rng(0)
        X, w_true = rng.standard_normal((300, 6)), rng.standard_normal(6)
        y = (X @ w_true > 0).astype(np.float64)
        from asyncframework_b200.data import ShardedDataset
        ds = ShardedDataset.from_arrays(sc.scheduler, X, y, 3).zip_with_index().cache()
        ctx = sc.async_context(3)
        ctx.blocking = False
        w = torch.from_numpy(rng.standard_normal(6))
        bc = sc.broadcast(w)
        sampled = ds.async_barrier(barriers.asp(), ctx.stat).sample(False, 0.5, 11)
        sampled.async_grad("logistic", bc, ctx)
        got = [ctx.collect_all(timeout=5) for _ in range(3)]
        g = sum(r.data[0] for r in got)
        n = sum(r.data[2] for r in got)
        # reference: same sample, plain formulas
        shards = sampled.shards()
        Xs = torch.cat([s.X[:, :6] for s in shards]); ys = torch.cat([s.y for s in shards])
        ref = Xs.T @ (torch.sigmoid(Xs @ w) - ys)
        assert n == Xs.shape[0] and torch.allclose(g, ref, atol=1e-10)
        assert sorted(r.worker_id for r in got) == [0, 1, 2] and all(r.staleness >= 0 for r in got)
        # SAGA: history-corrected gradient + new residuals + their row positions
        hist = {p: torch.full((s.num_rows,), 0.25, dtype=torch.float64) for p, s in enumerate(ds.shards())}
        sampled.async_grad("logistic", bc, ctx, history=hist)
        got = [ctx.collect_all(timeout=5) for _ in range(3)]
        g2 = sum(r.data[0] for r in got)
        assert torch.allclose(g2, Xs.T @ (torch.sigmoid(Xs @ w) - ys - 0.25), atol=1e-10)
        r0 = [r for r in got if r.worker_id == 0][0]
        assert r0.data[3].numel() == r0.data[2] == r0.data[4].numel()          # residuals and row positions line up


def test_result_handler_exception_fails_the_job_but_never_kills_an_executor_thread():
    """A throwing result handler (``async_aggregate(strict_empty=True)`` on an empty partition: the reference's ``None.get``,
    RDD.scala:1333) must fail THAT job, free the worker and leave the executor threads alive — the next blocking job still runs."""
    sched = Scheduler(2, block_first_async_job=False)
    try:
        ds = ShardedDataset.from_records(sched, list(range(8)), 2)
        empty = ds.filter(lambda x: x < 4)                      # partition 1 (records 4..7) becomes empty
        ctx = AsyncContext(2, sched)
        waiter = empty.async_aggregate(0, lambda a, r: a + r, lambda a, b: a + b, ctx, strict_empty=True)
        r = ctx.collect_all(timeout=5)
        assert r is not None and r.worker_id == 0 and r.data == 0 + 1 + 2 + 3
        t0 = time.time()
        while waiter.failure is None and time.time() - t0 < 5:
            time.sleep(0.01)
        assert isinstance(waiter.failure, ValueError)
        assert all(t.is_alive() for t in sched._threads)
        assert ds.count() == 8                                  # would hang forever if an executor thread had died
        assert ctx.lost and ctx.lost[0][0] == 1
        t0 = time.time()
        while not ctx.stat[1].available and time.time() - t0 < 5:
            time.sleep(0.01)
        assert ctx.stat[1].available
    finally:
        sched.stop()
