"""The five drivers on the CPU backend + oracle equivalence under scripted schedules."""
import math
import threading
import time

import numpy as np
import pytest
import torch

from asyncframework_b200 import algo
from asyncframework_b200.algo.delay import DelayInjector, straggler_sets
from asyncframework_b200.algo.oracle import (AsyncOracle, run_asaga_oracle, run_asgd_oracle, run_saga_sync_oracle,
                                             run_sgd_sync_oracle)
from asyncframework_b200.config import AsyncConfig
from asyncframework_b200.core import AsyncBroadcast, AsyncContext, Scheduler, barriers
from asyncframework_b200.data import LinearDataGenerator, ShardedDataset
from asyncframework_b200.optim import make_gradient


def _cfg(alg, **kw):
    base = dict(algorithm=alg, synthetic="linear", num_rows=1000, num_features=20, num_partitions=4, batch_rate=0.2,
                printer_freq=50, backend="cpu", seed=42)
    base.update(kw)
    return AsyncConfig(**base)


def _parts(ds):
    zi = ds.zip_with_index()
    return [(s.X[:, :s.num_features].numpy(), s.y.numpy(), s.gidx.numpy()) for s in zi.shards()]


@pytest.mark.parametrize("alg,kw", [
    ("sgd_sync", dict(num_iterations=60, step_size=1.0)),
    ("asgd", dict(num_iterations=240, step_size=1.0, bucket_ratio=0.5)),
    ("saga_sync", dict(num_iterations=60, step_size=0.3)),
    ("asaga", dict(num_iterations=240, step_size=0.05, bucket_ratio=0.5)),
    ("sgd_mllib", dict(num_iterations=60, step_size=1.0)),
])
def test_driver_converges_and_reports(alg, kw, capsys):
    rep = algo.run(_cfg(alg, **kw))
    objs = [o for _, o in rep.objective]
    assert objs[0] > 0.3 and objs[-1] < 0.05                              # noise floor is eps² = 0.01
    assert rep.elapsed_ms > 0 and rep.final_weights is not None
    if alg != "sgd_mllib":
        assert rep.iterations == kw["num_iterations"] and len(rep.waiting_ms) == 4
        assert len(rep.snapshots) >= 2 and rep.snapshots[0][0] == 0.0
    rep.print_reference_style()
    out = capsys.readouterr().out
    if alg == "sgd_mllib":                                                 # SparkSGDMLLIB prints only the table and "finished"
        assert out.strip().splitlines()[0] == "*" * 33 and out.strip().splitlines()[-1] == "finished"
    else:
        assert "Elapsed time(ms): " in out and "Individual waiting times:" in out
        assert "Average waiting time(ms) per worker and iteration:" in out and out.count("*" * 33) == 2


def test_sync_sgd_matches_oracle_exactly():
    sched = Scheduler(3)
    try:
        ds = LinearDataGenerator.generate_linear_dataset(sched, 600, 8, 0.1, 3, seed=5)
        cfg = _cfg("sgd_sync", num_partitions=3, num_iterations=25, step_size=0.7, batch_rate=0.3, eval_objective=False)
        rep = algo.run(cfg, points=ds, scheduler=sched)
        orc = AsyncOracle(_parts(ds), "least_squares", 0.3, 0.7, 10**9, 42)
        ref = run_sgd_sync_oracle(orc, 25)
        assert np.allclose(rep.final_weights.numpy(), ref.w, rtol=1e-10, atol=1e-12)
        cfg2 = _cfg("saga_sync", num_partitions=3, num_iterations=25, step_size=0.2, batch_rate=0.3, eval_objective=False)
        rep2 = algo.run(cfg2, points=ds, scheduler=sched)
        orc2 = AsyncOracle(_parts(ds), "least_squares", 0.3, 0.2, 10**9, 42)
        ref2 = run_saga_sync_oracle(orc2, 25)
        assert np.allclose(rep2.final_weights.numpy(), ref2.w, rtol=1e-9, atol=1e-11)
    finally:
        sched.stop()


class ScriptedRun:
    """Drives the PUBLIC API (async_barrier / sample / map_batch / async_reduce|aggregate /
    collect_all) with an explicit event list, releasing task completions by hand."""

    def __init__(self, ds, loss, b, gamma, taw, seed, saga=False):
        self.sched = ds.scheduler
        self.sched.block_first_async_job = False
        self.P = ds.num_partitions
        self.pts = ds.zip_with_index().cache()
        self.sizes = ds.partition_sizes()
        self.N = sum(self.sizes)
        self.starts = [sum(self.sizes[:p]) for p in range(self.P)]
        self.grad = make_gradient(loss)
        self.b, self.gamma, self.taw, self.seed, self.saga = b, gamma, taw, seed, saga
        d = ds.batch(0).num_features
        self.w = torch.zeros(d, dtype=torch.float64)
        self.abar = torch.zeros(d, dtype=torch.float64)
        self.hist = [torch.zeros(n, dtype=torch.float64) for n in self.sizes]
        self.ctx = AsyncContext(self.P, self.sched)
        self.k = 0
        self.gates = {p: [] for p in range(self.P)}
        self.log = []
        for p in range(self.P):                       # give every worker a STAT entry so predicates apply
            self.ctx.stat.set_available(p, True)

    def submit(self, workers):
        a = AsyncBroadcast(self.ctx, self.w)
        sel = self.pts.async_barrier(lambda st: st.worker_id in workers, self.ctx.stat)
        assert sel.selected_workers == sorted(workers)
        gate = {p: threading.Event() for p in workers}
        for p in workers:
            self.gates[p].append(gate[p])
        sampled = sel.sample(False, self.b, self.seed + self.k + 1)
        if self.saga:
            from asyncframework_b200.algo.asaga import saga_partition

            def task(s, a=a, gate=gate):
                gate[s.partition].wait(10)
                s.meta["rows"] = (s.gidx - self.starts[s.partition]).to(torch.int64)
                return saga_partition(self.grad, s, a.value, self.hist[s.partition])
            sampled.map_batch(task).async_aggregate(lambda: None, lambda acc, r: r, lambda x, y: x, self.ctx)
        else:
            def task(s, a=a, gate=gate):
                gate[s.partition].wait(10)
                return self.grad.batch(s, a.value)[0]
            sampled.map_batch(task).async_reduce(lambda x, y: x + y, self.ctx)

    def complete(self, p):
        before = self.ctx.current_time
        self.gates[p].pop(0).set()
        info = self.ctx.collect_all(timeout=10)
        assert info is not None and info.worker_id == p and self.ctx.current_time == before + 1
        par = self.b * self.N / self.P
        if self.saga:
            lag = self.k - info.timestamp
            ok = lag <= self.taw
            self.log.append((p, lag, ok))
            if ok:
                rows, r, g, _ = info.data
                self.hist[p][rows] = r
                self.w = self.w - self.gamma * (g / par + self.abar)
                self.abar = self.abar + g / self.N
                self.k += 1
        else:
            ok = info.staleness <= self.taw
            self.log.append((p, info.staleness, ok))
            if ok:
                self.w = self.w + (-self.gamma / math.sqrt(self.k // self.P + 1)) * info.data / par
                self.k += 1
        t0 = time.time()
        while not self.ctx.stat[p].available and time.time() - t0 < 5:
            time.sleep(0.001)

    def play(self, events):
        for ev, arg in events:
            self.submit(arg) if ev == "submit" else self.complete(arg)


EVENTS = [("submit", [0, 1, 2]), ("complete", 1), ("submit", [1]), ("complete", 0), ("complete", 1),
          ("submit", [0, 1]), ("complete", 2), ("complete", 1), ("submit", [1, 2]), ("complete", 0),
          ("complete", 2), ("complete", 1)]


@pytest.mark.parametrize("taw", [10**6, 1, 0])
def test_asgd_public_api_matches_oracle_under_script(taw):
    sched = Scheduler(3)
    try:
        ds = LinearDataGenerator.generate_linear_dataset(sched, 450, 6, 0.1, 3, seed=8)
        run = ScriptedRun(ds, "least_squares", 0.4, 0.9, taw, 42)
        run.play(EVENTS)
        orc = AsyncOracle(_parts(ds), "least_squares", 0.4, 0.9, taw, 42)
        ref = run_asgd_oracle(orc, EVENTS, 10**6)
        assert [s for _, s, _ in run.log] == ref.staleness and [o for _, _, o in run.log] == ref.accepted
        assert run.k == ref.k and run.ctx.current_time == ref.clock
        assert np.allclose(run.w.numpy(), ref.w, rtol=1e-10, atol=1e-12)
        if taw == 0:
            assert not all(ref.accepted)             # the staleness filter really dropped something
    finally:
        sched.stop()


@pytest.mark.parametrize("taw", [10**6, 1])
def test_asaga_public_api_matches_oracle_under_script(taw):
    sched = Scheduler(3)
    try:
        ds = LinearDataGenerator.generate_linear_dataset(sched, 450, 6, 0.1, 3, seed=9)
        run = ScriptedRun(ds, "least_squares", 0.4, 0.05, taw, 42, saga=True)
        run.play(EVENTS)
        orc = AsyncOracle(_parts(ds), "least_squares", 0.4, 0.05, taw, 42)
        ref = run_asaga_oracle(orc, EVENTS, 10**6)
        assert [o for _, _, o in run.log] == ref.accepted and run.k == ref.k
        assert np.allclose(run.w.numpy(), ref.w, rtol=1e-9, atol=1e-12)
        assert np.allclose(run.abar.numpy(), ref.alpha_bar, rtol=1e-9, atol=1e-12)
        for p in range(3):
            assert np.allclose(run.hist[p].numpy(), ref.hist[p], atol=1e-12)       # discarded results left history untouched
    finally:
        sched.stop()


def test_delay_injector_modes():
    sleeps = []
    inj = DelayInjector(8, 0.6, warmup_updates=4, sleep=sleeps.append)
    hook = inj.hook()
    list(hook(0, iter(())))
    assert sleeps == []                                           # no delay before calibration
    for k in range(4):
        inj.observe(k, 10.0)
    assert not inj.maybe_calibrate(4) and inj.maybe_calibrate(5) and inj.avg_delay_ms == 10.0
    list(hook(0, iter(())))
    list(hook(3, iter(())))
    assert sleeps == [pytest.approx(0.006)]                        # only worker 0, coeff * avgDelay
    normal, long_tail = straggler_sets(32)                         # README Fig.7: 32 workers, 8 stragglers
    assert len(normal) + len(long_tail) == 8 and long_tail == [0, 4] and normal == [8, 12, 16, 20, 24, 28]
    prod = DelayInjector(32, -1, 1, sleep=sleeps.append)
    prod.observe(0, 4.0)
    prod.maybe_calibrate(2)
    assert prod.mode == "production"
    assert 10.0 <= prod.delay_ms_for(0) < 40.0 and 6.0 <= prod.delay_ms_for(8) < 10.0 and prod.delay_ms_for(1) == 0.0


def test_asgd_with_controlled_delay_still_converges():
    rep = algo.run(_cfg("asgd", num_iterations=600, step_size=1.0, bucket_ratio=0.5, coeff=1.0,
                        warmup_updates_per_worker=20, printer_freq=200))
    assert rep.objective[-1][1] < 0.05
    assert rep.extra["avg_delay_ms"] > 0 and rep.extra["injected_delay_ms"] > 0
    # the straggler (worker 0) contributes fewer updates => it waits less than the others get to
    assert sum(rep.staleness_hist.values()) >= 600


def test_staleness_bound_drops_updates():
    rep = algo.run(_cfg("asgd", num_iterations=200, step_size=0.5, bucket_ratio=0.25, taw=0, printer_freq=100))
    assert rep.iterations == 200 and rep.dropped > 0


def test_metrics_jsonl(tmp_path):
    import json
    path = str(tmp_path / "m.jsonl")
    algo.run(_cfg("sgd_sync", num_iterations=20, step_size=1.0, metrics_path=path, printer_freq=5))
    kinds = [json.loads(l)["kind"] for l in open(path)]
    assert kinds.count("snapshot") == 4 and kinds[-1] == "report"


def test_async_driver_survives_a_hung_worker_and_sync_driver_aborts():
    """Failure detection wired into the drivers: a worker whose task never returns is declared dead after
    heartbeat_timeout_s; ASGD finishes with the remaining workers, synchronous SGD raises WorkerLost."""
    import threading
    from asyncframework_b200.algo.common import WorkerLost
    from asyncframework_b200.core import Scheduler
    release = threading.Event()
    hung = {"n": 0}

    def injector(tctx):
        if tctx.partition_id == 2:
            injector.calls += 1
            if injector.calls == 6:              # worker 2's sixth task never returns
                hung["n"] += 1
                release.wait(20)                 # its executor thread is stuck until the test ends

    try:
        for name in ("asgd", "sgd_sync"):
            sched = Scheduler(4, block_first_async_job=False)
            injector.calls = 0
            sched.fault_injector = injector
            cfg = _cfg(name, num_iterations=3000 if name == "asgd" else 30, step_size=0.5, bucket_ratio=0.5, quiet=True, printer_freq=1000)
            cfg.heartbeat_timeout_s = 0.4                 # generous: a loaded CI box must not declare HEALTHY workers dead
            cfg.eval_objective = False
            holder = {}
            runner = threading.Thread(target=lambda: holder.update(_try(lambda: algo.run(cfg, scheduler=sched))), daemon=True)
            runner.start()
            runner.join(60)
            assert not runner.is_alive()
            release.set()
            sched.stop()
            release.clear()
            if name == "asgd":
                rep = holder["ok"]
                assert rep.accepted == 3000 and rep.extra["dead_workers"] == [2] and hung["n"] >= 1
            else:
                assert isinstance(holder["err"], WorkerLost)
    finally:
        release.set()


def _try(fn):
    try:
        return {"ok": fn()}
    except BaseException as exc:  # noqa: BLE001
        return {"err": exc}


@pytest.mark.parametrize("name,step", [("sgd_sync", 0.5), ("saga_sync", 0.05)])
def test_host_driver_checkpoint_resume_is_exact_for_bsp(name, step, tmp_path):
    """Synchronous drivers are deterministic: 40 iterations == 20 iterations + checkpoint + resume for 20 more."""
    full = algo.run(_cfg(name, num_iterations=40, step_size=step, quiet=True))
    ckdir = str(tmp_path / "ck")
    algo.run(_cfg(name, num_iterations=20, step_size=step, quiet=True, checkpoint_dir=ckdir, checkpoint_every=10))
    from asyncframework_b200.utils import load_checkpoint
    ck = load_checkpoint(ckdir)
    assert ck["k"] == 20 and ck["config"]["algorithm"] == name
    if name == "saga_sync":
        assert ck["alpha_bar"] is not None and len(ck["history"]) == 4
    resumed = algo.run(_cfg(name, num_iterations=40, step_size=step, quiet=True, resume=ckdir))
    assert torch.allclose(resumed.final_weights, full.final_weights, rtol=0, atol=1e-12)


@pytest.mark.parametrize("name,step", [("asgd", 0.5), ("asaga", 0.05)])
def test_host_driver_checkpoint_resume_async(name, step, tmp_path):
    ckdir = str(tmp_path / "ck")
    first = algo.run(_cfg(name, num_iterations=150, step_size=step, quiet=True, checkpoint_dir=ckdir, checkpoint_every=50))
    from asyncframework_b200.utils import load_checkpoint
    ck = load_checkpoint(ckdir)
    assert ck["k"] in (100, 150) and ck["w"].numel() == 20
    resumed = algo.run(_cfg(name, num_iterations=300, step_size=step, quiet=True, resume=ckdir))
    assert resumed.accepted == 300
    assert resumed.objective[-1][1] <= first.objective[-1][1] * 1.5         # continued from the checkpoint (async runs are not bit-reproducible: loose factor)
    assert resumed.objective[0][1] < first.objective[0][1] * 0.5


@pytest.mark.parametrize("name", ["sgd_sync", "saga_sync"])
def test_bsp_driver_raises_when_a_task_exhausts_its_retries(name):
    """A BSP round whose task fails for good must raise on the driver (a blocking Spark job would), not wait forever."""
    import threading
    from asyncframework_b200.core import JobFailed, Scheduler
    sched = Scheduler(4, block_first_async_job=False, max_failures=2)
    sched.fault_injector = lambda t: (_ for _ in ()).throw(RuntimeError("disk on fire")) if (t.partition_id == 0 and t.job_id > 6) else None
    cfg = _cfg(name, num_iterations=30, step_size=0.05, quiet=True, printer_freq=1000)
    cfg.eval_objective = False
    holder = {}
    runner = threading.Thread(target=lambda: holder.update(_try(lambda: getattr(__import__("asyncframework_b200.algo", fromlist=[name]), name).run(cfg, scheduler=sched))), daemon=True)
    runner.start()
    runner.join(60)
    try:
        assert not runner.is_alive(), "driver hung on a permanently failed task"
        assert isinstance(holder.get("err"), JobFailed), holder
    finally:
        sched.stop()
