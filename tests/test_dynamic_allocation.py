"""Backlog-driven dynamic allocation (reference ExecutorAllocationManager.scala:228,292 and its suite
ExecutorAllocationManagerSuite: "add executors" ramps 1, 2, 4, ...; "remove executors" after the idle timeout; bounds)."""
import os
import time

import pytest

from asyncframework_b200.core.allocation import AllocationPolicy, ExecutorAllocationManager
from tests.test_scheduler_fake_backend import FakeTransport
from asyncframework_b200.core.cluster import ClusterScheduler
from asyncframework_b200.data.dataset import ShardedDataset


class ResizableTransport(FakeTransport):
    def __init__(self, n):
        super().__init__(n)
        self.requests = []

    def request_total_executors(self, total):
        with self.sched._lock:
            if self.sched._jobs:
                raise RuntimeError("request_total_executors while jobs are running would re-home their partitions")
        self.requests.append(total)
        self.sched.num_executors = total
        return True


class Clock:
    def __init__(self):
        self.t = 100.0

    def __call__(self):
        return self.t


def make(n=1, parts=16, **pol):
    t = ResizableTransport(n)
    s = ClusterScheduler(n, transport=t, use_env=False)
    clock = Clock()
    am = ExecutorAllocationManager(s, AllocationPolicy(**pol), clock=clock)
    s.allocation_manager = am                                    # no thread: the test drives tick()
    ds = ShardedDataset.from_records(s, list(range(4 * parts)), parts)
    return s, t, am, clock, ds


def test_policy_is_validated():
    with pytest.raises(ValueError):
        AllocationPolicy(min_executors=3, max_executors=2).validate()
    with pytest.raises(ValueError):
        AllocationPolicy(min_executors=1, max_executors=4, initial_executors=9).validate()


def test_sustained_backlog_ramps_the_target_exponentially_up_to_what_the_load_needs():
    s, t, am, clock, ds = make(1, 16, min_executors=1, max_executors=10, backlog_timeout=1.0, sustained_backlog_timeout=0.5)
    w = s.submit_job(ds, lambda tc, it: 0, result_handler=lambda i, r: None)
    t.started(0, w.job_id, 0)                                   # 16 pending, 1 running -> backlog 15, 16 executors wanted (cap 10)
    assert am.load() == {"pending": 16, "running": 1, "backlog": 15, "busy": {0}, "jobs": 1}
    assert am.tick() == 1                                        # backlog seen: the timer starts
    clock.t += 0.9
    assert am.tick() == 1                                        # not yet backlogged for schedulerBacklogTimeout
    clock.t += 0.2
    assert am.tick() == 2                                        # +1
    clock.t += 0.3
    assert am.tick() == 2                                        # sustained timeout not reached
    clock.t += 0.3
    assert am.tick() == 4                                        # +2
    clock.t += 0.6
    assert am.tick() == 8                                        # +4
    clock.t += 0.6
    assert am.tick() == 10                                       # +8 would overshoot: capped at maxExecutors
    clock.t += 0.6
    assert am.tick() == 10
    assert t.requests == []                                      # a job is in flight: slots cannot be re-homed now
    for p in range(16):                                          # the job drains
        t.started(0, w.job_id, p)
        t.ok(0, w.job_id, p, 0)
    assert w.done
    # the NEXT job picks the target up before it is routed
    w2 = s.submit_job(ds, lambda tc, it: 0, result_handler=lambda i, r: None)
    assert t.requests == [10] and s.num_executors == 10
    assert {e for e, _, _ in t.tasks(w2.job_id)} == set(range(10))
    assert [h["event"] for h in am.history].count("applied") == 1
    s.stop()


def test_target_never_exceeds_the_load_and_backlog_timer_resets():
    s, t, am, clock, ds = make(1, 3, min_executors=1, max_executors=50, backlog_timeout=1.0, sustained_backlog_timeout=1.0)
    w = s.submit_job(ds, lambda tc, it: 0, result_handler=lambda i, r: None)
    am.tick()
    clock.t += 1.1
    assert am.tick() == 2
    clock.t += 1.1
    assert am.tick() == 3                                        # +2 clipped to the 3 tasks that exist
    clock.t += 1.1
    assert am.tick() == 3
    for p in range(3):
        t.started(0, w.job_id, p)
    assert am.load()["backlog"] == 0
    am.tick()
    assert am._add_time is None and am._to_add == 1              # no backlog: the ramp starts over next time
    s.stop()


def test_idle_executors_are_released_down_to_the_minimum_when_the_cluster_is_idle():
    s, t, am, clock, ds = make(6, 6, min_executors=2, max_executors=8, idle_timeout=30.0)
    am.target = 6
    assert am.tick() == 6 and t.requests == []
    clock.t += 29.0
    assert am.tick() == 6
    clock.t += 2.0
    assert am.tick() == 2 and t.requests == [2] and s.num_executors == 2       # all six idle > 30 s: floor is minExecutors
    clock.t += 100.0
    assert am.tick() == 2 and t.requests == [2]
    s.stop()


def test_busy_executors_keep_their_slot_and_partial_idleness_shrinks_partially():
    s, t, am, clock, ds = make(4, 2, min_executors=1, max_executors=8, idle_timeout=10.0)
    am.target = 4
    w = s.submit_job(ds, lambda tc, it: 0, result_handler=lambda i, r: None)   # 2 partitions -> slots 0 and 1 busy, 2 and 3 idle
    t.started(0, w.job_id, 0)
    t.started(1, w.job_id, 1)
    am.tick()
    clock.t += 11.0
    assert am.tick() == 2                                        # two idle slots expired; the two busy ones stay
    assert t.requests == []                                      # applied only once the job is out of the way
    t.ok(0, w.job_id, 0, 0)
    t.ok(1, w.job_id, 1, 0)
    am.tick()
    assert t.requests == [2]
    s.stop()


def test_a_refused_request_is_recorded_and_retried():
    s, t, am, clock, ds = make(2, 2, min_executors=1, max_executors=4, idle_timeout=1.0)
    t.request_total_executors = lambda total: False
    clock.t += 2.0
    am.tick()
    clock.t += 2.0
    assert am.tick() == 1 and am.failed_requests >= 1 and s.num_executors == 2
    del t.request_total_executors
    am.tick()
    assert s.num_executors == 1 and am.applied == 1
    s.stop()


@pytest.mark.timeout(120)
def test_policy_loop_grows_a_standalone_application_under_backlog_and_shrinks_it_when_idle():
    """End to end on the standalone cluster manager: one executor, eight slow tasks per job -> the loop raises the target
    while the first job runs, the second job runs on more executor processes; after the idle timeout the app shrinks."""
    from asyncframework_b200.conf import AsyncConf
    from asyncframework_b200 import AsyncSession
    from asyncframework_b200.deploy import LocalStandaloneCluster
    with LocalStandaloneCluster(num_workers=4) as c:
        conf = (AsyncConf(load_defaults=False).set_master(c.master_url).set("spark.executor.instances", "1")
                .set("spark.dynamicAllocation.enabled", "true").set("spark.dynamicAllocation.maxExecutors", "4")
                .set("spark.dynamicAllocation.schedulerBacklogTimeout", "100ms")
                .set("spark.dynamicAllocation.sustainedSchedulerBacklogTimeout", "100ms")
                .set("spark.dynamicAllocation.executorIdleTimeout", "1500ms"))
        with AsyncSession(conf=conf) as sc:
            am = sc.scheduler.allocation_manager
            assert am is not None and sc.scheduler.num_executors == 1
            ds = sc.parallelize(range(32), 8)

            def slow(x):
                time.sleep(0.02)
                return os.getpid()
            first = set(ds.map(slow).collect())
            assert len(first) == 1 and am.target >= 3                       # ramped 1 -> 2 -> 4 while the job was backlogged
            second = set(ds.map(slow).collect())
            assert sc.scheduler.num_executors == am.target >= 3 and len(second) >= 3
            assert ds.map(lambda x: x).reduce(lambda a, b: a + b) == sum(range(32))
            t0 = time.time()
            while sc.scheduler.num_executors > 1 and time.time() - t0 < 20:
                time.sleep(0.1)
            assert sc.scheduler.num_executors == 1                          # idle: back to minExecutors
            assert len(set(ds.map(slow).collect())) == 1
            events = [h["event"] for h in am.history]
            assert events.count("applied") >= 2


def test_short_tasks_between_ticks_keep_executors_from_looking_idle():
    """Tasks that start and finish between two policy passes never appear in the job table when the policy looks; the idle
    timer must still be reset by their launches / status messages (ExecutorAllocationManager resets it in onTaskStart)."""
    t = ResizableTransport(3)
    s = ClusterScheduler(3, transport=t, use_env=False)
    am = ExecutorAllocationManager(s, AllocationPolicy(min_executors=1, max_executors=3, idle_timeout=0.4))     # real monotonic clock
    s.allocation_manager = am
    am.target = 3
    ds = ShardedDataset.from_records(s, list(range(6)), 3)
    t_end = time.monotonic() + 1.2
    while time.monotonic() < t_end:                                  # 1.2 s of activity, three times the idle timeout
        w = s.submit_job(ds, lambda tc, it: 0, result_handler=lambda i, r: None)
        for p in range(3):
            t.started(p, w.job_id, p)
            t.ok(p, w.job_id, p, 0)
        assert w.done
        am.tick()                                                    # the job table is empty every time the policy looks
        time.sleep(0.05)
    assert am.target == 3 and t.requests == []
    t0 = time.monotonic()
    while am.tick() == 3 and time.monotonic() - t0 < 3.0:            # activity stops: now the timer runs out
        time.sleep(0.05)
    assert am.target == 1 and t.requests == [1] and 0.3 < time.monotonic() - t0 < 2.0
    s.stop()
