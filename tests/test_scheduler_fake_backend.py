"""Cluster-scheduler logic driven by hand through a FAKE transport — no processes, no sockets, no sleeps.

The reference tests its DAGScheduler/TaskSetManager the same way: ``DAGSchedulerSuite`` swaps the TaskScheduler for a
stub that records TaskSets and drives completions with ``runEvent(makeCompletionEvent(...))``
(core/src/test/scala/org/apache/spark/scheduler/DAGSchedulerSuite.scala:134-160,291-330); ``SchedulerIntegrationSuite``
uses a ``MockBackend``.  Here ``FakeTransport`` records every message the scheduler sends to an executor slot and the
test plays the executors' status messages back in whatever order it wants."""
import pickle

import cloudpickle
import pytest

from asyncframework_b200.core.cluster import ClusterScheduler
from asyncframework_b200.core.scheduler import JobCancelled, JobFailed
from asyncframework_b200.data.dataset import ShardedDataset


class FakeTransport:
    def __init__(self, n):
        self.n, self.sent, self.killed, self.relaunched = n, [], [], []
        self.relaunch_ok = True

    def start(self, sched):
        self.sched = sched

    def send(self, eid, msg):
        self.sent.append((eid, msg))

    def relaunch(self, eid):
        self.relaunched.append(eid)
        return self.relaunch_ok

    def kill(self, eid):
        self.killed.append(eid)

    def stop(self):
        pass

    # ---- helpers for the tests ---------------------------------------------------------------------------------------
    def tasks(self, job_id=None):
        return [(e, m[2], m[3]) for e, m in self.sent if m[0] == "task" and (job_id is None or m[1] == job_id)]   # (executor, partition, attempt)

    def binaries(self):
        return [(e, m[1]) for e, m in self.sent if m[0] == "binary"]

    def ok(self, eid, job_id, p, value, attempt=0, run_ms=1.0):
        self.sched._on_message(eid, ("ok", job_id, p, attempt, cloudpickle.dumps((value, {})), {"run_ms": run_ms, "serialize_ms": 0.0}))

    def err(self, eid, job_id, p, attempt=0, text="Boom('x')\ntraceback"):
        self.sched._on_message(eid, ("err", job_id, p, attempt, text, 0.0))

    def started(self, eid, job_id, p, attempt=0):
        self.sched._on_message(eid, ("started", job_id, p, attempt, "", 0.0))


def make(n=2, **kw):
    t = FakeTransport(n)
    s = ClusterScheduler(n, transport=t, use_env=False, **kw)
    ds = ShardedDataset.from_records(s, list(range(8)), 4)
    return s, t, ds


def test_submission_sends_one_binary_per_executor_then_tasks_with_hard_affinity():
    s, t, ds = make()
    events = []
    s.listeners.add(events.append)
    got = {}
    w = s.submit_job(ds, lambda tc, it: sum(it), result_handler=got.__setitem__)
    assert sorted(t.binaries()) == [(0, w.job_id), (1, w.job_id)]
    assert t.tasks() == [(0, 0, 0), (1, 1, 0), (0, 2, 0), (1, 3, 0)]                      # partition p -> executor p % 2
    assert t.sent.index((0, t.sent[0][1])) < [i for i, (e, m) in enumerate(t.sent) if m[0] == "task"][0]   # binary before tasks
    for e, p, _ in t.tasks():
        t.ok(e, w.job_id, p, p * 10)
    assert w.done and got == {0: 0, 1: 10, 2: 20, 3: 30}
    assert [m for _, m in t.sent if m[0] == "drop"] == [("drop", w.job_id)] * 2           # executors forget the binary
    assert [e["event"] for e in events if e["event"].startswith("job")] == ["job_start", "job_end"]
    t.ok(1, w.job_id, 3, 999)                                                             # a straggling duplicate is ignored
    assert got[3] == 30
    s.stop()


def test_retry_on_same_executor_then_job_failure_after_max_failures():
    s, t, ds = make(max_failures=3)
    lost = []
    w = s.submit_job(ds, lambda tc, it: 0, result_handler=lambda i, r: None, on_task_lost=lambda p, exc: lost.append(p))
    t.err(1, w.job_id, 1, attempt=0)
    assert t.tasks()[-1] == (1, 1, 1)                                                      # retried where its data lives
    t.err(1, w.job_id, 1, attempt=1)
    assert t.tasks()[-1] == (1, 1, 2)
    t.err(1, w.job_id, 1, attempt=2)                                                       # third failure: give up
    with pytest.raises(JobFailed, match="failed 3 times"):
        w.wait(0)
    assert lost == [1] and s.failed_jobs == [w.job_id]
    s.stop()


def test_blacklisted_executor_loses_its_retries_and_new_jobs():
    s, t, ds = make(max_failures=4, blacklist=True, blacklist_max_failures=2, blacklist_timeout=1000.0)
    w = s.submit_job(ds, lambda tc, it: 0, result_handler=lambda i, r: None)
    t.err(1, w.job_id, 1)
    assert t.tasks()[-1] == (1, 1, 1) and not s.is_blacklisted(1)
    t.err(1, w.job_id, 3)                                                                  # second failed task on executor 1
    assert s.is_blacklisted(1) and t.tasks()[-1] == (0, 3, 1)                              # retry re-routed to executor 0
    w2 = s.submit_job(ds, lambda tc, it: 0, result_handler=lambda i, r: None)
    assert {e for e, _, _ in t.tasks(w2.job_id)} == {0} and t.binaries().count((1, w2.job_id)) == 0
    s._blacklisted[1] = 0.0                                                                # timeout elapses
    w3 = s.submit_job(ds, lambda tc, it: 0, result_handler=lambda i, r: None)
    assert {e for e, _, _ in t.tasks(w3.job_id)} == {0, 1}
    s.stop()


def test_speculation_needs_running_tasks_a_quantile_and_a_slow_outlier():
    s, t, ds = make(speculation=False, speculation_quantile=0.5, speculation_multiplier=2.0, speculation_min_ms=0.0)
    got = {}
    w = s.submit_job(ds, lambda tc, it: 0, result_handler=got.__setitem__)
    job = s._jobs[w.job_id]
    s._check_speculatable(w.job_id, job)
    assert s.speculated_tasks == 0                                                         # nothing finished yet
    for e, p in ((0, 0), (1, 1), (0, 2)):
        t.started(e, w.job_id, p)
        t.ok(e, w.job_id, p, p, run_ms=10.0)
    s._check_speculatable(w.job_id, job)
    assert s.speculated_tasks == 0                                                         # partition 3 is queued, not running
    t.started(1, w.job_id, 3)
    job["started"][3] -= 1.0                                                               # ... and has been running for 1 s >> 2 x 10 ms
    s._check_speculatable(w.job_id, job)
    assert s.speculated_tasks == 1 and t.tasks()[-1] == (0, 3, 1000)                       # copy on the OTHER executor
    s._check_speculatable(w.job_id, job)
    assert s.speculated_tasks == 1                                                         # at most one copy per task
    t.ok(0, w.job_id, 3, "copy", attempt=1000)
    assert w.done and got[3] == "copy" and s.speculative_wins == 1
    # the losing ORIGINAL is told to stop (TaskSetManager.scala:736 killTask on the other attempts)
    assert [(e, m[1:]) for e, m in t.sent if m[0] == "kill_task"] == [(1, (w.job_id, 3, 0))] and s.speculative_kills_sent == 1
    t.ok(1, w.job_id, 3, "original")                                                       # loser arrives late: dropped
    assert got[3] == "copy"
    seen = []
    s.listeners.add(seen.append)
    s._on_message(1, ("killed", w.job_id, 3, 0, "interrupted", 12.0))                      # ... or acknowledges the kill
    assert s.tasks_killed == 1 and seen[-1]["event"] == "task_killed" and seen[-1]["executor"] == 1
    s.stop()


def test_original_winning_kills_the_speculative_copy():
    s, t, ds = make(speculation=False, speculation_quantile=0.5, speculation_multiplier=2.0, speculation_min_ms=0.0)
    got = {}
    w = s.submit_job(ds, lambda tc, it: 0, result_handler=got.__setitem__)
    job = s._jobs[w.job_id]
    for e, p in ((0, 0), (1, 1), (0, 2)):
        t.started(e, w.job_id, p)
        t.ok(e, w.job_id, p, p, run_ms=10.0)
    t.started(1, w.job_id, 3)
    job["started"][3] -= 1.0
    s._check_speculatable(w.job_id, job)
    assert t.tasks()[-1] == (0, 3, 1000)
    t.ok(1, w.job_id, 3, "original")                                                       # the original finishes first after all
    assert got[3] == "original" and s.speculative_wins == 0
    assert [(e, m[1:]) for e, m in t.sent if m[0] == "kill_task"] == [(0, (w.job_id, 3, 1000))]
    s.stop()


def test_lost_executor_fails_or_recovers_depending_on_policy():
    s, t, ds = make()
    lost = []
    w = s.submit_job(ds, lambda tc, it: 0, result_handler=lambda i, r: None, on_task_lost=lambda p, e: lost.append(p))
    t.ok(1, w.job_id, 1, 0)
    s._executor_lost(1)                                                                    # policy: fail what was in flight there
    assert lost == [3] and w.done
    with pytest.raises(JobFailed, match="executor 1 lost"):
        w.wait(0)
    s.stop()

    s, t, ds = make(recover_lost_executors=True)
    got = {}
    w = s.submit_job(ds, lambda tc, it: 0, result_handler=got.__setitem__)
    t.ok(1, w.job_id, 1, "done-before-loss")
    n_bin = len(t.binaries())
    s._executor_lost(1)
    assert t.relaunched == [1] and len(t.binaries()) == n_bin + 1                          # fresh binary for the replacement
    assert t.tasks()[-1] == (1, 3, 0)                                                      # only the unfinished partition is re-run
    for e, p in ((0, 0), (0, 2), (1, 3)):
        t.ok(e, w.job_id, p, p)
    assert w.done and got == {1: "done-before-loss", 0: 0, 2: 2, 3: 3} and s.executors_recovered == 1
    t.relaunch_ok = False                                                                  # the cluster manager cannot replace it
    w2 = s.submit_job(ds, lambda tc, it: 0, result_handler=lambda i, r: None)
    s._executor_lost(0)
    with pytest.raises(JobFailed, match="executor 0 lost"):
        w2.wait(0)
    s.stop()


def test_cancel_drops_binaries_and_ignores_late_results():
    s, t, ds = make()
    lost, got = [], {}
    w = s.submit_job(ds, lambda tc, it: 0, result_handler=got.__setitem__, on_task_lost=lambda p, e: lost.append(p))
    t.ok(0, w.job_id, 0, "early")
    assert s.cancel_job(w.job_id)
    assert sorted(lost) == [1, 2, 3] and [m for _, m in t.sent if m[0] == "drop"] == [("drop", w.job_id)] * 2
    t.ok(1, w.job_id, 1, "late")
    assert got == {0: "early"}
    with pytest.raises(JobCancelled):
        w.wait(0)
    assert not s.cancel_job(w.job_id)
    s.stop()


def test_result_handler_exception_fails_the_job_not_the_scheduler():
    s, t, ds = make()

    def handler(i, r):
        raise ValueError("handler bug")
    w = s.submit_job(ds, lambda tc, it: 0, result_handler=handler)
    t.ok(0, w.job_id, 0, 1)
    with pytest.raises(JobFailed, match="handler bug"):
        w.wait(0)
    w2 = s.submit_job(ds, lambda tc, it: 0, result_handler=lambda i, r: None)            # the scheduler is still usable
    for e, p, _ in t.tasks(w2.job_id):
        t.ok(e, w2.job_id, p, 0)
    assert w2.wait(0)
    s.stop()


def test_unreachable_executor_fails_its_tasks_and_does_not_leave_the_job_dangling():
    s, t, ds = make()
    real_send = t.send

    def flaky_send(eid, msg):
        if eid == 1:
            raise BrokenPipeError("executor 1 pipe is gone")
        real_send(eid, msg)
    t.send = flaky_send
    lost = []
    w = s.submit_job(ds, lambda tc, it: 0, result_handler=lambda i, r: None, on_task_lost=lambda p, e: lost.append(p))
    assert sorted(lost) == [1, 3] and w.done
    with pytest.raises(JobFailed, match="unreachable"):
        w.wait(0)
    for e, p in ((0, 0), (0, 2)):
        t.ok(e, w.job_id, p, 0)                         # the healthy executor's results are simply late
    assert s.active_job_ids() == []                     # nothing dangling
    t.send = real_send
    w2 = s.submit_job(ds, lambda tc, it: 0, result_handler=lambda i, r: None)
    for e, p, _ in t.tasks(w2.job_id):
        t.ok(e, w2.job_id, p, 0)
    assert w2.wait(0)
    s.stop()


def test_failed_job_is_forgotten_once_nothing_is_in_flight_and_announces_its_end_once():
    s, t, ds = make(max_failures=1)
    events = []
    s.listeners.add(lambda e: events.append(e) if e["event"] == "job_end" else None)
    w = s.submit_job(ds, lambda tc, it: 0, result_handler=lambda i, r: None)
    for e, p in ((0, 0), (1, 1), (0, 2)):
        t.ok(e, w.job_id, p, 0)
    t.err(1, w.job_id, 3)                                   # the LAST outstanding task fails for good
    with pytest.raises(JobFailed):
        w.wait(0)
    assert w.job_id not in s._jobs                          # not leaked
    assert len(events) == 1 and events[0].get("failed") is True
    assert ("drop", w.job_id) in [m for _, m in t.sent]     # executors were told to forget the binary
    s.stop()
