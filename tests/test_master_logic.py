"""Master scheduling / lifecycle logic with fake workers and a fake driver — no processes (MasterSuite scenarios:
core/src/test/scala/org/apache/spark/deploy/master/MasterSuite.scala "basic scheduling - spread out / no spread out",
executor retry cap, worker re-registration)."""
import time

from asyncframework_b200.deploy.master import MAX_EXECUTOR_RETRIES, Master
from asyncframework_b200.rpc import RpcEndpoint, RpcEnv


class Recorder(RpcEndpoint):
    def __init__(self):
        self.msgs = []

    def receive(self, msg, sender):
        self.msgs.append(msg)


def _wait(pred, timeout=5.0):
    t0 = time.time()
    while time.time() - t0 < timeout:
        if pred():
            return True
        time.sleep(0.01)
    return False


def _setup(spread_out=True, workers=((4, 4096), (4, 4096), (4, 4096))):
    env = RpcEnv("m")
    master = Master(worker_timeout=60, spread_out=spread_out)
    mref = env.setup_endpoint("Master", master)
    recs = []
    for i, (cores, mem) in enumerate(workers):
        r = Recorder()
        ref = env.setup_endpoint(f"w{i}", r)
        assert mref.ask_sync(("register_worker", f"w{i}", "127.0.0.1", 10_000 + i, cores, mem, ref))[0] == "registered"
        recs.append(r)
    drv = Recorder()
    dref = env.setup_endpoint("drv", drv)
    return env, master, mref, recs, drv, dref


def _launches(rec):
    return [m for m in rec.msgs if m[0] == "launch_executor"]


def test_spread_out_vs_consolidated_placement_and_resource_limits():
    env, master, mref, recs, drv, dref = _setup(spread_out=True)
    try:
        app_id, granted = mref.ask_sync(("register_app", "a", 3, 2, 1024, dref, "127.0.0.1:1"))
        assert granted == 3 and _wait(lambda: sum(len(_launches(r)) for r in recs) == 3)
        assert [len(_launches(r)) for r in recs] == [1, 1, 1]                         # one per worker
        assert sorted(m[3] for r in recs for m in _launches(r)) == [0, 1, 2]          # slots 0..2
        st = mref.ask_sync(("state",))
        assert [w["cores_used"] for w in st["workers"]] == [2, 2, 2]
        # a second application only gets what is left: 3 workers x 2 free cores, executors of 2 cores -> 3 of the 5 asked
        app2, _ = mref.ask_sync(("register_app", "b", 5, 2, 1024, dref, "127.0.0.1:1"))
        assert _wait(lambda: sum(len(_launches(r)) for r in recs) == 6)
        time.sleep(0.1)
        assert sum(len(_launches(r)) for r in recs) == 6                              # 2 slots of app b wait for resources
        mref.send(("unregister_app", app_id))                                         # app a leaves: its cores go to app b
        assert _wait(lambda: sum(len(_launches(r)) for r in recs) == 8)
        assert all(any(m[0] == "kill_executor" and m[1] == app_id for m in r.msgs) for r in recs)
    finally:
        env.shutdown()
    env, master, mref, recs, drv, dref = _setup(spread_out=False)
    try:
        mref.ask_sync(("register_app", "a", 3, 2, 1024, dref, "127.0.0.1:1"))
        assert _wait(lambda: sum(len(_launches(r)) for r in recs) == 3)
        assert [len(_launches(r)) for r in recs] == [2, 1, 0]                         # pack worker 0 (4 cores) first
    finally:
        env.shutdown()


def test_failed_executor_is_relaunched_until_the_retry_cap_then_the_app_fails():
    env, master, mref, recs, drv, dref = _setup(workers=((2, 2048),))
    try:
        app_id, _ = mref.ask_sync(("register_app", "crashy", 1, 1, 512, dref, "127.0.0.1:1"))
        assert _wait(lambda: len(_launches(recs[0])) == 1)
        for n in range(1, MAX_EXECUTOR_RETRIES + 1):
            exec_id = _launches(recs[0])[-1][2]
            mref.send(("executor_state_changed", app_id, exec_id, "EXITED", "exit code 1", 1))     # it never reaches RUNNING
            if n < MAX_EXECUTOR_RETRIES:
                assert _wait(lambda: len(_launches(recs[0])) == n + 1), n
        assert _wait(lambda: ("application_removed", "FAILED") in drv.msgs)
        st = mref.ask_sync(("state",))
        assert st["active_apps"] == [] and st["completed_apps"][-1]["state"] == "FAILED"
        assert st["workers"][0]["cores_used"] == 0
        # a RUNNING report resets the counter: an executor that came up and died later is an ordinary loss
        app2, _ = mref.ask_sync(("register_app", "ok", 1, 1, 512, dref, "127.0.0.1:1"))
        assert _wait(lambda: any(m[1] == app2 for m in _launches(recs[0])))
        for _ in range(MAX_EXECUTOR_RETRIES + 3):
            exec_id = [m for m in _launches(recs[0]) if m[1] == app2][-1][2]
            mref.send(("executor_state_changed", app2, exec_id, "RUNNING", "", None))
            mref.send(("executor_state_changed", app2, exec_id, "EXITED", "exit code 137", 137))
            n_before = len([m for m in _launches(recs[0]) if m[1] == app2])
            assert _wait(lambda: len([m for m in _launches(recs[0]) if m[1] == app2]) == n_before + 1 or True)
        time.sleep(0.2)
        assert any(a["id"] == app2 for a in mref.ask_sync(("state",))["active_apps"])
    finally:
        env.shutdown()


def test_duplicate_worker_id_is_refused_and_lost_worker_executors_move():
    env, master, mref, recs, drv, dref = _setup(workers=((2, 2048), (2, 2048)))
    try:
        dup = mref.ask_sync(("register_worker", "w0", "127.0.0.1", 10_000, 2, 2048, dref))
        assert dup[0] == "failed" and "duplicate" in dup[1]
        app_id, _ = mref.ask_sync(("register_app", "a", 2, 1, 512, dref, "127.0.0.1:1"))
        assert _wait(lambda: [len(_launches(r)) for r in recs] == [1, 1])
        for r in recs:
            m = _launches(r)[0]
            mref.send(("executor_state_changed", app_id, m[2], "RUNNING", "", None))
        master.workers["w1"].last_heartbeat -= 1000                                   # w1 goes silent
        mref.send(("check_worker_timeouts",))
        assert _wait(lambda: len(_launches(recs[0])) == 2)                            # its slot is re-placed on w0
        assert _launches(recs[0])[-1][3] == _launches(recs[1])[0][3]                   # same SLOT number
        assert any(m[0] == "executor_updated" and m[3] == "LOST" for m in drv.msgs)
        assert [w["id"] for w in mref.ask_sync(("state",))["workers"]] == ["w0"]
    finally:
        env.shutdown()


def test_worker_that_comes_back_with_superseded_executors_is_told_to_kill_them():
    env, master, mref, recs, drv, dref = _setup(workers=((2, 2048), (2, 2048)))
    try:
        app_id, _ = mref.ask_sync(("register_app", "a", 2, 1, 512, dref, "127.0.0.1:1"))
        assert _wait(lambda: [len(_launches(r)) for r in recs] == [1, 1])
        old = _launches(recs[1])[0]
        mref.send(("executor_state_changed", app_id, old[2], "RUNNING", "", None))
        master.workers["w1"].last_heartbeat -= 1000
        mref.send(("check_worker_timeouts",))                                  # w1 is declared dead, its slot moves to w0
        assert _wait(lambda: len(_launches(recs[0])) == 2)
        # w1 was only partitioned away: it re-registers and reports the executor it is STILL running
        w1ref = env.endpoint_ref(None, "w1")
        reply = mref.ask_sync(("register_worker", "w1", "127.0.0.1", 10_001, 2, 2048, w1ref, [(app_id, old[2], old[3], 1, 512)]))
        assert reply[0] == "registered"
        assert _wait(lambda: ("kill_executor", app_id, old[2], False) in recs[1].msgs)       # the duplicate is killed ...
        st = mref.ask_sync(("state",))
        assert [w["cores_used"] for w in sorted(st["workers"], key=lambda w: w["id"])] == [2, 0]   # ... and not accounted for
    finally:
        env.shutdown()
