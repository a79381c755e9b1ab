"""Standalone cluster manager — the Master daemon.

Reference: ``Master`` (core/deploy/master/Master.scala:41): workers register and heartbeat; applications register
with a resource request; the master places executors on workers (``schedule`` / ``startExecutorsOnWorkers``,
spread-out by default), tells the worker to ``LaunchExecutor``, relays ``ExecutorStateChanged`` to the
application's driver, re-launches executors that exit abnormally (bounded by ``MAX_EXECUTOR_RETRIES``), times out
silent workers (``spark.worker.timeout`` = 60 s, ``CheckForWorkerTimeOut``) and cleans up when a driver
disconnects.  README.md:56 runs the paper's experiments on such a 1-master + 8-worker standalone cluster.

Run it: ``python -m asyncframework_b200.deploy.master --host 0.0.0.0 --port 7077``; applications connect with the
master URL ``spark://host:7077`` (``async://`` is accepted as an alias).

Executors have a *slot* (0..n-1): the driver binds partition p to slot p % n (hard shard↔worker affinity, SURVEY
S4), so when an executor dies the master re-fills the SAME slot and the driver re-ships that slot's data.
"""
from __future__ import annotations

import argparse
import itertools
import threading
import time
from dataclasses import dataclass, field
from typing import Dict, List, Optional

from ..rpc import RpcAddress, RpcCallContext, RpcEndpoint, RpcEndpointRef, RpcEnv, RpcError

MAX_EXECUTOR_RETRIES = 10


@dataclass
class WorkerInfo:
    id: str
    host: str
    port: int
    cores: int
    memory_mb: int
    ref: RpcEndpointRef
    cores_used: int = 0
    memory_used: int = 0
    last_heartbeat: float = field(default_factory=time.time)
    state: str = "ALIVE"                      # ALIVE | DEAD
    executors: Dict[str, int] = field(default_factory=dict)       # "app/exec" -> cores


@dataclass
class ExecutorDesc:
    exec_id: int
    slot: int
    worker_id: str
    cores: int
    memory_mb: int
    state: str = "LAUNCHING"                  # LAUNCHING | RUNNING | EXITED | KILLED | FAILED | LOST


@dataclass
class AppInfo:
    id: str
    name: str
    driver: RpcEndpointRef
    driver_url: str
    num_executors: int
    cores_per_executor: int
    memory_per_executor: int
    submit_time: float = field(default_factory=time.time)
    state: str = "RUNNING"                    # RUNNING | FINISHED | FAILED
    executors: Dict[int, ExecutorDesc] = field(default_factory=dict)
    retries: int = 0
    next_exec_id: int = 0


class Master(RpcEndpoint):
    NAME = "Master"

    def __init__(self, worker_timeout: float = 60.0, spread_out: bool = True, recovery_dir: Optional[str] = None,
                 recovery_grace: float = 2.0, ha: bool = False, lease_s: float = 3.0):
        # HA (the ZooKeeper mode of the reference, ``spark.deploy.recoveryMode=ZOOKEEPER`` + standby masters,
        # deploy/master/ZooKeeperLeaderElectionAgent.scala) with a shared directory instead of ZooKeeper: the leader holds a
        # LEASE file in ``recovery_dir`` (renewed every lease/3 under an flock); standbys answer registrations with
        # ("standby", leader url) and take over — RECOVERING, then ALIVE — when the lease has expired.  Workers and drivers are
        # given every master address (``spark://h1:p1,h2:p2``) and walk the list.
        self.ha = bool(ha and recovery_dir)
        self.lease_s = float(lease_s)
        self.leader_url: Optional[str] = None
        self.worker_timeout = worker_timeout
        self.spread_out = spread_out
        # FILESYSTEM recovery mode (spark.deploy.recoveryMode=FILESYSTEM, FileSystemPersistenceEngine): registrations are
        # persisted; a restarted master on the same address starts RECOVERING — workers and drivers notice the dropped
        # connection, re-register (workers report the executors they are still running, drivers their application id and
        # slots), the master adopts them, and only after the grace period does it schedule again, so nothing that is
        # still alive gets launched twice.
        self.recovery_dir = recovery_dir
        self.recovery_grace = recovery_grace
        self.recovering_until = 0.0
        self.status = "ALIVE"
        self._reported: Dict[tuple, tuple] = {}       # (app_id, exec_id) -> (worker_id, slot, cores, memory) reported by workers
        self.recovered = {"workers": 0, "apps": 0, "executors": 0}
        self.workers: Dict[str, WorkerInfo] = {}
        self.apps: Dict[str, AppInfo] = {}
        self.completed_apps: List[AppInfo] = []
        self._app_ids = itertools.count(0)
        self._next_app_id = 0
        self._stop = threading.Event()
        self.start_time = time.time()

    # ------------------------------------------------------------------ lifecycle
    def on_start(self) -> None:
        threading.Thread(target=self._timeout_loop, daemon=True, name="master-worker-timeout").start()
        if self.ha:
            self.status = "STANDBY"
            threading.Thread(target=self._election_loop, daemon=True, name="master-election").start()
            return
        self._become_leader()

    def _lease(self, acquire: bool) -> bool:
        """Read-check-write of the lease file under an exclusive flock.  Returns True when this master holds the lease after
        the call (``acquire=False`` only renews a lease that is already ours)."""
        import fcntl
        import json
        import os
        os.makedirs(self.recovery_dir, exist_ok=True)
        me = f"spark://{self.rpc_env.address}"
        with open(os.path.join(self.recovery_dir, "leader.lock"), "a+") as lock:
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                path = os.path.join(self.recovery_dir, "leader.json")
                cur = None
                try:
                    with open(path) as f:
                        cur = json.load(f)
                except (OSError, ValueError):
                    pass
                now = time.time()
                mine = cur is not None and cur.get("url") == me
                free = cur is None or cur.get("expires", 0) < now
                self.leader_url = cur.get("url") if cur and not free else None
                if mine or (acquire and free):
                    tmp = path + f".{os.getpid()}.tmp"
                    with open(tmp, "w") as f:
                        json.dump({"url": me, "expires": now + self.lease_s}, f)
                    os.replace(tmp, path)
                    self.leader_url = me
                    return True
                return False
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)

    def _election_loop(self) -> None:
        while not self._stop.wait(self.lease_s / 3):
            try:
                if self.status == "STANDBY":
                    if self._lease(acquire=True):
                        self.self_ref.send(("elected",))
                elif not self._lease(acquire=False):               # we lost the lease (paused too long): step down hard
                    self.self_ref.send(("revoked",))
            except (OSError, RpcError):
                continue

    def _become_leader(self) -> None:
        self.status = "ALIVE"
        prior = self._load_persisted()
        if prior is not None:
            self._app_ids = itertools.count(prior.get("next_app", 0))
            if prior.get("apps") or prior.get("workers"):
                self.status = "RECOVERING"
                self.recovering_until = time.time() + self.recovery_grace
                self.expected = {"workers": sorted(prior.get("workers", [])), "apps": sorted(prior.get("apps", []))}

                def done():
                    time.sleep(self.recovery_grace)
                    try:
                        self.self_ref.send(("complete_recovery",))
                    except RpcError:
                        pass
                threading.Thread(target=done, daemon=True, name="master-recovery").start()

    # ------------------------------------------------------------------ persistence
    def _state_file(self) -> Optional[str]:
        import os
        if not self.recovery_dir:
            return None
        os.makedirs(self.recovery_dir, exist_ok=True)
        return os.path.join(self.recovery_dir, "master-state.json")

    def _persist(self) -> None:
        path = self._state_file()
        if path is None:
            return
        import json
        import os
        snap = {"next_app": self._next_app_id, "workers": sorted(self.workers), "apps": sorted(self.apps), "t": time.time()}
        tmp = path + ".tmp"
        with open(tmp, "w") as f:
            json.dump(snap, f)
        os.replace(tmp, path)

    def _load_persisted(self) -> Optional[dict]:
        path = self._state_file()
        if path is None:
            return None
        import json
        import os
        if not os.path.exists(path):
            return None
        try:
            with open(path) as f:
                return json.load(f)
        except (OSError, ValueError):
            return None

    def on_stop(self) -> None:
        self._stop.set()

    def _timeout_loop(self) -> None:
        while not self._stop.wait(max(0.05, self.worker_timeout / 4)):
            try:
                self.self_ref.send(("check_worker_timeouts",))
            except RpcError:
                return

    # ------------------------------------------------------------------ messages
    def receive(self, msg, sender: Optional[RpcAddress]) -> None:
        op = msg[0]
        if self.status == "STANDBY" and op in ("heartbeat", "executor_state_changed", "unregister_app"):
            return
        if op == "heartbeat":
            w = self.workers.get(msg[1])
            if w is not None and w.state == "ALIVE":
                w.last_heartbeat = time.time()
            elif sender is not None:                     # a worker we forgot (master restart / expiry): ask it to re-register
                try:
                    self.rpc_env.endpoint_ref(sender, "Worker").send(("reconnect",))
                except RpcError:
                    pass
        elif op == "check_worker_timeouts":
            now = time.time()
            for w in [w for w in self.workers.values() if w.state == "ALIVE" and now - w.last_heartbeat > self.worker_timeout]:
                self._remove_worker(w, f"no heartbeat in {self.worker_timeout:.0f} s")
        elif op == "executor_state_changed":
            _, app_id, exec_id, state, message, exit_code = msg
            self._executor_state_changed(app_id, exec_id, state, message, exit_code)
        elif op == "unregister_app":
            self._finish_app(msg[1], "FINISHED")
        elif op == "complete_recovery":
            if self.status == "RECOVERING":
                self.status = "ALIVE"
            self.recovering_until = 0.0
            self._schedule()
        elif op == "elected":
            if self.status == "STANDBY":
                self._become_leader()
                if self.status == "ALIVE":                          # nothing persisted to wait for
                    self._schedule()
        elif op == "revoked":
            # another master holds the lease now: forget everything (it recovers the cluster from the workers / drivers)
            self.status = "STANDBY"
            self.workers.clear()
            self.apps.clear()
        else:
            raise RpcError(f"Master: unknown one-way message {op!r}")

    def receive_and_reply(self, msg, ctx: RpcCallContext) -> None:
        op = msg[0]
        if self.status == "STANDBY" and op in ("register_worker", "register_app", "reregister_app", "request_total_executors", "kill_executor"):
            ctx.reply(("standby", self.leader_url))
            return
        if op == "register_worker":
            _, wid, host, port, cores, memory_mb, ref = msg[:7]
            running = msg[7] if len(msg) > 7 else []
            old = self.workers.get(wid)
            if old is not None and old.state == "ALIVE":
                ctx.reply(("failed", f"duplicate worker id {wid}"))
                return
            self.workers[wid] = WorkerInfo(wid, host, port, int(cores), int(memory_mb), ref)
            ctx.reply(("registered", str(self.rpc_env.address)))
            for app_id, exec_id, slot, ecores, emem in running:       # executors that survived a master restart
                self._reported[(app_id, exec_id)] = (wid, slot, ecores, emem)
                app = self.apps.get(app_id)
                known = app.executors.get(exec_id) if app is not None else None
                if known is not None and known.state in ("LOST", "EXITED", "KILLED", "FAILED"):
                    # the worker dropped off, its executors were declared LOST and their slots re-filled elsewhere; now it is
                    # back with the old processes still alive: they are duplicates of a slot someone else owns
                    self._reported.pop((app_id, exec_id), None)
                    ref.send(("kill_executor", app_id, exec_id, False))
                elif app is not None:
                    self._adopt(app, exec_id)
                elif self.status == "ALIVE":                           # nobody owns it any more
                    ref.send(("kill_executor", app_id, exec_id, False))
            if running:
                self.recovered["workers"] += 1
            self._persist()
            self._schedule()
        elif op == "reregister_app":
            _, app_id, name, num_executors, cores_per_executor, memory_mb, driver_ref, driver_url = msg
            if app_id not in self.apps:
                self.apps[app_id] = AppInfo(app_id, name, driver_ref, driver_url, int(num_executors), int(cores_per_executor), int(memory_mb))
                self.recovered["apps"] += 1
            app = self.apps[app_id]
            app.driver = driver_ref
            for (a, exec_id) in [k for k in self._reported if k[0] == app_id]:
                self._adopt(app, exec_id)
            ctx.reply(("registered", app_id))
            self._persist()
            self._schedule()
        elif op == "request_total_executors":                          # dynamic allocation (ExecutorAllocationManager ->
            _, app_id, total = msg                                     #   StandaloneSchedulerBackend.doRequestTotalExecutors)
            app = self.apps.get(app_id)
            if app is None:
                ctx.reply(False)
                return
            app.num_executors = int(total)
            for ex in app.executors.values():                          # shrink: release the highest slots
                if ex.slot >= app.num_executors and ex.state in ("LAUNCHING", "RUNNING"):
                    w = self.workers.get(ex.worker_id)
                    if w is not None:
                        w.ref.send(("kill_executor", app_id, ex.exec_id, False))
            ctx.reply(True)
            self._schedule()
        elif op == "register_app":
            _, name, num_executors, cores_per_executor, memory_mb, driver_ref, driver_url = msg
            n = next(self._app_ids)
            self._next_app_id = n + 1
            app_id = f"app-{time.strftime('%Y%m%d%H%M%S')}-{n:04d}"
            if num_executors is None:                   # default: one executor on every alive worker
                num_executors = max(1, sum(1 for w in self.workers.values() if w.state == "ALIVE"))
            self.apps[app_id] = AppInfo(app_id, name, driver_ref, driver_url, int(num_executors),
                                        int(cores_per_executor), int(memory_mb))
            ctx.reply((app_id, int(num_executors)))
            self._persist()
            self._schedule()
        elif op == "state":
            ctx.reply(self.state())
        elif op == "kill_executor":                      # administrative / fault injection
            _, app_id, slot = msg
            app = self.apps.get(app_id)
            done = False
            if app is not None:
                for ex in app.executors.values():
                    if ex.slot == slot and ex.state in ("LAUNCHING", "RUNNING"):
                        w = self.workers.get(ex.worker_id)
                        if w is not None:
                            w.ref.send(("kill_executor", app_id, ex.exec_id, True))
                            done = True
            ctx.reply(done)
        else:
            ctx.send_failure(RpcError(f"Master: unknown request {op!r}"))

    def on_disconnected(self, address: RpcAddress) -> None:
        for w in [w for w in self.workers.values() if w.state == "ALIVE" and (w.host, w.port) == (address.host, address.port)]:
            self._remove_worker(w, "connection lost")
        for app in [a for a in self.apps.values() if a.driver.address == address]:
            self._finish_app(app.id, "FINISHED")       # a driver that goes away ends its application

    # ------------------------------------------------------------------ scheduling
    def _adopt(self, app: AppInfo, exec_id: int) -> None:
        """Take over an executor that a worker reports as running for ``app`` (master restart)."""
        rep = self._reported.pop((app.id, exec_id), None)
        if rep is None or exec_id in app.executors:
            return
        wid, slot, cores, mem = rep
        w = self.workers.get(wid)
        if w is None:
            return
        app.executors[exec_id] = ExecutorDesc(exec_id, slot, wid, cores, mem, "RUNNING")
        app.next_exec_id = max(app.next_exec_id, exec_id + 1)
        w.cores_used += cores
        w.memory_used += mem
        w.executors[f"{app.id}/{exec_id}"] = cores
        self.recovered["executors"] += 1

    def _schedule(self) -> None:
        """Fill every empty slot of every running app: spread-out = the alive worker with the most free cores first
        (``spark.deploy.spreadOut``), consolidate = pack the first worker that fits."""
        if time.time() < self.recovering_until:          # RECOVERING: wait for workers/drivers to report what is still alive
            return
        for app in sorted(self.apps.values(), key=lambda a: a.submit_time):          # FIFO across applications
            if app.state != "RUNNING":
                continue
            filled = {ex.slot for ex in app.executors.values() if ex.state in ("LAUNCHING", "RUNNING")}
            for slot in range(app.num_executors):
                if slot in filled:
                    continue
                usable = [w for w in self.workers.values() if w.state == "ALIVE"
                          and w.cores - w.cores_used >= app.cores_per_executor
                          and w.memory_mb - w.memory_used >= app.memory_per_executor]
                if not usable:
                    break
                if self.spread_out:
                    on_worker = {w.id: sum(1 for ex in app.executors.values()
                                           if ex.worker_id == w.id and ex.state in ("LAUNCHING", "RUNNING")) for w in usable}
                    w = min(usable, key=lambda w: (on_worker[w.id], -(w.cores - w.cores_used), w.id))
                else:
                    w = sorted(usable, key=lambda w: w.id)[0]
                self._launch(app, slot, w)

    def _launch(self, app: AppInfo, slot: int, w: WorkerInfo) -> None:
        exec_id = app.next_exec_id
        app.next_exec_id += 1
        ex = ExecutorDesc(exec_id, slot, w.id, app.cores_per_executor, app.memory_per_executor)
        app.executors[exec_id] = ex
        w.cores_used += ex.cores
        w.memory_used += ex.memory_mb
        w.executors[f"{app.id}/{exec_id}"] = ex.cores
        try:
            w.ref.send(("launch_executor", app.id, exec_id, slot, app.driver_url, ex.cores, ex.memory_mb))
            app.driver.send(("executor_added", exec_id, slot, w.id, f"{w.host}:{w.port}", ex.cores))
        except RpcError as exc:
            self._executor_state_changed(app.id, exec_id, "FAILED", f"cannot reach worker {w.id}: {exc}", None)

    def _release(self, app: AppInfo, ex: ExecutorDesc) -> None:
        w = self.workers.get(ex.worker_id)
        if w is not None and w.executors.pop(f"{app.id}/{ex.exec_id}", None) is not None:
            w.cores_used -= ex.cores
            w.memory_used -= ex.memory_mb

    def _executor_state_changed(self, app_id: str, exec_id: int, state: str, message: str, exit_code) -> None:
        app = self.apps.get(app_id)
        if app is None:
            return
        ex = app.executors.get(exec_id)
        if ex is None:
            return
        old, ex.state = ex.state, state
        if state == "RUNNING":
            app.retries = 0
        try:
            app.driver.send(("executor_updated", exec_id, ex.slot, state, message, exit_code))
        except RpcError:
            pass
        if state in ("EXITED", "KILLED", "FAILED", "LOST") and old not in ("EXITED", "KILLED", "FAILED", "LOST"):
            self._release(app, ex)
            normal = (state == "EXITED" and exit_code == 0) or (state == "KILLED" and message == "app finished") \
                or ex.slot >= app.num_executors
            if app.state == "RUNNING" and not normal:
                app.retries += 1
                if app.retries >= MAX_EXECUTOR_RETRIES and not any(e.state == "RUNNING" for e in app.executors.values()):
                    self._finish_app(app.id, "FAILED")        # Master.scala: "Application ... failed N times; removing it"
                    return
            self._schedule()

    def _remove_worker(self, w: WorkerInfo, reason: str) -> None:
        w.state = "DEAD"
        for key in list(w.executors):
            app_id, exec_id = key.rsplit("/", 1)
            self._executor_state_changed(app_id, int(exec_id), "LOST", f"worker {w.id} lost: {reason}", None)
        self.workers.pop(w.id, None)
        self._schedule()

    def _finish_app(self, app_id: str, state: str) -> None:
        app = self.apps.pop(app_id, None)
        if app is None:
            return
        app.state = state
        for ex in app.executors.values():
            if ex.state in ("LAUNCHING", "RUNNING"):
                w = self.workers.get(ex.worker_id)
                if w is not None:
                    try:
                        w.ref.send(("kill_executor", app_id, ex.exec_id, False))
                    except RpcError:
                        pass
                self._release(app, ex)
                ex.state = "KILLED"
        try:
            app.driver.send(("application_removed", state))
        except RpcError:
            pass
        self.completed_apps.append(app)
        del self.completed_apps[:-50]
        self._persist()
        self._schedule()

    # ------------------------------------------------------------------ status (MasterWebUI /json)
    def state(self) -> dict:
        def app_d(a: AppInfo) -> dict:
            return {"id": a.id, "name": a.name, "state": a.state, "submit_time": a.submit_time,
                    "num_executors": a.num_executors, "cores_per_executor": a.cores_per_executor,
                    "executors": [{"id": e.exec_id, "slot": e.slot, "worker": e.worker_id, "state": e.state, "cores": e.cores}
                                  for e in a.executors.values()]}
        return {"url": f"spark://{self.rpc_env.address}", "status": self.status, "recovered": dict(self.recovered), "uptime_s": time.time() - self.start_time,
                "workers": [{"id": w.id, "host": w.host, "port": w.port, "cores": w.cores, "cores_used": w.cores_used,
                             "memory_mb": w.memory_mb, "memory_used": w.memory_used, "state": w.state,
                             "last_heartbeat": w.last_heartbeat, "executors": sorted(w.executors)}
                            for w in self.workers.values()],
                "active_apps": [app_d(a) for a in self.apps.values()],
                "completed_apps": [app_d(a) for a in self.completed_apps]}


def start_master(host: str = "127.0.0.1", port: int = 7077, worker_timeout: float = 60.0, secret: Optional[str] = None,
                 spread_out: bool = True, recovery_dir: Optional[str] = None, recovery_grace: float = 2.0, encrypt: bool = False,
                 ha: bool = False, lease_s: float = 3.0):
    """Returns ``(rpc_env, master_endpoint)``; the URL is ``spark://<rpc_env.address>``."""
    env = RpcEnv("master", host, port, secret=secret, encrypt=encrypt)
    master = Master(worker_timeout, spread_out, recovery_dir, recovery_grace, ha, lease_s)
    env.setup_endpoint(Master.NAME, master)
    return env, master


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(description="asyncframework_b200 standalone master")
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=7077)
    ap.add_argument("--webui-port", type=int, default=0, help="serve the JSON status page on this port (0 = off)")
    ap.add_argument("--worker-timeout", type=float, default=60.0)
    ap.add_argument("--secret", default=None)
    ap.add_argument("--recovery-dir", default=None, help="persist registrations here and recover them after a restart")
    ap.add_argument("--encrypt", action="store_true", help="AES-GCM on every connection (needs --secret)")
    ap.add_argument("--ha", action="store_true", help="leader election through a lease file in --recovery-dir (run several masters; "
                    "give workers and applications spark://h1:p1,h2:p2)")
    a = ap.parse_args(argv)
    env, master = start_master(a.host, a.port, a.worker_timeout, a.secret, recovery_dir=a.recovery_dir, encrypt=a.encrypt, ha=a.ha)
    print(f"Master started at spark://{env.address}", flush=True)
    if a.webui_port:
        from ..utils.status_server import serve_json
        state = lambda: master.self_ref.ask_sync(("state",), 10)  # noqa: E731  (through the inbox: no torn reads)
        serve_json(a.host, a.webui_port, {"/json": state, "/": state})
    try:
        while True:
            time.sleep(3600)
    except KeyboardInterrupt:
        env.shutdown()


if __name__ == "__main__":
    main()
