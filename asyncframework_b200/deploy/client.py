"""Application side of the standalone deployment: the driver's scheduler backend + an in-process test cluster.

``StandaloneTransport`` is what ``StandaloneSchedulerBackend`` + ``StandaloneAppClient`` + the ``DriverEndpoint`` of
``CoarseGrainedSchedulerBackend`` are in the reference (core/scheduler/cluster/StandaloneSchedulerBackend.scala:37,
core/deploy/client/StandaloneAppClient.scala, CoarseGrainedSchedulerBackend.scala:121,239): it registers the
application with the Master, receives ``ExecutorAdded`` / ``ExecutorUpdated``, accepts ``RegisterExecutor`` from the
executors the workers fork, routes their ``StatusUpdate`` messages into :class:`ClusterScheduler` and turns a dropped
executor connection into ``_executor_lost`` — after which the Master re-fills the slot and the scheduler re-launches
the lost tasks on the replacement.

``LocalStandaloneCluster`` = ``LocalSparkCluster`` (core/deploy/LocalSparkCluster.scala:49-70): a Master and ``n``
Workers inside this process (each with its own RpcEnv), forking real executor processes — the multi-process,
real-TCP deployment on one box.
"""
from __future__ import annotations

import os
import signal
import threading
import time
from typing import Dict, List, Optional

from ..rpc import RpcAddress, RpcCallContext, RpcEndpoint, RpcEndpointRef, RpcError
from .master import start_master
from .worker import start_worker


class _DriverEndpoint(RpcEndpoint):
    NAME = "CoarseGrainedScheduler"

    def __init__(self, transport: "StandaloneTransport"):
        self.t = transport

    def receive(self, msg, sender) -> None:
        op = msg[0]
        if op == "status":
            _, slot, exec_id, status = msg
            if self.t.exec_ids.get(slot) == exec_id:         # ignore stragglers from a replaced executor
                self.t.sched._on_message(slot, status)
        elif op == "executor_added":
            self.t.events.append(msg)
        elif op == "executor_updated":
            self.t.events.append(msg)
            _, exec_id, slot, state, message, exit_code = msg
            if state in ("EXITED", "KILLED", "FAILED", "LOST"):
                self.t._slot_down(slot, exec_id, f"{state}: {message}")
        elif op == "application_removed":
            self.t.app_state = msg[1]
            if msg[1] == "FAILED":
                self.t._app_failed()
        else:
            raise RpcError(f"driver endpoint: unknown message {op!r}")

    def receive_and_reply(self, msg, ctx: RpcCallContext) -> None:
        if msg[0] == "register_executor":
            _, slot, exec_id, ref, cores, pid, host = msg
            if slot < 0:
                ctx.reply(("refused", f"slot {slot} out of range"))
                return
            with self.t.cv:
                self.t.refs[slot] = ref
                self.t.exec_ids[slot] = exec_id
                self.t.pids[slot] = (host, pid)
                self.t.addresses[slot] = ref.address
                self.t.cv.notify_all()
            ctx.reply(("registered", self.t.sched.env.limits))
        else:
            ctx.send_failure(RpcError(f"driver endpoint: unknown request {msg[0]!r}"))

    def on_disconnected(self, address: RpcAddress) -> None:
        for slot, addr in list(self.t.addresses.items()):
            if addr == address:
                self.t._slot_down(slot, self.t.exec_ids.get(slot), "connection lost")
        if address == self.t.master_address and self.t.sched is not None and not self.t.sched._stopped:
            threading.Thread(target=self.t._reregister_loop, daemon=True, name="app-client-reregister").start()


class StandaloneTransport:
    """Plug into ``ClusterScheduler(transport=...)``.  ``num_executors=None`` asks the master for one executor per
    alive worker (the standalone default: an application takes every worker)."""

    def __init__(self, master_url: str, num_executors: Optional[int] = None, cores_per_executor: int = 1,
                 memory_mb: int = 512, app_name: str = "asyncframework_b200", register_timeout: float = 60.0):
        from ..rpc import parse_master_urls
        self.master_addresses = parse_master_urls(master_url)
        self.master_address = self.master_addresses[0]
        self.num_executors = num_executors
        self.cores_per_executor, self.memory_mb, self.app_name = cores_per_executor, memory_mb, app_name
        self.register_timeout = register_timeout
        self.refs: Dict[int, RpcEndpointRef] = {}
        self.exec_ids: Dict[int, int] = {}
        self.pids: Dict[int, tuple] = {}
        self.addresses: Dict[int, RpcAddress] = {}
        self.cv = threading.Condition()
        self.events: List[tuple] = []
        self.app_id: Optional[str] = None
        self.app_state = "RUNNING"
        self.sched = None
        self.master: Optional[RpcEndpointRef] = None
        self.master_reconnects = 0
        self._reregistering = False

    # ---- ClusterScheduler transport interface -------------------------------------------------------------------------
    def start(self, sched) -> None:
        if sched.env is None:
            raise RuntimeError("the standalone deployment needs the scheduler's AsyncEnv (use_env=True)")
        self.sched = sched
        env = sched.env.rpc_env
        env.setup_endpoint(_DriverEndpoint.NAME, _DriverEndpoint(self))
        me = env.endpoint_ref(None, _DriverEndpoint.NAME)
        reply, last_err = None, None
        deadline = time.time() + self.register_timeout
        while reply is None:
            for addr in self.master_addresses:                       # HA: the first master that is not a standby
                try:
                    ref = env.endpoint_ref(addr, "Master")
                    r = ref.ask_sync(("register_app", self.app_name, self.num_executors, self.cores_per_executor, self.memory_mb, me,
                                      str(env.address)), 30)
                    if r[0] == "standby":
                        last_err = RpcError(f"{addr} is a standby (leader: {r[1]})")
                        continue
                    reply, self.master, self.master_address = r, ref, addr
                    break
                except RpcError as exc:
                    last_err = exc
            if reply is None:
                if len(self.master_addresses) == 1 or time.time() > deadline:
                    raise last_err
                time.sleep(0.3)
        self.app_id, granted = reply
        self.num_executors = granted
        sched.num_executors = granted
        deadline = time.time() + self.register_timeout
        with self.cv:                       # spark.scheduler.minRegisteredResourcesRatio = 1.0: wait for every slot
            while len(self.refs) < granted:
                left = deadline - time.time()
                if left <= 0 or self.app_state != "RUNNING":
                    raise RuntimeError(f"only {len(self.refs)} of {granted} executors registered with the driver within "
                                       f"{self.register_timeout:.0f} s (master {self.master_address}, app {self.app_id}, "
                                       f"state {self.app_state}) — are enough workers with free cores alive?")
                self.cv.wait(min(left, 0.5))

    def send(self, eid: int, msg: tuple) -> None:
        ref = self.refs.get(eid)
        if ref is None and self.sched is not None and self.sched.recover_lost_executors and not self.sched._stopped:
            self.relaunch(eid)                 # the Master is re-filling the slot: wait for the replacement to register
            ref = self.refs.get(eid)
        if ref is None:
            raise OSError(f"executor slot {eid} is down")
        ref.send(msg)

    def _slot_down(self, slot: int, exec_id, why: str) -> None:
        with self.cv:
            if self.exec_ids.get(slot) != exec_id or slot not in self.refs:
                return                                   # already handled (both the disconnect and the master report arrive)
            del self.refs[slot]
            self.addresses.pop(slot, None)
        if self.sched is not None and not self.sched._stopped:
            # off the inbox thread: recovery waits for the replacement's RegisterExecutor, which this inbox must serve
            threading.Thread(target=self.sched._executor_lost, args=(slot,), daemon=True,
                             name=f"executor-lost-{slot}").start()

    def _reregister_loop(self) -> None:
        """The master went away (restart / failover to a standby on the same address): keep the executors running and
        re-register the application under its existing id until a master answers (StandaloneAppClient ``MasterChanged``)."""
        with self.cv:
            if self._reregistering:
                return
            self._reregistering = True
        try:
            self._reregister()
        finally:
            self._reregistering = False

    def _reregister(self) -> None:
        env = self.sched.env.rpc_env
        me = env.endpoint_ref(None, _DriverEndpoint.NAME)
        delay = 0.2
        while not self.sched._stopped and self.app_state == "RUNNING":
            for addr in self.master_addresses:
                try:
                    ref = env.endpoint_ref(addr, "Master")
                    reply = ref.ask_sync(("reregister_app", self.app_id, self.app_name, self.num_executors, self.cores_per_executor,
                                          self.memory_mb, me, str(env.address)), 5)
                    if reply[0] == "registered":
                        self.master, self.master_address = ref, addr
                        self.master_reconnects += 1
                        return
                except RpcError:
                    pass
            time.sleep(delay)
            delay = min(delay * 2, 5.0)

    def request_total_executors(self, total: int, timeout: Optional[float] = None) -> bool:
        """Dynamic allocation: grow or shrink the application to ``total`` executor slots.  Partition p lives on slot
        p % total afterwards, so only call this between jobs (cached blocks of re-homed partitions are recomputed)."""
        with self.sched._lock:
            if self.sched._jobs:
                raise RuntimeError("request_total_executors while jobs are running would re-home their partitions")
        old = self.num_executors
        if not self.master.ask_sync(("request_total_executors", self.app_id, int(total)), 10):
            return False
        if total < old:
            with self.cv:
                for slot in range(total, old):
                    self.refs.pop(slot, None)
                    self.addresses.pop(slot, None)
                    self.exec_ids[slot] = -1              # late status / disconnect events of the released executors are stale
        self.num_executors = total
        deadline = time.time() + (timeout or self.register_timeout)
        with self.cv:
            while any(slot not in self.refs for slot in range(total)):
                left = deadline - time.time()
                if left <= 0:
                    return False
                self.cv.wait(min(left, 0.5))
        self.sched.num_executors = total
        with self.sched._lock:
            self.sched._blocks_held.clear()               # the partition -> slot map changed: re-ship source blocks on demand
        return True

    def relaunch(self, eid: int) -> bool:
        """The Master re-fills the slot on its own; wait for the replacement to register."""
        deadline = time.time() + self.register_timeout
        with self.cv:
            while eid not in self.refs:
                left = deadline - time.time()
                if left <= 0 or self.app_state != "RUNNING" or self.sched._stopped:
                    return False
                self.cv.wait(min(left, 0.5))
        return True

    def _app_failed(self) -> None:
        with self.cv:
            self.cv.notify_all()

    def kill(self, eid: int) -> None:
        """Fault injection: SIGKILL the executor process of a slot (same host) or ask the master to do it."""
        host, pid = self.pids.get(eid, (None, None))
        if pid is not None and host in ("127.0.0.1", "localhost"):
            try:
                os.kill(pid, signal.SIGKILL)
                return
            except ProcessLookupError:
                pass
        self.master.ask_sync(("kill_executor", self.app_id, eid), 10)

    def stop(self) -> None:
        # unregister FIRST: an executor that exits because of "stop" while the application is still registered would look
        # like a lost executor to the Master and be re-launched
        try:
            if self.master is not None and self.app_id is not None:
                self.master.send(("unregister_app", self.app_id))
        except RpcError:
            pass
        for ref in list(self.refs.values()):
            try:
                ref.send(("stop",))
            except (RpcError, OSError):
                pass
        time.sleep(0.05)                     # let the one-way messages leave before the env closes its sockets


class LocalStandaloneCluster:
    """Master + ``num_workers`` Workers in this process, real executor processes, real TCP."""

    def __init__(self, num_workers: int = 2, cores_per_worker: int = 1, memory_per_worker: int = 1024,
                 worker_timeout: float = 60.0, heartbeat_interval: float = 1.0, secret: Optional[str] = None,
                 recovery_dir: Optional[str] = None, encrypt: bool = False, gpus: Optional[List[int]] = None):
        """``gpus=[0, 1, ...]`` pins worker i's executors to CUDA device ``gpus[i]`` (one executor per GPU on a B200 node)."""
        self._master_args = dict(worker_timeout=worker_timeout, secret=secret, recovery_dir=recovery_dir, recovery_grace=1.5,
                                 encrypt=encrypt)
        self.master_env, self.master = start_master("127.0.0.1", 0, **self._master_args)
        self.master_url = f"spark://{self.master_env.address}"
        self.worker_envs, self.workers = [], []
        for i in range(num_workers):
            e, w = start_worker(self.master_url, cores_per_worker, memory_per_worker, heartbeat_interval=heartbeat_interval,
                                secret=secret, worker_id=f"worker-{i}", encrypt=encrypt,
                                gpu=(gpus[i] if gpus is not None and i < len(gpus) else None))
            self.worker_envs.append(e)
            self.workers.append(w)

    def add_worker(self, cores: int = 1, memory_mb: int = 1024, **kw):
        e, w = start_worker(self.master_url, cores, memory_mb, worker_id=f"worker-{len(self.workers)}", **kw)
        self.worker_envs.append(e)
        self.workers.append(w)
        return w

    def stop_worker(self, i: int) -> None:
        self.worker_envs[i].shutdown()

    def restart_master(self, down_for: float = 0.3) -> None:
        """Kill the master and bring a new one up on the SAME address (FaultToleranceTest's "kill leader")."""
        addr = self.master_env.address
        self.master_env.shutdown()
        time.sleep(down_for)
        self.master_env, self.master = start_master(addr.host, addr.port, **self._master_args)

    def state(self) -> dict:
        return self.master_env.endpoint_ref(None, "Master").ask_sync(("state",), 10)

    def stop(self) -> None:
        for e in self.worker_envs:
            e.shutdown()
        self.master_env.shutdown()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.stop()
