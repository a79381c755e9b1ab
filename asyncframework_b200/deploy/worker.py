"""Standalone cluster manager — the Worker daemon and its ExecutorRunner.

Reference: ``Worker`` (core/deploy/worker/Worker.scala:43) registers with the master (retrying), heartbeats every
``spark.worker.timeout / 4``, launches executors on ``LaunchExecutor`` through an ``ExecutorRunner``
(core/deploy/worker/ExecutorRunner.scala:39 — forks the executor JVM with the app's command line, redirects
stdout/stderr into ``work/<app>/<exec>/``, waits for exit and reports ``ExecutorStateChanged``), kills them on
``KillExecutor``, and cleans up everything when it stops or loses the master for good.

Run it: ``python -m asyncframework_b200.deploy.worker spark://master-host:7077 --cores 8 --memory 16384``.
Each executor is ``python -m asyncframework_b200.deploy.executor …`` — one OS process; on a GPU node give every
worker ``--gpu <index>`` so its executors see exactly that device (``CUDA_VISIBLE_DEVICES``), which reproduces the
"one executor per GPU" layout of the fused backend for closure-API jobs.
"""
from __future__ import annotations

import argparse
import os
import signal
import subprocess
import sys
import tempfile
import threading
import time
import uuid
from typing import Dict, Optional

from ..rpc import RpcAddress, RpcCallContext, RpcEndpoint, RpcEnv, RpcError


class ExecutorRunner:
    def __init__(self, worker: "Worker", app_id: str, exec_id: int, slot: int, driver_url: str, cores: int, memory_mb: int):
        self.worker, self.app_id, self.exec_id, self.slot = worker, app_id, exec_id, slot
        self.driver_url, self.cores, self.memory_mb = driver_url, cores, memory_mb
        self.dir = os.path.join(worker.work_dir, app_id, str(exec_id))
        self.proc: Optional[subprocess.Popen] = None
        self.killed: Optional[str] = None

    def start(self) -> None:
        os.makedirs(self.dir, exist_ok=True)
        cmd = [sys.executable, "-m", "asyncframework_b200.deploy.executor", "--driver-url", self.driver_url,
               "--executor-id", str(self.exec_id), "--slot", str(self.slot), "--cores", str(self.cores),
               "--app-id", self.app_id, "--worker-url", str(self.worker.rpc_env.address), "--host", self.worker.host]
        if self.worker.encrypt:
            cmd += ["--encrypt"]
        env = dict(os.environ)
        if self.worker.secret:
            env["AFB_RPC_SECRET"] = self.worker.secret           # never in argv: `ps` shows command lines to every local user
        root = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
        env["PYTHONPATH"] = root + (os.pathsep + env["PYTHONPATH"] if env.get("PYTHONPATH") else "")
        env["AFB_EXECUTOR_MEMORY_MB"] = str(self.memory_mb)
        env["OMP_NUM_THREADS"] = str(max(1, self.cores))
        if self.worker.gpu is not None:
            env["CUDA_VISIBLE_DEVICES"] = str(self.worker.gpu)
        self.stdout = open(os.path.join(self.dir, "stdout"), "wb")
        self.stderr = open(os.path.join(self.dir, "stderr"), "wb")
        try:
            self.proc = subprocess.Popen(cmd, stdout=self.stdout, stderr=self.stderr, env=env, cwd=self.dir,
                                         start_new_session=True)
        except OSError as exc:
            self.worker.report(self, "FAILED", f"cannot launch: {exc}", None)
            return
        threading.Thread(target=self._wait, daemon=True, name=f"executor-runner-{self.app_id}-{self.exec_id}").start()
        self.worker.report(self, "RUNNING", "", None)

    def _wait(self) -> None:
        code = self.proc.wait()
        for f in (self.stdout, self.stderr):
            try:
                f.close()
            except OSError:
                pass
        if self.killed is not None:
            self.worker.report(self, "KILLED", self.killed, code)
        else:
            tail = ""
            try:
                with open(os.path.join(self.dir, "stderr"), "rb") as f:
                    tail = f.read()[-400:].decode(errors="replace")
            except OSError:
                pass
            self.worker.report(self, "EXITED", f"exit code {code}" + (f": {tail.strip().splitlines()[-1]}" if tail.strip() else ""), code)

    def kill(self, reason: str = "app finished", hard: bool = False) -> None:
        if self.proc is None or self.proc.poll() is not None:
            return
        if not hard:
            self.killed = reason
        try:
            os.killpg(self.proc.pid, signal.SIGKILL if hard else signal.SIGTERM)
        except (ProcessLookupError, PermissionError):
            pass
        if not hard:
            def reap():
                try:
                    self.proc.wait(5)
                except subprocess.TimeoutExpired:
                    try:
                        os.killpg(self.proc.pid, signal.SIGKILL)
                    except (ProcessLookupError, PermissionError):
                        pass
            threading.Thread(target=reap, daemon=True).start()


class Worker(RpcEndpoint):
    NAME = "Worker"

    def __init__(self, master_url: str, cores: int, memory_mb: int, work_dir: Optional[str] = None,
                 heartbeat_interval: float = 15.0, host: str = "127.0.0.1", gpu: Optional[int] = None,
                 secret: Optional[str] = None, worker_id: Optional[str] = None, encrypt: bool = False):
        self.encrypt = encrypt
        from ..rpc import parse_master_urls
        self.master_addresses = parse_master_urls(master_url)      # spark://h1:p1,h2:p2 — walked until one is the leader
        self.master_address = self.master_addresses[0]
        self.cores, self.memory_mb = int(cores), int(memory_mb)
        self.host, self.gpu, self.secret = host, gpu, secret
        self.heartbeat_interval = heartbeat_interval
        self._own_work_dir = work_dir is None        # a temp dir we made up is removed when the worker stops
        self.work_dir = work_dir or os.path.join(tempfile.gettempdir(), f"afb-worker-{uuid.uuid4().hex[:8]}")
        self.id = worker_id or f"worker-{time.strftime('%Y%m%d%H%M%S')}-{host}-{uuid.uuid4().hex[:6]}"
        self.master = None
        self.registered = threading.Event()
        self.executors: Dict[str, ExecutorRunner] = {}
        self.finished: Dict[str, str] = {}
        self._stop = threading.Event()

    def on_start(self) -> None:
        os.makedirs(self.work_dir, exist_ok=True)
        threading.Thread(target=self._register_loop, daemon=True, name=f"{self.id}-register").start()

    def _register_loop(self) -> None:
        """Register (retrying with back-off, Worker.scala ``registerWithMaster``), then heartbeat; re-register when the
        master asks for it or the connection drops."""
        delay = 0.2
        while not self._stop.is_set():
            if not self.registered.is_set():
                try:
                    self.master = self.rpc_env.endpoint_ref(self.master_address, "Master")
                    me = self.self_ref                  # pickles with this env's listening address
                    running = [(r.app_id, r.exec_id, r.slot, r.cores, r.memory_mb) for r in list(self.executors.values())
                               if r.proc is not None and r.proc.poll() is None]
                    reply = self.master.ask_sync(("register_worker", self.id, self.rpc_env.address.host, self.rpc_env.address.port,
                                                  self.cores, self.memory_mb, me, running), 10)
                    if reply[0] == "registered":
                        self.registered.set()
                        delay = 0.2
                    else:
                        raise RpcError(str(reply[1]))
                except RpcError:
                    # dead, unreachable or a standby: try the next master of the list (with back-off once the list is exhausted)
                    i = self.master_addresses.index(self.master_address)
                    self.master_address = self.master_addresses[(i + 1) % len(self.master_addresses)]
                    if (i + 1) % len(self.master_addresses) == 0 or len(self.master_addresses) == 1:
                        self._stop.wait(delay)
                        delay = min(delay * 2, 10.0)
                    continue
            try:
                self.master.send(("heartbeat", self.id))
            except RpcError:
                self.registered.clear()
                continue
            self._stop.wait(self.heartbeat_interval)

    def receive(self, msg, sender) -> None:
        op = msg[0]
        if op == "launch_executor":
            _, app_id, exec_id, slot, driver_url, cores, memory_mb = msg
            runner = ExecutorRunner(self, app_id, exec_id, slot, driver_url, cores, memory_mb)
            self.executors[f"{app_id}/{exec_id}"] = runner
            runner.start()
        elif op == "kill_executor":
            _, app_id, exec_id, hard = msg
            runner = self.executors.get(f"{app_id}/{exec_id}")
            if runner is not None:
                runner.kill(hard=hard)
        elif op == "reconnect":
            self.registered.clear()
        else:
            raise RpcError(f"Worker: unknown message {op!r}")

    def receive_and_reply(self, msg, ctx: RpcCallContext) -> None:
        if msg[0] == "state":
            ctx.reply({"id": self.id, "cores": self.cores, "memory_mb": self.memory_mb, "work_dir": self.work_dir,
                       "registered": self.registered.is_set(),
                       "executors": {k: ("RUNNING" if r.proc is not None and r.proc.poll() is None else "DONE")
                                     for k, r in self.executors.items()}, "finished": dict(self.finished)})
        else:
            ctx.send_failure(RpcError(f"Worker: unknown request {msg[0]!r}"))

    def report(self, runner: ExecutorRunner, state: str, message: str, exit_code) -> None:
        key = f"{runner.app_id}/{runner.exec_id}"
        if state in ("EXITED", "KILLED", "FAILED"):
            self.executors.pop(key, None)
            self.finished[key] = state
        try:
            self.master.send(("executor_state_changed", runner.app_id, runner.exec_id, state, message, exit_code))
        except (RpcError, AttributeError):
            pass

    def on_disconnected(self, address: RpcAddress) -> None:
        if address == self.master_address:
            self.registered.clear()            # the register loop dials again; executors keep running meanwhile

    def on_stop(self) -> None:
        self._stop.set()
        for r in list(self.executors.values()):
            r.kill("worker stopped")
        if self._own_work_dir:
            import shutil

            def cleanup():                              # after the executors had a moment to exit and close their log files
                time.sleep(1.0)
                shutil.rmtree(self.work_dir, ignore_errors=True)
            threading.Thread(target=cleanup, daemon=False, name=f"{self.id}-cleanup").start()


def start_worker(master_url: str, cores: int = 1, memory_mb: int = 1024, host: str = "127.0.0.1", port: int = 0,
                 wait_registered: float = 10.0, **kw):
    """Returns ``(rpc_env, worker_endpoint)``."""
    env = RpcEnv("worker", host, port, secret=kw.get("secret"), encrypt=kw.get("encrypt", False))
    worker = Worker(master_url, cores, memory_mb, host=host, **kw)
    env.setup_endpoint(Worker.NAME, worker)
    if wait_registered:
        worker.registered.wait(wait_registered)
    return env, worker


def main(argv=None) -> None:
    ap = argparse.ArgumentParser(description="asyncframework_b200 standalone worker")
    ap.add_argument("master", help="spark://host:port")
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--port", type=int, default=0)
    ap.add_argument("--cores", type=int, default=os.cpu_count() or 1)
    ap.add_argument("--memory", type=int, default=1024, help="MB offered to executors")
    ap.add_argument("--work-dir", default=None)
    ap.add_argument("--gpu", type=int, default=None, help="CUDA device index given to this worker's executors")
    ap.add_argument("--secret", default=None)
    ap.add_argument("--encrypt", action="store_true")
    a = ap.parse_args(argv)
    env, worker = start_worker(a.master, a.cores, a.memory, a.host, a.port, work_dir=a.work_dir, gpu=a.gpu, secret=a.secret,
                               encrypt=a.encrypt)
    print(f"Worker {worker.id} started at {env.address}, registered={worker.registered.is_set()}", flush=True)
    try:
        while True:
            time.sleep(3600)
    except KeyboardInterrupt:
        env.shutdown()


if __name__ == "__main__":
    main()
