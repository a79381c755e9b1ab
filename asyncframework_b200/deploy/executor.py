"""Executor process of a standalone deployment (``CoarseGrainedExecutorBackend``,
core/executor/CoarseGrainedExecutorBackend.scala:92,143 + ``Executor.TaskRunner``, core/executor/Executor.scala:290-560).

Started by a Worker's ExecutorRunner.  It builds the per-process :class:`AsyncEnv` (RpcEnv + BlockManager registered
with the driver's BlockManagerMaster), registers with the driver's ``CoarseGrainedScheduler`` endpoint
(``RegisterExecutor``), then feeds the driver's ``binary`` / ``task`` / ``drop`` / ``stop`` messages to the
transport-agnostic :class:`~asyncframework_b200.core.cluster.ExecutorBackend`; status messages (direct result,
indirect-result block reference, failure) go back as one-way ``StatusUpdate`` messages.  If the driver's connection
drops the process exits (``Driver … disassociated! Shutting down``).
"""
from __future__ import annotations

import argparse
import os
import sys


def main(argv=None) -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--driver-url", required=True)
    ap.add_argument("--executor-id", type=int, required=True)
    ap.add_argument("--slot", type=int, required=True)
    ap.add_argument("--cores", type=int, default=1)
    ap.add_argument("--app-id", default="")
    ap.add_argument("--worker-url", default="")
    ap.add_argument("--host", default="127.0.0.1")
    ap.add_argument("--secret", default=None)
    ap.add_argument("--encrypt", action="store_true")
    a = ap.parse_args(argv)

    import torch
    torch.set_num_threads(1)          # parallelism comes from a.cores concurrent task threads, not from intra-op threads
    from ..core.cluster import ExecutorBackend
    from ..core.env import AsyncEnv, RuntimeLimits
    from ..rpc import RpcAddress, RpcEndpoint, RpcError

    driver_addr = RpcAddress.parse(a.driver_url)
    # the driver's limits arrive with the registration reply; start with defaults sized by the worker's grant
    limits = RuntimeLimits(auth_secret=a.secret, encrypt=a.encrypt)
    mem_mb = int(os.environ.get("AFB_EXECUTOR_MEMORY_MB", "0") or 0)
    if mem_mb:
        limits.storage_memory = int(mem_mb * (1 << 20) * 0.6)          # spark.memory.fraction of the executor's grant
    env = AsyncEnv.create_executor(str(a.slot), driver_addr, limits, host=a.host)
    driver = env.rpc_env.endpoint_ref(driver_addr, "CoarseGrainedScheduler")
    backend = ExecutorBackend(a.slot, lambda msg: driver.send(("status", a.slot, a.executor_id, msg)), a.cores)

    class ExecutorEndpoint(RpcEndpoint):
        def receive(self, msg, sender):
            if not backend.handle(msg):
                return

        def on_disconnected(self, address):
            if address == driver_addr:
                backend.close()

    env.rpc_env.setup_endpoint("Executor", ExecutorEndpoint())
    me = env.rpc_env.endpoint_ref(None, "Executor")
    try:
        reply = driver.ask_sync(("register_executor", a.slot, a.executor_id, me, a.cores, os.getpid(), a.host), 30)
    except RpcError as exc:
        print(f"cannot register with the driver at {driver_addr}: {exc}", file=sys.stderr, flush=True)
        env.stop()
        return 1
    if reply[0] != "registered":
        print(f"driver refused the registration: {reply[1]}", file=sys.stderr, flush=True)
        env.stop()
        return 1
    driver_limits = reply[1]
    if driver_limits is not None:                       # result / broadcast limits are the application's, not the worker's
        for k in ("max_direct_result_size", "max_result_size", "broadcast_block_size", "broadcast_compress",
                  "broadcast_checksum", "broadcast_inline_threshold", "network_timeout", "heartbeat_interval",
                  "shuffle_memory", "shuffle_spill_elements", "shuffle_spill_batch", "shuffle_spill_compress"):
            setattr(env.limits, k, getattr(driver_limits, k))
    backend.start_heartbeats(env.limits.heartbeat_interval)
    try:
        backend.run_tasks()
    finally:
        env.stop()
    return 0


if __name__ == "__main__":
    sys.exit(main())
